# Builds deftet_b200/libdeftet_b200.so (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude -Ideftet_b200/csrc --expt-relaxed-constexpr
SRC := $(wildcard deftet_b200/csrc/*.cu)
OBJ := $(patsubst deftet_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := deftet_b200/libdeftet_b200.so

SHIMS := tet_point_adj tet_adj_share tet_face_adj colaps_v
SHIM_SO := $(foreach n,$(SHIMS),deftet_b200/dropin/utils/lib/$(n)/run.so)

all: $(LIB) shims oracle

shims: $(SHIM_SO)

deftet_b200/dropin/utils/lib/%/run.so: deftet_b200/csrc/shims/run_shim.c $(LIB)
	gcc -O2 -fPIC -shared -Iinclude -DSHIM_$(shell echo $* | tr a-z A-Z) -o $@ $< -Ldeftet_b200 -ldeftet_b200 -Wl,-rpath,'$$ORIGIN/../../../..'


$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

build/%.o: deftet_b200/csrc/%.cu deftet_b200/csrc/*.cuh include/deftet_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean shims

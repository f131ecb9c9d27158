# Builds deftet_b200/libdeftet_b200.so (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude -Ideftet_b200/csrc --expt-relaxed-constexpr
SRC := $(wildcard deftet_b200/csrc/*.cu)
OBJ := $(patsubst deftet_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := deftet_b200/libdeftet_b200.so

all: $(LIB) oracle

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

.PHONY: all oracle clean

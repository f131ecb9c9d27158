#!/bin/sh
# Dev tool: build A/B variants of ONE kernel file with different -D flags (same ABI), into build/variants/lib_<name>.so.
#   sh tools/ab_variants.sh tridist "base:@HEAD" "t64:" "t128:-DPFD_THREADS_N=128 -DPFD_MIN_CTAS=6"
# "@HEAD" compiles the committed version of the file instead of the working copy.
set -e
FILE=$1; shift
NVFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -Ideftet_b200/csrc --expt-relaxed-constexpr"
mkdir -p build/variants
OTHERS=$(ls build/*.o | grep -v "build/$FILE.o")
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  src=deftet_b200/csrc/$FILE.cu
  if [ "$flags" = "@HEAD" ]; then git show HEAD:$src > build/variants/${FILE}_head.cu; src=build/variants/${FILE}_head.cu; flags=""; fi
  nvcc $NVFLAGS $flags -c $src -o build/variants/${FILE}_$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/lib_$name.so build/variants/${FILE}_$name.o $OTHERS -lcudart
  echo "built build/variants/lib_$name.so ($flags)"
done

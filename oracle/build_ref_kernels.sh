#!/bin/sh
# Compile the reference's own __host__ __device__ kernel helper functions FOR THE HOST (test infrastructure).
# The CUDA extensions of the reference cannot run here (no GPU) and do not build against torch >= 1.11, but their per-element
# math lives in `template<typename scalar_t> __host__ __device__` functions.  This script cuts each reference .cu at its
# ATen-dependent host launcher (into a temp dir, nothing from the reference is copied into the repository), appends a small
# extern "C" driver (oracle/ref_kernels_driver_*.inc, ours) that replays the kernel's control flow on the CPU, and builds
# oracle/_ref/kernels/<name>.so with nvcc.  tests/test_golden.py checks oracle/deftet_oracle.c against these bit for bit.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref/kernels"
[ -d "$REF/layers/DefTet" ] || { echo "build_ref_kernels: $REF not mounted; keeping prebuilt oracle/_ref/kernels"; exit 0; }
command -v nvcc >/dev/null 2>&1 || { echo "build_ref_kernels: nvcc not found"; exit 0; }
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
cut_body() {  # $1 = reference .cu, $2 = first line of the host launcher (regex), $3 = output
    awk -v pat="$2" '$0 ~ pat {exit} {print}' "$1" | grep -v '#include <ATen' | grep -v '#include <THC' > "$3"
}
build() {  # $1 = name, $2 = body file, $3 = driver
    { echo '#include <cuda_runtime.h>'; echo '#include <math.h>'; echo '#include <stdint.h>'; cat "$2"; cat "$3"; } > "$TMP/$1.cu"
    nvcc -O2 -w -std=c++14 -Xcompiler -fPIC -Xcompiler -ffp-contract=off -shared -o "$OUT/$1.so" "$TMP/$1.cu"
}
cut_body "$REF/layers/DefTet/check_condition_tetrahedron_base/check_condition_tet_for.cu" '^void dr_cuda_forward_batch' "$TMP/a1_body.cuh"
build point_in_tet "$TMP/a1_body.cuh" "$HERE/ref_kernels_driver_a1.inc"
cut_body "$REF/layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_for.cu" '^void dr_cuda_forward_batch' "$TMP/a4f_body.cuh"
build face_distance_fwd "$TMP/a4f_body.cuh" "$HERE/ref_kernels_driver_a4f.inc"
cut_body "$REF/layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_back.cu" '^void dr_cuda_backward_batch' "$TMP/a4b_body.cuh"
build face_distance_bwd "$TMP/a4b_body.cuh" "$HERE/ref_kernels_driver_a4b.inc"
cut_body "$REF/layers/DefTet/tet_face_adj_m_idx/tet_face_adj_m_for.cu" '^void dr_cuda_forward_batch' "$TMP/a5_body.cuh"
build face_adj "$TMP/a5_body.cuh" "$HERE/ref_kernels_driver_a5.inc"
echo "build_ref_kernels: built $(ls "$OUT" | tr '\n' ' ')"

# ---- the same reference kernels compiled FOR THE DEVICE (sm_100a): the reference's own __global__ kernels, unmodified, behind
# small extern "C" launchers (oracle/ref_kernels_cuda_driver_*.inc, ours) that take raw device pointers instead of at::Tensor.
# This is the "reference's own CUDA extensions" arm of the measurement: default nvcc code generation
# (FMA contraction on, like the reference's torch cpp_extension build), the reference's launch geometry.  Runs only on the GPU box.
OUTD="$HERE/_ref/kernels_cuda"
mkdir -p "$OUTD"
build_dev() {  # $1 = name, $2 = body file, $3 = driver
    { echo '#include <cuda_runtime.h>'; echo '#include <math.h>'; echo '#include <stdint.h>'; echo '#include <stdio.h>'; cat "$2"; cat "$3"; } > "$TMP/$1_dev.cu"
    nvcc -O3 -w -std=c++14 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o "$OUTD/$1.so" "$TMP/$1_dev.cu" -lcudart
}
build_dev point_in_tet "$TMP/a1_body.cuh" "$HERE/ref_kernels_cuda_driver_a1.inc"
build_dev face_distance_fwd "$TMP/a4f_body.cuh" "$HERE/ref_kernels_cuda_driver_a4f.inc"
build_dev face_distance_bwd "$TMP/a4b_body.cuh" "$HERE/ref_kernels_cuda_driver_a4b.inc"
build_dev face_adj "$TMP/a5_body.cuh" "$HERE/ref_kernels_cuda_driver_a5.inc"
grep -v '#include <ATen' "$REF/layers/nearest_neighbor/nearest_neighbor_cuda.cu" > "$TMP/a2_body.cuh"
build_dev nearest_neighbor "$TMP/a2_body.cuh" "$HERE/ref_kernels_cuda_driver_a2.inc"
echo "build_ref_kernels: built (device) $(ls "$OUTD" | tr '\n' ' ')"

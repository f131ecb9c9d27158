#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_surface.py tests/test_gpu_scale_parity.py tests/test_gpu_robustness.py tests/test_gpu_reference_cuda.py -q -x > $OUT/s21_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s21_tests.log; tail -3 $OUT/s21_tests.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfd_forward_tiled|pit_tet_kernel|nn_query_group" --launch-skip 3 -c 3 -f -o $OUT/s21_search python tools/r2_ncu_search.py > $OUT/s21_ncu.log 2>&1; echo "ncu rc=$?"; ls -la $OUT/s21_search.ncu-rep

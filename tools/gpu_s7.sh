#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
X1_DIAG=1 timeout 600 python -m pytest tests/test_gpu_reference_callers.py -q -s -k two_train > $OUT/s7_tests.log 2>&1
grep -n "DIAG\|bwd max\|   pt\|per-point" $OUT/s7_tests.log | cut -c1-600 | head -40

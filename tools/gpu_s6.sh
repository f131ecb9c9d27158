#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_reference_callers.py tests/test_gpu_scale_parity.py -q -s > $OUT/s6_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s6_tests.log
grep -n "X1a\|X1b\|parity res\|passed\|failed\|Error" $OUT/s6_tests.log | cut -c1-2500 | head -40

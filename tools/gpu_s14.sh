#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_second_oracle.py -q -s > $OUT/s14_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s14_tests.log
grep -n "second oracle\|passed\|failed\|^E " $OUT/s14_tests.log | cut -c1-400 | head -20

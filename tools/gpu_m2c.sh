#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_reference_callers.py tests/test_gpu_dropin.py tests/test_gpu_energies.py tests/test_gpu_scale_parity.py -q > $OUT/m2c_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/m2c_tests.log
tail -12 $OUT/m2c_tests.log | cut -c1-400

#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_robustness.py tests/test_gpu_reference_cuda.py tests/test_gpu_scale_parity.py tests/test_gpu_golden.py -q -x > $OUT/s11_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s11_tests.log
tail -5 $OUT/s11_tests.log | cut -c1-300
timeout 300 python tools/r2_time.py a4 > $OUT/s11_time.jsonl 2> $OUT/s11_time.err; cut -c1-200 $OUT/s11_time.jsonl

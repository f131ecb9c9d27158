#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_reference_callers.py tests/test_gpu_surface.py tests/test_gpu_search.py tests/test_gpu_energies.py tests/test_gpu_dropin.py -x -q > $OUT/s3_tests.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/s3_tests.log | cut -c1-300
timeout 300 python tools/diag_a4.py 40 2000 100000 > $OUT/s3_diag.log 2>&1; echo "diag rc=$?"; head -60 $OUT/s3_diag.log | cut -c1-400
timeout 300 python tools/r2_time.py nn > $OUT/s3_time.jsonl 2> $OUT/s3_time.err; echo "time rc=$?"; cut -c1-200 $OUT/s3_time.jsonl; tail -3 $OUT/s3_time.err

#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/s8_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s8_tests.log
tail -6 $OUT/s8_tests.log | cut -c1-300
timeout 300 python tools/r2_time.py pit > $OUT/s8_time.jsonl 2> $OUT/s8_time.err; cut -c1-200 $OUT/s8_time.jsonl

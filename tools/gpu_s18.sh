#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
python tools/ab_run.py a4 t_base t_u2 t_u4 > $OUT/s18_ab.jsonl 2>&1; cut -c1-200 $OUT/s18_ab.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $OUT/s18_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s18_tests.log
tail -6 $OUT/s18_tests.log | cut -c1-300

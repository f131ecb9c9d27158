#!/bin/sh
# round-2 GPU session 1: full GPU test suite (new kernels + scale parity), A/B timings, one bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/s1_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s1_tests.log
tail -15 $OUT/s1_tests.log
timeout 600 python tools/r2_time.py energies nn pit a4 > $OUT/s1_time.jsonl 2> $OUT/s1_time.err; echo "time rc=$?"
cat $OUT/s1_time.jsonl | cut -c1-300
tail -5 $OUT/s1_time.err
timeout 400 python bench.py --skip-cpu > $OUT/s1_bench.json 2> $OUT/s1_bench.err; echo "bench rc=$?"
head -c 3000 $OUT/s1_bench.json; tail -5 $OUT/s1_bench.err

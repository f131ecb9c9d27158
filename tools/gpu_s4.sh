#!/bin/sh
# round-2 GPU session 4: full GPU suite (all failures listed), A4 res-40 diag after the reach clamp, bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/s4_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s4_tests.log
tail -40 $OUT/s4_tests.log | cut -c1-400
timeout 300 python tools/diag_a4.py 40 2000 100000 > $OUT/s4_diag.log 2>&1; echo "diag rc=$?"; head -30 $OUT/s4_diag.log | cut -c1-300
timeout 400 python bench.py --skip-cpu > $OUT/s4_bench.json 2> $OUT/s4_bench.err; echo "bench rc=$?"
head -c 3500 $OUT/s4_bench.json; tail -5 $OUT/s4_bench.err

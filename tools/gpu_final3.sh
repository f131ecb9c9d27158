#!/bin/sh
# last validation of the round: full GPU suite + smoke + the default bench line with the final code
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/h_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/h_tests.log; tail -3 $OUT/h_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/h_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $OUT/h_bench_n1.json 2> $OUT/h_bench_n1.err; echo "bench rc=$?"; grep '"value"' $OUT/h_bench_n1.json | head -c 260; echo

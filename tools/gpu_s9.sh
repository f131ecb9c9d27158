#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_surface.py tests/test_gpu_search.py -q -x > $OUT/s9_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s9_tests.log
tail -6 $OUT/s9_tests.log | cut -c1-300
timeout 300 python tools/r2_time.py pit 2>/dev/null | head -3 | cut -c1-200
for c in 2 3s 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > $OUT/s9_bench_c$c.json 2> $OUT/s9_bench_c$c.err; echo "config $c rc=$?"; head -c 1500 $OUT/s9_bench_c$c.json; echo; tail -3 $OUT/s9_bench_c$c.err | cut -c1-300
done

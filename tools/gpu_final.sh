#!/bin/sh
# final evidence run of the round: smoke, the default bench line (all legs), the CPU reference arm, ncu launch list of one eager step
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/f_smoke.log | cut -c1-200
timeout 900 python bench.py > $OUT/f_bench.json 2> $OUT/f_bench.err; echo "bench rc=$?"; head -c 400 $OUT/f_bench.json; echo
timeout 900 python bench.py --impl reference --steps 200 --warmup 10 > $OUT/f_bench_ref.json 2> $OUT/f_bench_ref.err; echo "ref rc=$?"; head -c 300 $OUT/f_bench_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/f_launches.csv python bench.py --steps 2 --warmup 1 --load-steps 0 --no-graph --serial --skip-cpu --skip-ref-cuda --no-verify --skip-dropin > $OUT/f_launches.log 2>&1; echo "ncu rc=$?"; wc -l $OUT/f_launches.csv
timeout 600 python bench.py --shapes 40,48 --steps 100 --skip-cpu --skip-ref-cuda --no-verify --skip-dropin > $OUT/f_bench_fb12k.json 2> $OUT/f_bench_fb12k.err; echo "fb12k rc=$?"; head -c 300 $OUT/f_bench_fb12k.json; echo

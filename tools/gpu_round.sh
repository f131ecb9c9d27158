#!/bin/sh
# Dev tool: everything a round needs from ONE gpurun call (1 GPU), outputs under gpurun_out/<tag>_*:
#   gpurun --timeout 600 -- 'sh tools/gpu_round.sh r2'
# then, back in the authoring container:
#   python tools/summarize_ncu.py launches gpurun_out/r2_launches.csv profiles/r2_launches_bench_step.md "<command>" "<note>"
#   python tools/summarize_ncu.py full gpurun_out/r2_top.ncu-rep profiles/r2_ncu_full_top_kernels.md profiles/traffic.json "<command>"
# Numbers printed under ncu are never bench values; the bench line is the plain run in step 2.
TAG=${1:-r}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
# 1. parity
timeout 240 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_tests.log
# 2. bench line (default flags = what the driver runs)
timeout 240 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
# 3. launch list of one eager, serial step
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --serial --skip-cpu --skip-ref-cuda > $OUT/${TAG}_ncu_l.log 2>&1
# 4. full set of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"pfd_forward_tiled|nn_query_thread|pit_tet_kernel|energies_bwd_kernel|energies_fwd_kernel" -s 5 -c 5 -o $OUT/${TAG}_top \
    python bench.py --steps 1 --warmup 1 --no-graph --serial --skip-cpu --skip-ref-cuda > $OUT/${TAG}_ncu_f.log 2>&1
# 5. component rows (SURVEY.md section 8 rows + widening rows)
timeout 300 python tools/bench_components.py > $OUT/${TAG}_components.jsonl 2> $OUT/${TAG}_components.err
tail -3 $OUT/${TAG}_tests.log
head -c 400 $OUT/${TAG}_bench.json

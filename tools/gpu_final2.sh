#!/bin/sh
# final validation on a 2-GPU box: full GPU suite (incl. nn.DataParallel X1), the driver's N=2 and N=1 default bench commands
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/g_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/g_tests.log; tail -3 $OUT/g_tests.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 200 --warmup 10 > $OUT/g_bench_n2.json 2> $OUT/g_bench_n2.err; echo "N=2 default bench rc=$?"; grep '"value"' $OUT/g_bench_n2.json | head -c 300; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --impl reference --gpus 2 --steps 200 --warmup 10 > $OUT/g_bench_ref_n2.json 2> $OUT/g_bench_ref_n2.err; echo "N=2 reference arm rc=$?"; head -c 200 $OUT/g_bench_ref_n2.json; echo
timeout 900 python bench.py > $OUT/g_bench_n1.json 2> $OUT/g_bench_n1.err; echo "N=1 default bench rc=$?"; grep '"value"' $OUT/g_bench_n1.json | head -c 300; echo

#!/bin/sh
# 2-GPU session: full GPU suite (incl. the nn.DataParallel X1 test), N=2 bench 10x in a row, gradient check
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q > $OUT/m2_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/m2_tests.log
tail -8 $OUT/m2_tests.log | cut -c1-300
: > $OUT/m2_bench_n2_x10.jsonl
for i in 1 2 3 4 5 6 7 8 9 10; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus 2 --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda --no-verify >> $OUT/m2_bench_n2_x10.jsonl 2>> $OUT/m2_bench_n2.err
  echo "run $i rc=$?" >> $OUT/m2_bench_n2_x10.jsonl
done
grep -c '"value"' $OUT/m2_bench_n2_x10.jsonl; grep "rc=" $OUT/m2_bench_n2_x10.jsonl | tr '\n' ' '
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 tools/check_multigpu_grad.py > $OUT/m2_gradcheck.log 2>&1; echo "gradcheck rc=$?"; tail -5 $OUT/m2_gradcheck.log

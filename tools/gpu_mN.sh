#!/bin/sh
# multi-GPU session: N ranks (argument), configs 3 (weak), 3s (strong), 4, 5 + gradient check
N=$1
OUT=gpurun_out
mkdir -p $OUT
P=29700
for c in 3 3s 4 5; do
  P=$((P+1))
  EXTRA="--skip-cpu --skip-ref-cuda --no-verify --skip-dropin"
  STEPS="--steps 100 --warmup 5"
  [ "$c" = "4" ] && STEPS="--steps 20 --warmup 3"
  [ "$c" = "5" ] && STEPS="--steps 4 --warmup 2"
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --config $c $STEPS $EXTRA > $OUT/m${N}_c$c.json 2> $OUT/m${N}_c$c.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $c $STEPS $EXTRA > $OUT/m${N}_c$c.json 2> $OUT/m${N}_c$c.err
  fi
  echo "N=$N config $c rc=$?"; grep '"value"' $OUT/m${N}_c$c.json | cut -c1-330; tail -2 $OUT/m${N}_c$c.err | cut -c1-200
done
if [ "$N" != "1" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29790 tools/check_multigpu_grad.py > $OUT/m${N}_gradcheck.log 2>&1; echo "gradcheck rc=$?"; tail -1 $OUT/m${N}_gradcheck.log
fi

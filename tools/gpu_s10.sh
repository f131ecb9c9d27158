#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pfd_forward_tiled|pit_tet_kernel|nn_query_group" --launch-skip 3 -c 3 -f -o $OUT/s10_search python tools/r2_ncu_search.py > $OUT/s10_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/s10_ncu.log
ls -la $OUT/s10_search.ncu-rep
timeout 600 python bench.py --skip-cpu > $OUT/s10_bench.json 2> $OUT/s10_bench.err; echo "bench rc=$?"; head -c 600 $OUT/s10_bench.json; echo; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s10_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('e2e_dropin'))[:1500]); print(json.dumps(d.get('parity'))[:1200]); print(d['config'].get('graph'))
PY

#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_dropin.py tests/test_gpu_second_oracle.py tests/test_gpu_reference_callers.py -q -x > $OUT/s15_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/s15_tests.log
tail -4 $OUT/s15_tests.log | cut -c1-300
timeout 600 python bench.py --skip-cpu --skip-ref-cuda --no-verify --steps 50 > $OUT/s15_bench.json 2> $OUT/s15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d.get('e2e_dropin'))[:600])
PY

#!/bin/sh
# memcheck of the nn.DataParallel X1 harness on 2 GPUs
OUT=gpurun_out
mkdir -p $OUT
REF=/tmp/x1ref; rm -rf $REF; mkdir -p $REF; python -c "import zipfile; zipfile.ZipFile('oracle/_ref/reference_py.zip').extractall('$REF')"
ROOT=$(pwd)
cd $REF
PYTHONPATH=$ROOT/deftet_b200/dropin:$ROOT:$ROOT/tests/x1/stubs:$REF DEFTET_REFERENCE_ROOT=$REF DEFTET_B200_REPO=$ROOT \
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python $ROOT/tests/x1/harness_parallel.py 10 2 /tmp/x1_dp2.json > $ROOT/$OUT/m2b_sanitizer.log 2>&1
echo "rc=$?"
grep -n "Invalid\|misaligned\|at 0x\|by thread\|Address\|in .*kernel\|ERROR SUMMARY" $ROOT/$OUT/m2b_sanitizer.log | head -30 | cut -c1-300
tail -5 $ROOT/$OUT/m2b_sanitizer.log | cut -c1-300

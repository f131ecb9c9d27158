#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/diag_a4b.py 70 3000 > $OUT/s5_diag70.log 2>&1; echo "diag rc=$?"; head -80 $OUT/s5_diag70.log | cut -c1-400

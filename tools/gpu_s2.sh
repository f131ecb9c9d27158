#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/diag_a4.py 40 2000 > $OUT/s2_diag.log 2>&1; echo "diag rc=$?"; head -60 $OUT/s2_diag.log
timeout 300 python tools/r2_time.py nn > $OUT/s2_time.jsonl 2> $OUT/s2_time.err; echo "time rc=$?"; cut -c1-200 $OUT/s2_time.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"energies_tiled|nn_query_brick|nn_query_thread|energies_fwd_kernel|energies_bwd_kernel" -c 12 -o $OUT/s2_ncu python tools/r2_ncu_driver.py > $OUT/s2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/s2_ncu.log
ls -la $OUT/s2_ncu.ncu-rep

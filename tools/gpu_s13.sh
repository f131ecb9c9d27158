#!/bin/sh
OUT=gpurun_out
mkdir -p $OUT
K="bf_flag|bf_compact|chamfer_fwd|chamfer_bwd|pfd_backward_indexed|bary_backward|fa_neighbour|energies_fwd|energies_bwd|rc_forward|rc_backward|cs_query|spmm_csr|radix_scatter"
timeout 420 ncu --set full --clock-control none -k regex:"$K" -c 20 -f -o $OUT/s13_others python tools/r2_ncu_others.py > $OUT/s13_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/s13_ncu.log | cut -c1-200
ls -la $OUT/s13_others.ncu-rep

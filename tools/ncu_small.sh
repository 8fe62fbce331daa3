#!/bin/bash
# ncu --set full capture of one launch each of the small kernels on the minibatch chain (diagnostic)
OUT=${1:-gpurun_out/ncu_small}
T=8192 EPOCHS=1 ncu --set full --clock-control none --import-source on \
  -k regex:"enc_fwd_kernel|enc_bwd_kernel|gather_kernel|sort_scan_kernel|jsmlp_presplit_kernel|adam_apply_kernel|concat_sorted_split|unsort_grad|tower_rows" \
  --launch-skip 60 --launch-count 18 -o $OUT python tools/kernel_probe.py > $OUT.log 2>&1

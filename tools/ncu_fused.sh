#!/bin/bash
# launch list + one `ncu --set full` capture of the six fused tcgen05 kernels (run only after the plain command exits 0)
OUT=${1:-gpurun_out/r02}
T=8192 EPOCHS=1 python tools/kernel_probe.py > ${OUT}_plain.log 2>&1 || exit 1
T=8192 EPOCHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file ${OUT}_launches.csv python tools/kernel_probe.py > ${OUT}_ncu_launches.log 2>&1
T=8192 EPOCHS=1 ncu --set full --clock-control none --import-source on \
  -k regex:"tower_fwd_tc_kernel|tower_bwd_tc_kernel|tower_wgrad_tc_kernel|jsmlp_fwd_tc_kernel|jsmlp_bwd_tc_kernel|jsmlp_wgrad_tc_kernel|tower_wgrad_reduce|jsmlp_wgrad_reduce" \
  --launch-skip 40 --launch-count 16 -o ${OUT}_ncu_fused python tools/kernel_probe.py > ${OUT}_ncu_fused.log 2>&1

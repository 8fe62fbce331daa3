#!/bin/bash
# launch list + one `ncu --set full` capture of the fused tcgen05 kernels, their ordered reductions and the HBM-class
# kernels around them (run only after the plain command exits 0)
OUT=${1:-gpurun_out/r02}
T=8192 EPOCHS=1 python tools/kernel_probe.py > ${OUT}_plain.log 2>&1 || exit 1
T=8192 EPOCHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${OUT}_launches.csv python tools/kernel_probe.py > ${OUT}_ncu_launches.log 2>&1
T=8192 EPOCHS=1 ncu --set full --clock-control none --import-source on \
  -k regex:"tower_fwd_tc_kernel|tower_bwd_tc_kernel|tower_wgrad_tc_kernel|jsmlp_fwd_tc_kernel|jsmlp_bwd_tc_kernel|jsmlp_wgrad_tc_kernel|tower_wgrad_reduce|jsmlp_wgrad_reduce|critic_mlp_fwd_tc_kernel|critic_mlp_bwd_tc_kernel|gather_kernel|concat_sorted_split|unsort_grad|adam_apply_kernel|norm_moments_kernel|gauss_loss_kernel|value_head_kernel" \
  --launch-skip 60 --launch-count 36 -o ${OUT}_ncu_fused python tools/kernel_probe.py > ${OUT}_ncu_fused.log 2>&1

#!/bin/bash
# multi-GPU bench lines of one box: N GPUs, weak + strong for locomotion_ft, weak for the other BASELINE configs
N=$1; shift
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/scale_n${N}_${tag}.json 2> gpurun_out/scale_n${N}_${tag}.err; tail -1 gpurun_out/scale_n${N}_${tag}.json | cut -c1-200; }
run ft_weak --env locomotion_ft
run ft_strong --env locomotion_ft --scaling strong --no-e2e
if [ "$1" == "all" ]; then
  run vt_weak --env locomotion_vt --no-e2e
  run vt_strong --env locomotion_vt --scaling strong --no-e2e
  run box_weak --env manipulation_box --no-e2e
fi

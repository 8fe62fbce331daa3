"""Data-parallel update under torchrun: device time and host enqueue time per update on every rank (diagnostic).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dp_probe.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.engine import Comm  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        comm = Comm()
    T = int(os.environ.get('T', 50000))
    shape = synthetic.ENV_SHAPES[os.environ.get('ENV', 'locomotion_ft')]
    d = dict(DERL_YML)
    d.update(min_batch_size=T, mini_batch_size=2048, num_optim_epoch=int(os.environ.get('EPOCHS', 10)))
    torch.manual_seed(0)
    np.random.seed(1234)
    agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape), comm=comm)
    roll = synthetic.generate_rollout(shape, T, seed=1234 + rank, dtype=np.float32)
    arena = agent.pack(roll)
    agent._prepare(arena)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def step():
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        return agent.update_arena(arena)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(int(os.environ.get('REPS', 4))):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        hs.append(1e3 * agent.last_update_stats['host_enqueue_s'])
    if os.environ.get('PROFILE') and rank == 0:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats('tottime').print_stats(16)
    elif os.environ.get('PROFILE'):
        step()
        torch.cuda.synchronize()
    print(f'{os.environ.get("TAG", "")} rank {rank}/{world}: median {np.median(ts):.1f} ms, host enqueue {np.median(hs):.1f} ms', flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

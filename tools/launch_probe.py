"""Is the update bound by the host's launch rate or by the GPU?  (diagnostic, run on the GPU box)

Times one resident PPO update three ways: host enqueue time (no sync), device time, and device time
with the critic and the policy on one stream (overlap_nets=False).
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.rl_core import estimate_advantages  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    T = int(os.environ.get('T', 50000))
    shape = synthetic.ENV_SHAPES['locomotion_ft']
    d = dict(DERL_YML)
    d.update(min_batch_size=T, mini_batch_size=int(os.environ.get('M', 2048)), num_optim_epoch=10)
    torch.manual_seed(0)
    np.random.seed(1234)
    agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape))
    roll = synthetic.generate_rollout(shape, T, seed=1234, dtype=np.float32)
    arena = agent.pack(roll)
    agent._prepare(arena)
    gen = torch.Generator(device=dev).manual_seed(99)

    def step():
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        agent.update_arena(arena)

    only = os.environ.get('ONLY', '')
    if only:                      # run one net alone: how long is each stream's own chain?
        skip = agent.value_net.engine if only == 'policy' else agent.policy_net.engine
        skip.train_step = lambda *a, **k: None
        skip.train_step_prefetched = lambda *a, **k: None
    for overlap in (True, False):
        agent.overlap_nets = overlap
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        for _ in range(2):
            t0 = time.perf_counter()
            step()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f'overlap_nets={overlap}: host enqueue {1e3 * agent.last_update_stats["host_enqueue_s"]:.1f} ms of the '
                  f'PPO loop, update wall {1e3 * (t1 - t0):.1f} ms, until GPU idle {1e3 * (t2 - t0):.1f} ms', flush=True)


if __name__ == '__main__':
    main()

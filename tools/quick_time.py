"""Device time of resident PPO updates under the production schedule (diagnostic; env knobs are read by the package).
T / M / EPOCHS / ENV select the workload, REPS the number of timed updates; prints the median and the host enqueue time."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    T = int(os.environ.get('T', 50000))
    shape = synthetic.ENV_SHAPES[os.environ.get('ENV', 'locomotion_ft')]
    d = dict(DERL_YML)
    d.update(min_batch_size=T, mini_batch_size=int(os.environ.get('M', 2048)), num_optim_epoch=int(os.environ.get('EPOCHS', 10)))
    torch.manual_seed(0)
    np.random.seed(1234)
    agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape))
    roll = synthetic.generate_rollout(shape, T, seed=1234, dtype=np.float32)
    arena = agent.pack(roll)
    agent._prepare(arena)
    gen = torch.Generator(device=dev).manual_seed(99)

    def step():
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        return agent.update_arena(arena)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(int(os.environ.get('REPS', 5))):
        t0 = time.perf_counter()
        log = step()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        hs.append(1e3 * agent.last_update_stats['host_enqueue_s'])
    print(f'{os.environ.get("TAG", "")}: median {np.median(ts):.1f} ms (min {min(ts):.1f}), host enqueue {np.median(hs):.1f} ms, '
          f'valid {agent.last_update_stats["valid_steps"]}/{agent.last_update_stats["steps"]}, surr {log["surr_loss"]:.6f}', flush=True)


if __name__ == '__main__':
    main()

"""Poor man's timeline (no nsys in the image): CUDA events around every site-labelled launch of one update,
printed per stream for a few minibatches in the middle.  Diagnostic; events perturb the schedule slightly."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import _lib, synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.rl_core import estimate_advantages  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402

KINDS = ('in_fc', 'gconv', 'il', 'enc', 'mlp', 'head', 'agg')
PASS = ('fwd', 'bwd', 'wgrad')


def main():
    dev = torch.device('cuda', 0)
    T = int(os.environ.get('T', 50000))
    shape = synthetic.ENV_SHAPES['locomotion_ft']
    d = dict(DERL_YML)
    d.update(min_batch_size=T, mini_batch_size=2048, num_optim_epoch=int(os.environ.get('EPOCHS', 2)))
    torch.manual_seed(0); np.random.seed(1234)
    agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape))
    roll = synthetic.generate_rollout(shape, T, seed=1234, dtype=np.float32)
    arena = agent.pack(roll)
    agent._prepare(arena)
    gen = torch.Generator(device=dev).manual_seed(99)
    lib = _lib.lib

    def step():
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        agent.update_arena(arena)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    cap = 40000
    lib.sard_prof_enable(0xFFFFFFFF, cap)
    step()
    torch.cuda.synchronize()
    buf = (C.c_double * (4 * cap))()
    n = lib.sard_prof_timeline(buf, cap)
    rows = np.frombuffer(buf, dtype=np.float64)[:4 * n].reshape(n, 4).copy()
    sums = (C.c_double * (4 * 24))()
    lib.sard_prof_collect(sums, 24)
    lib.sard_prof_enable(0, 0)
    print(f'{n} records; span {rows[:, 3].max():.1f} ms')
    # a window in the middle
    t_mid = rows[:, 3].max() * 0.6
    win = float(os.environ.get('WIN_MS', 2.2))
    sel = rows[(rows[:, 2] >= t_mid) & (rows[:, 2] < t_mid + win)]
    sel = sel[np.argsort(sel[:, 2])]
    for sid in sorted(set(sel[:, 1].astype(int))):
        print(f'--- stream {sid}')
        prev_end = None
        for site, _, t0, t1 in sel[sel[:, 1] == sid]:
            gap = (t0 - prev_end) * 1e3 if prev_end is not None else 0.0
            print(f'  {(t0 - t_mid) * 1e3:8.1f} us  +{(t1 - t0) * 1e3:6.1f} us  gap {gap:6.1f}  {KINDS[int(site) // 3]}.{PASS[int(site) % 3]}')
            prev_end = t1
    # busy fraction per stream over the whole update
    for sid in sorted(set(rows[:, 1].astype(int))):
        r = rows[rows[:, 1] == sid]
        print(f'stream {sid}: {len(r)} labelled launches, busy {np.sum(r[:, 3] - r[:, 2]):.1f} ms of {rows[:, 3].max():.1f} ms')


if __name__ == '__main__':
    main()

"""Every kernel of one PPO update by name (sard_prof_kernels_*): per-kernel launch count / mean duration / share, and a
per-stream timeline of a window in the middle.  Diagnostic: events around each launch switch off the programmatic
overlap between consecutive kernels, so the sums are kernel time, not step time.

  MODE=serial   one stream, no side streams (kernel durations without co-scheduling)  [default]
  MODE=overlap  the production schedule (durations while sharing the SMs)
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import _lib, synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402


def collect(lib, cap):
    names = C.create_string_buffer(1 << 16)
    rows = (C.c_double * (6 * cap))()
    n = lib.sard_prof_kernels_collect(names, len(names), rows, cap)
    arr = np.frombuffer(rows, dtype=np.float64)[:6 * n].reshape(n, 6).copy()
    return [s for s in names.value.decode().split('\n') if s], arr


def main():
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    comm = None
    if world > 1:                                          # under torchrun: the data-parallel schedule, rank 0 reports
        import torch.distributed as dist
        from sard_b200.engine import Comm
        dist.init_process_group('nccl', device_id=dev)
        comm = Comm()
    T = int(os.environ.get('T', 50000))
    env = os.environ.get('ENV', 'locomotion_ft')
    mode = os.environ.get('MODE', 'serial')
    shape = synthetic.ENV_SHAPES[env]
    d = dict(DERL_YML)
    d.update(min_batch_size=T, mini_batch_size=int(os.environ.get('M', 2048)), num_optim_epoch=int(os.environ.get('EPOCHS', 2)))
    torch.manual_seed(0); np.random.seed(1234)
    agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape), comm=comm)
    roll = synthetic.generate_rollout(shape, T, seed=1234 + rank, dtype=np.float32)
    arena = agent.pack(roll)
    agent._prepare(arena)
    gen = torch.Generator(device=dev).manual_seed(99)
    lib = _lib.lib

    def step():
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        agent.update_arena(arena)

    if mode == 'serial':
        agent.overlap_nets = False
        lib.sard_set_wgrad_overlap(0)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    cap = 60000
    lib.sard_prof_kernels_enable(cap)
    step()
    torch.cuda.synchronize()
    names, rows = collect(lib, cap)
    lib.sard_prof_kernels_enable(0)
    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    span = rows[:, 3].max()
    print(f'{len(rows)} launches; span {span:.1f} ms; mode {mode}; env {env}')
    dur = rows[:, 3] - rows[:, 2]
    tot = dur.sum()
    agg = []
    for i, nm in enumerate(names):
        m = rows[:, 0] == i
        if not m.any():
            continue
        by = rows[m, 4]
        agg.append((nm, int(m.sum()), dur[m].sum(), dur[m].mean() * 1e3, float(np.median(dur[m])) * 1e3, by.mean()))
    agg.sort(key=lambda r: -r[2])
    print(f'{"kernel":34s} {"n":>6s} {"total ms":>9s} {"mean us":>8s} {"med us":>8s} {"share":>6s} {"MB/launch":>10s} {"GB/s":>8s}')
    for nm, n, t, mean, med, by in agg:
        gbs = by / (mean * 1e-6) / 1e9 if by > 0 else 0.0
        print(f'{nm:34s} {n:6d} {t:9.2f} {mean:8.1f} {med:8.1f} {t / tot:6.3f} {by / 1e6:10.2f} {gbs:8.0f}')
    print(f'sum of kernel time {tot:.1f} ms')
    out = os.environ.get('OUT')
    if out:
        json.dump({'mode': mode, 'env': env, 'span_ms': span, 'kernels': [
            {'kernel': nm, 'launches': n, 'total_ms': t, 'mean_us': mean, 'median_us': med, 'bytes_per_launch': by} for nm, n, t, mean, med, by in agg]},
            open(out, 'w'), indent=1)
    # timeline window in the middle of the minibatch loop
    t_mid = span * 0.6
    win = float(os.environ.get('WIN_MS', 1.2))
    sel = rows[(rows[:, 2] >= t_mid) & (rows[:, 2] < t_mid + win)]
    sel = sel[np.argsort(sel[:, 2])]
    for sid in sorted(set(sel[:, 1].astype(int))):
        print(f'--- stream {sid}')
        prev = None
        for nid, _, t0, t1, _b, _f in sel[sel[:, 1] == sid]:
            gap = (t0 - prev) * 1e3 if prev is not None else 0.0
            print(f'  {(t0 - t_mid) * 1e3:8.1f} us  +{(t1 - t0) * 1e3:6.1f} us  gap {gap:6.1f}  {names[int(nid)]}')
            prev = t1


if __name__ == '__main__':
    main()

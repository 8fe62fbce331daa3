#!/usr/bin/env python
"""Benchmark of the SARD PPO-update hot path (BASELINE.json metric: PPO transitions/sec,
policy+value fwd/bwd, and kernel roofline).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU update on host cores
  python bench.py --scaling strong ...                     # N>1: the reference's 50 000 / 2048 update split 1/N per rank
  python bench.py --env locomotion_vt|manipulation_box|...  # the other BASELINE.json configs (same contract)

Workload (config.workload): locomotion_ft.yml PPO update -- synthetic rollouts of the env's
observation/graph shape, T = 50 000 transitions per GPU, minibatch 2048 per GPU, 10 optimisation
epochs (design_opt/cfg/derl.yml:65-69).  One *step* = one whole ``update_params``
(value pre-pass, GAE, fixed log-prob pre-pass, 240 minibatch updates of value fwd/bwd/Adam +
policy fwd/bwd/clip/Adam) = 491 520 transition-updates per GPU.

``value``  : transition-updates/s with the rollout arena already resident in HBM.
``e2e``    : the same through ``Transform2ActAgent.update_params(batch)`` with the reference's
             host-side batch (list of per-transition states, TrajBatchDisc layout): host packing,
             host->device copies and the device->host log read are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'PPO transitions/sec (policy+value fwd/bwd)'
UNIT = 'transition-updates/s'
SITE_KINDS = ('in_fc', 'graphconv', 'index_linear', 'obs_encoder', 'critic_mlp', 'value_head', 'csr_aggregate')
SITE_PASS = ('fwd', 'bwd_data', 'bwd_weight')
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FFMA lanes x 2 flop x boost clock
# dram__bytes_read.sum + dram__bytes_write.sum per launch of a site's kernels, from the committed `ncu --set full`
# captures profiles/r01_ncu_full_wgrad_all_layers.csv and r01_ncu_full_final_gemm_wgrad_reduce.csv (cold caches: ncu
# flushes between kernels; in the real run a minibatch's working set is L2-resident).  A site launch is one
# sard_linear_* call, i.e. partial-product kernel + ordered reduction for the weight gradients.  Sites without an
# unambiguous capture report null.
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernels, from the committed `ncu --set full`
    capture (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep of this round; ncu flushes
    the caches between kernels, so this is the cold-cache figure)."""
    p = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


SITE_KERNEL = {          # call site -> the kernel that serves it on the default engine
    'graphconv.fwd': 'tower_fwd_tc_kernel', 'graphconv.bwd_data': 'tower_bwd_tc_kernel', 'graphconv.bwd_weight': 'tower_wgrad_tc_kernel',
    'index_linear.fwd': 'jsmlp_fwd_tc_kernel', 'index_linear.bwd_data': 'jsmlp_bwd_tc_kernel', 'index_linear.bwd_weight': 'jsmlp_wgrad_tc_kernel',
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--env', default='locomotion_ft')
    ap.add_argument('--transitions', type=int, default=50000)
    ap.add_argument('--mini-batch', type=int, default=2048)
    ap.add_argument('--epochs', type=int, default=10)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --transitions / --mini-batch per GPU; strong: the global update split 1/N per rank (SURVEY 8e)')
    ap.add_argument('--sweep-nodes', type=int, default=0,
                    help='BASELINE config 5: graphs of this many nodes (random trees, execution stage), hidden width 256 '
                         '(--env names the observation shape); 0 = the env\'s own ant morphologies')
    ap.add_argument('--sweep-hidden', type=int, default=256)
    ap.add_argument('--ref-sample', type=int, default=2048, help='transitions of the bounded reference / CPU-baseline sample')
    ap.add_argument('--ref-epochs', type=int, default=4)
    ap.add_argument('--kernels-out', default='', help='write the per-kernel table (JSON) here')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--sites-out', default='', help='write the per-call-site kernel time table (JSON) here')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_burst': d['bf16_tflops'], 'bf16_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:                 # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def reference_sample(env_name, n_trans, seed=1234):
    """A bounded sample of the bench workload: whole synthetic episodes of the env's shape (all three stages present)."""
    from sard_b200 import synthetic
    shape = synthetic.ENV_SHAPES[env_name]
    return shape, synthetic.generate_rollout(shape, n_trans, seed=seed, min_exec=50, max_exec=400)


def reference_time(env_name, steps, warmup, n_trans, mini_batch, epochs, dtype='f64'):
    """One step = the UNMODIFIED reference's ``update_params`` (oracle/_ref, staged by oracle/stage_ref.py) on the sample:
    value pre-pass, GAE, fixed log-prob pre-pass, ``epochs`` x (n_trans // mini_batch) minibatch updates with the full
    derl.yml nets.  Falls back to the oracle port when the staged reference is absent."""
    import torch
    from oracle import ref_runner
    from sard_b200.config import DERL_YML
    torch.set_num_threads(os.cpu_count() or 1)
    tdt = torch.float64 if dtype == 'f64' else torch.float32
    mini_batch = min(mini_batch, n_trans)
    upd = (n_trans // mini_batch) * mini_batch * epochs
    shape, roll = reference_sample(env_name, n_trans)
    if ref_runner.available():
        hyper = dict(DERL_YML)
        hyper.update(mini_batch_size=mini_batch, num_optim_epoch=epochs)
        prev = torch.get_default_dtype()
        try:
            rr = ref_runner.ReferenceRunner(shape, hyper, dtype=tdt)
            batch = rr.make_batch(roll)
            for _ in range(warmup):
                rr.update_params(batch)
            t0 = time.perf_counter()
            for _ in range(steps):
                rr.update_params(batch)
            dt = (time.perf_counter() - t0) / steps
        finally:
            torch.set_default_dtype(prev)
        kind = 'reference'
    else:
        dt = oracle_port_time(env_name, roll, shape, steps, warmup, mini_batch, epochs)
        kind = 'port'
    sample = (f'{kind}: update_params on {n_trans} synthetic {env_name} transitions (value pre-pass, GAE, log-prob pre-pass, '
              f'{epochs} epochs x {n_trans // mini_batch} minibatches of {mini_batch}; full derl.yml nets, {dtype}), '
              f'{steps} timed after {warmup} warm-up, {dt:.2f} s each')
    return upd / dt, dt, torch.get_num_threads(), kind, sample


def oracle_port_time(env_name, roll, shape, steps, warmup, mini_batch, epochs):
    """The oracle port (oracle/sard_oracle.py, fp64) on the same sample: used only when oracle/_ref is absent."""
    import torch
    from oracle import sard_oracle as O
    from sard_b200.config import derl_config
    from sard_b200.transform2act_agent import Transform2ActAgent
    from sard_b200 import synthetic
    torch.manual_seed(0)
    holder = Transform2ActAgent(derl_config(), torch.float32, 'cpu', 0, 1, env=synthetic.SyntheticEnv(shape))

    def sd64(m):
        out = {}
        for k, v in m.state_dict().items():
            if k.endswith('.inv'):
                continue
            t = v.detach().double().clone() if v.is_floating_point() else v.clone()
            if t.is_floating_point() and not any(k.endswith(x) for x in ('.mean', '.var', '.std')):
                t.requires_grad_(True)
            out[k] = t
        return out

    psd, vsd = sd64(holder.policy_net), sd64(holder.value_net)
    spec = O.Spec(attr_fixed_dim=shape.attr_fixed_dim, index_base=shape.index_base,
                  embeds=holder.policy_net.group.get_cur_subgroup_embeds().double(),
                  orbits={int(k): [int(x) for x in v] for k, v in holder.policy_net.group.get_cur_orbits().items()})
    b = roll.to_traj_batch()
    states = [[torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(st)] for st in b.states]
    actions = [torch.tensor(a) for a in b.actions]
    agent = O.OracleAgent(psd, vsd, spec, mini_batch_size=mini_batch, opt_num_epochs=epochs)
    rew, msk = torch.from_numpy(np.asarray(roll.rewards, dtype=np.float64)), torch.from_numpy(np.asarray(roll.masks, dtype=np.float64))
    for _ in range(warmup):
        agent.update_params(states, actions, rew, msk)
    t0 = time.perf_counter()
    for _ in range(steps):
        agent.update_params(states, actions, rew, msk)
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    dt = 'f64'                              # the reference trains in float64 (design_opt/train.py:51)
    val, per_step, threads, kind, sample = reference_time(args.env, args.steps, max(args.warmup, 1), args.ref_sample,
                                                          args.mini_batch, args.ref_epochs, dt)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': dt, 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': ('the unmodified reference (design_opt + khrylib staged verbatim into oracle/_ref by oracle/stage_ref.py) driven '
                 'through its own Transform2ActAgent.update_params on host cores; torch_geometric.GraphConv and design_opt.envs '
                 'are the only shims (oracle/ref_runner.py)') if kind == 'reference' else
                'oracle port of the reference PPO update (oracle/_ref not staged on this machine)',
    }
    print(json.dumps(line), flush=True)


def per_rank(args, world):
    """(transitions, minibatch) one rank works on."""
    if args.scaling == 'strong' and world > 1:
        return args.transitions // world, args.mini_batch // world
    return args.transitions, args.mini_batch


def workload_config(args, world):
    t, m = per_rank(args, world)
    wl = f'{args.env}.yml PPO update_params, synthetic rollouts of the env obs/graph shape'
    if getattr(args, 'sweep_nodes', 0) > 0:
        wl = (f'scale sweep: {args.sweep_nodes}-node random-tree graphs (execution stage), hidden {args.sweep_hidden}, '
              f'{args.env} observation shape')
    return {'workload': wl, 'sweep_nodes': getattr(args, 'sweep_nodes', 0),
            'env': args.env, 'transitions_per_gpu': t, 'mini_batch_per_gpu': m,
            'optim_epochs': args.epochs, 'global_mini_batch': m * world, 'global_transitions': t * world,
            'transition_updates_per_step_per_gpu': (t // m) * m * args.epochs,
            'parallelism': f'dp{world} (transition-sharded, NCCL allreduce)' if world > 1 else 'single GPU',
            'l2': 'arena + activations exceed L2 (obs arena > 100 MB per GPU at 50 000 transitions + 240 different minibatches '
                  'per step); 256 MiB flush written between steps'}


# ------------------------------------------------------------------------------------------ native arm
def kernel_table(lib, agent, arena, mini_batch, pk):
    """Per-kernel rows of the serialised pass: launches, mean duration, algorithmic bytes / flops per launch where the
    wrapper states them (sard_prof_bytes / ProfScope) or where they follow from the workload, achieved GB/s (TFLOP/s)
    and the fraction of the measured HBM (tensor) peak.  Durations are cudaEvent-bracketed launches, so every kernel
    carries the ~5 us event + launch floor: for the microsecond-scale kernels this is an upper bound on their time."""
    import ctypes as C
    cap = 120000
    names = C.create_string_buffer(1 << 16)
    rows = (C.c_double * (6 * cap))()
    n = lib.sard_prof_kernels_collect(names, len(names), rows, cap)
    arr = np.frombuffer(rows, dtype=np.float64)[:6 * n].reshape(n, 6)
    names = [x for x in names.value.decode().split('\n') if x]
    # workload-derived byte counts for kernels whose wrapper does not know the row count
    n_nodes = arena.n_total * mini_batch / max(arena.T, 1)
    peng = agent.policy_net.engine
    p_active = float(peng.store.size + sum(int(peng.ever_live[2].sum()) * (t.out_dim * t.input_dim + t.out_dim) for t in peng._tables(2)))
    derived = {
        'unsort_grad_kernel': 4.0 * (3 * n_nodes * 64 + 2 * mini_batch * 192),          # dxs + h read, dh written; ext read, dext written
        'adam_apply_kernel': 28.0 * (p_active + agent.value_net.engine.store.size) / 2,  # p, m, v read + written, g read (+ zeroed); mean of the two nets
        'norm_apply_kernel': None,
    }
    out = []
    dur = arr[:, 3] - arr[:, 2]
    tot = float(dur.sum()) if n else 0.0
    for i, nm in enumerate(names):
        m = arr[:, 0] == i
        if not m.any():
            continue
        mean_us = float(dur[m].mean() * 1e3)
        by, fl = float(arr[m, 4].mean()), float(arr[m, 5].mean())
        if by <= 0 and derived.get(nm):
            by = derived[nm]
        row = {'kernel': nm, 'launches': int(m.sum()), 'mean_us': mean_us, 'share_of_kernel_time': float(dur[m].sum() / tot) if tot else 0.0}
        if by > 0:
            gbs = by / (mean_us * 1e-6) / 1e9
            row.update({'algorithmic_bytes_per_launch': by, 'gbs': gbs, 'frac_hbm': gbs / pk['hbm_gbs']})
        if fl > 0:
            tf = fl / (mean_us * 1e-6) / 1e12
            row.update({'algorithmic_flops_per_launch': fl, 'tflops': tf, 'frac_tensor': tf / pk['bf16_sustained']})
        if by > 0:
            row['bound'] = 'hbm' if (fl <= 0 or fl / by < pk['bf16_sustained'] * 1e3 / pk['hbm_gbs']) else 'tensor'
        out.append(row)
    out.sort(key=lambda r: -r['share_of_kernel_time'])
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    from sard_b200 import _lib, synthetic
    from sard_b200.config import DERL_YML, Config
    from sard_b200.engine import Comm
    from sard_b200.rl_core import estimate_advantages
    from sard_b200.transform2act_agent import Transform2ActAgent

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the SARD B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        comm = Comm()
    lib = _lib.lib

    shape = synthetic.ENV_SHAPES[args.env]
    n_trans, n_mini = per_rank(args, world)          # strong scaling: the global update split 1/N per rank
    d = dict(DERL_YML)
    if args.sweep_nodes > 0:
        from sard_b200.config import sweep_yml
        d = sweep_yml(args.sweep_hidden)
    d.update(min_batch_size=n_trans, mini_batch_size=n_mini, num_optim_epoch=args.epochs)
    cfg = Config(d)
    torch.manual_seed(0)                       # identical initial weights on every rank
    np.random.seed(1234)                       # identical shuffle stream on every rank
    agent = Transform2ActAgent(cfg, torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape), comm=comm)
    if args.sweep_nodes > 0:
        roll = synthetic.generate_sweep_rollout(shape, n_trans, args.sweep_nodes, seed=1234 + rank)
    else:
        roll = synthetic.generate_rollout(shape, n_trans, seed=1234 + rank, dtype=np.float32)
    upd_per_step = (n_trans // n_mini) * n_mini * args.epochs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    host_rng = np.random.default_rng(99 + rank)

    def resident_step(arena):
        flush.fill_(1)
        # fresh synthetic rewards every step: re-using one batch would let the policy over-fit its
        # advantages and trip the KL gate (transform2act_agent.py:404-407), which real rollouts do not
        arena.rewards.normal_(0.01, 0.1, generator=gen)
        agent.update_arena(arena)          # value pre-pass, GAE, fixed log-prob pre-pass, 240 minibatch updates

    arena = agent.pack(roll)
    agent._prepare(arena)
    for _ in range(args.warmup):
        resident_step(arena)
    barrier()

    # ---- per-call-site table (diagnostic pass, not part of any reported throughput) ----------------
    site_rows = None
    top_site = None
    # run serialised (one stream, no side streams) so a site's time is its kernels' own time, comparable with
    # the ncu launch list; the timed region below runs the normal overlapped schedule
    agent.overlap_nets = False
    lib.sard_set_wgrad_overlap(0)
    if rank == 0:
        lib.sard_prof_enable(0xFFFFFFFF, 200000)
    resident_step(arena)                   # every rank takes part (the step contains collectives)
    torch.cuda.synchronize()
    agent.overlap_nets = True
    lib.sard_set_wgrad_overlap(1)
    if rank == 0:
        import ctypes as C
        buf = (C.c_double * (4 * 24))()
        _lib.check(lib.sard_prof_collect(buf, 24))
        lib.sard_prof_enable(0, 0)
        site_rows = []
        for s in range(21):
            n, ms, fl, by = buf[4 * s:4 * s + 4]
            if n > 0:
                site_rows.append({'site': f'{SITE_KINDS[s // 3]}.{SITE_PASS[s % 3]}', 'id': s, 'launches': int(n), 'ms': ms,
                                  'tflops': fl / ms / 1e9 if ms > 0 else 0.0, 'gbs': by / ms / 1e6 if ms > 0 else 0.0})
        site_rows.sort(key=lambda r: -r['ms'])
        top_site = site_rows[0]['id'] if site_rows else None
        site_total_ms = sum(r['ms'] for r in site_rows)
    if world > 1:
        t = torch.tensor([top_site if top_site is not None else -1], device=dev)
        dist.broadcast(t, 0)
        top_site = int(t.item())
    barrier()

    # ---- per-kernel table (every launch by name, serialised like the site pass; diagnostic) --------------
    kernel_rows = None
    agent.overlap_nets = False
    lib.sard_set_wgrad_overlap(0)
    if rank == 0:
        lib.sard_prof_kernels_enable(120000)
    resident_step(arena)
    torch.cuda.synchronize()
    agent.overlap_nets = True
    lib.sard_set_wgrad_overlap(1)
    if rank == 0:
        kernel_rows = kernel_table(lib, agent, arena, n_mini, peaks())
        lib.sard_prof_kernels_enable(0)
    barrier()

    # ---- timed region: K steps on resident data -----------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if top_site is not None and top_site >= 0 and rank == 0:
        lib.sard_prof_enable(1 << top_site, 400000)
    launches0 = lib.sard_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        resident_step(arena)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.sard_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    roof = None
    if rank == 0 and top_site is not None and top_site >= 0:
        import ctypes as C
        buf = (C.c_double * (4 * 24))()
        _lib.check(lib.sard_prof_collect(buf, 24))
        lib.sard_prof_enable(0, 0)
        n, tms, fl, by = buf[4 * top_site:4 * top_site + 4]
        pk = peaks()
        kind = SITE_KINDS[top_site // 3]
        site_name = f'{kind}.{SITE_PASS[top_site % 3]}'
        tc = lib.sard_get_gemm_engine() == 4 and site_name in SITE_KERNEL
        # Which roof: the fused kernels move ~40 flop per algorithmic byte (3xTF32: ~120 issued), far left of the ridge
        # (measured bf16 peak / measured HBM = ~200 flop/B): they are HBM-class; the dense-GEMM sites of the SIMT engine
        # are reported against the tensor roof as before.
        ach_gbs, ach_tf = by / tms / 1e6, fl / tms / 1e9
        if kind == 'csr_aggregate' or tc:
            roof = {'bound': 'hbm', 'achieved': ach_gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach_gbs / pk['hbm_gbs'],
                    'traffic': None, 'algorithmic_tflops': ach_tf, 'tensor_peak_tflops': pk['bf16_sustained'],
                    'frac_of_tensor_peak': ach_tf / pk['bf16_sustained'],
                    'pipe': 'tcgen05.mma kind::tf32 x3 (error-compensated hi/lo split), operands by TMA, accumulators in TMEM' if tc else 'SIMT'}
        else:
            roof = {'bound': 'tensor', 'achieved': ach_tf, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                    'frac': ach_tf / pk['bf16_sustained'], 'traffic': None,
                    'pipe': 'fp32 FFMA (SIMT); parity target rtol 1e-4 rules out single-pass TF32/bf16',
                    'fp32_peak_tflops': FP32_PEAK_TFLOPS, 'frac_of_fp32_peak': ach_tf / FP32_PEAK_TFLOPS,
                    'algorithmic_gbs': ach_gbs}
        iso = next(r for r in site_rows if r['id'] == top_site)
        kname = SITE_KERNEL.get(site_name, site_name) if tc else site_name
        roof['traffic'] = ncu_traffic().get(kname)
        roof.update({'kernel': f'{kname} ({site_name})' if tc else site_name, 'launches_timed': int(n),
                     'avg_launch_us': tms / n * 1e3 if n else None, 'algorithmic_bytes_per_launch': by / n if n else None,
                     'algorithmic_flops_per_launch': fl / n if n else None,
                     'share_of_step': tms / ms, 'peak_source': pk['source'] + ' (HBM: copy bandwidth; tensor: sustained figure, kernel timed inside a long step)',
                     'timing': 'CUDA events on the launching stream inside the timed region; up to eight streams run concurrently there '
                               '(critic, policy, their side and gather streams), so this is the launch duration while '
                               'sharing the SMs; the isolated_* keys are the same site timed in a serialised pass before the timed region',
                     'isolated_avg_launch_us': iso['ms'] / iso['launches'] * 1e3,
                     'isolated_achieved': iso['gbs'] if roof['bound'] == 'hbm' else iso['tflops'],
                     'isolated_share_of_site_time': iso['ms'] / site_total_ms})
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * upd_per_step * args.steps / (ms / 1e3)

    # ---- e2e: public API with the reference's host batch --------------------------------------------
    e2e = None
    if not args.no_e2e:
        batch = roll.to_traj_batch()
        agent.update_params(batch)            # warm-up of the host path
        barrier()
        k = max(1, min(args.steps, 3))
        e0.record()
        for _ in range(k):
            batch.rewards = host_rng.normal(0.01, 0.1, size=n_trans)
            agent.update_params(batch)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        st = agent.last_update_stats
        e2e = {'value': world * upd_per_step * k / (ems / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': int(st['h2d_bytes']),
               'd2h_bytes_per_step': int(st['d2h_bytes']), 'ms_per_step': ems / k, 'steps': k,
               'api': 'Transform2ActAgent.update_params(TrajBatchDisc-style list of states)'}
        # the same with the batch already flattened on the host (HostRollout): isolates the Python list walk
        agent.update_params(roll)
        barrier()
        e0.record()
        for _ in range(k):
            roll.rewards = host_rng.normal(0.01, 0.1, size=n_trans).astype(np.float32)
            agent.update_params(roll)
        e1.record()
        barrier()
        e2e['packed_host_value'] = world * upd_per_step * k / (e0.elapsed_time(e1) / 1e3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, per, thr, kind, sample = reference_time(args.env, 1, 1, args.ref_sample, args.mini_batch, args.ref_epochs, 'f64')
        cpu = {'value': v, 'unit': UNIT, 'cores': thr, 'kind': kind, 'sample': sample}
        try:                                   # the same sample in float32 (what the CUDA path computes in)
            v32, per32, _, _, _ = reference_time(args.env, 1, 1, args.ref_sample, args.mini_batch, args.ref_epochs, 'f32')
            cpu['value_f32'] = v32
        except Exception as e:                 # noqa: BLE001
            cpu['value_f32'] = None
            cpu['f32_error'] = str(e)[:200]

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling if world > 1 else 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, world), 'clocks': clocks, 'e2e': e2e,
            'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu,
            'roofline_kernels': [r for r in (kernel_rows or []) if r.get('bound')],
            'valid_policy_steps_last': int(agent.last_update_stats.get('valid_steps', -1)),
            'policy_steps_per_update': int(agent.last_update_stats.get('steps', -1)),
            'note': 'a KL-gated policy step still runs its whole forward + backward on the device; only the Adam apply is skipped',
        }
        print(json.dumps(line), flush=True)
        if args.kernels_out and kernel_rows is not None:
            with open(args.kernels_out, 'w') as f:
                json.dump({'ms_per_update': ms / args.steps, 'kernels': kernel_rows}, f, indent=1)
        if args.sites_out and site_rows is not None:
            with open(args.sites_out, 'w') as f:
                json.dump({'ms_per_update': ms / args.steps, 'sites': site_rows}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()

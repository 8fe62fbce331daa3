"""GPU x2: the transition-sharded data-parallel update (NCCL) equals the single-GPU update on the
union of the shards: RunningNorm statistics, gradients, gate decisions and Adam steps are global."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
T_LOCAL, M_LOCAL, EPOCHS = 128, 32, 2


def _setup(device, comm=None):
    from sard_b200 import synthetic
    from sard_b200.transform2act_agent import Transform2ActAgent
    from tests.util import small_config
    shape = synthetic.EnvShape('dp', sim_obs_dim=7, hfield_dim=9, obj_dim=1, goal_dim=3)
    torch.manual_seed(3)
    cfg = small_config('H_2>0>H_2>4', M_LOCAL if comm is not None else 2 * M_LOCAL, EPOCHS, T_LOCAL)
    cfg.agent_specs = {'batch_design': False}
    ag = Transform2ActAgent(cfg, torch.float32, device, 0, 1, env=synthetic.SyntheticEnv(shape), comm=comm)
    with torch.no_grad():
        for nm in ('skel', 'attr', 'control'):
            lin = getattr(ag.policy_net, f'{nm}_ind_mlp').linear
            lin.b.add_(0.05 * torch.randn_like(lin.b))
    rolls = [synthetic.generate_rollout(shape, T_LOCAL, seed=40 + r, orbits=synthetic.SYMMETRIC_ORBITS['H_2'], min_exec=5,
                                        max_exec=14, dtype=np.float32) for r in range(2)]
    rng = np.random.default_rng(1)
    perms = [[rng.permutation(T_LOCAL) for _ in range(EPOCHS)] for _ in range(2)]
    return ag, rolls, perms


def _run(ag, roll, perms):
    from sard_b200.rl_core import estimate_advantages
    arena = ag.pack(roll)
    ag._prepare(arena)
    for m in ag.update_modules:
        m.train(False)
    ag._eval_values(arena)
    a, r = estimate_advantages(arena.rewards, arena.masks, arena.values.unsqueeze(1), 0.995, 0.95)
    arena.advantages.copy_(a.reshape(-1)); arena.returns.copy_(r.reshape(-1))
    return arena


def _worker(rank, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=2, device_id=torch.device('cuda', rank))
    from sard_b200.engine import Comm
    ag, rolls, perms = _setup(f'cuda:{rank}', Comm())
    arena = _run(ag, rolls[rank], perms[rank])
    # advantages must be the single-GPU ones (GAE is per rollout); they are injected by the parent
    adv = torch.load(os.path.join(out, 'adv.pt'))
    arena.advantages.copy_(adv['adv'][rank].cuda()); arena.returns.copy_(adv['ret'][rank].cuda())
    ag.update_policy_packed(arena, perms=perms[rank])
    if rank == 0:
        sd = {('p.' + k): v.cpu() for k, v in ag.policy_net.state_dict().items()}
        sd.update({('v.' + k): v.cpu() for k, v in ag.value_net.state_dict().items()})
        sd['valid'] = torch.tensor(ag.last_update_stats['valid_steps'])
        torch.save(sd, os.path.join(out, 'dp.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_update_equals_single_gpu_update_on_the_union(tmp_path):
    import torch.multiprocessing as mp
    from sard_b200.synthetic import HostRollout
    torch.cuda.set_device(0)
    ag, rolls, perms = _setup('cuda:0')
    # union arena: rank0's transitions then rank1's; global minibatch k = rank0's k-th + rank1's k-th local minibatch
    cat = lambda f: np.concatenate([getattr(rolls[0], f), getattr(rolls[1], f)], axis=(1 if f == 'edges' else 0))
    n0, e0 = rolls[0].node_ptr[-1], rolls[0].edge_ptr[-1]
    union = HostRollout(shape=rolls[0].shape, obs=cat('obs'), node_ptr=np.concatenate([rolls[0].node_ptr, rolls[1].node_ptr[1:] + n0]),
                        edges=cat('edges'), edge_ptr=np.concatenate([rolls[0].edge_ptr, rolls[1].edge_ptr[1:] + e0]),
                        stage=cat('stage'), body_ind=cat('body_ind'), hfield=cat('hfield'), obj=cat('obj'), goal=cat('goal'),
                        actions=cat('actions'), rewards=cat('rewards'), masks=cat('masks'), exps=cat('exps'))
    arena = _run(ag, union, None)
    # per-rollout GAE (each rank normalises its own advantages in the DP run; use the same numbers in both runs)
    adv = {'adv': [arena.advantages[:T_LOCAL].cpu(), arena.advantages[T_LOCAL:].cpu()],
           'ret': [arena.returns[:T_LOCAL].cpu(), arena.returns[T_LOCAL:].cpu()]}
    torch.save(adv, os.path.join(tmp_path, 'adv.pt'))
    # global order for epoch e: the local perms compose like the agent composes them
    cur = [np.arange(T_LOCAL), np.arange(T_LOCAL)]
    gperms = []
    prev_global = np.arange(2 * T_LOCAL)
    for e in range(EPOCHS):
        cur = [cur[r][perms[r][e]] for r in range(2)]
        order = []
        for k in range(T_LOCAL // M_LOCAL):
            order += list(cur[0][k * M_LOCAL:(k + 1) * M_LOCAL]) + list(cur[1][k * M_LOCAL:(k + 1) * M_LOCAL] + T_LOCAL)
        order = np.array(order)
        inv_prev = np.empty_like(prev_global); inv_prev[prev_global] = np.arange(len(prev_global))
        gperms.append(inv_prev[order])           # update_policy_packed applies perms to the previous order
        prev_global = order
    ag.update_policy_packed(arena, perms=gperms)
    single = {('p.' + k): v.cpu() for k, v in ag.policy_net.state_dict().items()}
    single.update({('v.' + k): v.cpu() for k, v in ag.value_net.state_dict().items()})
    port = 29600 + (os.getpid() % 300)
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=_worker, args=(r, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    dp = torch.load(os.path.join(tmp_path, 'dp.pt'))
    assert int(dp['valid']) == ag.last_update_stats['valid_steps'] == EPOCHS * (T_LOCAL // M_LOCAL)
    worst = 0.0
    for k, v in single.items():
        if k.endswith('.inv'):
            continue
        if k.endswith('.n'):
            assert int(dp[k]) == int(v), k
            continue
        d = float((dp[k].double() - v.double()).abs().max())
        worst = max(worst, d)
        lr = 3e-4 if k.startswith('v.') else 5e-5
        frac = float(((dp[k].double() - v.double()).abs() <= 0.1 * lr + 1e-4 * v.double().abs()).double().mean())
        assert frac >= 0.995 and d <= 14 * lr, (k, frac, d)
    print('max |dp - single| =', worst)

"""GPU: the whole CUDA path (modules -> C++ schedules -> kernels) against the reference's golden
vectors (tests/golden, made from the unmodified reference) and against the oracle on fresh inputs."""
import numpy as np
import pytest
import torch

from tests.util import CASES, Case, assert_close, assert_mostly_close, build_agent

pytestmark = pytest.mark.gpu
FWD = dict(rtol=1e-4, atol=2e-5)        # forward outputs / log-probs: rtol 1e-4 (BASELINE.json)
GRAD = dict(rtol=1e-3)                  # gradients: rtol 1e-3


@pytest.fixture(scope='module', params=CASES)
def case(request, cuda_device):
    torch.cuda.set_device(cuda_device)
    return Case(request.param)


def load_norms(case, agent, which):
    for net, key in ((agent.policy_net, f'{which}.policy.'), (agent.value_net, f'{which}.value.')):
        sd = net.state_dict()
        for k, v in case.z.items():
            if k.startswith(key) and k[len(key):] in sd and ('norm.' in k):
                sd[k[len(key):]].copy_(torch.from_numpy(np.array(v)).to(sd[k[len(key):]].dtype))


def test_eval_forward_matches_reference_goldens(case):
    ag = build_agent(case, 'cuda')
    ag.policy_net.eval(); ag.value_net.eval()
    st, ac = case.states(torch.float32), case.actions(torch.float32)
    cd, ad, sk, ndm, dm, total, cc, ca, cs, dev = ag.policy_net.forward(st)
    assert_close(cd.loc.cpu(), case.get('eval.control_mean'), what='control mean', **FWD)
    assert_close(ad.loc.cpu(), case.get('eval.attr_mean'), what='attr mean', **FWD)
    assert_close(sk.logits.cpu(), case.get('eval.skel_logits'), what='skel logits', **FWD)
    assert total == case.roll.obs.shape[0]
    assert np.array_equal(ndm['execution'].numpy(), np.repeat(case.roll.stage == 2, case.roll.num_nodes))
    assert np.array_equal(cc, np.cumsum(case.roll.num_nodes[case.roll.stage == 2]))
    lp = ag.policy_net.get_log_prob(st, ac)
    assert_close(lp.cpu(), case.get('eval.log_prob'), what='log_prob', rtol=1e-4, atol=1e-4)
    v = ag.value_net(st)
    assert_close(v.cpu(), case.get('eval.value'), what='value', **FWD)


def test_train_mode_forward_updates_running_norm_like_reference(case):
    ag = build_agent(case, 'cuda')
    st, ac = case.states(torch.float32), case.actions(torch.float32)
    T = len(st)
    ag.policy_net.train(); ag.value_net.train()
    ag.policy_net.get_log_prob(st[:T // 2], ac[:T // 2])
    ag.value_net(st[:T // 2])
    psd, vsd = ag.policy_net.state_dict(), ag.value_net.state_dict()
    for k, v in case.z.items():
        if k.startswith('warm.policy.'):
            assert_close(psd[k[len('warm.policy.'):]].cpu(), v, rtol=1e-4, atol=1e-6, what=k)
        if k.startswith('warm.value.'):
            assert_close(vsd[k[len('warm.value.'):]].cpu(), v, rtol=1e-4, atol=1e-6, what=k)
    ag.policy_net.eval(); ag.value_net.eval()
    cd, ad, sk, *_ = ag.policy_net.forward(st)
    assert_close(cd.loc.cpu(), case.get('warm.control_mean'), what='control mean', rtol=1e-4, atol=5e-5)
    assert_close(ad.loc.cpu(), case.get('warm.attr_mean'), what='attr mean', rtol=1e-4, atol=5e-5)
    assert_close(sk.logits.cpu(), case.get('warm.skel_logits'), what='skel', rtol=1e-4, atol=5e-5)
    assert_close(ag.policy_net.get_log_prob(st, ac).cpu(), case.get('warm.log_prob'), what='lp', rtol=1e-4, atol=2e-4)
    assert_close(ag.value_net(st).cpu(), case.get('warm.value'), what='value', rtol=1e-4, atol=5e-5)


def test_loss_and_gradients_match_reference_goldens(case):
    from sard_b200.packed import Ordering, PackedRollout, stage_partition
    ag = build_agent(case, 'cuda')
    load_norms(case, ag, 'warm')
    arena = PackedRollout(case.roll, 'cuda')
    ag._prepare(arena)
    T = arena.T
    arena.advantages.copy_(torch.from_numpy(case.get('gae.advantages').reshape(-1)).float())
    arena.returns.copy_(torch.from_numpy(case.get('gae.returns').reshape(-1)).float())
    arena.fixed_log_probs.copy_(torch.from_numpy(case.get('grad.fixed_log_prob').reshape(-1)).float())
    ag.policy_net.train(); ag.value_net.train()
    peng, veng = ag.policy_net.engine, ag.value_net.engine
    order = Ordering(arena, stage_partition(arena.stage_np))
    runs = order.stage_runs(0, T)
    peng.GA.zero_()
    for s in (2, 1, 0):
        peng.run_stage(s, arena, runs[s], 1, 0.2, 1.0 / T)
    GA = peng.GA.cpu().numpy().astype(np.float64)
    n_skel = int(arena.num_nodes_np[arena.stage_np == 0].sum())
    assert_close(-GA[0] / T, case.get('grad.surr_loss'), rtol=1e-4, atol=1e-6, what='surr loss')
    assert_close(GA[1] / T, case.get('grad.approx_kl'), rtol=1e-4, atol=1e-6, what='approx kl')
    assert_close(GA[4] + GA[2] / n_skel, case.get('grad.entropy'), rtol=1e-4, atol=1e-6, what='entropy')
    checked = 0
    for name, p, off, grp in peng.store.entries:
        want = case.get('grad.policy.' + name)
        got = GA[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 1e-5 * float(np.abs(want).max()), what=name, **GRAD)
        checked += 1
    for s, nm in enumerate(('skel', 'attr', 'control')):
        live = np.flatnonzero(peng.ever_live[s])
        st = peng.net.stage[s]
        tabs = ag.policy_net.__getattr__(f'{nm}_ind_mlp').layers()
        base = peng.gt_off[s]
        for j, t in enumerate(tabs):
            pre = f'{nm}_ind_mlp.' + (f'affine_layers.{j}.' if j < len(tabs) - 1 else 'linear.')
            wW, wb = case.get('grad.policy.' + pre + 'W'), case.get('grad.policy.' + pre + 'b')
            rw = t.out_dim * t.input_dim
            gW = GA[base + st.il[j].gW: base + st.il[j].gW + len(live) * rw].reshape(len(live), t.out_dim, t.input_dim)
            gb = GA[base + st.il[j].gb: base + st.il[j].gb + len(live) * t.out_dim].reshape(len(live), t.out_dim)
            assert_close(gW, wW[live], atol=1e-6 + 1e-5 * float(np.abs(wW).max()), what=pre + 'W', **GRAD)
            assert_close(gb, wb[live], atol=1e-6 + 1e-5 * float(np.abs(wb).max()), what=pre + 'b', **GRAD)
            dead = np.setdiff1d(np.arange(wW.shape[0]), live)
            assert float(np.abs(wW[dead]).max()) == 0.0
            checked += 2
    assert checked > 40
    psd = ag.policy_net.state_dict()
    for k, v in case.z.items():
        if k.startswith('grad.policy_after.'):
            assert_close(psd[k[len('grad.policy_after.'):]].cpu(), v, rtol=1e-4, atol=1e-6, what=k)
    # critic
    veng.GA.zero_()
    vorder = Ordering(arena, np.arange(T))
    veng.run(arena, vorder.run(0, T), 1, 1.0 / T)
    GV = veng.GA.cpu().numpy().astype(np.float64)
    assert_close(GV[5] / T, case.get('grad.value_loss'), rtol=1e-4, atol=1e-6, what='value loss')
    for name, p, off, grp in veng.store.entries:
        want = case.get('grad.value.' + name)
        got = GV[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 1e-5 * float(np.abs(want).max()), what='value ' + name, **GRAD)


def test_gae_through_agent_matches_reference(case):
    from sard_b200.rl_core import estimate_advantages
    r = torch.from_numpy(case.roll.rewards).float().cuda(); m = torch.from_numpy(case.roll.masks).float().cuda()
    v = torch.from_numpy(case.get('warm.value')).float().cuda()
    adv, ret = estimate_advantages(r, m, v, 0.995, 0.95)
    assert_close(adv.cpu(), case.get('gae.advantages'), rtol=0, atol=2e-6, what='adv')   # golden stored as fp32
    assert_close(ret.cpu(), case.get('gae.returns'), rtol=0, atol=2e-6, what='ret')


@pytest.mark.parametrize('tag', ['trivial', 'h2_alpha2'])
def test_update_policy_run_matches_reference_final_weights(tag, cuda_device):
    """a18: fixed log-prob pre-pass + 2 epochs x 3 minibatches of value step, PPO loss, gate, clip, Adam."""
    from sard_b200.packed import PackedRollout
    case = Case(tag)
    ag = build_agent(case, 'cuda')
    load_norms(case, ag, 'warm')
    load_norms(case, ag, 'grad.policy_after'.replace('.policy_after', '')) if False else None
    for net, key in ((ag.policy_net, 'grad.policy_after.'), (ag.value_net, 'grad.value_after.')):
        sd = net.state_dict()
        for k, v in case.z.items():
            if k.startswith(key):
                sd[k[len(key):]].copy_(torch.from_numpy(np.array(v)).to(sd[k[len(key):]].dtype))
    arena = PackedRollout(case.roll, 'cuda')
    ag._prepare(arena)
    arena.advantages.copy_(torch.from_numpy(case.get('gae.advantages').reshape(-1)).float())
    arena.returns.copy_(torch.from_numpy(case.get('gae.returns').reshape(-1)).float())
    log = ag.update_policy_packed(arena, perms=list(case.z['upd.perms']))
    assert ag.last_update_stats['steps'] == 6 and ag.last_update_stats['valid_steps'] == 6
    assert_close(log['surr_loss'], case.get('upd.surr_loss'), rtol=2e-3, atol=2e-5, what='surr')
    assert_close(log['approx_kl'], case.get('upd.approx_kl'), rtol=2e-3, atol=2e-5, what='kl')
    assert_close(log['entropy'], case.get('upd.entropy'), rtol=1e-4, atol=1e-5, what='entropy')
    psd, vsd = ag.policy_net.state_dict(), ag.value_net.state_dict()
    n = 0
    for k, v in case.z.items():
        for pre, sd, lr in (('upd.policy.', psd, 5e-5), ('upd.value.', vsd, 3e-4)):
            if k.startswith(pre):
                name = k[len(pre):]
                if name.endswith('.n'):
                    assert int(sd[name]) == int(v), name
                else:
                    assert_mostly_close(sd[name].cpu(), v, rtol=1e-4, atol=0.1 * lr, frac=0.995, hard_atol=14 * lr, what=k)
                n += 1
    assert n > 80


@pytest.mark.parametrize('zero_grad', ['none', 'zeros'])
def test_full_update_params_against_oracle_on_fresh_rollout(cuda_device, zero_grad):
    """End to end on new data: pack -> value pre-pass -> GAE -> update_policy, vs the oracle (fp64 CPU).
    zero_grad='zeros' is the reference's pinned torch 1.8 behaviour (absent stages keep zero gradient tensors and are
    stepped from momentum); 'none' is torch >= 2 (they are skipped)."""
    from oracle import sard_oracle as O
    from sard_b200 import synthetic
    from tests.util import small_config, load_sd
    from sard_b200.transform2act_agent import Transform2ActAgent
    shape = synthetic.EnvShape('mix', sim_obs_dim=9, hfield_dim=33, obj_dim=5, goal_dim=1)
    T, M = 160, 64
    roll = synthetic.generate_rollout(shape, T, seed=77, orbits=synthetic.SYMMETRIC_ORBITS['K_1'], min_exec=6, max_exec=18,
                                      dtype=np.float32)
    for f in ('obs', 'hfield', 'obj', 'goal', 'actions', 'rewards', 'masks'):
        setattr(roll, f, getattr(roll, f).astype(np.float64))
    cfg = small_config('K_1>0>K_1>4', M, 2, T)
    torch.manual_seed(5)
    ag = Transform2ActAgent(cfg, torch.float32, 'cuda', 0, 1, env=synthetic.SyntheticEnv(shape))
    ag.zero_grad_semantics = zero_grad
    with torch.no_grad():
        for nm in ('skel', 'attr', 'control'):
            lin = getattr(ag.policy_net, f'{nm}_ind_mlp').linear
            lin.b.add_(0.05 * torch.randn_like(lin.b))
    psd = {k: v.detach().cpu().double().clone() for k, v in ag.policy_net.state_dict().items()}
    vsd = {k: v.detach().cpu().double().clone() for k, v in ag.value_net.state_dict().items()}
    for sd in (psd, vsd):
        for k, v in sd.items():
            if v.is_floating_point() and not any(k.endswith(s) for s in ('.mean', '.var', '.std')):
                v.requires_grad_(True)
    spec = O.Spec(attr_fixed_dim=shape.attr_fixed_dim, index_base=5, embeds=ag.policy_net.group.get_cur_subgroup_embeds().double(),
                  orbits={int(k): [int(x) for x in v] for k, v in ag.policy_net.group.get_cur_orbits().items()})
    oag = O.OracleAgent(psd, vsd, spec, mini_batch_size=M, opt_num_epochs=2, zero_grad_to_none=(zero_grad == 'none'))
    rng = np.random.default_rng(0)
    perms = [rng.permutation(T) for _ in range(2)]
    batch = roll.to_traj_batch()
    states = [[torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(s)] for s in batch.states]
    actions = [torch.tensor(a) for a in batch.actions]
    values, adv, ret, logs = oag.update_params(states, actions, torch.from_numpy(roll.rewards), torch.from_numpy(roll.masks), perms)
    # product path
    arena = ag.pack(roll)
    ag._prepare(arena)
    for m in ag.update_modules:
        m.train(False)
    ag._eval_values(arena)
    assert_close(arena.values.cpu(), values.reshape(-1), rtol=1e-4, atol=2e-5, what='values')
    from sard_b200.rl_core import estimate_advantages
    a, r = estimate_advantages(arena.rewards, arena.masks, arena.values.unsqueeze(1), 0.995, 0.95)
    arena.advantages.copy_(a.reshape(-1)); arena.returns.copy_(r.reshape(-1))
    assert_close(a.cpu(), adv, rtol=1e-4, atol=1e-4, what='advantages')
    log = ag.update_policy_packed(arena, perms=perms)
    surr = np.array([l['surr_loss'] for l in logs])
    assert_close(log['surr_loss'], surr.sum() / (np.count_nonzero(surr) + 1e-6), rtol=5e-3, atol=5e-5, what='surr')
    for sd, net, lr in ((psd, ag.policy_net, 5e-5), (vsd, ag.value_net, 3e-4)):
        own = net.state_dict()
        for k, v in sd.items():
            if k.endswith('.n'):
                assert int(own[k]) == int(v)
            elif k.endswith('.inv'):
                continue
            else:
                assert_mostly_close(own[k].cpu(), v.detach(), rtol=1e-4, atol=0.1 * lr, frac=0.995, hard_atol=10 * lr, what=k)


def test_update_is_deterministic(cuda_device):
    """Two identical runs give bit-identical weights (no floating-point atomics anywhere)."""
    case = Case('k1')
    from sard_b200.packed import PackedRollout
    outs = []
    for _ in range(2):
        ag = build_agent(case, 'cuda')
        arena = PackedRollout(case.roll, 'cuda')
        ag._prepare(arena)
        ag._eval_values(arena)
        from sard_b200.rl_core import estimate_advantages
        a, r = estimate_advantages(arena.rewards, arena.masks, arena.values.unsqueeze(1), 0.995, 0.95)
        arena.advantages.copy_(a.reshape(-1)); arena.returns.copy_(r.reshape(-1))
        ag.update_policy_packed(arena, perms=[np.arange(96)[::-1].copy(), np.arange(96)])
        outs.append({k: v.detach().cpu().clone() for k, v in list(ag.policy_net.state_dict().items()) + list(ag.value_net.state_dict().items())})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.fixture
def gemm_engine(request):
    from sard_b200 import _lib
    prev = _lib.lib.sard_get_gemm_engine()
    _lib.lib.sard_set_gemm_engine(int(request.param))
    yield int(request.param)
    _lib.lib.sard_set_gemm_engine(prev)


# engine 4 (default): fused tcgen05 tower + JSMLP; 0: fp32 FFMA tiles; 3: mma.sync 3xTF32; 1, 2: per-layer tcgen05 (SS / TS)
@pytest.mark.parametrize('env_name,gemm_engine', [('manipulation_box', 4), ('locomotion_vt', 4), ('locomotion_ft', 4),
                                                  ('point_nav', 4), ('manipulation_box', 0), ('locomotion_vt', 0),
                                                  ('manipulation_box', 3), ('manipulation_box', 1), ('manipulation_box', 2)],
                         indirect=['gemm_engine'])
def test_full_width_nets_loss_and_gradients_match_oracle_autograd(cuda_device, env_name, gemm_engine):
    """derl.yml network sizes (GNN 64x3, JSMLP 128x2 over 256 body indices, critic MLP 512/256): this is the
    configuration the aligned fast paths serve (fused neighbour-sum GEMMs, 64x128 weight-gradient tiles, masked in_fc
    tile, skinny layers, fused encoders), which the 16-wide golden nets never reach.  Gradients of one stage-partitioned
    minibatch against fp64 autograd of the oracle, policy (all three stages) and critic."""
    from oracle import sard_oracle as O
    from sard_b200 import synthetic
    from sard_b200.config import derl_config
    from sard_b200.packed import Ordering, PackedRollout, stage_partition
    from sard_b200.transform2act_agent import Transform2ActAgent
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES[env_name]
    T = 300
    roll = synthetic.generate_rollout(shape, T, seed=31, min_exec=6, max_exec=20, dtype=np.float32)
    for f in ('obs', 'hfield', 'obj', 'goal', 'actions', 'rewards', 'masks'):
        setattr(roll, f, getattr(roll, f).astype(np.float64))
    torch.manual_seed(3)
    ag = Transform2ActAgent(derl_config(), torch.float32, 'cuda', 0, 1, env=synthetic.SyntheticEnv(shape))
    with torch.no_grad():
        for nm in ('skel', 'attr', 'control'):
            lin = getattr(ag.policy_net, f'{nm}_ind_mlp').linear
            lin.b.add_(0.05 * torch.randn_like(lin.b))
    psd = {k: v.detach().cpu().double().clone() for k, v in ag.policy_net.state_dict().items()}
    vsd = {k: v.detach().cpu().double().clone() for k, v in ag.value_net.state_dict().items()}
    for sd in (psd, vsd):
        for k, v in sd.items():
            if v.is_floating_point() and not any(k.endswith(s) for s in ('.mean', '.var', '.std')):
                v.requires_grad_(True)
    spec = O.Spec(attr_fixed_dim=shape.attr_fixed_dim, index_base=shape.index_base,
                  embeds=ag.policy_net.group.get_cur_subgroup_embeds().double(),
                  orbits={int(k): [int(x) for x in v] for k, v in ag.policy_net.group.get_cur_orbits().items()})
    batch = roll.to_traj_batch()
    states = [[torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(s)] for s in batch.states]
    actions = [torch.tensor(a) for a in batch.actions]
    g = torch.Generator().manual_seed(1)
    adv = torch.randn(T, 1, dtype=torch.float64, generator=g)
    ret = torch.randn(T, 1, dtype=torch.float64, generator=g)
    perm = O.stage_partition(states)
    st_p = [states[i] for i in perm]; ac_p = [actions[i] for i in perm]
    with torch.no_grad():
        fixed = O.policy_log_prob(psd, spec, st_p, ac_p, training=False) + 0.05 * torch.randn(T, 1, dtype=torch.float64, generator=g)
    lp = O.policy_log_prob(psd, spec, st_p, ac_p, training=True)
    loss, kl, _ = O.ppo_loss(lp, fixed, adv[perm], 0.2)
    loss.backward()
    vloss = (O.value_forward(vsd, spec, states, training=True) - ret).pow(2).mean()
    vloss.backward()

    arena = PackedRollout(roll, 'cuda')
    ag._prepare(arena)
    arena.advantages.copy_(adv.reshape(-1).float()); arena.returns.copy_(ret.reshape(-1).float())
    fl = torch.empty(T, dtype=torch.float64); fl[torch.from_numpy(perm)] = fixed.reshape(-1)
    arena.fixed_log_probs.copy_(fl.float())
    ag.policy_net.train(); ag.value_net.train()
    peng, veng = ag.policy_net.engine, ag.value_net.engine
    order = Ordering(arena, stage_partition(arena.stage_np))
    runs = order.stage_runs(0, T)
    assert all(r is not None for r in runs), 'the rollout must contain all three stages'
    peng.GA.zero_()
    for s in (2, 1, 0):
        peng.run_stage(s, arena, runs[s], 1, 0.2, 1.0 / T)
    GA = peng.GA.cpu().numpy().astype(np.float64)
    assert_close(-GA[0] / T, float(loss), rtol=1e-4, atol=1e-6, what='surr loss')
    assert_close(GA[1] / T, kl, rtol=1e-4, atol=1e-6, what='approx kl')
    for name, p, off, grp in peng.store.entries:
        want = psd[name].grad.numpy()
        got = GA[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 2e-5 * float(np.abs(want).max()), what='policy ' + name, **GRAD)
    for s, nm in enumerate(('skel', 'attr', 'control')):
        live = np.flatnonzero(peng.ever_live[s])
        st = peng.net.stage[s]
        tabs = ag.policy_net.__getattr__(f'{nm}_ind_mlp').layers()
        base = peng.gt_off[s]
        for j, t in enumerate(tabs):
            pre = f'{nm}_ind_mlp.' + (f'affine_layers.{j}.' if j < len(tabs) - 1 else 'linear.')
            wW, wb = psd[pre + 'W'].grad.numpy(), psd[pre + 'b'].grad.numpy()
            rw = t.out_dim * t.input_dim
            gW = GA[base + st.il[j].gW: base + st.il[j].gW + len(live) * rw].reshape(len(live), t.out_dim, t.input_dim)
            gb = GA[base + st.il[j].gb: base + st.il[j].gb + len(live) * t.out_dim].reshape(len(live), t.out_dim)
            assert_close(gW, wW[live], atol=1e-6 + 2e-5 * float(np.abs(wW).max()), what=pre + 'W', **GRAD)
            assert_close(gb, wb[live], atol=1e-6 + 2e-5 * float(np.abs(wb).max()), what=pre + 'b', **GRAD)
    veng.GA.zero_()
    veng.run(arena, Ordering(arena, np.arange(T)).run(0, T), 1, 1.0 / T)
    GV = veng.GA.cpu().numpy().astype(np.float64)
    assert_close(GV[5] / T, float(vloss), rtol=1e-4, atol=1e-6, what='value loss')
    for name, p, off, grp in veng.store.entries:
        want = vsd[name].grad.numpy()
        got = GV[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 2e-5 * float(np.abs(want).max()), what='value ' + name, **GRAD)


@pytest.mark.parametrize('n_nodes,gemm_engine', [(256, 4), (300, 0)], indirect=['gemm_engine'])
def test_sweep_shape_loss_and_gradients_match_oracle_autograd(cuda_device, n_nodes, gemm_engine):
    """BASELINE.json config 5 shape: graphs of >= 256 nodes (random trees, degree up to ~8), hidden width 256 everywhere
    (GNN 256x3, JSMLP 256x2): outside the fused 64-wide tensor-core kernels, so this pins the general tiles (fused
    neighbour-sum GEMMs with long neighbour lists, grouped IndexLinear over all 256 body indices).  Execution-stage
    minibatch, loss and every gradient against fp64 autograd of the oracle."""
    from oracle import sard_oracle as O
    from sard_b200 import synthetic
    from sard_b200.config import Config, sweep_yml
    from sard_b200.packed import Ordering, PackedRollout
    from sard_b200.transform2act_agent import Transform2ActAgent
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['locomotion_ft']
    T = 24
    roll = synthetic.generate_sweep_rollout(shape, T, n_nodes, seed=5, n_designs=3)
    for f in ('obs', 'hfield', 'obj', 'goal', 'actions', 'rewards', 'masks'):
        setattr(roll, f, getattr(roll, f).astype(np.float64))
    d = sweep_yml(256); d.update(min_batch_size=T, mini_batch_size=T, num_optim_epoch=1)
    torch.manual_seed(11)
    ag = Transform2ActAgent(Config(d), torch.float32, 'cuda', 0, 1, env=synthetic.SyntheticEnv(shape))
    with torch.no_grad():
        lin = ag.policy_net.control_ind_mlp.linear
        lin.b.add_(0.05 * torch.randn_like(lin.b))
    psd = {k: v.detach().cpu().double().clone() for k, v in ag.policy_net.state_dict().items()}
    vsd = {k: v.detach().cpu().double().clone() for k, v in ag.value_net.state_dict().items()}
    for sd in (psd, vsd):
        for k, v in sd.items():
            if v.is_floating_point() and not any(k.endswith(s) for s in ('.mean', '.var', '.std')):
                v.requires_grad_(True)
    spec = O.Spec(attr_fixed_dim=shape.attr_fixed_dim, index_base=shape.index_base,
                  embeds=ag.policy_net.group.get_cur_subgroup_embeds().double(),
                  orbits={int(k): [int(x) for x in v] for k, v in ag.policy_net.group.get_cur_orbits().items()})
    batch = roll.to_traj_batch()
    states = [[torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(s)] for s in batch.states]
    actions = [torch.tensor(a) for a in batch.actions]
    g = torch.Generator().manual_seed(1)
    adv = torch.randn(T, 1, dtype=torch.float64, generator=g)
    ret = torch.randn(T, 1, dtype=torch.float64, generator=g)
    with torch.no_grad():
        fixed = O.policy_log_prob(psd, spec, states, actions, training=False) + 0.05 * torch.randn(T, 1, dtype=torch.float64, generator=g)
    lp = O.policy_log_prob(psd, spec, states, actions, training=True)
    loss, kl, _ = O.ppo_loss(lp, fixed, adv, 0.2)
    loss.backward()
    vloss = (O.value_forward(vsd, spec, states, training=True) - ret).pow(2).mean()
    vloss.backward()

    arena = PackedRollout(roll, 'cuda')
    ag._prepare(arena)
    arena.advantages.copy_(adv.reshape(-1).float()); arena.returns.copy_(ret.reshape(-1).float())
    arena.fixed_log_probs.copy_(fixed.reshape(-1).float())
    ag.policy_net.train(); ag.value_net.train()
    peng, veng = ag.policy_net.engine, ag.value_net.engine
    order = Ordering(arena, np.arange(T))
    runs = order.stage_runs(0, T)
    assert runs[2] is not None and runs[0] is None and runs[1] is None
    peng.GA.zero_()
    peng.run_stage(2, arena, runs[2], 1, 0.2, 1.0 / T)
    GA = peng.GA.cpu().numpy().astype(np.float64)
    assert_close(-GA[0] / T, float(loss), rtol=1e-4, atol=1e-6, what='surr loss')
    assert_close(GA[1] / T, kl, rtol=1e-4, atol=1e-6, what='approx kl')
    for name, p, off, grp in peng.store.entries:
        if psd[name].grad is None:
            continue                                   # skel / attr parameters: no gradient in an execution-only batch
        want = psd[name].grad.numpy()
        got = GA[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 2e-5 * float(np.abs(want).max()), what='policy ' + name, **GRAD)
    live = np.flatnonzero(peng.ever_live[2])
    assert len(live) == min(n_nodes, 256)
    st = peng.net.stage[2]
    tabs = ag.policy_net.control_ind_mlp.layers()
    base = peng.gt_off[2]
    for j, t in enumerate(tabs):
        pre = 'control_ind_mlp.' + (f'affine_layers.{j}.' if j < len(tabs) - 1 else 'linear.')
        wW, wb = psd[pre + 'W'].grad.numpy(), psd[pre + 'b'].grad.numpy()
        rw = t.out_dim * t.input_dim
        gW = GA[base + st.il[j].gW: base + st.il[j].gW + len(live) * rw].reshape(len(live), t.out_dim, t.input_dim)
        gb = GA[base + st.il[j].gb: base + st.il[j].gb + len(live) * t.out_dim].reshape(len(live), t.out_dim)
        assert_close(gW, wW[live], atol=1e-6 + 2e-5 * float(np.abs(wW).max()), what=pre + 'W', **GRAD)
        assert_close(gb, wb[live], atol=1e-6 + 2e-5 * float(np.abs(wb).max()), what=pre + 'b', **GRAD)
    veng.GA.zero_()
    veng.run(arena, order.run(0, T), 1, 1.0 / T)
    GV = veng.GA.cpu().numpy().astype(np.float64)
    assert_close(GV[5] / T, float(vloss), rtol=1e-4, atol=1e-6, what='value loss')
    for name, p, off, grp in veng.store.entries:
        want = vsd[name].grad.numpy()
        got = GV[16 + off:16 + off + p.numel()].reshape(want.shape)
        assert_close(got, want, atol=1e-6 + 2e-5 * float(np.abs(want).max()), what='value ' + name, **GRAD)

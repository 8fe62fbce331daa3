"""GPU: drop-in API behaviour around the hot path -- every env's observation shape, batched
select_action, device round trips of the modules (the reference moves the nets to the CPU for
sampling) and checkpoint interchange."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests.util import assert_close, small_config

pytestmark = pytest.mark.gpu


def make_agent(shape, subgroup='H_4>0>H_4>4', seed=0, T=96, M=32):
    from sard_b200 import synthetic
    from sard_b200.transform2act_agent import Transform2ActAgent
    torch.manual_seed(seed)
    ag = Transform2ActAgent(small_config(subgroup, M, 2, T), torch.float32, 'cuda', 0, 1, env=synthetic.SyntheticEnv(shape))
    with torch.no_grad():
        for nm in ('skel', 'attr', 'control'):
            lin = getattr(ag.policy_net, f'{nm}_ind_mlp').linear
            lin.b.add_(0.05 * torch.randn_like(lin.b))
    return ag


def oracle_of(ag, shape):
    from oracle import sard_oracle as O
    psd = {k: v.detach().cpu().double().clone() for k, v in ag.policy_net.state_dict().items()}
    vsd = {k: v.detach().cpu().double().clone() for k, v in ag.value_net.state_dict().items()}
    spec = O.Spec(attr_fixed_dim=shape.attr_fixed_dim, index_base=shape.index_base,
                  embeds=ag.policy_net.group.get_cur_subgroup_embeds().double(),
                  orbits={int(k): [int(x) for x in v] for k, v in ag.policy_net.group.get_cur_orbits().items()})
    return O, psd, vsd, spec


def lists_of(roll):
    b = roll.to_traj_batch()
    states = [[torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(s)] for s in b.states]
    return b, states, [torch.tensor(a) for a in b.actions]


@pytest.mark.parametrize('env', ['point_nav', 'patrol', 'locomotion_ft', 'locomotion_vt', 'escape', 'manipulation_box'])
def test_all_six_env_shapes_forward_parity(env, cuda_device):
    """Values and log-probs for every SARD env's observation layout (hfield 1410, obj 12, goal 3, S 41/54)."""
    from sard_b200 import synthetic
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES[env]
    ag = make_agent(shape)
    roll = synthetic.generate_rollout(shape, 96, seed=21, min_exec=8, max_exec=25, dtype=np.float32)
    for f in ('obs', 'hfield', 'obj', 'goal', 'actions', 'rewards', 'masks'):
        setattr(roll, f, getattr(roll, f).astype(np.float64))
    O, psd, vsd, spec = oracle_of(ag, shape)
    b, states, actions = lists_of(roll)
    ag.policy_net.train(); ag.value_net.train()        # train mode: RunningNorm update, then normalised forward
    with torch.no_grad():
        want_lp = O.policy_log_prob(psd, spec, states, actions, training=True)
        want_v = O.value_forward(vsd, spec, states, training=True)
    got_lp = ag.policy_net.get_log_prob(b.states, b.actions)
    got_v = ag.value_net(b.states)
    assert_close(got_lp.cpu(), want_lp, rtol=1e-4, atol=3e-4, what=f'{env} log_prob')
    assert_close(got_v.cpu(), want_v, rtol=1e-4, atol=5e-5, what=f'{env} value')
    for k in ('control_norm.mean', 'control_norm.var'):
        assert_close(ag.policy_net.state_dict()[k].cpu(), psd[k], rtol=1e-4, atol=1e-6, what=k)


def test_select_action_batched_equals_per_state(cuda_device):
    from sard_b200 import synthetic
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['point_nav']
    ag = make_agent(shape, 'K_1>0>K_1>4')
    roll = synthetic.generate_rollout(shape, 40, seed=8, orbits=synthetic.SYMMETRIC_ORBITS['K_1'], min_exec=4, max_exec=10)
    b = roll.to_traj_batch()
    ag.policy_net.eval()
    act, act_env = ag.policy_net.select_action(b.states, mean_action=True)
    assert act.shape == (roll.obs.shape[0], 8) and act.dtype == np.float64
    # oracle side (not the CUDA path against itself): the batched mean actions are the oracle's distribution parameters of
    # policy_forward -- control means, clamped attribute means (columns the subgroup symmetrizer leaves alone) and the
    # arg-max of the tied skeleton logits with the root forced to the no-op (policy.py:339-381)
    O, psd, _, spec = oracle_of(ag, shape)
    roll64 = [[torch.tensor(np.asarray(x, dtype=np.float64)) if i in (0, 5, 6, 7) else (torch.tensor(x) if i == 1 else x) for i, x in enumerate(s)]
              for s in b.states]
    with torch.no_grad():
        fw = O.policy_forward(psd, spec, roll64, training=False)
    starts = np.concatenate([[0], np.cumsum(fw['num_nodes'])])
    rows_of = lambda sid: np.concatenate([np.arange(starts[t], starts[t + 1]) for t in range(len(b.states)) if fw['stage_of'][t] == sid] or [np.zeros(0, int)]).astype(int)
    cad = ag.policy_net.control_action_dim
    if fw['control'] is not None:
        assert_close(act[rows_of(2), :cad], fw['control']['mean'].numpy(), rtol=1e-4, atol=2e-5, what='control mean vs oracle')
    if fw['attr'] is not None:
        want = fw['attr']['mean'].clamp(-1, 1).numpy()
        assert_close(act[rows_of(1), cad + 2:cad + 5], want[:, 2:5], rtol=1e-4, atol=2e-5, what='tied attribute means vs oracle')
    if fw['skel'] is not None:
        want = fw['skel']['logits'].argmax(dim=1).numpy().astype(np.float64)
        want[fw['skel']['body'].numpy() == 0] = 0
        got = act_env[rows_of(0), -1]
        lg = fw['skel']['logits'].numpy()
        top2 = np.sort(lg, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-4                  # rows without a near-tie of the two largest logits
        assert np.array_equal(got[clear], want[clear]), 'skeleton arg-max vs oracle'
    at = 0
    for t, s in enumerate(b.states):
        a1, e1 = ag.policy_net.select_action([s], mean_action=True)
        n = a1.shape[0]
        assert_close(act[at:at + n], a1, rtol=1e-5, atol=1e-6, what=f'state {t} action')
        assert_close(act_env[at:at + n], e1, rtol=1e-5, atol=1e-6, what=f'state {t} action_env')
        st = int(s[2][0])
        if st == 2:
            assert np.all(a1[:, 1:] == 0)
        elif st == 1:
            assert np.all(np.abs(a1[:, 1:7]) <= 1) and np.all(a1[:, 0] == 0) and np.all(a1[:, 7] == 0)
            # tied columns (offset_z, gear, size) are equal inside an orbit of K_1: legs {1,2} and {3,4}
            body = np.asarray(s[4])
            for lvl_legs in ((1, 2), (3, 4), (6, 7), (8, 9)):
                rows = [int(np.flatnonzero(body == x)[0]) for x in lvl_legs if np.any(body == x)]
                if len(rows) == 2:
                    assert_close(a1[rows[0], 3:6], a1[rows[1], 3:6], rtol=0, atol=1e-7, what='orbit tie')
        at += n


def test_modules_survive_cpu_round_trip(cuda_device):
    from sard_b200 import synthetic
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['manipulation_box']
    ag = make_agent(shape)
    roll = synthetic.generate_rollout(shape, 48, seed=2, min_exec=5, max_exec=12)
    b = roll.to_traj_batch()
    ag.policy_net.eval(); ag.value_net.eval()
    lp0, v0 = ag.policy_net.get_log_prob(b.states, b.actions).clone(), ag.value_net(b.states).clone()
    ag.policy_net.to('cpu'); ag.value_net.to('cpu')            # khrylib/utils/torch.py:15-29 (to_cpu for samplers)
    with pytest.raises(Exception):
        ag.policy_net.get_log_prob(b.states, b.actions)        # no CPU implementation: must fail loudly
    ag.policy_net.to('cuda'); ag.value_net.to('cuda')
    assert torch.equal(ag.policy_net.get_log_prob(b.states, b.actions), lp0)
    assert torch.equal(ag.value_net(b.states), v0)
    secs, log = ag.update_params(roll)                          # and the training path still works afterwards
    assert np.isfinite(log['surr_loss'])


def test_checkpoint_interchange(tmp_path, cuda_device):
    from sard_b200 import synthetic
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['point_nav']
    ag = make_agent(shape, seed=1)
    roll = synthetic.generate_rollout(shape, 96, seed=5, min_exec=6, max_exec=20)
    ag.update_params(roll)
    path = os.path.join(tmp_path, 'cp.p')
    ag.save_checkpoint_to(path, epoch=3)
    cp = pickle.load(open(path, 'rb'))
    assert set(cp) == {'policy_dict', 'value_dict', 'running_state', 'loss_iter', 'best_rewards', 'epoch', 'symmetry',
                       'optimizer_state'}                      # the reference's keys + the optimiser
    keys = set(cp['policy_dict'])
    for k in ('control_action_log_std', 'attr_action_log_std', 'hfield_enc.enc.0.weight', 'goal_enc.enc.2.bias', 'skel_norm.n',
              'control_gnn.in_fc.weight', 'attr_gnn.gconv_layers.2.lin_l.bias', 'attr_gnn.gconv_layers.0.lin_r.weight',
              'control_ind_mlp.affine_layers.1.W', 'skel_ind_mlp.linear.b'):
        assert k in keys, k
    assert not any(k.endswith('.inv') for k in keys)            # scratch buffers are not persisted
    assert set(cp['value_dict']) >= {'norm.mean', 'gnn.in_fc.bias', 'mlp.affine_layers.1.weight', 'value_head.weight'}
    ag2 = make_agent(shape, seed=9)
    ag2.load_checkpoint(path)
    b = roll.to_traj_batch()
    for a in (ag, ag2):
        a.policy_net.eval(); a.value_net.eval()
    assert torch.equal(ag.policy_net.get_log_prob(b.states, b.actions), ag2.policy_net.get_log_prob(b.states, b.actions))
    assert torch.equal(ag.value_net(b.states), ag2.value_net(b.states))
    # optimiser state: sparse table moments, and a resumed run continues bit-identically
    tabs = cp['optimizer_state']['policy']['tables']
    rows, m_rows, _, tshape = tabs['control_ind_mlp.affine_layers.0.W']
    assert 0 < len(rows) <= 13 and m_rows.shape[0] == len(rows) and tshape[0] == 40
    roll2 = synthetic.generate_rollout(shape, 96, seed=6, min_exec=6, max_exec=20)
    for a in (ag, ag2):
        a.policy_net.train(); a.value_net.train()
        np.random.seed(123)
        a.update_params(roll2)
    for k, v in ag.policy_net.state_dict().items():
        assert torch.equal(v, ag2.policy_net.state_dict()[k]), k
    for k, v in ag.value_net.state_dict().items():
        assert torch.equal(v, ag2.value_net.state_dict()[k]), k


def test_ppo_loss_then_policy_step_equals_train_step(cuda_device):
    """The reference's single-step API: ppo_loss() + policy_step_from_loss() must leave the same weights as one
    engine train_step on the same (stage-partitioned) batch."""
    from sard_b200 import synthetic
    from sard_b200.packed import Ordering, PackedRollout, stage_partition
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['patrol']
    roll = synthetic.generate_rollout(shape, 80, seed=13, min_exec=5, max_exec=15)
    b = roll.to_traj_batch()
    g = torch.Generator().manual_seed(2)
    adv = torch.randn(80, 1, generator=g)
    a1, a2 = make_agent(shape, seed=4), make_agent(shape, seed=4)
    for ag in (a1, a2):
        ag.policy_net.train()
    fixed = a1.policy_net.get_log_prob(b.states, b.actions).detach().clone()
    for ag in (a1, a2):        # get_log_prob in train mode updated a1's RunningNorm: bring both to the same state
        ag.policy_net.load_state_dict(a1.policy_net.state_dict())
    before = {k: v.detach().clone() for k, v in a1.policy_net.state_dict().items()}
    loss, log, valid = a1.ppo_loss(b.states, b.actions, adv, fixed)
    assert valid and np.isfinite(float(loss))
    a1.policy_step_from_loss()
    arena = PackedRollout.from_lists(b.states, b.actions, device='cuda')
    a2._prepare(arena)
    arena.advantages.copy_(adv.reshape(-1).cuda()); arena.fixed_log_probs.copy_(fixed.reshape(-1).cuda())
    order = Ordering(arena, stage_partition(arena.stage_np))
    runs = order.stage_runs(0, arena.T)
    n_skel = int(arena.num_nodes_np[arena.stage_np == 0].sum())
    a2.policy_net.engine.train_step(arena, runs, arena.T, n_skel, [r is not None for r in runs], a2.clip_epsilon, 40.0,
                                    a2.clip_first_step_only)
    s1, s2 = a1.policy_net.state_dict(), a2.policy_net.state_dict()
    for k in s1:
        assert torch.equal(s1[k], s2[k]), k
    assert any(not torch.equal(s1[k], before[k]) for k in s1 if k.endswith('in_fc.weight')), 'the step must move the weights'


def test_action_server_serves_the_batched_policy(cuda_device):
    """Rollout workers (threads here) get, through the action server, exactly the rows a direct per-state
    select_action(mean_action=True) gives -- from batched forwards."""
    import threading
    from sard_b200 import synthetic
    from sard_b200.action_server import ActionServer
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['point_nav']
    ag = make_agent(shape, 'K_1>0>K_1>4')
    roll = synthetic.generate_rollout(shape, 48, seed=9, orbits=synthetic.SYMMETRIC_ORBITS['K_1'], min_exec=4, max_exec=10)
    b = roll.to_traj_batch()
    ag.policy_net.eval()
    want = [ag.policy_net.select_action([s], True) for s in b.states]
    srv = ActionServer(ag.policy_net, max_batch=16, max_wait_ms=20.0)
    clients = [srv.client() for _ in range(4)]
    got = [None] * len(b.states)

    def work(w):
        for i in range(w, len(b.states), 4):
            got[i] = clients[w].select_action([b.states[i]], True)

    srv.start()
    th = [threading.Thread(target=work, args=(w,)) for w in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    srv.stop()
    assert srv.served == len(b.states) and srv.batches < len(b.states), 'requests must have been batched'
    for i, ((a, ae), (wa, wae)) in enumerate(zip(got, want)):
        assert_close(a, wa, rtol=1e-5, atol=1e-6, what=f'action {i}')
        assert_close(ae, wae, rtol=1e-5, atol=1e-6, what=f'action_env {i}')


def test_keyed_action_server_serves_each_subgroup_like_the_oracle(cuda_device):
    """the neighbour subgroups of optimize_policy (agent :244-256) are evaluated concurrently through one
    action server, every request keyed by its subgroup name.  The rows a keyed client gets must be the oracle's
    distribution parameters under THAT subgroup (embedding and orbit tie both follow the key), never another one's."""
    import threading
    from sard_b200 import synthetic
    from sard_b200.action_server import ActionServer
    torch.cuda.set_device(cuda_device)
    shape = synthetic.ENV_SHAPES['point_nav']
    names = ['H_4>0>H_4>4', 'K_1>0>K_1>4', 'H_2>0>H_2>4']
    ag = make_agent(shape, names[0])
    # all four legs alike: symmetric under every subgroup above
    roll = synthetic.generate_rollout(shape, 36, seed=5, orbits=synthetic.SYMMETRIC_ORBITS['H_1_0'], min_exec=3, max_exec=6)
    b = roll.to_traj_batch()
    ag.policy_net.eval()
    roll64 = [[torch.tensor(np.asarray(x, dtype=np.float64)) if i in (0, 5, 6, 7) else (torch.tensor(x) if i == 1 else x) for i, x in enumerate(s)]
              for s in b.states]
    group = ag.policy_net.group
    start = group.cur_subgroup_name
    want = {}
    for nm in names:                                       # oracle side, one subgroup at a time
        group.set_cur_subgroup_name(nm)
        O, psd, _, spec = oracle_of(ag, shape)
        with torch.no_grad():
            want[nm] = O.policy_forward(psd, spec, roll64, training=False)
    group.set_cur_subgroup_name(start)
    srv = ActionServer(ag.policy_net, max_batch=12, max_wait_ms=10.0)
    clients = {nm: srv.client(key=nm) for nm in names}
    got = {nm: [None] * len(b.states) for nm in names}

    def work(nm):
        for i, s in enumerate(b.states):
            got[nm][i] = clients[nm].select_action([s], True)

    srv.start()
    th = [threading.Thread(target=work, args=(nm,)) for nm in names]
    for t in th:
        t.start()
    for t in th:
        t.join()
    srv.stop()
    assert group.cur_subgroup_name == start, 'the server must leave the policy on its own subgroup'
    cad = ag.policy_net.control_action_dim
    for nm in names:
        fw = want[nm]
        act = np.concatenate([a for a, _ in got[nm]])
        starts = np.concatenate([[0], np.cumsum(fw['num_nodes'])])
        rows_of = lambda sid: np.concatenate([np.arange(starts[t], starts[t + 1]) for t in range(len(b.states)) if fw['stage_of'][t] == sid]).astype(int)
        assert fw['control'] is not None and fw['attr'] is not None
        assert_close(act[rows_of(2), :cad], fw['control']['mean'].numpy(), rtol=1e-4, atol=2e-5, what=f'{nm}: control mean vs oracle')
        assert_close(act[rows_of(1), cad + 2:cad + 5], fw['attr']['mean'].clamp(-1, 1).numpy()[:, 2:5], rtol=1e-4, atol=2e-5,
                     what=f'{nm}: tied attribute means vs oracle')
    # the subgroups really differ on the design stage (different embedding and tie), so a mixed-up key would show
    a0 = np.concatenate([a for a, _ in got[names[0]]])
    a1 = np.concatenate([a for a, _ in got[names[1]]])
    assert np.abs(a0 - a1).max() > 1e-4

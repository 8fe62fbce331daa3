"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only where ``/root/reference`` exists (the build container); the GPU box replays the
committed ``*.npz`` / ``*.json`` files.  The reference is pure Python; two ``sys.modules`` shims
make its hot-path modules importable without the un-vendored wheels:

* ``torch_geometric.nn.GraphConv`` -> PyG-1.6.1 semantics ``lin_l(sum_{j->i} x_j) + lin_r(x_i)``
  (the only arithmetic on the path that lives outside the reference repo; "parity unpinned");
* ``design_opt.envs`` -> ``env_dict = {}`` (MuJoCo envs are out of scope).

Usage:  python tests/golden/make_golden.py            (rewrites the fixtures)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SARD_REFERENCE', '/root/reference')
sys.path.insert(0, REPO)


def install_shims():
    import torch.nn as nn

    class GraphConv(nn.Module):
        def __init__(self, in_channels, out_channels, aggr='add', bias=True, **kwargs):
            super().__init__()
            assert aggr == 'add'
            self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
            self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

        def forward(self, x, edge_index):
            agg = torch.zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])
            return self.lin_l(agg) + self.lin_r(x)

    tg = types.ModuleType('torch_geometric')
    tgnn = types.ModuleType('torch_geometric.nn')
    for nm in ('GCNConv', 'GATConv', 'SAGEConv', 'AGNNConv'):
        setattr(tgnn, nm, None)
    tgnn.GraphConv = GraphConv
    tg.nn = tgnn
    sys.modules['torch_geometric'] = tg
    sys.modules['torch_geometric.nn'] = tgnn
    envs = types.ModuleType('design_opt.envs')
    envs.env_dict = {}
    sys.modules['design_opt.envs'] = envs
    sys.path.insert(0, REF)


def small_cfg(max_index=40):
    gnn = lambda: {'layer_type': 'graph_conv', 'hdims': [16, 16, 16], 'aggr': 'add', 'bias': True}
    imlp = lambda: {'hdims': [16, 16], 'rescale_linear': True, 'max_index': max_index}
    policy = {'htype': 'tanh', 'skel_gnn_specs': gnn(), 'skel_index_mlp': imlp(),
              'control_gnn_specs': gnn(), 'control_index_mlp': imlp(),
              'attr_gnn_specs': gnn(), 'attr_index_mlp': imlp(),
              'control_log_std': 0, 'attr_log_std': -2.3, 'fix_control_std': False, 'fix_attr_std': False}
    value = {'htype': 'tanh', 'design_flag_in_state': True, 'onehot_design_flag': True,
             'mlp': [32, 16], 'gnn_specs': gnn()}
    return policy, value


class StubCfg:
    enable_infer = True
    num_legs = [4]
    struct_subgroup = True
    updating_subgroup = True
    subgroup_alpha_gap = 3
    robot_cfg = {'joint_params': {}}
    agent_specs = {'batch_design': True}

    def __init__(self, init_subgroup_name, policy_specs, value_specs):
        self.init_subgroup_name = init_subgroup_name
        self.policy_specs, self.value_specs = policy_specs, value_specs


class StubEnv:
    index_base = 5


class StubAgent:
    def __init__(self, cfg, shape):
        self.cfg = cfg
        self.env = StubEnv()
        self.attr_fixed_dim = shape.attr_fixed_dim
        self.sim_obs_dim = shape.sim_obs_dim
        self.attr_design_dim = 6
        self.state_dim = shape.state_dim
        self.control_action_dim = 1
        self.skel_num_action = 3
        self.sim_obs_hfield_dim = shape.hfield_dim
        self.sim_obs_obj_dim = shape.obj_dim
        self.sim_obs_goal_dim = shape.goal_dim


def tensorfy_states(roll):
    # same conversion as transform2act_agent.py:20-24 (slots 0,1,5,6,7 -> tensors)
    out = []
    for t in range(roll.num_transitions):
        s = roll.state(t)
        out.append([torch.tensor(x) if i in (0, 1, 5, 6, 7) else x for i, x in enumerate(s)])
    return out


def zero_dead_rows(sd, live):
    dead = np.setdiff1d(np.arange(next(v for k, v in sd.items() if k.endswith('.W')).shape[0]), live)
    for k, v in sd.items():
        if k.endswith('ind_mlp.linear.W') or k.endswith('.W') or (k.endswith('.b') and 'ind_mlp' in k):
            v.data[dead] = 0


def sd_to_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def group_tables():
    from design_opt.models.group_action import LearnedGroup
    out = {}
    for ns in ([4], [3, 4], [5], [6]):
        np.random.seed(0)
        n0 = ns[0]
        G = LearnedGroup(ns, init_subgroup_name=f'H_{n0}>0>H_{n0}>{n0}')
        key = ','.join(map(str, ns))
        entry = {'groups': {}}
        for n, g in G.Gs.items():
            ge = {}
            for nm, sg in g.subgroups.items():
                ge[nm] = {
                    'elements': sorted(sg.elements_set), 'size': sg.size,
                    'orbits': {str(k): [int(x) for x in v] for k, v in sg.orbits.items()},
                    'indicator': sg.elements_indicator.reshape(-1).tolist(),
                    'sub': sorted(g.subgroup_struct[nm]['sub']), 'super': sorted(g.subgroup_struct[nm]['super']),
                    'set_defaults': {str(k): [bool(v[0]), None if v[1] is None else v[1].tolist()]
                                     for k, v in sg.reprentative_set_defaults.items()},
                    'orbits_trans_matrix': {str(k): v.tolist() for k, v in sg.orbits_trans_matrix.items()},
                }
            entry['groups'][str(n)] = ge
            entry.setdefault('elements', {})[str(n)] = {
                nm: {'map_to': e.map_to.tolist(), 'matrix': e.matrix.tolist(),
                     'permutation_inv': e.permutation_inv.tolist()} for nm, e in g.group_elements.items()}
        names = [f'H_{n0}>0>H_{n0}>{n0}']
        if ns == [4]:
            names += ['H_2>1>H_1>4', 'H_2>2>H_1>4', 'K_1>0>K_1>4', 'H_4>1>H_2>4', 'H_2_0>0>H_2_0>4', 'H_1_0>0>H_1_0>4']
        per_name = {}
        for nm in names:
            G.set_cur_subgroup_name(nm)
            def neighbours(count):
                G.count = count
                try:
                    return sorted(G.get_neibor_subgroup_name())
                except AssertionError:
                    return 'AssertionError'
                finally:
                    G.count = 1
            nb_coarse, nb_fine = neighbours(1), neighbours(25)
            p = G._get_parsed_names()
            per_name[nm] = {
                'parsed': {k: (v if isinstance(v, (int, str)) else float(v)) for k, v in p.items()},
                'orbits': {str(k): [int(x) for x in v] for k, v in G.get_cur_orbits().items()},
                'embeds': G.get_cur_subgroup_embeds().tolist(),
                'symmetrizer': [{'matrix': e['matrix'].tolist(), 'permutation_inv': e['permutation_inv'].tolist(),
                                 'ratio': float(e['ratio'])} for e in G.get_cur_subgroup_symmetrizer()],
                'neighbours_count1': nb_coarse, 'neighbours_count25': nb_fine,
            }
        entry['names'] = per_name
        out[key] = entry
    with open(os.path.join(HERE, 'group_tables.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print('group_tables.json written')


def make_case(tag, subgroup_name, orbits_key, seed, shape, T=96, mini_batch=32, n_epochs=2, keep_update=True):
    """One small end-to-end case: forwards, log-probs, grads, GAE and a short update_policy run."""
    from design_opt.models.transform2act_policy import Transform2ActPolicy
    from design_opt.models.transform2act_critic import Transform2ActValue
    from design_opt.agents.transform2act_agent import Transform2ActAgent
    from khrylib.rl.core import estimate_advantages
    from khrylib.utils.torch import to_test
    from sard_b200 import synthetic

    torch.manual_seed(seed)
    np.random.seed(seed)
    pol_specs, val_specs = small_cfg()
    cfg = StubCfg(subgroup_name, pol_specs, val_specs)
    agent = StubAgent(cfg, shape)
    policy = Transform2ActPolicy(pol_specs, agent)
    agent.policy_net = policy
    value = Transform2ActValue(val_specs, agent)
    roll = synthetic.generate_rollout(shape, T, seed=seed, orbits=synthetic.SYMMETRIC_ORBITS[orbits_key],
                                      min_exec=8, max_exec=20)
    live = np.unique(roll.body_ind)
    zero_dead_rows(policy.state_dict(), live)
    # perturb the zero-initialised output layers so the goldens exercise them
    with torch.no_grad():
        for st in ('skel', 'attr', 'control'):
            lin = getattr(policy, f'{st}_ind_mlp').linear
            lin.b.data[live] = 0.05 * torch.randn(len(live), lin.b.shape[1])
        value.value_head.bias.data.fill_(0.1)
    # weights are rounded to fp32-representable values so the fp32 CUDA path sees identical inputs
    for m in (policy, value):
        for p in m.parameters():
            p.data = p.data.float().double()

    out = {}
    out.update(sd_to_np(policy.state_dict(), 'init.policy.'))
    out.update(sd_to_np(value.state_dict(), 'init.value.'))
    for k in ('obs', 'node_ptr', 'edges', 'edge_ptr', 'stage', 'body_ind', 'hfield', 'obj', 'goal',
              'actions', 'rewards', 'masks'):
        v = getattr(roll, k)
        if v.dtype == np.float64:
            v = v.astype(np.float32).astype(np.float64)
            setattr(roll, k, v)
        out['roll.' + k] = v
    states = tensorfy_states(roll)
    actions = [torch.tensor(roll.action(t)) for t in range(T)]

    # ---- eval-mode forwards --------------------------------------------------------
    policy.eval(); value.eval()
    with torch.no_grad():
        cd, ad, sk, *_ = policy.forward(states)
        out['eval.control_mean'] = cd.loc.numpy()
        out['eval.attr_mean'] = ad.loc.numpy()
        out['eval.skel_logits'] = sk.logits.numpy()      # normalised logits of torch Categorical
        out['eval.log_prob'] = policy.get_log_prob(states, actions).numpy()
        out['eval.value'] = value(states).numpy()
    # warm the running norms with one train-mode pass so normalisation is active, then re-evaluate
    policy.train(); value.train()
    with torch.no_grad():
        policy.get_log_prob(states[:T // 2], actions[:T // 2])
        value(states[:T // 2])
    out.update(sd_to_np({k: v for k, v in policy.state_dict().items() if '_norm.' in k}, 'warm.policy.'))
    out.update(sd_to_np({k: v for k, v in value.state_dict().items() if k.startswith('norm.')}, 'warm.value.'))
    policy.eval(); value.eval()
    with torch.no_grad():
        cd, ad, sk, *_ = policy.forward(states)
        out['warm.control_mean'] = cd.loc.numpy()
        out['warm.attr_mean'] = ad.loc.numpy()
        out['warm.skel_logits'] = sk.logits.numpy()
        fixed_lp = policy.get_log_prob(states, actions)
        values = value(states)
        out['warm.log_prob'] = fixed_lp.numpy()
        out['warm.value'] = values.numpy()

    # ---- GAE -------------------------------------------------------------------------
    rewards = torch.from_numpy(roll.rewards); masks = torch.from_numpy(roll.masks)
    adv, ret = estimate_advantages(rewards, masks, values, 0.995, 0.95)
    out['gae.advantages'] = adv.numpy(); out['gae.returns'] = ret.numpy()

    # ---- train-mode loss + gradients on one minibatch (reference ppo_loss / update_value math) ----
    ag = object.__new__(Transform2ActAgent)
    ag.cfg = cfg; ag.device = torch.device('cpu'); ag.dtype = torch.float64
    ag.policy_net, ag.value_net = policy, value
    ag.update_modules = [policy, value]
    ag.gamma, ag.tau, ag.clip_epsilon = 0.995, 0.95, 0.2
    ag.opt_num_epochs, ag.use_mini_batch, ag.mini_batch_size = n_epochs, True, mini_batch
    ag.value_opt_niter = 1
    ag.policy_grad_clip = [(policy.parameters(), 40)]
    ag.trans_policy = lambda s: s
    ag.trans_value = lambda s: s
    ag.optimizer_policy = torch.optim.Adam(policy.parameters(), lr=5e-5)
    ag.optimizer_value = torch.optim.Adam(value.parameters(), lr=3e-4)
    policy.train(); value.train()
    # perturb the fixed log-probs so the ratio is not identically 1 and clipping is exercised
    fixed_pert = fixed_lp + 0.1 * torch.randn_like(fixed_lp)
    out['grad.fixed_log_prob'] = fixed_pert.numpy()
    loss, log, valid = ag.ppo_loss(states, actions, adv, fixed_pert)
    policy.zero_grad()
    loss.backward()
    out['grad.surr_loss'] = np.array(log['surr_loss']); out['grad.entropy'] = np.array(log['entropy'])
    out['grad.approx_kl'] = np.array(log['approx_kl']); out['grad.valid'] = np.array(valid)
    for k, p in policy.named_parameters():
        if p.grad is not None:
            out['grad.policy.' + k] = p.grad.numpy().copy()
    out.update(sd_to_np({k: v for k, v in policy.state_dict().items() if '_norm.' in k}, 'grad.policy_after.'))
    vpred = value(states)
    vloss = (vpred - ret).pow(2).mean()
    value.zero_grad(); vloss.backward()
    out['grad.value_loss'] = vloss.detach().numpy()
    for k, p in value.named_parameters():
        out['grad.value.' + k] = p.grad.numpy().copy()
    out.update(sd_to_np({k: v for k, v in value.state_dict().items() if k.startswith('norm.')}, 'grad.value_after.'))
    policy.zero_grad(set_to_none=True); value.zero_grad(set_to_none=True)

    # ---- a short reference update_policy run (a18): perms recorded, final weights stored ----
    np.random.seed(seed + 7)
    st0 = np.random.get_state()
    perms = []
    for _ in range(n_epochs):
        p = np.arange(T); np.random.shuffle(p); perms.append(p)
    np.random.set_state(st0)
    out['upd.perms'] = np.stack(perms)
    exps = torch.ones(T)
    log_train = ag.update_policy(states, actions, ret, adv, exps)
    out['upd.surr_loss'] = np.array(log_train['surr_loss']); out['upd.approx_kl'] = np.array(log_train['approx_kl'])
    out['upd.entropy'] = np.array(log_train['entropy'])
    if keep_update:
        out.update(sd_to_np(policy.state_dict(), 'upd.policy.'))
        out.update(sd_to_np(value.state_dict(), 'upd.value.'))
    out['meta.embeds'] = policy.group.get_cur_subgroup_embeds().numpy()
    out['meta.orbit_rep'] = np.array([[k, int(x)] for k, v in policy.group.get_cur_orbits().items() for x in v])
    meta = {'tag': tag, 'subgroup': subgroup_name, 'orbits': orbits_key, 'seed': seed, 'T': T,
            'mini_batch': mini_batch, 'n_epochs': n_epochs, 'shape': shape.__dict__}
    out['meta.json'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    # inputs are fp32-exact; reference outputs are rounded to fp32 for storage (6e-8 relative,
    # far below the 1e-4 / 1e-3 parity tolerances)
    out = {k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f'case_{tag}.npz'), **out)
    print(f'case_{tag}.npz written: {len(out)} arrays, valid={valid}, kl={log["approx_kl"]:.4f}')


def symmetrizer_cases():
    """transform_policy_output with attr_select=True (orbit tie + subgroup symmetrizer, policy.py:421-498)."""
    from design_opt.models.transform2act_policy import Transform2ActPolicy
    from sard_b200 import synthetic
    shape = synthetic.EnvShape('golden_b', sim_obs_dim=4, hfield_dim=1, obj_dim=1, goal_dim=3)
    pol_specs, _ = small_cfg()
    out = {}
    names = ['H_4>0>H_4>4', 'H_2>1>H_1>4', 'H_2>2>H_1>4', 'K_1>0>K_1>4', 'H_1_0>0>H_1_0>4', 'K_0>1>H_2_0>4']
    for i, name in enumerate(names):
        torch.manual_seed(100 + i)
        np.random.seed(100 + i)
        agent = StubAgent(StubCfg(name, pol_specs, None), shape)
        policy = Transform2ActPolicy(pol_specs, agent)
        _, body, _, _ = synthetic.make_ant_graph((2, 2, 2, 2), shape)       # symmetric under every subgroup of D_4
        x = torch.randn(len(body), 6)
        env_out, orbit_out = policy.transform_policy_output(x.clone(), body, attr=True, attr_select=True)
        tie_only, _ = policy.transform_policy_output(x.clone(), body, attr=True)
        sk = torch.randint(0, 3, (len(body),))
        sk_out, _ = policy.transform_policy_output(sk.clone(), body)
        out[f'{i}.x'] = x.numpy(); out[f'{i}.body'] = body; out[f'{i}.env'] = env_out.numpy(); out[f'{i}.orbit'] = orbit_out.numpy()
        out[f'{i}.tie'] = tie_only.numpy(); out[f'{i}.skel_in'] = sk.numpy(); out[f'{i}.skel_out'] = sk_out.numpy()
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'symmetrizer.npz'), **out)
    print('symmetrizer.npz written')


def main():
    install_shims()
    torch.set_default_dtype(torch.float64)      # design_opt/train.py:51
    torch.set_num_threads(1)
    from sard_b200 import synthetic
    group_tables()
    symmetrizer_cases()
    shp_a = synthetic.EnvShape('golden_a', sim_obs_dim=5, hfield_dim=7, obj_dim=3, goal_dim=2)
    shp_b = synthetic.EnvShape('golden_b', sim_obs_dim=4, hfield_dim=1, obj_dim=1, goal_dim=3)
    make_case('trivial', 'H_4>0>H_4>4', 'H_4', 11, shp_a)
    make_case('h2_alpha2', 'H_2>2>H_1>4', 'H_2', 12, shp_b)
    make_case('k1', 'K_1>0>K_1>4', 'K_1', 13, shp_b, keep_update=False)


if __name__ == '__main__':
    main()

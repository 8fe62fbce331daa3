"""Drive the UNMODIFIED reference (staged by oracle/stage_ref.py into oracle/_ref/) on CPU: test infrastructure.

``ReferenceRunner`` builds the reference's own ``Transform2ActPolicy`` / ``Transform2ActValue`` / ``Transform2ActAgent``
classes (design_opt/models, design_opt/agents) with the derl.yml network sizes and calls the reference's own
``update_params`` (agent :274-303: value pre-pass, GAE, fixed log-prob pre-pass, the epoch / minibatch loop of
update_value + ppo_loss + clip + Adam).  Nothing of sard_b200's networks, kernels or engine is on this path; the only
repo code it touches is the synthetic rollout generator (the input data) and the hyper-parameter dictionary.

Two ``sys.modules`` shims make the reference importable without its un-vendored wheels (same as
tests/golden/make_golden.py): ``torch_geometric.nn.GraphConv`` (PyG 1.6.1 semantics) and ``design_opt.envs``.
"""
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'design_opt', 'agents'))


def install_shims():
    import torch.nn as nn

    class GraphConv(nn.Module):                      # torch-geometric 1.6.1: lin_l(sum_{j->i} x_j) + lin_r(x_i)
        def __init__(self, in_channels, out_channels, aggr='add', bias=True, **kwargs):
            super().__init__()
            assert aggr == 'add'
            self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
            self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

        def forward(self, x, edge_index):
            agg = torch.zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])
            return self.lin_l(agg) + self.lin_r(x)

    if 'torch_geometric' not in sys.modules:
        tg = types.ModuleType('torch_geometric')
        tgnn = types.ModuleType('torch_geometric.nn')
        for nm in ('GCNConv', 'GATConv', 'SAGEConv', 'AGNNConv'):
            setattr(tgnn, nm, None)
        tgnn.GraphConv = GraphConv
        tg.nn = tgnn
        sys.modules['torch_geometric'] = tg
        sys.modules['torch_geometric.nn'] = tgnn
    if 'design_opt.envs' not in sys.modules:
        envs = types.ModuleType('design_opt.envs')
        envs.env_dict = {}
        sys.modules['design_opt.envs'] = envs
    if REF not in sys.path:
        sys.path.insert(0, REF)


class _Cfg:
    enable_infer = True
    num_legs = [4]
    struct_subgroup = True
    updating_subgroup = True
    subgroup_alpha_gap = 3
    robot_cfg = {'joint_params': {}}
    agent_specs = {'batch_design': True}
    init_subgroup_name = 'H_4>0>H_4>4'          # the trivial subgroup: every synthetic morphology is admissible

    def __init__(self, policy_specs, value_specs):
        self.policy_specs, self.value_specs = policy_specs, value_specs


class _Env:
    def __init__(self, shape):
        self.index_base = shape.index_base


class _AgentStub:
    """What the reference's policy / critic constructors read from the agent (agent :111-124)."""

    def __init__(self, cfg, shape):
        self.cfg = cfg
        self.env = _Env(shape)
        self.attr_fixed_dim = shape.attr_fixed_dim
        self.sim_obs_dim = shape.sim_obs_dim
        self.attr_design_dim = 6
        self.state_dim = shape.state_dim
        self.control_action_dim = 1
        self.skel_num_action = 3
        self.sim_obs_hfield_dim = shape.hfield_dim
        self.sim_obs_obj_dim = shape.obj_dim
        self.sim_obs_goal_dim = shape.goal_dim


class ReferenceRunner:
    def __init__(self, shape, hyper: dict, dtype=torch.float64, seed=0):
        install_shims()
        from design_opt.agents.transform2act_agent import Transform2ActAgent
        from design_opt.models.transform2act_critic import Transform2ActValue
        from design_opt.models.transform2act_policy import Transform2ActPolicy
        torch.set_default_dtype(dtype)               # design_opt/train.py:51-52
        torch.manual_seed(seed)
        np.random.seed(seed)
        cfg = _Cfg(hyper['policy_specs'], hyper['value_specs'])
        stub = _AgentStub(cfg, shape)
        policy = Transform2ActPolicy(cfg.policy_specs, stub)
        stub.policy_net = policy
        value = Transform2ActValue(cfg.value_specs, stub)
        ag = object.__new__(Transform2ActAgent)      # the env / logger / sampler set-up of __init__ needs MuJoCo
        ag.cfg = cfg; ag.device = torch.device('cpu'); ag.dtype = dtype
        ag.policy_net, ag.value_net = policy, value
        ag.update_modules = [policy, value]
        ag.gamma, ag.tau, ag.clip_epsilon = hyper['gamma'], hyper['tau'], hyper['clip_epsilon']
        ag.opt_num_epochs, ag.use_mini_batch, ag.mini_batch_size = hyper['num_optim_epoch'], True, hyper['mini_batch_size']
        ag.value_opt_niter = 1
        ag.policy_grad_clip = [(policy.parameters(), 40)]            # agent :48 (a generator, as in the reference)
        ag.trans_policy = lambda s: s
        ag.trans_value = lambda s: s
        ag.optimizer_policy = torch.optim.Adam(policy.parameters(), lr=hyper['policy_lr'])
        ag.optimizer_value = torch.optim.Adam(value.parameters(), lr=hyper['value_lr'])
        self.agent = ag
        self.dtype = dtype

    def make_batch(self, roll):
        """HostRollout -> the TrajBatchDisc fields update_params reads (states, actions, rewards, masks, exps)."""
        np_dt = np.float64 if self.dtype == torch.float64 else np.float32
        b = roll.to_traj_batch()
        conv = lambda x: x.astype(np_dt) if isinstance(x, np.ndarray) and x.dtype.kind == 'f' else x
        b.states = [[conv(x) for x in s] for s in b.states]
        b.actions = [conv(a) for a in b.actions]
        b.rewards, b.masks, b.exps = conv(np.asarray(b.rewards)), conv(np.asarray(b.masks)), conv(np.asarray(b.exps))
        return b

    def update_params(self, batch):
        t0 = time.perf_counter()
        _, log = self.agent.update_params(batch)
        return time.perf_counter() - t0, log

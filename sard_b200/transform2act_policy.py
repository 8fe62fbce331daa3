"""Drop-in ``Transform2ActPolicy`` (reference: design_opt/models/transform2act_policy.py).

Same constructor, ``state_dict`` layout and method contracts (``forward``, ``select_action``,
``get_log_prob``, ``transform_policy_output``, ``update_symmetry``, ``get_info``); the tensor work
of the three stage networks runs in libsard_b200.so.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .engine import PolicyEngine
from .group_action import LearnedGroup
from .nets import GNNSimple, JSMLP, ObsEncoder, RunningNorm, STAGE_NAMES
from .packed import Ordering, PackedRollout, stage_partition
from .rl_core import Categorical, DiagGaussian

STAGE_KEYS = ('skel_trans', 'attr_trans', 'execution')


def body_levels_and_legs(body_ind, index_base):
    """level = number of base-``index_base`` digits (0 for the root), leg = last digit (policy.py:422-426)."""
    b = np.asarray(body_ind, dtype=np.int64)
    levels = np.zeros_like(b)
    t = b.copy()
    while np.any(t > 0):
        levels += (t > 0)
        t //= index_base
    return levels, b % index_base


class Transform2ActPolicy(nn.Module):
    def __init__(self, cfg, agent):
        super().__init__()
        self.type = 'gaussian'
        self.cfg = cfg
        self.agent = agent
        self.enable_infer = agent.cfg.enable_infer
        self.attr_fixed_dim = agent.attr_fixed_dim
        self.sim_obs_dim = agent.sim_obs_dim
        self.sim_obs_hfield_dim = agent.sim_obs_hfield_dim
        self.sim_obs_obj_dim = agent.sim_obs_obj_dim
        self.sim_obs_goal_dim = agent.sim_obs_goal_dim
        self.attr_design_dim = agent.attr_design_dim
        self.control_state_dim = self.attr_fixed_dim + self.sim_obs_dim + self.attr_design_dim
        self.sym_embed_dim = sum(agent.cfg.num_legs) * 2
        extra = self.sym_embed_dim if self.enable_infer else 0
        self.skel_state_dim = self.attr_fixed_dim + self.attr_design_dim + extra
        self.attr_state_dim = self.attr_fixed_dim + self.attr_design_dim + extra
        self.skel_action_dim = agent.skel_num_action
        self.control_action_dim = agent.control_action_dim
        self.attr_action_dim = agent.attr_design_dim
        self.action_dim = self.control_action_dim + self.attr_action_dim + 1
        self.skel_uniform_prob = cfg.get('skel_uniform_prob', 0.0)
        if self.skel_uniform_prob > 0:
            # distributions.py:46-50 mixes a uniform term into the skeleton log-prob; the fused categorical loss kernel
            # computes the plain log-softmax, so a non-zero value would silently train on the wrong ratios
            raise NotImplementedError('skel_uniform_prob > 0 is not on the CUDA training path (derl.yml leaves it at 0)')
        self.shared_std_att = 1
        self.shared_std_ctl = 1
        if self.control_action_dim != 1 or self.action_dim != _lib.ACTION_DIM:
            raise NotImplementedError('action layout [control(1) | attr(6) | skel(1)] is assumed by the kernels')
        self.index_base = agent.env.index_base
        self.max_index = 256
        if self.enable_infer:
            self.group = LearnedGroup(ns=agent.cfg.num_legs, struct_subgroup=agent.cfg.struct_subgroup,
                                      updating_subgroup=agent.cfg.updating_subgroup,
                                      init_subgroup_name=agent.cfg.init_subgroup_name,
                                      subgroup_alpha_gap=agent.cfg.subgroup_alpha_gap)
            self.is_new_morphology = False
        for unsupported in ('skel_pre_mlp', 'skel_mlp', 'attr_pre_mlp', 'attr_mlp', 'control_pre_mlp', 'control_mlp'):
            if unsupported in cfg:
                raise NotImplementedError(f'{unsupported} is not on the CUDA path (absent from design_opt/cfg/derl.yml)')
        for need in ('skel_gnn_specs', 'attr_gnn_specs', 'control_gnn_specs', 'skel_index_mlp', 'attr_index_mlp',
                     'control_index_mlp'):
            if need not in cfg:
                raise NotImplementedError(f'{need} is required by the CUDA path (design_opt/cfg/derl.yml)')
        if not cfg.get('attr_norm', True):
            raise NotImplementedError('attr_norm=False is not on the CUDA path')

        # construction order follows the reference so a shared torch seed gives identical initial weights
        self.hfield_enc = ObsEncoder(self.sim_obs_hfield_dim)
        self.obj_enc = ObsEncoder(self.sim_obs_obj_dim)
        self.goal_enc = ObsEncoder(self.sim_obs_goal_dim)
        shared = 3 * 64

        def jsmlp(spec_key, in_dim, out_dim):
            c = cfg[spec_key]
            j = JSMLP(in_dim, c['hdims'], out_dim, c.get('max_index', 256), c.get('htype', 'tanh'),
                      c['rescale_linear'], c.get('zero_init', False))
            return j

        self.skel_norm = RunningNorm(self.attr_state_dim)
        self.skel_gnn = GNNSimple(self.attr_state_dim, cfg['skel_gnn_specs'])
        self.skel_ind_mlp = jsmlp('skel_index_mlp', self.skel_gnn.out_dim + shared, self.skel_action_dim)
        self.attr_norm = RunningNorm(self.skel_state_dim)
        self.attr_gnn = GNNSimple(self.skel_state_dim, cfg['attr_gnn_specs'])
        self.attr_ind_mlp = jsmlp('attr_index_mlp', self.attr_gnn.out_dim + shared, self.attr_action_dim)
        self.attr_action_log_std = nn.Parameter(torch.ones(1, self.attr_action_dim) * cfg['attr_log_std'],
                                                requires_grad=not cfg['fix_attr_std'])
        self.control_norm = RunningNorm(self.control_state_dim)
        self.control_gnn = GNNSimple(self.control_state_dim, cfg['control_gnn_specs'])
        self.control_ind_mlp = jsmlp('control_index_mlp', self.control_gnn.out_dim + shared, self.control_action_dim)
        self.control_action_log_std = nn.Parameter(torch.ones(1, self.control_action_dim) * cfg['control_log_std'],
                                                   requires_grad=not cfg['fix_control_std'])
        if cfg['fix_attr_std'] or cfg['fix_control_std']:
            raise NotImplementedError('fixed action std is not on the CUDA path (derl.yml trains both)')
        mi = {self.skel_ind_mlp.max_index, self.attr_ind_mlp.max_index, self.control_ind_mlp.max_index}
        if len(mi) != 1:
            raise NotImplementedError('index MLPs must share max_index')
        self.max_index = mi.pop()
        object.__setattr__(self, '_engine', None)

    # ---- engine plumbing ----------------------------------------------------------------
    @property
    def engine(self) -> PolicyEngine:
        if self._engine is None:
            object.__setattr__(self, '_engine', PolicyEngine(self))
        return self._engine

    def tie_columns(self, stage: int):
        """Output columns overwritten by the orbit representative (policy.py:452-473)."""
        if not self.enable_infer or stage == 2:
            return 0, 0
        if stage == 0:
            return 0, self.skel_action_dim
        if 'axis' in self.agent.cfg.robot_cfg.get('joint_params', {}):
            return 2, 3                      # slices [2:3] and the empty [5:-1]
        return 2, self.attr_action_dim - 1   # slice(2, -1)

    def prepare(self, arena: Optional[PackedRollout] = None, opt=None, comm=None) -> PolicyEngine:
        eng = self.engine
        live = [arena.live_indices(s) for s in range(3)] if arena is not None else None
        if live is not None and comm is not None and comm.world > 1:
            # data parallel: the compact table-gradient layout (slot_of_index) must be the same on every rank, so the
            # live body indices are the union over the ranks' rollouts
            mask = np.zeros((3, self.max_index), dtype=np.int64)
            for s in range(3):
                mask[s, np.asarray(live[s], dtype=np.int64)] = 1
            live = [np.flatnonzero(m) for m in comm.max_array(mask)]
        eng.sync(opt if opt is not None else eng.opt, live)
        if self.enable_infer:
            eng.set_embeds(self.group.get_cur_subgroup_embeds())
            if arena is not None and not arena.has_orbits:
                arena.set_orbits(self.group.leg_representatives(self.index_base), self.index_base)
        return eng

    # ---- reference API ------------------------------------------------------------------
    def _run_stages(self, x, want_outputs: bool, actions=None):
        dev = self.control_action_log_std.device
        with torch.cuda.device(dev):
            arena = PackedRollout.from_lists(x, actions, device=dev)
            eng = self.prepare(arena)
            part = stage_partition(arena.stage_np)
            order = Ordering(arena, part)
            runs = order.stage_runs(0, arena.T)
            mode = 2 if self.training else 0
            outs, lps = [None] * 3, [None] * 3
            for s in (2, 1, 0):
                r = runs[s]
                if r is None:
                    continue
                ldo = (eng.net.stage[s].out_dim + 3) // 4 * 4
                o = torch.zeros(r.n, ldo, dtype=torch.float32, device=dev) if want_outputs else None
                lp = torch.zeros(r.B, dtype=torch.float32, device=dev)
                eng.run_stage(s, arena, r, mode, outputs=o, log_prob_run=lp)
                outs[s] = o[:, :eng.net.stage[s].out_dim] if want_outputs else None
                lps[s] = lp
        return arena, order, runs, outs, lps

    def forward(self, x):
        """Same 10-tuple as the reference (policy.py:215-337)."""
        arena, order, runs, outs, _ = self._run_stages(x, True)
        dev = self.control_action_log_std.device
        stage_of = arena.stage_np
        node_stage = np.repeat(stage_of, arena.num_nodes_np)
        node_design_mask = {k: torch.from_numpy(node_stage == s) for s, k in enumerate(STAGE_KEYS)}
        design_mask = {k: torch.from_numpy(stage_of == s) for s, k in enumerate(STAGE_KEYS)}
        cums = [np.cumsum(arena.num_nodes_np[stage_of == s]) if runs[s] is not None else None for s in range(3)]
        control_dist = attr_dist = skel_dist = None
        if outs[2] is not None:
            control_dist = DiagGaussian(outs[2], self.control_action_log_std.detach().expand_as(outs[2]).exp())
            self.body_ind_execution = torch.from_numpy(arena.body_np[node_stage == 2])
        if outs[1] is not None:
            attr_dist = DiagGaussian(outs[1], self.attr_action_log_std.detach().expand_as(outs[1]).exp())
            self.body_ind_attribute = torch.from_numpy(arena.body_np[node_stage == 1])
        if outs[0] is not None:
            skel_dist = Categorical(logits=outs[0], uniform_prob=self.skel_uniform_prob)
            self.body_ind_skeleleton = torch.from_numpy(arena.body_np[node_stage == 0])
        return (control_dist, attr_dist, skel_dist, node_design_mask, design_mask, arena.n_total,
                cums[2], cums[1], cums[0], dev)

    def get_log_prob(self, x, action, get_entropy=False):
        """Per-graph log-prob ``[B,1]`` in the order of ``x`` (policy.py:383-419)."""
        arena, order, runs, outs, lps = self._run_stages(x, get_entropy, actions=action)
        dev = self.control_action_log_std.device
        out = torch.zeros(arena.T, dtype=torch.float32, device=dev)
        entropy = 0
        for s in (2, 1, 0):
            r = runs[s]
            if r is None:
                continue
            out[order.ids[r.lo:r.hi].long()] = lps[s]
            if get_entropy:
                if s == 0:
                    logp = torch.log_softmax(outs[0], dim=-1)
                    entropy = entropy + (-(logp.exp() * logp).sum(-1)).mean()
                else:
                    ls = (self.attr_action_log_std if s == 1 else self.control_action_log_std).detach()
                    entropy = entropy + (0.5 + 0.5 * float(np.log(2 * np.pi)) + ls).mean()
        out = out.unsqueeze(1)
        return (out, entropy) if get_entropy else out

    def select_action(self, x, mean_action=False):
        """Rollout-side sampling (policy.py:339-381); returns ``(action, action_env)`` as float64 numpy."""
        control_dist, attr_dist, skel_dist, node_design_mask, _, total_num_nodes, _, cum_attr, _, device = self.forward(x)
        action = torch.zeros(total_num_nodes, self.action_dim, device=device)
        action_env = torch.zeros(total_num_nodes, self.action_dim, device=device)
        cad = self.control_action_dim
        if control_dist is not None:
            a = control_dist.mean_sample() if mean_action else control_dist.sample()
            m = node_design_mask['execution'].to(device)
            action[m, :cad] = a
            action_env[m, :cad] = a
        if attr_dist is not None:
            a = attr_dist.mean_sample() if mean_action else attr_dist.sample()
            a = torch.clamp(a, -1, 1)
            a_env = a
            if self.enable_infer:
                a_env, a = self.transform_policy_output(a, self.body_ind_attribute, attr=True, attr_select=True,
                                                        num_nodes_cum=cum_attr)
                self.is_new_morphology = False
            m = node_design_mask['attr_trans'].to(device)
            action[m, cad:-1] = a
            action_env[m, cad:-1] = a_env
        if skel_dist is not None:
            a = skel_dist.mean_sample() if mean_action else skel_dist.sample()
            # the reference zeroes a[0] (policy.py:362) and is only ever called with one graph; for a batch the
            # root of every graph is forced to the no-op
            a[(self.body_ind_skeleleton == 0).to(a.device)] = 0
            a_env = a
            if self.enable_infer:
                a_env, a = self.transform_policy_output(a, self.body_ind_skeleleton)
                self.is_new_morphology = True
            m = node_design_mask['skel_trans'].to(device)
            action[m, -1] = a.to(action.dtype)
            action_env[m, -1] = a_env.to(action.dtype)
        return action.cpu().numpy().astype(np.float64), action_env.cpu().numpy().astype(np.float64)

    def orbit_rep_rows(self, body_ind) -> np.ndarray:
        """Row each row is overwritten with under the current orbits (rows of one or more whole graphs)."""
        levels, legs = body_levels_and_legs(body_ind, self.index_base)
        rep = np.arange(len(levels), dtype=np.int64)
        for r, orbit in self.group.get_cur_orbits().items():
            members = [int(o) for o in orbit]
            in_orbit = np.isin(legs, members)
            for l in range(1, int(levels.max(initial=0)) + 1):
                rows_rep = np.flatnonzero((legs == int(r)) & (levels == l))
                if len(rows_rep) == 0:
                    continue
                rows_orb = np.flatnonzero(in_orbit & (levels == l))
                if len(rows_orb) != len(rows_rep) * len(members):
                    raise RuntimeError('orbit tie: morphology is not symmetric under the current subgroup')
                rep[rows_orb] = np.repeat(rows_rep, len(members))
        return rep

    def transform_policy_output(self, outputs, body_ind, attr=False, attr_select=False, num_nodes_cum=None):
        """Orbit tie (+ subgroup symmetrizer on the xy offsets when ``attr_select``), policy.py:421-498.

        Host-side utility for rollouts and inspection; the training path applies the tie inside the
        fused loss kernels."""
        body = np.asarray(body_ind.cpu() if isinstance(body_ind, torch.Tensor) else body_ind)
        rep = torch.from_numpy(self.orbit_rep_rows(body)).to(outputs.device)
        outputs = outputs.clone()
        if attr:
            c0, c1 = self.tie_columns(1)
            outputs[:, c0:c1] = outputs[rep][:, c0:c1]
        else:
            outputs = outputs[rep]
        orbit_outputs = outputs.clone()
        if attr_select:
            levels, legs = body_levels_and_legs(body, self.index_base)
            sym = self.group.get_cur_subgroup_symmetrizer()
            # the reference symmetrises all rows of a level at once and is only called with one graph;
            # with several graphs (num_nodes_cum given) each graph is symmetrised on its own
            bounds = [0, len(body)] if num_nodes_cum is None else [0] + [int(c) for c in num_nodes_cum]
            graph_of = np.repeat(np.arange(len(bounds) - 1), np.diff(bounds))
            work = [(g, l) for g in range(len(bounds) - 1) for l in range(1, int(levels.max(initial=0)) + 1)]
            for g, l in work:
                rows = np.flatnonzero((levels == l) & (graph_of == g))
                if len(rows) == 0:
                    continue
                leg0 = legs[rows] - 1
                xy = outputs[rows, 0:2].t().double()                     # [2, k]
                acc = torch.zeros_like(xy)
                for e in sym:
                    mat = torch.tensor(e['matrix'], dtype=torch.float64, device=outputs.device)
                    pinv = torch.tensor(e['permutation_inv'][leg0][:, leg0], dtype=torch.float64, device=outputs.device)
                    acc = acc + (mat @ xy @ pinv) * e['ratio']
                outputs[rows, 0:2] = (acc.t() * (np.sqrt(2) / 2)).to(outputs.dtype)
        return outputs, orbit_outputs

    def update_symmetry(self, episode_reward):
        self.group.update_subgroup(episode_reward)

    def get_info(self):
        return self.group.get_info()

"""CPU oracle for the SARD PPO-update hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch on CPU (fp64 by default), the arithmetic of the
reference's PPO update over batched robot-morphology graphs, so that the CUDA path can be
checked on machines where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it; nothing under ``sard_b200/`` does.

Parity pin: every function below is checked against the reference's own code, imported
unmodified from ``/root/reference`` by ``tests/golden/make_golden.py``; the resulting vectors
are committed under ``tests/golden/`` and replayed by ``tests/test_oracle_golden.py``.
The one piece whose arithmetic lives outside the reference repo is
``torch_geometric.nn.GraphConv`` (torch-geometric==1.6.1 per the reference ``README.md:27``,
not vendored): its published 1.6.x algorithm ``lin_l(sum_{j->i} x_j) + lin_r(x_i)``
(aggr='add', bias on ``lin_l`` only) is restated in :func:`graph_conv`; that boundary is
"parity unpinned" by any reference test and is fixed here by declaration.

All functions take a flat ``state_dict``-style mapping with the reference's parameter names
(SURVEY.md section 8b), so reference checkpoints can be fed in directly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

STAGES = ('skel', 'attr', 'control')          # stage ids 0, 1, 2
_ACT = {'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid}


# --------------------------------------------------------------------------------------
# a4  RunningNorm   (khrylib/rl/core/running_norm.py:22-42)
# --------------------------------------------------------------------------------------
def running_norm(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, training: bool,
                 clip: float = 5.0) -> torch.Tensor:
    n, mean, var, std = (sd[prefix + k] for k in ('n', 'mean', 'var', 'std'))
    if training:
        with torch.no_grad():
            xd = x.detach()
            m = xd.shape[0]
            mean_x = xd.mean(dim=0)
            var_x = ((xd - mean_x) ** 2).mean(dim=0)          # biased, running_norm.py:23
            w = n.to(x.dtype) / (m + n).to(x.dtype)
            var.copy_(w * var + (1 - w) * var_x + w * (1 - w) * (mean_x - mean) ** 2)
            mean.copy_(w * mean + (1 - w) * mean_x)
            std.copy_(var.sqrt())
            n += m
    if int(n) > 0:
        x = torch.clamp((x - mean) / (std + 1e-8), -clip, clip)
    return x


# --------------------------------------------------------------------------------------
# a6  GraphConv (torch-geometric 1.6.1, restated)  /  a5 GNNSimple (design_opt/models/gnn.py:54-68)
# --------------------------------------------------------------------------------------
def graph_conv(sd, prefix: str, x: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    src, dst = edges[0], edges[1]
    agg = torch.zeros_like(x).index_add_(0, dst, x[src])
    out = agg @ sd[prefix + 'lin_l.weight'].t() + x @ sd[prefix + 'lin_r.weight'].t()
    if prefix + 'lin_l.bias' in sd:
        out = out + sd[prefix + 'lin_l.bias']
    return out


def gnn_simple(sd, prefix: str, x, edges, act: str = 'relu'):
    f = _ACT[act]
    x = x @ sd[prefix + 'in_fc.weight'].t() + sd[prefix + 'in_fc.bias']
    layer = 0
    while f'{prefix}gconv_layers.{layer}.lin_r.weight' in sd:
        x = f(graph_conv(sd, f'{prefix}gconv_layers.{layer}.', x, edges))
        layer += 1
    return x


# --------------------------------------------------------------------------------------
# a7  ObsEncoder (jsmlp.py:7-20), a8 IndexLinear / JSMLP (jsmlp.py:32-40, 74-79), MLP (mlp.py:22-25)
# --------------------------------------------------------------------------------------
def obs_encoder(sd, prefix: str, x):
    h = torch.relu(x @ sd[prefix + 'enc.0.weight'].t() + sd[prefix + 'enc.0.bias'])
    return torch.relu(h @ sd[prefix + 'enc.2.weight'].t() + sd[prefix + 'enc.2.bias'])


def index_linear(W, b, x, ind):
    """Row i uses the weight slab of its body index: ``b[ind_i] + W[ind_i] @ x_i``.

    Same schedule as the reference (one masked addmm per distinct index, jsmlp.py:33-40) so the
    CPU-baseline timing of this port is faithful; rows of absent indices are never touched.
    """
    out = x.new_zeros(x.shape[0], W.shape[1])
    for u in torch.unique(ind):
        rows = ind == u
        out[rows] = torch.addmm(b[u], x[rows], W[u].t())
    return out


def jsmlp(sd, prefix: str, x, ind, act: str = 'tanh'):
    f = _ACT[act]
    layer = 0
    while f'{prefix}affine_layers.{layer}.W' in sd:
        x = f(index_linear(sd[f'{prefix}affine_layers.{layer}.W'], sd[f'{prefix}affine_layers.{layer}.b'], x, ind))
        layer += 1
    return index_linear(sd[prefix + 'linear.W'], sd[prefix + 'linear.b'], x, ind)


def mlp(sd, prefix: str, x, act: str = 'tanh'):
    f = _ACT[act]
    layer = 0
    while f'{prefix}affine_layers.{layer}.weight' in sd:
        x = f(x @ sd[f'{prefix}affine_layers.{layer}.weight'].t() + sd[f'{prefix}affine_layers.{layer}.bias'])
        layer += 1
    return x


# --------------------------------------------------------------------------------------
# a2  batch_data (transform2act_policy.py:183-202, transform2act_critic.py:58-76)
# --------------------------------------------------------------------------------------
def batch_graphs(states: Sequence[list]):
    """Concatenate per-graph tensors; edges get each graph's node offset added."""
    num_nodes = np.concatenate([np.asarray(s[3]).reshape(1) for s in states]).astype(np.int64)
    cum = np.cumsum(num_nodes)
    start = cum - num_nodes
    obs = torch.cat([torch.as_tensor(s[0]) for s in states])
    edges = torch.cat([torch.as_tensor(s[1]) + int(o) for s, o in zip(states, start)], dim=1)
    stage = np.concatenate([np.asarray(s[2]).reshape(1) for s in states]).astype(np.int64)
    body = torch.from_numpy(np.concatenate([np.asarray(s[4]) for s in states]).astype(np.int64))
    ext = [torch.stack([torch.as_tensor(s[k]) for s in states]) for k in (5, 6, 7)]
    return obs, edges, stage, num_nodes, cum, body, ext


def ext_features(sd, prefix: str, ext, num_nodes):
    """a7: three ObsEncoders, concatenated per graph and repeated per node."""
    z = torch.cat([obs_encoder(sd, prefix + 'hfield_enc.', ext[0]),
                   obs_encoder(sd, prefix + 'obj_enc.', ext[1]),
                   obs_encoder(sd, prefix + 'goal_enc.', ext[2])], dim=-1)
    return torch.repeat_interleave(z, torch.as_tensor(num_nodes), dim=0)


# --------------------------------------------------------------------------------------
# a10  orbit tying (transform2act_policy.py:421-475, training mode: attr_select=False)
# --------------------------------------------------------------------------------------
def body_levels_and_legs(body_ind: np.ndarray, index_base: int):
    reps = [np.base_repr(int(b), index_base) for b in body_ind]
    levels = np.array([0 if r == '0' else len(r) for r in reps], dtype=np.int64)
    legs = np.array([int(r[-1], 36) for r in reps], dtype=np.int64)
    return levels, legs


def orbit_rep_index(body_ind: np.ndarray, orbits: Dict[int, Sequence[int]], index_base: int) -> np.ndarray:
    """Row map equivalent to the reference's in-place orbit overwrite.

    The reference, per orbit and per level l, selects in row order all rows whose leg digit is
    the representative (R rows) and all rows whose leg digit is in the orbit (R*|orbit| rows),
    and overwrites the k-th group of |orbit| selected rows with the k-th representative row
    (``policy.py:430-473``).  The returned ``rep[i]`` is the row that row i is overwritten with
    (``rep[i] == i`` when untouched).  Raises like the reference's reshape when counts mismatch.
    """
    levels, legs = body_levels_and_legs(body_ind, index_base)
    rep = np.arange(len(body_ind), dtype=np.int64)
    for r, orbit in orbits.items():
        orbit = [int(o) for o in orbit]
        in_orbit = np.isin(legs, orbit)
        for l in range(1, int(levels.max(initial=0)) + 1):
            rows_rep = np.flatnonzero((legs == int(r)) & (levels == l))
            if len(rows_rep) == 0:
                continue
            rows_orb = np.flatnonzero(in_orbit & (levels == l))
            if len(rows_orb) != len(rows_rep) * len(orbit):
                raise RuntimeError('orbit tie: morphology is not symmetric under the current subgroup')
            rep[rows_orb] = np.repeat(rows_rep, len(orbit))
    return rep


def orbit_tie(out: torch.Tensor, rep: np.ndarray, cols: slice) -> torch.Tensor:
    rep_t = torch.from_numpy(rep)
    tied = out.clone()
    tied[:, cols] = out[rep_t][:, cols]
    return tied


# --------------------------------------------------------------------------------------
# policy / value forward
# --------------------------------------------------------------------------------------
class Spec:
    """The few hyper-parameters the forward needs (mirrors agent/cfg attributes)."""

    def __init__(self, attr_fixed_dim: int, attr_design_dim: int = 6, index_base: int = 5,
                 gnn_act: str = 'relu', htype: str = 'tanh', enable_infer: bool = True,
                 embeds: Optional[torch.Tensor] = None, orbits: Optional[dict] = None,
                 control_action_dim: int = 1, tie_axis: bool = False,
                 design_flag_in_state: bool = True, onehot_design_flag: bool = True):
        self.attr_fixed_dim, self.attr_design_dim, self.index_base = attr_fixed_dim, attr_design_dim, index_base
        self.gnn_act, self.htype, self.enable_infer = gnn_act, htype, enable_infer
        self.embeds, self.orbits = embeds, orbits
        self.control_action_dim, self.tie_axis = control_action_dim, tie_axis
        self.design_flag_in_state, self.onehot_design_flag = design_flag_in_state, onehot_design_flag


def _design_obs(obs, spec: Spec):
    parts = [obs[:, :spec.attr_fixed_dim], obs[:, -spec.attr_design_dim:]]
    if spec.enable_infer:
        parts.append(spec.embeds.to(obs.dtype).unsqueeze(0).repeat(obs.shape[0], 1))
    return torch.cat(parts, dim=-1)


def policy_forward(sd, spec: Spec, states: Sequence[list], training: bool):
    """transform2act_policy.py:215-337.  Returns per-stage dicts (or None when the stage is absent)."""
    stage_of = np.array([int(np.asarray(s[2]).reshape(-1)[0]) for s in states])
    nn_all = np.array([int(np.asarray(s[3]).reshape(-1)[0]) for s in states])
    out = {'stage_of': stage_of, 'num_nodes': nn_all}
    # the reference evaluates execution, then attribute, then skeleton (RunningNorm order is per-stage anyway)
    for sid in (2, 1, 0):
        name = STAGES[sid]
        sel = [s for s, k in zip(states, stage_of) if k == sid]
        if not sel:
            out[name] = None
            continue
        obs, edges, _, num_nodes, cum, body, ext = batch_graphs(sel)
        x = obs if sid == 2 else _design_obs(obs, spec)
        x = running_norm(sd, f'{name}_norm.', x, training)
        x = gnn_simple(sd, f'{name}_gnn.', x, edges, spec.gnn_act)
        x = torch.cat([x, ext_features(sd, '', ext, num_nodes)], dim=-1)
        y = jsmlp(sd, f'{name}_ind_mlp.', x, body, spec.htype)
        res = {'num_nodes': num_nodes, 'cum': cum, 'body': body, 'raw': y}
        if sid == 2:
            res['mean'] = y
            res['std'] = sd['control_action_log_std'].expand_as(y).exp()
        elif sid == 1:
            mean = torch.sin(y * (0.5 * np.pi))
            if spec.enable_infer:
                rep = orbit_rep_index(body.numpy(), spec.orbits, spec.index_base)
                cols = slice(2, 3) if spec.tie_axis else slice(2, spec.attr_design_dim - 1)
                # with 'axis' in joint_params the reference ties cols [2:3] and [5:-1] (empty for d=6)
                mean = orbit_tie(mean, rep, cols)
                res['rep'] = rep
            res['mean'] = mean
            res['std'] = sd['attr_action_log_std'].expand_as(mean).exp()
        else:
            logits = y
            if spec.enable_infer:
                rep = orbit_rep_index(body.numpy(), spec.orbits, spec.index_base)
                logits = orbit_tie(logits, rep, slice(None))
                res['rep'] = rep
            res['logits'] = logits
        out[name] = res
    return out


def _segment_sum(v: torch.Tensor, cum: np.ndarray) -> torch.Tensor:
    """Per-graph sum of a per-node column (the reference's cumsum-and-difference, policy.py:393-395)."""
    c = torch.cumsum(v, dim=0)[torch.as_tensor(cum) - 1]
    return torch.cat([c[:1], c[1:] - c[:-1]])


def policy_log_prob(sd, spec: Spec, states, actions, training: bool, get_entropy: bool = False):
    """transform2act_policy.py:383-419."""
    fw = policy_forward(sd, spec, states, training)
    action = torch.cat([torch.as_tensor(a) for a in actions])
    stage_of, nn_all = fw['stage_of'], fw['num_nodes']
    node_stage = np.repeat(stage_of, nn_all)
    dtype = action.dtype
    lp = torch.zeros(len(states), 1, dtype=dtype)
    entropy = 0
    cad = spec.control_action_dim
    for sid in (2, 1, 0):
        r = fw[STAGES[sid]]
        if r is None:
            continue
        nmask = torch.from_numpy(node_stage == sid)
        gmask = torch.from_numpy(stage_of == sid)
        if sid == 0:
            a = action[nmask][:, -1].long()
            logp = torch.log_softmax(r['logits'], dim=-1)
            node_lp = logp.gather(1, a.unsqueeze(1))
            ent = -(logp.exp() * logp).sum(-1).mean()
        else:
            a = action[nmask][:, :cad] if sid == 2 else action[nmask][:, cad:-1]
            mean, std = r['mean'], r['std']
            node_lp = (-((a - mean) ** 2) / (2 * std ** 2) - std.log() - 0.5 * math.log(2 * math.pi)).sum(1, keepdim=True)
            ent = (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).mean()
        lp[gmask] = _segment_sum(node_lp, r['cum'])
        entropy = entropy + ent
    return (lp, entropy) if get_entropy else lp


def value_forward(sd, spec: Spec, states, training: bool):
    """transform2act_critic.py:78-108."""
    obs, edges, stage, num_nodes, cum, _, ext = batch_graphs(states)
    x = obs
    if spec.enable_infer:
        x = torch.cat([x, spec.embeds.to(obs.dtype).unsqueeze(0).repeat(obs.shape[0], 1)], dim=-1)
    if spec.design_flag_in_state:
        flag = torch.from_numpy(np.repeat(stage, num_nodes))
        if spec.onehot_design_flag:
            x = torch.cat([x, torch.nn.functional.one_hot(flag, 3).to(obs.dtype)], dim=-1)
        else:
            x = torch.cat([x, flag.unsqueeze(1).to(obs.dtype)], dim=-1)
    x = running_norm(sd, 'norm.', x, training)
    x = gnn_simple(sd, 'gnn.', x, edges, spec.gnn_act)
    x = mlp(sd, 'mlp.', x, spec.htype)
    x = torch.cat([x, ext_features(sd, '', ext, num_nodes)], dim=-1)
    v = x @ sd['value_head.weight'].t() + sd['value_head.bias']
    roots = torch.from_numpy(cum - num_nodes)
    return v[roots]


# --------------------------------------------------------------------------------------
# a15  GAE (khrylib/rl/core/common.py:5-25)   a17  PPO loss (transform2act_agent.py:390-409)
# --------------------------------------------------------------------------------------
def estimate_advantages(rewards, masks, values, gamma: float, tau: float):
    r = np.asarray(rewards.detach().cpu().numpy(), dtype=np.float64).reshape(-1)
    m = np.asarray(masks.detach().cpu().numpy(), dtype=np.float64).reshape(-1)
    v = np.asarray(values.detach().cpu().numpy(), dtype=np.float64).reshape(-1)
    adv = np.zeros_like(r)
    nxt_v, nxt_a = 0.0, 0.0
    for i in range(len(r) - 1, -1, -1):
        delta = r[i] + gamma * nxt_v * m[i] - v[i]
        adv[i] = delta + gamma * tau * nxt_a * m[i]
        nxt_v, nxt_a = v[i], adv[i]
    ret = v + adv
    adv = (adv - adv.mean()) / adv.std(ddof=1)            # torch.std is the unbiased estimator
    to = lambda a: torch.from_numpy(a).to(values.dtype).unsqueeze(1)
    return to(adv), to(ret)


def ppo_loss(log_probs, fixed_log_probs, advantages, clip_epsilon: float):
    ratio = torch.exp(log_probs - fixed_log_probs)
    surr = torch.min(ratio * advantages, torch.clamp(ratio, 1 - clip_epsilon, 1 + clip_epsilon) * advantages)
    loss = -surr.mean()
    kl = float((fixed_log_probs - log_probs).detach().abs().mean())
    valid = not (kl > 0.1 or float(loss) > 10)
    return loss, kl, valid


# --------------------------------------------------------------------------------------
# a16/a18/a19/a20  the update driver (transform2act_agent.py:274-387, agent_pg.py:18-25, agent_ppo.py:53-56)
# --------------------------------------------------------------------------------------
def stage_partition(states) -> np.ndarray:
    """``get_perm_batch_design``: stable partition of positions by stage 0|1|2 (agent :305-311)."""
    st = np.array([int(np.asarray(s[2]).reshape(-1)[0]) for s in states])
    return np.argsort(st, kind='stable')


class OracleAgent:
    """Reference-equivalent PPO update on CPU with autograd + torch.optim.Adam.

    ``policy_sd`` / ``value_sd`` are mutated in place like the reference's modules.  Floating
    tensors that the reference registers as ``nn.Parameter`` must have ``requires_grad=True``.
    """

    def __init__(self, policy_sd, value_sd, spec: Spec, *, gamma=0.995, tau=0.95, clip_epsilon=0.2,
                 policy_lr=5e-5, value_lr=3e-4, mini_batch_size=2048, opt_num_epochs=10,
                 batch_design=True, policy_grad_clip=40.0, clip_first_step_only=True,
                 zero_grad_to_none=True):
        self.policy_sd, self.value_sd, self.spec = policy_sd, value_sd, spec
        self.gamma, self.tau, self.clip_epsilon = gamma, tau, clip_epsilon
        self.mini_batch_size, self.opt_num_epochs, self.batch_design = mini_batch_size, opt_num_epochs, batch_design
        self.policy_grad_clip = policy_grad_clip
        # The reference hands clip_grad_norm_ a *generator* (policy_net.parameters(), agent :48),
        # which is exhausted by the first call: only the first policy step of a run is clipped
        # (agent_ppo.py:53-56).  Mirrored here; set False to clip every step.
        self.clip_first_step_only = clip_first_step_only
        # torch >= 2.0 zero_grad() drops the gradients (a stage absent from a minibatch is skipped by Adam); the
        # reference pins torch 1.8 (README.md:14) where they stay zero tensors: zero_grad_to_none=False restates that
        self.zero_grad_to_none = zero_grad_to_none
        self._clip_calls = 0
        self.policy_params = [p for p in policy_sd.values() if p.requires_grad]
        self.value_params = [p for p in value_sd.values() if p.requires_grad]
        self.optimizer_policy = torch.optim.Adam(self.policy_params, lr=policy_lr)
        self.optimizer_value = torch.optim.Adam(self.value_params, lr=value_lr)

    def update_value(self, states, returns):
        pred = value_forward(self.value_sd, self.spec, states, training=True)
        loss = (pred - returns).pow(2).mean()
        self.optimizer_value.zero_grad()
        loss.backward()
        self.optimizer_value.step()
        return float(loss)

    def policy_step(self, states, actions, advantages, fixed_log_probs):
        lp, ent = policy_log_prob(self.policy_sd, self.spec, states, actions, training=True, get_entropy=True)
        loss, kl, valid = ppo_loss(lp, fixed_log_probs, advantages, self.clip_epsilon)
        log = {'surr_loss': float(loss), 'entropy': float(ent), 'approx_kl': kl}
        if valid:
            self.optimizer_policy.zero_grad(set_to_none=self.zero_grad_to_none)
            loss.backward()
            if not (self.clip_first_step_only and self._clip_calls > 0):
                torch.nn.utils.clip_grad_norm_(self.policy_params, self.policy_grad_clip)
            self._clip_calls += 1
            self.optimizer_policy.step()
        return log, valid

    def eval_values(self, states, chunk=10000):
        with torch.no_grad():
            return torch.cat([value_forward(self.value_sd, self.spec, states[i:i + chunk], training=False)
                              for i in range(0, len(states), chunk)])

    def eval_log_probs(self, states, actions, chunk=10000):
        with torch.no_grad():
            return torch.cat([policy_log_prob(self.policy_sd, self.spec, states[i:i + chunk], actions[i:i + chunk], training=False)
                              for i in range(0, len(states), chunk)])

    def update_policy(self, states, actions, returns, advantages, perms: Optional[List[np.ndarray]] = None):
        fixed = self.eval_log_probs(states, actions)
        T = len(states)
        logs = []
        for ep in range(self.opt_num_epochs):
            perm = perms[ep] if perms is not None else np.random.permutation(T)
            if self.batch_design:
                perm = perm[stage_partition([states[i] for i in perm])]
            sts = [states[i] for i in perm]
            acts = [actions[i] for i in perm]
            pt = torch.from_numpy(np.asarray(perm))
            ret, adv, fx = returns[pt], advantages[pt], fixed[pt]
            # like the reference, later epochs permute the already permuted lists
            states, actions, returns, advantages, fixed = sts, acts, ret, adv, fx
            for i in range(T // self.mini_batch_size):
                sl = slice(i * self.mini_batch_size, (i + 1) * self.mini_batch_size)
                self.update_value(states[sl], returns[sl])
                log, valid = self.policy_step(states[sl], actions[sl], advantages[sl], fixed[sl])
                if valid:
                    logs.append(log)
        return logs

    def update_params(self, states, actions, rewards, masks, perms=None):
        values = self.eval_values(states)
        advantages, returns = estimate_advantages(rewards, masks, values, self.gamma, self.tau)
        logs = self.update_policy(states, actions, returns, advantages, perms)
        return values, advantages, returns, logs

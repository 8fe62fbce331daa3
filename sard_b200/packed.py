"""Rollout arena in HBM and the per-epoch minibatch plan.

Replaces, for one ``update_params`` call, the reference's per-state ``tensorfy``
(``design_opt/agents/transform2act_agent.py:20-24``, 5*T small host->device copies), the
per-forward ``batch_data`` (``design_opt/models/transform2act_policy.py:183-202``,
``transform2act_critic.py:58-76``) and the Python list shuffles of ``update_policy``
(``transform2act_agent.py:331-342``, ``khrylib/utils/tools.py:70-71``): the whole batch is packed
once into flat arrays, uploaded with a handful of copies from pinned memory, and minibatches
become index ranges over a device-resident permutation.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import os

import numpy as np
import torch

from . import _lib
from ._lib import SardArena, SardSel, lib, check, ptr, stream
from .synthetic import HostRollout, ACTION_DIM


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def host_rollout_from_lists(states: Sequence[list], actions: Optional[Sequence] = None,
                            rewards=None, masks=None, exps=None) -> HostRollout:
    """Flatten the reference's list-of-8-slot-states format (derl_env.py:377-396) into flat arrays."""
    T = len(states)
    num_nodes = np.fromiter((int(_np(s[3]).reshape(-1)[0]) for s in states), dtype=np.int64, count=T)
    stage = np.fromiter((int(_np(s[2]).reshape(-1)[0]) for s in states), dtype=np.int64, count=T)
    obs = np.concatenate([_np(s[0]) for s in states], axis=0)
    edge_list = [_np(s[1]) for s in states]
    ne = np.fromiter((e.shape[1] for e in edge_list), dtype=np.int64, count=T)
    edges = np.concatenate(edge_list, axis=1).astype(np.int64, copy=False)
    body = np.concatenate([_np(s[4]).reshape(-1) for s in states]).astype(np.int64, copy=False)
    hf = np.stack([_np(s[5]).reshape(-1) for s in states])
    ob = np.stack([_np(s[6]).reshape(-1) for s in states])
    go = np.stack([_np(s[7]).reshape(-1) for s in states])
    n_tot = int(num_nodes.sum())
    if actions is not None:
        act = np.concatenate([_np(a) for a in actions], axis=0)
    else:
        act = np.zeros((n_tot, ACTION_DIM), dtype=obs.dtype)
    z = lambda v: np.zeros(T) if v is None else _np(v).reshape(-1)
    return HostRollout(
        shape=None, obs=obs, node_ptr=np.concatenate([[0], np.cumsum(num_nodes)]).astype(np.int64),
        edges=edges, edge_ptr=np.concatenate([[0], np.cumsum(ne)]).astype(np.int64), stage=stage,
        body_ind=body, hfield=hf, obj=ob, goal=go, actions=act, rewards=z(rewards), masks=z(masks),
        exps=z(exps))


try:                                    # native list walker (sard_b200/csrc/hostpack.cpp); pure-Python packing otherwise
    from . import _hostpack
except ImportError:                     # pragma: no cover
    _hostpack = None

_F32 = ('obs', 'actions', 'hfield', 'obj', 'goal', 'rewards', 'masks')
_I32 = ('node_ptr', 'edge_ptr', 'esrc', 'edst', 'stage', 'body_ind')


def host_tensors_from_rollout(host: HostRollout, pin: bool = True) -> dict:
    """fp32 / int32 host tensors (pinned when asked) of a flat ``HostRollout``."""
    src = {'obs': host.obs, 'actions': host.actions, 'hfield': host.hfield, 'obj': host.obj, 'goal': host.goal,
           'rewards': host.rewards, 'masks': host.masks, 'node_ptr': host.node_ptr, 'edge_ptr': host.edge_ptr,
           'esrc': host.edges[0], 'edst': host.edges[1], 'stage': host.stage, 'body_ind': host.body_ind}
    out = {}
    for k, a in src.items():
        a = np.asarray(a)
        dt = torch.float32 if k in _F32 else torch.int32
        t = torch.empty(a.shape, dtype=dt, pin_memory=pin)
        np.copyto(t.numpy(), a, casting='unsafe')          # one pass: cast straight into the (pinned) staging buffer
        out[k] = t
    return out


class _PinnedPool:
    """Pinned staging memory reused from one ``update_params`` to the next: page-locking 150 MB per call
    (cudaHostAlloc) cost 5-10 ms, sometimes hundreds.  One upload is in flight at most: ``wait()`` blocks until the
    event recorded after the previous upload has passed before the buffers are overwritten."""

    def __init__(self):
        self.buf = {}
        self.event = None
        self.seq = 0
        self.owner = None          # id() of the host dict whose tensors live in the pool

    def wait(self):
        if self.event is not None:
            self.event.synchronize()
            self.event = None
        self.seq = 0

    def take(self, shape, dtype):
        n = int(np.prod(shape)) if len(shape) else 1
        key = (self.seq, dtype)
        self.seq += 1
        b = self.buf.get(key)
        if b is None or b.numel() < n:
            b = torch.empty(max(int(n * 1.2), 16), dtype=dtype, pin_memory=True)
            self.buf[key] = b
        return b[:n].view(shape)

    def uploaded(self, device):
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(device))


_PINNED = _PinnedPool()


def host_tensors_from_lists(states, actions=None, rewards=None, masks=None, pin: bool = True, nthreads: int = 0):
    """Same from the reference's list-of-states batch; uses the native packer when it is built."""
    if _hostpack is None or isinstance(states[0][0], torch.Tensor):
        return host_tensors_from_rollout(host_rollout_from_lists(states, actions, rewards, masks), pin)
    if nthreads <= 0:
        nthreads = int(os.environ.get('SARD_PACK_THREADS', min(16, os.cpu_count() or 8)))
    handle, T, n, e, d, H, O, G = _hostpack.collect(states, actions)       # one walk over the Python objects
    _PINNED.wait()                                                          # the previous upload from these buffers is over
    mk = (lambda shape, dt: _PINNED.take(shape, dt)) if pin else (lambda shape, dt: torch.empty(shape, dtype=dt))
    out = {'obs': mk((n, d), torch.float32), 'actions': mk((n, 8), torch.float32), 'node_ptr': mk((T + 1,), torch.int32),
           'edge_ptr': mk((T + 1,), torch.int32), 'esrc': mk((e,), torch.int32), 'edst': mk((e,), torch.int32),
           'stage': mk((T,), torch.int32), 'body_ind': mk((n,), torch.int32), 'hfield': mk((T, H), torch.float32),
           'obj': mk((T, O), torch.float32), 'goal': mk((T, G), torch.float32)}
    _hostpack.fill_collected(handle, *[out[k].numpy() for k in ('obs', 'actions', 'node_ptr', 'edge_ptr', 'esrc', 'edst', 'stage',
                                                                'body_ind', 'hfield', 'obj', 'goal')], nthreads)
    for k, v in (('rewards', rewards), ('masks', masks)):
        t = mk((T,), torch.float32)
        if v is None:
            t.zero_()
        else:
            np.copyto(t.numpy(), _np(v).reshape(-1), casting='unsafe')
        out[k] = t
    _PINNED.owner = id(out) if pin else None
    return out


class PackedRollout:
    """One rollout batch resident on the device (the ``SardArena`` of include/sard_b200.h)."""

    def __init__(self, host, device, pin: bool = True):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.SardError('PackedRollout needs a CUDA device: the SARD B200 path has no CPU implementation')
        self.device = device
        ht = host if isinstance(host, dict) else host_tensors_from_rollout(host, pin)
        pooled = isinstance(ht, dict) and _PINNED.owner == id(ht)
        self.T = int(ht['stage'].shape[0])
        self.n_total = int(ht['node_ptr'][-1])
        self.e_total = int(ht['edge_ptr'][-1])
        self.d_obs = int(ht['obs'].shape[1])
        # host-side integer copies used for planning (no device round trips in the update loop)
        self.stage_np = ht['stage'].numpy().astype(np.int64)
        self.node_ptr_np = ht['node_ptr'].numpy().astype(np.int64)
        self.num_nodes_np = np.diff(self.node_ptr_np)
        self.body_np = ht['body_ind'].numpy().astype(np.int64)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in ht.values())
        up = lambda k: ht[k].to(device, non_blocking=True)
        self._host = ht                                     # keep the pinned staging alive until the copies finish
        self.obs, self.actions = up('obs'), up('actions')
        self.node_ptr, self.edge_ptr = up('node_ptr'), up('edge_ptr')
        self.esrc, self.edst = up('esrc'), up('edst')
        self.stage, self.body_ind = up('stage'), up('body_ind')
        self.ext = [up('hfield'), up('obj'), up('goal')]
        self.rewards, self.masks = up('rewards'), up('masks')
        if pooled:
            _PINNED.uploaded(device)                        # the pool's buffers may be rewritten once this has passed
        i32 = dict(dtype=torch.int32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.in_ptr = torch.empty(self.n_total + 1, **i32)
        self.in_nbr = torch.empty(max(self.e_total, 1), **i32)
        self.out_ptr = torch.empty(self.n_total + 1, **i32)
        self.out_nbr = torch.empty(max(self.e_total, 1), **i32)
        self.rep_local = torch.zeros(self.n_total, **i32)
        self.values = torch.zeros(self.T, **f32)
        self.returns = torch.zeros(self.T, **f32)
        self.advantages = torch.zeros(self.T, **f32)
        self.fixed_log_probs = torch.zeros(self.T, **f32)
        self._err = torch.zeros(1, **i32)
        self.has_orbits = False
        check(lib.sard_build_csr(ptr(self.node_ptr), ptr(self.edge_ptr), ptr(self.esrc), ptr(self.edst), self.T,
                                 ptr(self.in_ptr), ptr(self.in_nbr), ptr(self.out_ptr), ptr(self.out_nbr), stream()))
        self.c = self._make_struct()

    @classmethod
    def from_lists(cls, states, actions=None, rewards=None, masks=None, exps=None, device='cuda', pin=False):
        return cls(host_tensors_from_lists(states, actions, rewards, masks, pin=pin), device, pin=pin)

    def _make_struct(self) -> SardArena:
        a = SardArena()
        a.T, a.n_total, a.e_total, a.d_obs = self.T, self.n_total, self.e_total, self.d_obs
        a.obs, a.actions = ptr(self.obs), ptr(self.actions)
        a.node_ptr, a.stage, a.body_ind = ptr(self.node_ptr), ptr(self.stage), ptr(self.body_ind)
        a.in_ptr, a.in_nbr, a.out_ptr, a.out_nbr = ptr(self.in_ptr), ptr(self.in_nbr), ptr(self.out_ptr), ptr(self.out_nbr)
        a.rep_local = ptr(self.rep_local)
        for k in range(3):
            a.ext[k] = ptr(self.ext[k])
            a.ext_dim[k] = int(self.ext[k].shape[1])
        a.values, a.returns = ptr(self.values), ptr(self.returns)
        a.advantages, a.fixed_log_probs = ptr(self.advantages), ptr(self.fixed_log_probs)
        a.max_nodes = int(self.num_nodes_np.max()) if self.T > 0 else 0
        return a

    def set_orbits(self, leg_rep: np.ndarray, index_base: int, verify: bool = True):
        """Orbit representatives for the current design subgroup (``sard_orbit_rep``)."""
        lr = torch.from_numpy(np.ascontiguousarray(leg_rep, dtype=np.int32)).to(self.device)
        self._err.zero_()
        check(lib.sard_orbit_rep(ptr(self.node_ptr), ptr(self.body_ind), self.T, ptr(lr), int(index_base),
                                 ptr(self.rep_local), ptr(self._err), stream()))
        self.has_orbits = True
        if verify and int(self._err.item()) != 0:
            raise _lib.SardError('orbit tie: a morphology is not symmetric under the current subgroup '
                                 '(the reference would fail to reshape, transform2act_policy.py:449)')

    def live_indices(self, stage: int) -> np.ndarray:
        if getattr(self, '_live', None) is None:          # one histogram over (stage, body index) for the three stages
            node_stage = np.repeat(self.stage_np, self.num_nodes_np)
            hi = int(self.body_np.max()) + 1 if len(self.body_np) else 1
            if len(self.body_np) and int(self.body_np.min()) < 0:
                raise _lib.SardError('negative body index')
            cnt = np.bincount(node_stage * hi + self.body_np, minlength=3 * hi).reshape(3, hi)
            self._live = [np.flatnonzero(cnt[s]) for s in range(3)]
        return self._live[stage]


class Run:
    """A contiguous range of an ordering: graphs ``[lo, hi)`` of ``ids`` with node prefix ``cum``."""
    __slots__ = ('sel', 'B', 'n', 'lo', 'hi', '_keep')

    def __init__(self, ids_t: torch.Tensor, cum_t: torch.Tensor, cum_np: np.ndarray, lo: int, hi: int):
        self.lo, self.hi = lo, hi
        self._keep = (ids_t, cum_t)          # the selection holds raw pointers into these tensors
        self.B = hi - lo
        self.n = int(cum_np[hi] - cum_np[lo])
        s = SardSel()
        s.ids = ids_t.data_ptr() + 4 * lo
        s.cum = cum_t.data_ptr() + 4 * lo
        s.B, s.n, s.node_base = self.B, self.n, int(cum_np[lo])
        self.sel = s


class Ordering:
    """A device-resident ordering of arena graphs plus its node prefix sums."""

    def __init__(self, arena: PackedRollout, ids_np: np.ndarray):
        self.ids_np = np.ascontiguousarray(ids_np, dtype=np.int64)
        self.cum_np = np.concatenate([[0], np.cumsum(arena.num_nodes_np[self.ids_np])]).astype(np.int64)
        both = np.concatenate([self.ids_np, self.cum_np]).astype(np.int32)
        t = torch.from_numpy(both).to(arena.device, non_blocking=True)
        n = len(self.ids_np)
        self.ids, self.cum = t[:n], t[n:]
        self.stage_np = arena.stage_np[self.ids_np]
        self.h2d_bytes = both.nbytes

    def run(self, lo: int, hi: int) -> Run:
        return Run(self.ids, self.cum, self.cum_np, lo, hi)

    def stage_runs(self, lo: int, hi: int) -> List[Optional[Run]]:
        """Sub-runs of [lo,hi) per stage; requires the range to be stage-sorted (0|1|2)."""
        st = self.stage_np[lo:hi]
        b1 = lo + int(np.searchsorted(st, 1, side='left'))
        b2 = lo + int(np.searchsorted(st, 2, side='left'))
        bounds = [(lo, b1), (b1, b2), (b2, hi)]
        return [self.run(a, b) if b > a else None for a, b in bounds]


def stage_partition(stage: np.ndarray) -> np.ndarray:
    """``get_perm_batch_design`` (transform2act_agent.py:305-311): stable partition by stage."""
    return np.argsort(stage, kind='stable')


class EpochPlan:
    """Minibatch index ranges of one optimisation epoch (transform2act_agent.py:329-347).

    ``perm`` is the epoch's shuffle (``np.random.shuffle`` order); with ``batch_design`` it is
    stably partitioned by stage (:338-342).  The value net consumes minibatch ``k`` as the range
    ``[k*M,(k+1)*M)`` of that order; the policy consumes the same graphs stably re-sorted by stage
    inside the minibatch (the split its ``forward`` does at policy.py:216-231).
    """

    def __init__(self, arena: PackedRollout, perm: np.ndarray, mini_batch_size: int, batch_design: bool,
                 n_mb: Optional[int] = None):
        """``n_mb`` caps the number of minibatches (data parallel: the minimum over ranks, so that every rank issues
        the same collectives when the rank-local rollouts differ in length)."""
        perm = np.asarray(perm, dtype=np.int64)
        if batch_design:
            perm = perm[stage_partition(arena.stage_np[perm])]
        self.perm = perm
        T = len(perm)
        self.M = int(mini_batch_size)
        self.n_mb = T // self.M if n_mb is None else min(int(n_mb), T // self.M)
        self.value_order = Ordering(arena, perm)
        st = arena.stage_np[perm]
        used = self.n_mb * self.M
        if self._stage_sorted_per_mb(st):
            self.policy_order = self.value_order
        else:
            key = (np.arange(used) // self.M) * 4 + st[:used]
            pperm = np.concatenate([perm[:used][np.argsort(key, kind='stable')], perm[used:]])
            self.policy_order = Ordering(arena, pperm)
        self.h2d_bytes = self.value_order.h2d_bytes + (0 if self.policy_order is self.value_order else self.policy_order.h2d_bytes)
        # host-known per-minibatch counts: graphs per stage and skel nodes (entropy normaliser)
        pst = self.policy_order.stage_np
        self.stage_graphs = np.zeros((self.n_mb, 3), dtype=np.int64)
        self.stage_nodes = np.zeros((self.n_mb, 3), dtype=np.int64)
        nn = arena.num_nodes_np[self.policy_order.ids_np]
        for k in range(self.n_mb):
            sl = slice(k * self.M, (k + 1) * self.M)
            for s in range(3):
                m = pst[sl] == s
                self.stage_graphs[k, s] = int(m.sum())
                self.stage_nodes[k, s] = int(nn[sl][m].sum())

    def _stage_sorted_per_mb(self, st: np.ndarray) -> bool:
        for k in range(self.n_mb):
            s = st[k * self.M:(k + 1) * self.M]
            if np.any(np.diff(s) < 0):
                return False
        return True

    def value_run(self, k: int) -> Run:
        return self.value_order.run(k * self.M, (k + 1) * self.M)

    def policy_runs(self, k: int) -> List[Optional[Run]]:
        return self.policy_order.stage_runs(k * self.M, (k + 1) * self.M)

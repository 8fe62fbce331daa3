"""Parameter holders with the reference's ``state_dict`` layout, flat parameter storage, and the
descriptors the C++ schedules consume.

The holder modules mirror the reference's module tree name for name so checkpoints interchange
(SURVEY.md section 8b): ``RunningNorm`` (khrylib/rl/core/running_norm.py), ``GNNSimple`` with
PyG-1.6.1 ``GraphConv`` children ``lin_l``/``lin_r`` (design_opt/models/gnn.py), ``ObsEncoder``,
``IndexLinear``, ``JSMLP`` (design_opt/models/jsmlp.py), ``MLP`` (khrylib/models/mlp.py).
They only *hold* parameters: all arithmetic happens in libsard_b200.so, reached through
:mod:`sard_b200._lib`.  Dense parameters of a net are views into one flat fp32 buffer so that the
gradient bucket, the clip norm and Adam are single launches.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from . import _lib
from ._lib import (ACT, EXT_DIM, MAX_IL, MAX_LAYERS, MAX_MLP, SardAdamSeg, SardDense, SardIndexLinear,
                   SardPolicyNet, SardPolicyStage, SardTower, SardValueNet, check, lib, ptr, stream)

STAGE_NAMES = ('skel', 'attr', 'control')
SEG_MAX = 4096        # elements per Adam segment: one block row of adam_apply_kernel covers it in one float4 pass


class RunningNorm(nn.Module):
    """Buffers of the reference's RunningNorm; updated on device by ``sard_norm_update``."""

    def __init__(self, dim, demean=True, destd=True, clip=5.0):
        super().__init__()
        if not (demean and destd and clip == 5.0):
            raise NotImplementedError('only the default RunningNorm configuration is on the CUDA path')
        self.dim = dim
        self.register_buffer('n', torch.tensor(0, dtype=torch.long))
        self.register_buffer('mean', torch.zeros(dim))
        self.register_buffer('var', torch.zeros(dim))
        self.register_buffer('std', torch.zeros(dim))
        self.register_buffer('inv', torch.zeros(dim), persistent=False)


class GraphConv(nn.Module):
    """Parameter layout of torch_geometric 1.6.1 GraphConv: ``lin_l`` (bias) on the aggregate, ``lin_r`` on x."""

    def __init__(self, in_channels, out_channels, aggr='add', bias=True, **kwargs):
        super().__init__()
        if aggr != 'add':
            raise NotImplementedError("GraphConv aggr must be 'add' (design_opt/cfg/derl.yml)")
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)


class GNNSimple(nn.Module):
    def __init__(self, in_dim, cfg, node_dim=0):
        super().__init__()
        self.cfg = cfg
        self.hdims = hdims = list(cfg['hdims'])
        self.num_layers = len(hdims)
        if cfg.get('residual', False) or cfg.get('cat_input', False) or cfg.get('num_layer_update', 1) != 1:
            raise NotImplementedError('GNNSimple residual / cat_input / num_layer_update are not on the CUDA path')
        if cfg['layer_type'] != 'graph_conv':
            raise NotImplementedError("only layer_type 'graph_conv' is on the CUDA path")
        if len(hdims) > MAX_LAYERS:
            raise NotImplementedError('too many GNN layers')
        self.act = cfg.get('act', 'relu')
        self.in_fc = nn.Linear(in_dim, hdims[0])
        self.out_dim = hdims[-1]
        self.gconv_layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.gconv_layers.append(GraphConv(hdims[0] if i == 0 else hdims[i - 1], hdims[i],
                                               aggr=cfg['aggr'], bias=cfg['bias']))


class ObsEncoder(nn.Module):
    def __init__(self, input_dim, out_dim=EXT_DIM):
        super().__init__()
        if out_dim != EXT_DIM:
            raise NotImplementedError('ObsEncoder out_dim is fixed at 64')
        self.input_dim, self.out_dim = input_dim, out_dim
        self.enc = nn.Sequential(nn.Linear(input_dim, 64), nn.ReLU(), nn.Linear(64, out_dim), nn.ReLU())


class IndexLinear(nn.Module):
    def __init__(self, input_dim, out_dim, max_index=256, zero_init=False):
        super().__init__()
        self.input_dim, self.out_dim, self.max_index = input_dim, out_dim, max_index
        self.W = nn.Parameter(torch.zeros(max_index, out_dim, input_dim))
        self.b = nn.Parameter(torch.zeros(max_index, out_dim))
        if not zero_init:
            init.kaiming_uniform_(self.W, a=math.sqrt(5))
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.W)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.b, -bound, bound)


class JSMLP(nn.Module):
    def __init__(self, input_dim, hidden_dims=(128, 128), linear_dim=None, max_index=256, activation='tanh',
                 rescale_linear=False, zero_init=False):
        super().__init__()
        self.activation = activation
        self.affine_layers = nn.ModuleList()
        cur = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(IndexLinear(cur, nh, max_index, zero_init))
            cur = nh
        if linear_dim is None:
            raise NotImplementedError('JSMLP without the final linear layer is not on the CUDA path')
        self.linear = IndexLinear(cur, linear_dim, max_index, zero_init)
        if rescale_linear:
            self.linear.W.data.mul_(0.1)
            self.linear.b.data.mul_(0.0)
        self.out_dim = linear_dim
        self.max_index = max_index
        if len(hidden_dims) + 1 > MAX_IL:
            raise NotImplementedError('too many JSMLP layers')

    def layers(self) -> List[IndexLinear]:
        return list(self.affine_layers) + [self.linear]


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dims=(128, 128), activation='tanh'):
        super().__init__()
        self.activation = activation
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh
        if len(hidden_dims) > MAX_MLP:
            raise NotImplementedError('too many MLP layers')


def init_fc_weights(fc):
    fc.weight.data.mul_(0.1)
    fc.bias.data.mul_(0.0)


# ------------------------------------------------------------------------------------------
class FlatStore:
    """Dense parameters of a net laid out in one fp32 buffer, grouped (one Adam step counter per group)."""

    def __init__(self, groups: Sequence[Sequence[Tuple[str, nn.Parameter]]]):
        self.entries: List[Tuple[str, nn.Parameter, int, int]] = []   # name, param, offset, group
        self.group_range: List[Tuple[int, int]] = []
        off = 0
        for g, plist in enumerate(groups):
            start = off
            for name, p in plist:
                self.entries.append((name, p, off, g))
                off += (p.numel() + 3) // 4 * 4
            self.group_range.append((start, off))
        self.size = off
        self.flat = None
        self.offset: Dict[int, int] = {id(p): o for _, p, o, _ in self.entries}

    def adopt(self, device):
        """(Re)point every parameter at its slice of the flat buffer, keeping current values."""
        flat = torch.zeros(max(self.size, 4), dtype=torch.float32, device=device)
        for _, p, off, _ in self.entries:
            flat[off:off + p.numel()].copy_(p.data.detach().reshape(-1).to(device=device, dtype=torch.float32))
        for _, p, off, _ in self.entries:
            p.data = flat[off:off + p.numel()].view(p.shape)
        self.flat = flat

    def is_adopted(self) -> bool:
        if self.flat is None or not self.entries:
            return self.flat is not None
        return all(p.data.data_ptr() == self.flat.data_ptr() + 4 * off and p.data.dtype == torch.float32
                   for _, p, off, _ in self.entries)

    def off(self, p: nn.Parameter) -> int:
        return self.offset[id(p)]


def _dense(store: FlatStore, lin: nn.Linear, act: str = 'none') -> SardDense:
    d = SardDense()
    d.out, d.in_, d.act = lin.out_features, lin.in_features, ACT[act]
    d.w = store.off(lin.weight)
    d.b = store.off(lin.bias) if lin.bias is not None else -1
    return d


def _tower(store: FlatStore, norm: RunningNorm, gnn: GNNSimple) -> SardTower:
    t = SardTower()
    t.D, t.L, t.act = norm.dim, gnn.num_layers, ACT[gnn.act]
    t.has_bias = 1 if gnn.gconv_layers[0].lin_l.bias is not None else 0
    t.h[0] = gnn.hdims[0]
    for l, hd in enumerate(gnn.hdims):
        t.h[l + 1] = hd
    t.in_w, t.in_b = store.off(gnn.in_fc.weight), store.off(gnn.in_fc.bias)
    for l, gc in enumerate(gnn.gconv_layers):
        t.wl[l] = store.off(gc.lin_l.weight)
        t.bl[l] = store.off(gc.lin_l.bias) if gc.lin_l.bias is not None else -1
        t.wr[l] = store.off(gc.lin_r.weight)
    t.mean, t.var, t.std, t.inv, t.count = ptr(norm.mean), ptr(norm.var), ptr(norm.std), ptr(norm.inv), ptr(norm.n)
    return t


def _encoders(store: FlatStore, encs: Sequence[ObsEncoder], dst):
    for k, e in enumerate(encs):
        dst[k][0] = _dense(store, e.enc[0], 'relu')
        dst[k][1] = _dense(store, e.enc[2], 'relu')


def _params_of(mod: nn.Module, prefix: str):
    return [(prefix + n, p) for n, p in mod.named_parameters()]


class AdamState:
    """Adam moments for a net: flat buffers for the dense part, table-shaped buffers for IndexLinear."""

    def __init__(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.param_groups = [{'lr': self.lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': 0.0}]
        self.m = self.v = None
        self.table_m: Dict[int, torch.Tensor] = {}
        self.table_v: Dict[int, torch.Tensor] = {}
        self.ctl_f = self.ctl_i = None
        self.segs = None
        self.n_seg = 0
        self.max_seg = 0

    def ensure(self, flat: torch.Tensor):
        if self.m is None or self.m.shape != flat.shape or self.m.device != flat.device:
            self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
        if self.ctl_f is None or self.ctl_f.device != flat.device:
            self.ctl_f = torch.zeros(_lib.CTL_FLOATS, dtype=torch.float32, device=flat.device)
            self.ctl_i = torch.zeros(_lib.CTL_INTS, dtype=torch.int32, device=flat.device)
            self.ctl_i[1] = 1     # clip armed (see sard_adam_prepare)

    def table_state(self, p: nn.Parameter):
        k = id(p)
        if k not in self.table_m or self.table_m[k].device != p.device:
            self.table_m[k], self.table_v[k] = torch.zeros_like(p.data), torch.zeros_like(p.data)
        return self.table_m[k], self.table_v[k]

    def set_segments(self, seg_list: List[Tuple[int, int, int, int, int, int]], device):
        """seg_list rows: (p_ptr, m_ptr, v_ptr, g_ptr, n, group)."""
        arr = (SardAdamSeg * max(len(seg_list), 1))()
        for i, (p, m, v, g, n, grp) in enumerate(seg_list):
            arr[i].p, arr[i].m, arr[i].v, arr[i].g, arr[i].n, arr[i].group = p, m, v, g, n, grp
        raw = np.frombuffer(arr, dtype=np.uint8).copy()
        self.segs = torch.from_numpy(raw).to(device)
        self.n_seg = len(seg_list)
        self.max_seg = max([s[4] for s in seg_list], default=0)

    def zero_grad(self, set_to_none=True):   # API compatibility; gradients live in the engine's bucket
        pass

    # -- checkpointing (the reference omits optimiser state, transform2act_agent.py:192-199) -----------
    def state_dict(self, module: nn.Module, store: 'FlatStore') -> dict:
        """Adam moments by parameter name.  IndexLinear tables are stored sparsely: only rows that ever received a
        gradient have non-zero moments (38 M table entries, 13 live body indices in a 4-leg run)."""
        out = {'dense': {}, 'tables': {}, 'ctl_i': None if self.ctl_i is None else self.ctl_i.cpu().clone()}
        if self.m is not None and store is not None and store.flat is not None and self.m.shape == store.flat.shape:
            for name, p, off, _ in store.entries:
                out['dense'][name] = (self.m[off:off + p.numel()].cpu().clone().view(p.shape),
                                      self.v[off:off + p.numel()].cpu().clone().view(p.shape))
        for name, p in module.named_parameters():
            k = id(p)
            if k in self.table_m:
                m, v = self.table_m[k], self.table_v[k]
                rows = torch.nonzero((m.reshape(m.shape[0], -1) != 0).any(1) | (v.reshape(v.shape[0], -1) != 0).any(1)).reshape(-1)
                out['tables'][name] = (rows.cpu(), m[rows].cpu().clone(), v[rows].cpu().clone(), tuple(m.shape))
        return out

    def load_state_dict(self, state: dict, module: nn.Module, store: 'FlatStore'):
        """Inverse of :meth:`state_dict`; call after the engine has built its flat store (``prepare``)."""
        if store is None or store.flat is None:
            raise RuntimeError('load_state_dict: the parameter store is not built yet (run prepare / one update first)')
        self.ensure(store.flat)
        for name, p, off, _ in store.entries:
            if name in state['dense']:
                m, v = state['dense'][name]
                self.m[off:off + p.numel()].copy_(m.reshape(-1).to(self.m.device))
                self.v[off:off + p.numel()].copy_(v.reshape(-1).to(self.v.device))
        params = dict(module.named_parameters())
        for name, (rows, m_rows, v_rows, shape) in state['tables'].items():
            p = params[name]
            if tuple(p.shape) != tuple(shape):
                raise ValueError(f'optimizer state of {name}: table shape {tuple(shape)} != {tuple(p.shape)}')
            m, v = self.table_state(p)
            m.zero_(); v.zero_()
            if len(rows):
                m[rows.to(m.device)] = m_rows.to(m.device)
                v[rows.to(v.device)] = v_rows.to(v.device)
        if state.get('ctl_i') is not None:
            self.ctl_i.copy_(state['ctl_i'].to(self.ctl_i.device))


def dense_segments(flat, m, v, grad_base_ptr: int, ranges: Sequence[Tuple[int, int]]) -> list:
    """Split each group's [start,end) of the flat buffer into <= SEG_MAX-element Adam segments."""
    out = []
    for g, (a, b) in enumerate(ranges):
        o = a
        while o < b:
            n = min(SEG_MAX, b - o)
            out.append((flat.data_ptr() + 4 * o, m.data_ptr() + 4 * o, v.data_ptr() + 4 * o, grad_base_ptr + 4 * o, n, g))
            o += n
    return out

"""Device-side state of the two nets: flat parameters, gradient buckets, workspaces, descriptors
for the C++ schedules, and the per-minibatch step (forward, loss, backward, clip, Adam) that the
agent drives.  One instance per net; everything here only enqueues work on the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (ACT, EXT_DIM, PHASE_ALL, PHASE_COMPUTE, PHASE_GATHER, POST_NONE, POST_SIN, SardIndexLinear,
                   SardPolicyNet, SardValueNet, check, lib, ptr, stream)
from .nets import (AdamState, FlatStore, SEG_MAX, STAGE_NAMES, _dense, _encoders, _params_of, _tower, dense_segments)
from .packed import PackedRollout, Run

N_STATS = 16


class Comm:
    """Collectives used by the data-parallel step (torch.distributed / NCCL); None = single GPU."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def allgather(self, local: torch.Tensor, out: torch.Tensor):
        self.dist.all_gather_into_tensor(out, local, group=self.group)

    def allreduce(self, t: torch.Tensor):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def lanes(self, n: int):
        """``n`` more communicators over the same ranks (collective: every rank calls it, in the same order).  NCCL runs the
        collectives of ONE communicator in issue order on one internal stream; the update issues collectives from four
        independent chains (critic / policy x moments all-gather of minibatch k+1 / gradient all-reduce of minibatch k),
        and on a single communicator an all-reduce queues behind an all-gather that is still waiting for its gather."""
        if getattr(self, '_lanes', None) is None or len(self._lanes) < n:
            ranks = list(range(self.world)) if self.group is None else self.dist.get_process_group_ranks(self.group)
            self._lanes = [Comm(self.dist.new_group(ranks=ranks)) for _ in range(n)]
        return self._lanes[:n]

    def symmetric_bucket(self, n_floats: int, device):
        """A zeroed fp32 buffer of ``n_floats`` allocated as symmetric memory and mapped into every rank of the group
        (``torch.distributed._symmetric_memory``: allocation + handle exchange only), wrapped for the library's
        peer-memory all-reduce.  Collective.  Returns ``(tensor, P2PAllReduce)`` or ``(tensor, None)`` when symmetric
        memory is unavailable (the gradient all-reduce then goes through NCCL)."""
        import os
        if os.environ.get('SARD_P2P_ALLREDUCE', '1') == '0' or self.dist.get_backend(self.group) != 'nccl' or self.world > 8:
            return torch.zeros(n_floats, dtype=torch.float32, device=device), None
        try:
            import torch.distributed._symmetric_memory as symm
            g = self.group if self.group is not None else self.dist.group.WORLD
            name = g.group_name
            import warnings
            with warnings.catch_warnings():      # needed by torch < 2.8, a deprecated no-op afterwards
                warnings.simplefilter('ignore')
                try:
                    symm.enable_symm_mem_for_group(name)
                except Exception:                # noqa: BLE001
                    pass
            t = symm.empty(n_floats, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(t, group=name)
            t.zero_()
            p2p = P2PAllReduce(t, hdl, self)
            ok = torch.ones(1, device=device)
        except Exception as e:                 # noqa: BLE001
            import warnings
            warnings.warn(f'sard_b200: symmetric memory unavailable ({e!r}); gradient all-reduce through NCCL')
            t, p2p, ok = torch.zeros(n_floats, dtype=torch.float32, device=device), None, torch.zeros(1, device=device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)       # all ranks or none
        if float(ok.item()) < 1.0:
            p2p = None
        return t, p2p

    def symmetric_board(self, n_doubles: int, device):
        """A zeroed fp64 symmetric buffer for the peer-memory moments all-gather (``P2PAllGather``), or None when symmetric
        memory is unavailable (the all-gathers then go through NCCL).  Collective."""
        import os
        if os.environ.get('SARD_P2P_ALLGATHER', '1') == '0' or self.dist.get_backend(self.group) != 'nccl' or self.world > 8:
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            g = self.group if self.group is not None else self.dist.group.WORLD
            t = symm.empty(n_doubles, dtype=torch.float64, device=device)
            hdl = symm.rendezvous(t, group=g.group_name)
            t.zero_()
            board = P2PAllGather(t, hdl, self)
            ok = torch.ones(1, device=device)
        except Exception as e:                 # noqa: BLE001
            import warnings
            warnings.warn(f'sard_b200: symmetric memory unavailable ({e!r}); moments all-gather through NCCL')
            board, ok = None, torch.zeros(1, device=device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)       # all ranks or none
        return board if float(ok.item()) >= 1.0 else None

    def _host_group(self):
        """Group for small HOST arrays (counts, live indices): gloo over the same ranks, so that host metadata never
        takes a device round trip through the training streams.  Created on first use (collective); None when gloo is
        unavailable (the arrays then go through the device group)."""
        if getattr(self, '_hg', None) is None:
            if self.dist.get_backend(self.group) != 'nccl':
                self._hg = (self.group,)
            else:
                try:
                    ranks = list(range(self.world)) if self.group is None else self.dist.get_process_group_ranks(self.group)
                    self._hg = (self.dist.new_group(ranks=ranks, backend='gloo'),)
                except Exception:               # noqa: BLE001
                    self._hg = ('device',)
        return self._hg[0]

    def _host_allreduce(self, a: np.ndarray, op) -> np.ndarray:
        g = self._host_group()
        if isinstance(g, str):                  # no gloo: device round trip
            t = torch.from_numpy(np.ascontiguousarray(a)).to('cuda')
            self.dist.all_reduce(t, op=op, group=self.group)
            return t.cpu().numpy()
        t = torch.from_numpy(np.array(a, copy=True))
        self.dist.all_reduce(t, op=op, group=g)
        return t.numpy()

    def sum_array(self, a: np.ndarray) -> np.ndarray:
        """Element-wise sum of a small host integer array over the ranks (host-synchronous)."""
        return self._host_allreduce(a, self.dist.ReduceOp.SUM)

    def max_array(self, a: np.ndarray) -> np.ndarray:
        """Element-wise maximum of a small host integer array over the ranks (host-synchronous)."""
        return self._host_allreduce(a, self.dist.ReduceOp.MAX)

    def min_int(self, v: int, device=None) -> int:
        """Minimum of a host integer over the ranks (one small collective; host-synchronous)."""
        return int(self._host_allreduce(np.array([int(v)], dtype=np.int64), self.dist.ReduceOp.MIN)[0])


class P2PAllReduce:
    """The library's peer-memory all-reduce (csrc/p2p.cu) bound to one symmetric gradient bucket."""
    PAD_OFFSET = 1024          # bytes into the signal pad (the first KB is left to torch's own barrier channels)

    def __init__(self, tensor, hdl, comm: Comm):
        need = self.PAD_OFFSET + lib.sard_allreduce_p2p_pad_bytes()
        if hdl.signal_pad_size < need:
            raise RuntimeError(f'signal pad of {hdl.signal_pad_size} bytes < {need}')
        self.tensor, self.hdl = tensor, hdl
        self.rank, self.world = comm.rank, comm.world
        dev = tensor.device
        self.bufs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.pads = torch.tensor([int(p) + self.PAD_OFFSET for p in hdl.signal_pad_ptrs], dtype=torch.int64, device=dev)
        self.seq = 0

    def allreduce(self, n_floats: int, st: int):
        check(lib.sard_allreduce_p2p(ptr(self.bufs), ptr(self.pads), self.rank, self.world, int(n_floats),
                                     (2 * self.seq + 1) & 0xffffffff, st))
        self.seq += 1


class P2PAllGather:
    """The library's peer-memory all-gather of RunningNorm moment records (csrc/p2p.cu) over one symmetric board:
    ``[slot][key][world][rec_max]`` doubles, ``key`` < N_KEYS (the critic's tower, or the policy's three stages)."""
    PAD_OFFSET = 1024
    N_KEYS, N_SLOTS = 4, 2

    def __init__(self, tensor, hdl, comm: Comm):
        need = self.PAD_OFFSET + lib.sard_allgather_p2p_pad_bytes()
        if hdl.signal_pad_size < need:
            raise RuntimeError(f'signal pad of {hdl.signal_pad_size} bytes < {need}')
        self.tensor, self.hdl = tensor, hdl
        self.rank, self.world = comm.rank, comm.world
        self.rec_max = tensor.numel() // (self.N_KEYS * self.N_SLOTS * self.world)
        dev = tensor.device
        self.bufs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.pads = torch.tensor([int(p) + self.PAD_OFFSET for p in hdl.signal_pad_ptrs], dtype=torch.int64, device=dev)
        self.seq = 0

    @staticmethod
    def doubles(world: int, d_max: int) -> int:
        return P2PAllGather.N_KEYS * P2PAllGather.N_SLOTS * world * (1 + 2 * d_max)

    def offset(self, key: int, slot: int) -> int:
        return (slot * self.N_KEYS + key) * self.world * self.rec_max

    def chunk(self, key: int, slot: int, rec: int) -> torch.Tensor:
        """This rank's copy of the gathered records of (key, slot): ``[world * rec]`` doubles."""
        assert rec <= self.rec_max
        o = self.offset(key, slot)
        return self.tensor[o:o + self.world * rec]

    def allgather(self, local: torch.Tensor, key: int, slot: int, st: int):
        check(lib.sard_allgather_p2p(ptr(self.bufs), ptr(self.pads), self.rank, self.world, self.offset(key, slot), ptr(local),
                                     local.numel(), (2 * self.seq + 1) & 0xffffffff, st))
        self.seq += 1


class _EngineBase:
    def __init__(self, module):
        self.module = module
        self.device = None
        self.store: Optional[FlatStore] = None
        self.opt: Optional[AdamState] = None
        self.ws = {}
        self.mom_local = {}
        self.mom_all = {}
        # 'none': zero_grad(set_to_none=True) of torch >= 2 (a stage absent from a minibatch is skipped by Adam);
        # 'zeros': torch < 2.0, the reference's pin (its gradients stay zero tensors, Adam keeps stepping from momentum)
        self.zero_grad_semantics = 'none'
        self.comm: Optional[Comm] = None       # data parallel: set before the first sync (symmetric gradient bucket)
        self.p2p: Optional[P2PAllReduce] = None
        self.board: Optional[P2PAllGather] = None   # data parallel: moments all-gather of the prefetched gather phase
        self._board_keys = set()

    def _alloc_board(self, dev, d_max: int = 256):
        """Collective (every rank, same place: right after the gradient bucket); once per device."""
        if self.board is not None and self.board.tensor.device == dev:
            return
        self.board = None
        self._board_keys.clear()
        if self.comm is not None and self.comm.world > 1:
            self.board = self.comm.symmetric_board(P2PAllGather.doubles(self.comm.world, d_max), dev)

    def _board_key(self, key):
        """(board key, slot) of a prefetch-path moments key ``('v', slot)`` / ``(stage, slot)``, else None."""
        if self.board is None or not isinstance(key, tuple):
            return None
        k, slot = key
        return (0 if k == 'v' else int(k)), int(slot)

    def _allgather(self, key, ml, ma, comm, st):
        if key in self._board_keys:                    # `ma` is this rank's chunk of the board
            bk = self._board_key(key)
            self.board.allgather(ml, bk[0], bk[1], st)
        else:
            comm.allgather(ml, ma)

    def _alloc_bucket(self, n: int, dev) -> torch.Tensor:
        n = (n + 3) // 4 * 4
        if self.comm is not None and self.comm.world > 1:
            t, self.p2p = self.comm.symmetric_bucket(n, dev)
            return t
        self.p2p = None
        return torch.zeros(n, dtype=torch.float32, device=dev)

    def _workspace(self, key, need: int) -> torch.Tensor:
        w = self.ws.get(key)
        if w is None or w.numel() < need:
            w = torch.empty(int(need * 1.25) + 1024, dtype=torch.float32, device=self.device)
            self.ws[key] = w
        return w

    def _moments(self, key, D: int, world: int):
        rec = 1 + 2 * D
        ml = self.mom_local.get(key)
        if ml is None or ml.numel() != rec:
            ml = torch.zeros(rec, dtype=torch.float64, device=self.device)
            self.mom_local[key] = ml
        ma = None
        if world > 1:
            ma = self.mom_all.get(key)
            if ma is None or ma.numel() != rec * world:
                bk = self._board_key(key)
                if bk is not None and rec <= self.board.rec_max:
                    ma = self.board.chunk(bk[0], bk[1], rec)       # the peers store their records straight into it
                    self._board_keys.add(key)
                else:
                    self._board_keys.discard(key)
                    ma = torch.zeros(rec * world, dtype=torch.float64, device=self.device)
                self.mom_all[key] = ma
        return ml, ma


class ValueEngine(_EngineBase):
    """Transform2ActValue on the device (transform2act_critic.py:13-108)."""

    def __init__(self, module):
        super().__init__(module)
        m = module
        params = (_params_of(m.hfield_enc, 'hfield_enc.') + _params_of(m.obj_enc, 'obj_enc.') +
                  _params_of(m.goal_enc, 'goal_enc.') + _params_of(m.gnn, 'gnn.') + _params_of(m.mlp, 'mlp.') +
                  _params_of(m.value_head, 'value_head.'))
        self.store = FlatStore([params])
        self.net = None
        self.GA = None
        self.consts = None

    def sync(self, opt: Optional[AdamState] = None):
        dev = self.module.value_head.weight.device
        if dev.type != 'cuda':
            raise _lib.SardError('Transform2ActValue must be on a CUDA device: there is no CPU implementation')
        rebuilt = False
        if self.device != dev or not self.store.is_adopted():
            self.device = dev
            self.store.adopt(dev)
            self.GA = self._alloc_bucket(N_STATS + self.store.size, dev)
            self.ws.clear(); self.mom_local.clear(); self.mom_all.clear()
            self._alloc_board(dev)
            rebuilt = True
        m = self.module
        if rebuilt or self.net is None:
            net = SardValueNet()
            net.tower = _tower(self.store, m.norm, m.gnn)
            _encoders(self.store, (m.hfield_enc, m.obj_enc, m.goal_enc), net.enc)
            net.n_mlp = len(m.mlp.affine_layers)
            for j, lin in enumerate(m.mlp.affine_layers):
                net.mlp[j] = _dense(self.store, lin, m.mlp.activation)
            net.head = _dense(self.store, m.value_head)
            self.consts = torch.zeros(max(m.sym_embed_dim, 1), dtype=torch.float32, device=dev)
            net.nconst = m.sym_embed_dim if m.enable_infer else 0
            net.consts = ptr(self.consts)
            net.onehot = 1
            net.P, net.G = ptr(self.store.flat), self.GA.data_ptr() + 4 * N_STATS
            self.net = net
        if opt is not None and (rebuilt or opt is not self.opt or opt.segs is None):
            self.opt = opt
            opt.ensure(self.store.flat)
            opt.set_segments(dense_segments(self.store.flat, opt.m, opt.v, self.GA.data_ptr() + 4 * N_STATS,
                                            self.store.group_range), dev)

    def set_embeds(self, embeds: Optional[torch.Tensor]):
        if embeds is not None and self.net.nconst > 0:
            self.consts.copy_(embeds.to(device=self.device, dtype=torch.float32))

    def run(self, arena: PackedRollout, run: Run, train, inv_B: float = 0.0, comm: Optional[Comm] = None):
        """train: 0/False eval, 1/True update norm + loss + backward, 2 update norm + forward only."""
        mode = int(train)
        net = self.net
        need = lib.sard_value_workspace(C.byref(net), run.n, run.B)
        ws = self._workspace('v', need)
        world = comm.world if comm is not None else 1
        ml, ma = self._moments('v', net.tower.D, world)
        st = stream()
        sq = self.GA.data_ptr() + 4 * 5
        if train and world > 1:
            check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), mode, PHASE_GATHER, inv_B, ptr(ml),
                                     None, 1, sq, ptr(ws), ws.numel(), st))
            comm.allgather(ml, ma)
            check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), mode, PHASE_COMPUTE, inv_B, ptr(ml),
                                     ptr(ma), world, sq, ptr(ws), ws.numel(), st))
        else:
            check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), mode, PHASE_ALL,
                                     inv_B, ptr(ml), None, 1, sq, ptr(ws), ws.numel(), st))

    # -- data-parallel phases: the agent interleaves them with the policy's so that the two nets' collectives
    #    are issued back to back (NCCL executes them in issue order) and their compute overlaps on two streams
    def _call(self, arena, run, mode, phase, inv_B, ma, world):
        net = self.net
        ws = self._workspace('v', lib.sard_value_workspace(C.byref(net), run.n, run.B))
        ml, _ = self._moments('v', net.tower.D, world)
        check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), mode, phase, inv_B, ptr(ml), ptr(ma), world,
                                 self.GA.data_ptr() + 4 * 5, ptr(ws), ws.numel(), stream()))

    def dp_gather(self, arena, run, world):
        self.GA.zero_()
        self._call(arena, run, 1, PHASE_GATHER, 0.0, None, 1)

    def dp_allgather(self, comm):
        ml, ma = self._moments('v', self.net.tower.D, comm.world)
        comm.allgather(ml, ma)

    def dp_compute(self, arena, run, B_global, world):
        _, ma = self._moments('v', self.net.tower.D, world)
        self._call(arena, run, 1, PHASE_COMPUTE, 1.0 / B_global, ma, world)

    def dp_allreduce(self, comm, st: Optional[int] = None):
        if self.p2p is not None:
            self.p2p.allreduce(self.GA.numel(), st if st is not None else stream())
        else:
            comm.allreduce(self.GA)

    def train_step(self, arena: PackedRollout, run: Run, B_global: int, log_row: Optional[torch.Tensor] = None,
                   comm: Optional[Comm] = None):
        """AgentPG.update_value (agent_pg.py:18-25): forward, MSE, backward, Adam."""
        self.GA.zero_()
        self.run(arena, run, True, 1.0 / B_global, comm)
        if comm is not None and comm.world > 1:
            self.dp_allreduce(comm)
        self.adam(B_global, log_row)

    # -- software-pipelined single-GPU step: the gather + RunningNorm moments of minibatch k+1 (they read only the
    #    arena) are issued on a gather stream into the other of two workspaces while minibatch k computes
    def prefetch(self, arena: PackedRollout, run: Run, slot: int, comm: Optional[Comm] = None, st: Optional[int] = None):
        """GATHER phase (+ the moments all-gather when data parallel) into workspace `slot`.  ``st``: raw stream handle
        (default: torch's current stream; the update loop passes handles to keep the per-launch host cost down)."""
        net = self.net
        world = comm.world if comm is not None else 1
        ws = self._workspace(('v', slot), lib.sard_value_workspace(C.byref(net), run.n, run.B))
        ml, ma = self._moments(('v', slot), net.tower.D, world)
        check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), 1, PHASE_GATHER, 0.0, ptr(ml), None, 1,
                                 self.GA.data_ptr() + 4 * 5, ptr(ws), ws.numel(), st if st is not None else stream()))
        if world > 1:
            self._allgather(('v', slot), ml, ma, comm, st if st is not None else stream())

    def compute_prefetched(self, arena: PackedRollout, run: Run, B_global: int, slot: int, world: int = 1,
                           st: Optional[int] = None):
        """COMPUTE phase on a prefetched workspace.  The gradient bucket must be zero on entry: ``adam(fused=True)``
        leaves it zeroed, the agent clears it once before the first step."""
        net = self.net
        ws, ml, ma = self.ws[('v', slot)], self.mom_local[('v', slot)], self.mom_all.get(('v', slot))
        check(lib.sard_value_run(C.byref(net), C.byref(arena.c), C.byref(run.sel), 1, PHASE_COMPUTE, 1.0 / B_global, ptr(ml),
                                 ptr(ma) if world > 1 else None, world, self.GA.data_ptr() + 4 * 5, ptr(ws), ws.numel(),
                                 st if st is not None else stream()))

    def train_step_prefetched(self, arena: PackedRollout, run: Run, B_global: int, log_row, slot: int, st: Optional[int] = None):
        self.compute_prefetched(arena, run, B_global, slot, st=st)
        self.adam(B_global, log_row, fused=True, st=st)

    def adam(self, B_global: int, log_row: Optional[torch.Tensor] = None, fused: bool = False, st: Optional[int] = None):
        opt = self.opt
        st = st if st is not None else stream()
        if fused:          # one call: gate + Adam, and the gradient bucket is left zeroed for the next step
            check(lib.sard_adam_update(ptr(self.GA), 1.0 / B_global, 1.0, 0.0, None, 0, None, -1.0, 0, 0, 1, opt.betas[0],
                                       opt.betas[1], ptr(opt.ctl_f), ptr(opt.ctl_i), ptr(log_row), ptr(opt.segs), opt.n_seg,
                                       opt.max_seg, opt.lr, opt.eps, st))
            return
        check(lib.sard_adam_prepare(ptr(self.GA), 1.0 / B_global, 1.0, 0.0, None, -1.0, 0, 0, 1, opt.betas[0], opt.betas[1],
                                    ptr(opt.ctl_f), ptr(opt.ctl_i), ptr(log_row), st))
        check(lib.sard_adam_apply(ptr(opt.segs), opt.n_seg, opt.max_seg, opt.lr, opt.betas[0], opt.betas[1], opt.eps, 1,
                                  ptr(opt.ctl_f), ptr(opt.ctl_i), st))


class PolicyEngine(_EngineBase):
    """Transform2ActPolicy on the device (transform2act_policy.py:43-419)."""

    def __init__(self, module):
        super().__init__(module)
        m = module
        shared = (_params_of(m.hfield_enc, 'hfield_enc.') + _params_of(m.obj_enc, 'obj_enc.') +
                  _params_of(m.goal_enc, 'goal_enc.'))
        groups = [shared]
        for s, name in enumerate(STAGE_NAMES):
            plist = _params_of(getattr(m, f'{name}_gnn'), f'{name}_gnn.')
            if name == 'attr':
                plist.append(('attr_action_log_std', m.attr_action_log_std))
            if name == 'control':
                plist.append(('control_action_log_std', m.control_action_log_std))
            groups.append(plist)
        self.store = FlatStore(groups)
        self.max_index = m.max_index
        self.ever_live = [np.zeros(self.max_index, dtype=bool) for _ in range(3)]
        self.slot_of_index = None
        self.net = None
        self.GA = None
        self.gt_off = [0, 0, 0]
        self.consts = None
        self.sq_scratch = None
        self.sq_out = None
        self.skip_sqnorm = False       # set per update by the agent: the first-step-only clip is already disarmed
        self._table_ptrs = None

    # -- layout ---------------------------------------------------------------------------
    def _tables(self, s):
        return getattr(self.module, f'{STAGE_NAMES[s]}_ind_mlp').layers()

    def sync(self, opt: Optional[AdamState] = None, new_live: Optional[Sequence[np.ndarray]] = None):
        m = self.module
        dev = m.control_action_log_std.device
        if dev.type != 'cuda':
            raise _lib.SardError('Transform2ActPolicy must be on a CUDA device: there is no CPU implementation')
        rebuilt = False
        if self.device != dev or not self.store.is_adopted():
            self.device = dev
            self.store.adopt(dev)
            self.ws.clear(); self.mom_local.clear(); self.mom_all.clear()
            self.consts = torch.zeros(max(m.sym_embed_dim, 1), dtype=torch.float32, device=dev)
            self.sq_scratch = torch.zeros(1024, dtype=torch.float32, device=dev)
            self.sq_out = torch.zeros(1, dtype=torch.float32, device=dev)
            rebuilt = True
        grew = False
        if new_live is not None:
            for s in range(3):
                idx = np.asarray(new_live[s], dtype=np.int64)
                if len(idx) and (idx.max() >= self.max_index or idx.min() < 0):
                    raise _lib.SardError(f'body index out of range [0,{self.max_index})')
                if len(idx) and not self.ever_live[s][idx].all():
                    self.ever_live[s][idx] = True
                    grew = True
        tp = tuple(t.W.data_ptr() for s in range(3) for t in self._tables(s))
        if tp != self._table_ptrs:
            self._table_ptrs = tp
            rebuilt = True
        if rebuilt or grew or self.net is None:
            self._build(dev)
            rebuilt = True
        if opt is not None and (rebuilt or opt is not self.opt or opt.segs is None):
            self.opt = opt
            opt.ensure(self.store.flat)
            self._build_segments(opt)

    def _build(self, dev):
        m = self.module
        slot = np.full((3, self.max_index), -1, dtype=np.int32)
        sizes = []
        for s in range(3):
            live = np.flatnonzero(self.ever_live[s])
            slot[s, live] = np.arange(len(live), dtype=np.int32)
            S = len(live)
            sizes.append(sum(S * (t.out_dim * t.input_dim + t.out_dim) + 8 for t in self._tables(s)))
        self.slot_of_index = torch.from_numpy(slot).to(dev)
        # bucket order: statistics | dense | control tables | attr tables | skel tables.  With batch_design ~23 of 24
        # minibatches hold execution-stage graphs only, so the data-parallel all-reduce covers the prefix that ends
        # with the last ACTIVE stage's tables (a third of the bucket) instead of all of it
        total = (N_STATS + self.store.size + 3) // 4 * 4
        self.gt_end = [0, 0, 0]
        for s in (2, 1, 0):
            self.gt_off[s] = total
            total += (sizes[s] + 3) // 4 * 4
            self.gt_end[s] = total
        self.GA = self._alloc_bucket(total, dev)
        self._alloc_board(dev)
        net = SardPolicyNet()
        _encoders(self.store, (m.hfield_enc, m.obj_enc, m.goal_enc), net.enc)
        net.nconst = m.sym_embed_dim if m.enable_infer else 0
        net.consts = ptr(self.consts)
        net.P, net.G = ptr(self.store.flat), self.GA.data_ptr() + 4 * N_STATS
        for s, name in enumerate(STAGE_NAMES):
            st = net.stage[s]
            st.stage = s
            st.tower = _tower(self.store, getattr(m, f'{name}_norm'), getattr(m, f'{name}_gnn'))
            if s == 2:
                st.lead, st.tail = m.control_state_dim, 0
            else:
                st.lead, st.tail = m.attr_fixed_dim, m.attr_design_dim
            tabs = self._tables(s)
            st.n_il = len(tabs)
            S = int(self.ever_live[s].sum())
            off = 0
            for j, t in enumerate(tabs):
                il = SardIndexLinear()
                il.out, il.in_ = t.out_dim, t.input_dim
                il.W, il.b = ptr(t.W.data), ptr(t.b.data)
                il.gW = off; off += S * t.out_dim * t.input_dim
                il.gb = off; off += (S * t.out_dim + 3) // 4 * 4
                st.il[j] = il
            st.il_act = ACT[getattr(m, f'{name}_ind_mlp').activation]
            st.out_dim = tabs[-1].out_dim
            st.post = POST_SIN if s == 1 else POST_NONE
            st.log_std = -1 if s == 0 else self.store.off(m.attr_action_log_std if s == 1 else m.control_action_log_std)
            st.tie_c0, st.tie_c1 = m.tie_columns(s)
            st.max_index = self.max_index
            st.n_live = max(int(self.ever_live[s].sum()), 1)
            st.slot_of_index = self.slot_of_index.data_ptr() + 4 * self.max_index * s
            st.GT = self.GA.data_ptr() + 4 * self.gt_off[s]
        self.net = net

    def _build_segments(self, opt: AdamState):
        segs = dense_segments(self.store.flat, opt.m, opt.v, self.GA.data_ptr() + 4 * N_STATS, self.store.group_range)
        for s in range(3):
            live = np.flatnonzero(self.ever_live[s])
            st = self.net.stage[s]
            for j, t in enumerate(self._tables(s)):
                mW, vW = opt.table_state(t.W)
                mb, vb = opt.table_state(t.b)
                rw, rb = t.out_dim * t.input_dim, t.out_dim
                for k, u in enumerate(live):
                    u = int(u)
                    for o in range(0, rw, SEG_MAX):          # <= SEG_MAX elements per segment (one block row each)
                        n = min(SEG_MAX, rw - o)
                        segs.append((t.W.data.data_ptr() + 4 * (u * rw + o), mW.data_ptr() + 4 * (u * rw + o),
                                     vW.data_ptr() + 4 * (u * rw + o), st.GT + 4 * (st.il[j].gW + k * rw + o), n, 1 + s))
                    segs.append((t.b.data.data_ptr() + 4 * u * rb, mb.data_ptr() + 4 * u * rb, vb.data_ptr() + 4 * u * rb,
                                 st.GT + 4 * (st.il[j].gb + k * rb), rb, 1 + s))
        opt.set_segments(segs, self.device)

    def set_embeds(self, embeds: Optional[torch.Tensor]):
        if embeds is not None and self.net.nconst > 0:
            self.consts.copy_(embeds.to(device=self.device, dtype=torch.float32))

    # -- runs -----------------------------------------------------------------------------
    def run_stage(self, s: int, arena: PackedRollout, run: Run, train: bool, clip_eps: float = 0.2, inv_B: float = 0.0,
                  outputs: Optional[torch.Tensor] = None, log_prob_run: Optional[torch.Tensor] = None,
                  comm: Optional[Comm] = None, phase: int = PHASE_ALL):
        net = self.net
        if net.stage[s].tie_c1 > net.stage[s].tie_c0 and not arena.has_orbits:
            raise _lib.SardError('orbit representatives missing: call PackedRollout.set_orbits first')
        need = lib.sard_policy_workspace(C.byref(net), s, run.n, run.B)
        ws = self._workspace(s, need)
        world = comm.world if comm is not None else 1
        ml, ma = self._moments(s, net.stage[s].tower.D, world)
        args = (C.byref(net), s, C.byref(arena.c), C.byref(run.sel), int(train))
        tail = (ptr(self.GA), ptr(outputs), ptr(log_prob_run), ptr(ws), ws.numel(), stream())
        if train and world > 1:
            check(lib.sard_policy_run(*args, PHASE_GATHER, clip_eps, inv_B, ptr(ml), None, 1, *tail))
            comm.allgather(ml, ma)
            check(lib.sard_policy_run(*args, PHASE_COMPUTE, clip_eps, inv_B, ptr(ml), ptr(ma), world, *tail))
        else:
            check(lib.sard_policy_run(*args, phase, clip_eps, inv_B, ptr(ml), None, 1, *tail))

    def dp_gather(self, arena, runs, active_global, world):
        self.GA.zero_()
        for s in (2, 1, 0):
            if runs[s] is not None:
                self.run_stage(s, arena, runs[s], True, phase=PHASE_GATHER)
            elif active_global[s]:
                ml, _ = self._moments(s, self.net.stage[s].tower.D, world)
                ml.zero_()                # empty record: another rank has graphs of this stage

    def dp_allgather(self, active_global, comm):
        for s in (2, 1, 0):
            if active_global[s]:
                ml, ma = self._moments(s, self.net.stage[s].tower.D, comm.world)
                comm.allgather(ml, ma)

    def dp_compute(self, arena, runs, active_global, B_global, clip_eps, world):
        net = self.net
        for s in (2, 1, 0):
            if runs[s] is not None:
                run = runs[s]
                ws = self._workspace(s, lib.sard_policy_workspace(C.byref(net), s, run.n, run.B))
                ml, ma = self._moments(s, net.stage[s].tower.D, world)
                check(lib.sard_policy_run(C.byref(net), s, C.byref(arena.c), C.byref(run.sel), 1, PHASE_COMPUTE, clip_eps,
                                          1.0 / B_global, ptr(ml), ptr(ma), world, ptr(self.GA), None, None, ptr(ws),
                                          ws.numel(), stream()))
            elif active_global[s]:
                _, ma = self._moments(s, net.stage[s].tower.D, world)
                self._empty_norm_update(s, ma, world)

    def dp_allreduce(self, comm, active_global: Optional[Sequence[bool]] = None, st: Optional[int] = None):
        """SUM all-reduce of the gradient bucket: the prefix up to the last active stage's tables (the rest is zero on
        every rank: stages absent from the global minibatch wrote nothing)."""
        if active_global is None or active_global[0] or not any(active_global):
            end = self.GA.numel()
        else:
            end = max(self.gt_end[s] for s in range(3) if active_global[s])
        if self.p2p is not None:
            self.p2p.allreduce(end, st if st is not None else stream())
        else:
            comm.allreduce(self.GA[:end])

    def train_step(self, arena: PackedRollout, runs: List[Optional[Run]], B_global: int, skel_nodes_global: int,
                   active_global: Sequence[bool], clip_eps: float, max_norm: float, clip_first_only: bool,
                   log_row: Optional[torch.Tensor] = None, comm: Optional[Comm] = None):
        """ppo_loss + backward + clip + Adam (transform2act_agent.py:351-358), gated on device (:404-407)."""
        self.GA.zero_()
        for s in (2, 1, 0):                       # the reference evaluates execution, attribute, skeleton
            if runs[s] is not None:
                self.run_stage(s, arena, runs[s], True, clip_eps, 1.0 / B_global, comm=comm)
            elif comm is not None and comm.world > 1 and active_global[s]:
                # another rank has graphs of this stage: contribute an empty moments record
                ml, ma = self._moments(s, self.net.stage[s].tower.D, comm.world)
                ml.zero_()
                comm.allgather(ml, ma)
                self._empty_norm_update(s, ma, comm.world)
        world = 1
        if comm is not None and comm.world > 1:
            self.dp_allreduce(comm, active_global)
            world = comm.world
        self.adam(B_global, skel_nodes_global, active_global, max_norm, clip_first_only, world, log_row)

    def prefetch(self, arena: PackedRollout, runs: List[Optional[Run]], slot: int, comm: Optional[Comm] = None,
                 active_global: Optional[Sequence[bool]] = None, st: Optional[int] = None):
        """GATHER phase of every present stage (+ the moments all-gathers when data parallel) into workspace `slot`."""
        net = self.net
        world = comm.world if comm is not None else 1
        st = st if st is not None else stream()
        for s in (2, 1, 0):
            run = runs[s]
            if run is None:
                if world > 1 and active_global[s]:         # another rank has graphs of this stage: empty record
                    ml, ma = self._moments((s, slot), net.stage[s].tower.D, world)
                    with torch.cuda.stream(torch.cuda.ExternalStream(st, device=self.device)):
                        ml.zero_()
                    self._allgather((s, slot), ml, ma, comm, st)
                continue
            if net.stage[s].tie_c1 > net.stage[s].tie_c0 and not arena.has_orbits:
                raise _lib.SardError('orbit representatives missing: call PackedRollout.set_orbits first')
            ws = self._workspace((s, slot), lib.sard_policy_workspace(C.byref(net), s, run.n, run.B))
            ml, ma = self._moments((s, slot), net.stage[s].tower.D, world)
            check(lib.sard_policy_run(C.byref(net), s, C.byref(arena.c), C.byref(run.sel), 1, PHASE_GATHER, 0.0, 0.0, ptr(ml),
                                      None, 1, ptr(self.GA), None, None, ptr(ws), ws.numel(), st))
            if world > 1:
                self._allgather((s, slot), ml, ma, comm, st)

    def compute_prefetched(self, arena: PackedRollout, runs: List[Optional[Run]], B_global: int, clip_eps: float, slot: int,
                           world: int = 1, active_global: Optional[Sequence[bool]] = None, st: Optional[int] = None):
        net = self.net                  # gradient bucket zero on entry (see ValueEngine.compute_prefetched)
        st = st if st is not None else stream()
        for s in (2, 1, 0):
            run = runs[s]
            if run is None:
                if world > 1 and active_global[s]:
                    self._empty_norm_update(s, self.mom_all[(s, slot)], world, st)
                continue
            ws, ml, ma = self.ws[(s, slot)], self.mom_local[(s, slot)], self.mom_all.get((s, slot))
            check(lib.sard_policy_run(C.byref(net), s, C.byref(arena.c), C.byref(run.sel), 1, PHASE_COMPUTE, clip_eps,
                                      1.0 / B_global, ptr(ml), ptr(ma) if world > 1 else None, world, ptr(self.GA), None, None,
                                      ptr(ws), ws.numel(), st))

    def train_step_prefetched(self, arena: PackedRollout, runs: List[Optional[Run]], B_global: int, skel_nodes_global: int,
                              active_global: Sequence[bool], clip_eps: float, max_norm: float, clip_first_only: bool,
                              log_row, slot: int, st: Optional[int] = None):
        self.compute_prefetched(arena, runs, B_global, clip_eps, slot, st=st)
        self.adam(B_global, skel_nodes_global, active_global, max_norm, clip_first_only, 1, log_row, fused=True, st=st)

    def adam(self, B_global, skel_nodes_global, active_global, max_norm, clip_first_only, world, log_row=None,
             fused: bool = False, st: Optional[int] = None):
        opt = self.opt
        mask = (1 if any(active_global) else 0) | sum((1 << (1 + s)) for s in range(3) if active_global[s])
        if self.zero_grad_semantics == 'zeros':
            mask |= _lib.ADAM_STICKY
        st = st if st is not None else stream()
        if fused:          # grad-norm partials + gate / clip + Adam in one call; leaves the gradient bucket zeroed
            inv_cat = 1.0 / skel_nodes_global if skel_nodes_global > 0 else 0.0
            # once the reference's one-shot clip is spent (agent :48, agent_ppo.py:53-56) the gradient norm is not needed
            g_sq = None if (self.skip_sqnorm and clip_first_only) else self.GA.data_ptr() + 4 * N_STATS
            check(lib.sard_adam_update(ptr(self.GA), 1.0 / B_global, 1.0 / world, inv_cat, g_sq,
                                       self.GA.numel() - N_STATS, ptr(self.sq_scratch), float(max_norm),
                                       1 if clip_first_only else 0, 1, mask, opt.betas[0], opt.betas[1], ptr(opt.ctl_f),
                                       ptr(opt.ctl_i), ptr(log_row), ptr(opt.segs), opt.n_seg, opt.max_seg, opt.lr, opt.eps, st))
            return
        check(lib.sard_sqnorm(self.GA.data_ptr() + 4 * N_STATS, self.GA.numel() - N_STATS, ptr(self.sq_scratch),
                              ptr(self.sq_out), st))
        inv_cat = 1.0 / skel_nodes_global if skel_nodes_global > 0 else 0.0
        check(lib.sard_adam_prepare(ptr(self.GA), 1.0 / B_global, 1.0 / world, inv_cat, ptr(self.sq_out), float(max_norm),
                                    1 if clip_first_only else 0, 1, mask, opt.betas[0], opt.betas[1], ptr(opt.ctl_f),
                                    ptr(opt.ctl_i), ptr(log_row), st))
        check(lib.sard_adam_apply(ptr(opt.segs), opt.n_seg, opt.max_seg, opt.lr, opt.betas[0], opt.betas[1], opt.eps, mask,
                                  ptr(opt.ctl_f), ptr(opt.ctl_i), st))

    def _empty_norm_update(self, s, ma, world, st=None):
        tw = self.net.stage[s].tower
        check(lib.sard_norm_update(ptr(ma), world, tw.D, tw.mean, tw.var, tw.std, tw.count, tw.inv,
                                   st if st is not None else stream()))

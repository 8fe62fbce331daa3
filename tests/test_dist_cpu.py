"""CPU, world_size 2 over gloo: the host side of the data-parallel step -- the collectives wrapper,
the per-epoch minibatch plan and the global count reduction that makes every rank take the same
gate / Adam decisions.  (The kernels themselves need a GPU; see tests/test_engine_gpu.py.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


class HostArena:
    """The host-side fields EpochPlan / Ordering read from a PackedRollout."""

    def __init__(self, roll):
        self.device = torch.device('cpu')
        self.stage_np = roll.stage.astype(np.int64)
        self.num_nodes_np = np.diff(roll.node_ptr).astype(np.int64)
        self.T = len(self.stage_np)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sard_b200 import synthetic
    from sard_b200.engine import Comm
    from sard_b200.packed import EpochPlan
    comm = Comm()
    assert comm.world == world and comm.rank == rank
    # 1. moments all-gather: every rank sees the records in rank order
    local = torch.full((7,), float(rank + 1), dtype=torch.float64)
    allr = torch.zeros(7 * world, dtype=torch.float64)
    comm.allgather(local, allr)
    assert allr.reshape(world, 7)[:, 0].tolist() == [1.0, 2.0]
    # 2. gradient bucket all-reduce (SUM)
    g = torch.arange(5, dtype=torch.float32) * (rank + 1)
    comm.allreduce(g)
    assert g.tolist() == [0.0, 3.0, 6.0, 9.0, 12.0]
    # 3. per-epoch plan: same shuffle stream on every rank, rank-local rollouts, global counts by all-reduce
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES['point_nav'], 400, seed=100 + rank, min_exec=5, max_exec=40)
    arena = HostArena(roll)
    np.random.seed(7)
    perm = np.arange(arena.T); np.random.shuffle(perm)
    plan = EpochPlan(arena, perm, 64, batch_design=True)
    counts = torch.from_numpy(np.concatenate([plan.stage_graphs, plan.stage_nodes], axis=1))
    local_counts = counts.clone()
    comm.allreduce(counts)
    assert plan.n_mb == 6
    assert np.all(counts[:, :3].sum(1).numpy() == 64 * world)          # global minibatch = sum of the shards
    # every minibatch is stage-sorted and its stage sub-runs tile it exactly
    for k in range(plan.n_mb):
        runs = plan.policy_runs(k)
        assert sum(r.B for r in runs if r is not None) == 64
        st = plan.policy_order.stage_np[k * 64:(k + 1) * 64]
        assert np.all(np.diff(st) >= 0)
        for s, r in enumerate(runs):
            if r is not None:
                assert np.all(st[r.lo - k * 64:r.hi - k * 64] == s)
                assert r.n == int(local_counts[k, 3 + s])
    # 4. unequal rollout lengths (rollouts end on episode boundaries): T // M differs between the ranks, every rank
    #    must plan the same number of minibatch steps = the same sequence of collectives (the minimum)
    T_r = 400 + 70 * rank                                   # 6 vs 7 minibatches of 64
    roll2 = synthetic.generate_rollout(synthetic.ENV_SHAPES['point_nav'], T_r, seed=300 + rank, min_exec=5, max_exec=40)
    arena2 = HostArena(roll2)
    n_mb = comm.min_int(arena2.T // 64)
    assert n_mb == 6
    perm2 = np.random.default_rng(rank).permutation(arena2.T)
    plan2 = EpochPlan(arena2, perm2, 64, batch_design=True, n_mb=n_mb)
    assert plan2.n_mb == 6 and plan2.stage_graphs.shape == (6, 3)
    c2 = torch.from_numpy(np.concatenate([plan2.stage_graphs, plan2.stage_nodes], axis=1))
    comm.allreduce(c2)                                      # same shape on every rank: no hang, no corruption
    assert np.all(c2[:, :3].sum(1).numpy() == 64 * world)
    # 5. host metadata over the host group (the agent's per-epoch stage counts, the union of live body indices): the input
    #    array is left untouched, every rank gets the same result
    mine = np.concatenate([plan2.stage_graphs, plan2.stage_nodes], axis=1)
    keep = mine.copy()
    tot = comm.sum_array(mine)
    assert np.array_equal(mine, keep) and np.array_equal(tot, c2.numpy())
    mask = np.zeros((3, 16), dtype=np.int64)
    mask[rank, [rank, 5 + rank]] = 1
    assert [np.flatnonzero(m).tolist() for m in comm.max_array(mask)] == [[0, 5], [1, 6], []]
    out.put((rank, perm[:8].tolist(), counts.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_host_logic_world_size_2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1]          # identical shuffle on both ranks
    assert res[0][2] == res[1][2]          # identical global counts on both ranks


def test_epoch_plan_without_batch_design_resorts_each_minibatch():
    sys.path.insert(0, ROOT)
    from sard_b200 import synthetic
    from sard_b200.packed import EpochPlan
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES['point_nav'], 300, seed=5, min_exec=5, max_exec=30)
    arena = HostArena(roll)
    perm = np.random.default_rng(0).permutation(arena.T)
    plan = EpochPlan(arena, perm, 50, batch_design=False)
    assert np.array_equal(plan.perm, perm)                       # the value net sees the plain shuffle
    for k in range(plan.n_mb):
        a = np.sort(plan.value_order.ids_np[k * 50:(k + 1) * 50])
        b = np.sort(plan.policy_order.ids_np[k * 50:(k + 1) * 50])
        assert np.array_equal(a, b)                              # same graphs, stage-sorted for the policy
        st = plan.policy_order.stage_np[k * 50:(k + 1) * 50]
        assert np.all(np.diff(st) >= 0)
        # stable: within a stage the shuffle order is kept (policy.py:216-231)
        for s in range(3):
            want = [g for g in plan.value_order.ids_np[k * 50:(k + 1) * 50] if arena.stage_np[g] == s]
            got = [g for g in plan.policy_order.ids_np[k * 50:(k + 1) * 50] if arena.stage_np[g] == s]
            assert want == got
    plan2 = EpochPlan(arena, perm, 50, batch_design=True)
    assert np.all(np.diff(arena.stage_np[plan2.perm]) >= 0)      # agent :305-311 stable partition
    assert plan2.policy_order is plan2.value_order

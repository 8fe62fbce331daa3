"""GPU: every kernel of libsard_b200.so, called through the C-ABI, against the oracle / plain
fp64 torch on the same seeded inputs.  Integer layouts are compared bit-exactly; floating point
within the tolerances BASELINE.json states (forward rtol 1e-4, gradients rtol 1e-3, GAE 1e-6)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def L(cuda_device):
    from sard_b200 import _lib
    torch.cuda.set_device(cuda_device)
    return _lib


def dev(a, dtype=None):
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def make_roll(T=200, seed=5, shape_name='manipulation_box', orbits='H_4'):
    from sard_b200 import synthetic
    return synthetic.generate_rollout(synthetic.ENV_SHAPES[shape_name], T, seed=seed,
                                      orbits=synthetic.SYMMETRIC_ORBITS[orbits], min_exec=5, max_exec=30)


# ------------------------------------------------------------------------------------------ packing
def test_build_csr_matches_edge_order(L):
    from sard_b200.packed import PackedRollout
    roll = make_roll()
    ar = PackedRollout(roll, 'cuda')
    torch.cuda.synchronize()
    in_ptr, in_nbr = ar.in_ptr.cpu().numpy(), ar.in_nbr.cpu().numpy()
    out_ptr, out_nbr = ar.out_ptr.cpu().numpy(), ar.out_nbr.cpu().numpy()
    for t in range(roll.num_transitions):
        a, b = roll.node_ptr[t], roll.node_ptr[t + 1]
        e = roll.edges[:, roll.edge_ptr[t]:roll.edge_ptr[t + 1]]
        for v in range(b - a):
            want_in = e[0][e[1] == v]
            want_out = e[1][e[0] == v]
            assert np.array_equal(in_nbr[in_ptr[a + v]:in_ptr[a + v + 1]], want_in)
            assert np.array_equal(out_nbr[out_ptr[a + v]:out_ptr[a + v + 1]], want_out)
    assert in_ptr[-1] == roll.edge_ptr[-1] and out_ptr[-1] == roll.edge_ptr[-1]


@pytest.mark.parametrize('orbits,sub', [('H_2', 'H_2>0>H_2>4'), ('K_1', 'K_1>0>K_1>4'), ('H_1_0', 'H_1_0>0>H_1_0>4'),
                                        ('H_4', 'H_4>0>H_4>4')])
def test_orbit_rep_bit_exact_with_reference_semantics(L, orbits, sub):
    from oracle import sard_oracle as O
    from sard_b200.group_action import LearnedGroup
    from sard_b200.packed import PackedRollout
    roll = make_roll(orbits=orbits, seed=9)
    G = LearnedGroup([4], init_subgroup_name=sub)
    ar = PackedRollout(roll, 'cuda')
    ar.set_orbits(G.leg_representatives(5), 5)
    got = ar.rep_local.cpu().numpy()
    # the reference's positional overwrite over the whole concatenated batch
    want_rows = O.orbit_rep_index(roll.body_ind, G.get_cur_orbits(), 5)
    want_local = want_rows - np.repeat(roll.node_ptr[:-1], roll.num_nodes)
    assert np.array_equal(got, want_local)


def test_orbit_rep_flags_asymmetric_morphology(L):
    from sard_b200.group_action import LearnedGroup
    from sard_b200.packed import PackedRollout
    roll = make_roll(orbits='H_4', seed=3)      # generic (asymmetric) morphologies
    G = LearnedGroup([4], init_subgroup_name='H_1_0>0>H_1_0>4')
    ar = PackedRollout(roll, 'cuda')
    with pytest.raises(L.SardError):
        ar.set_orbits(G.leg_representatives(5), 5)


def test_gather_nodes_layout_bit_exact(L):
    from sard_b200.packed import Ordering, PackedRollout
    roll = make_roll()
    ar = PackedRollout(roll, 'cuda')
    rng = np.random.default_rng(0)
    ids = rng.permutation(roll.num_transitions)[:77]
    order = Ordering(ar, ids)
    run = order.run(5, 70)
    D = roll.obs.shape[1]
    lead, tail, consts = 24, 6, np.arange(8, dtype=np.float32) / 7
    width = lead + tail + 8 + 3
    ldx = (width + 3) // 4 * 4
    x = torch.full((run.n, ldx), -7.0, device='cuda')
    ndg = torch.zeros(run.n, dtype=torch.int32, device='cuda'); nda = torch.zeros_like(ndg); body = torch.zeros_like(ndg)
    root = torch.zeros(run.B, dtype=torch.int32, device='cuda')
    cst = dev(consts)
    out = L.SardGatherOut()
    out.lead, out.tail, out.nconst, out.onehot, out.consts = lead, tail, 8, 1, L.ptr(cst)
    out.x, out.ldx, out.nd_graph, out.nd_arena, out.body, out.root_row = L.ptr(x), ldx, L.ptr(ndg), L.ptr(nda), L.ptr(body), L.ptr(root)
    L.check(L.lib.sard_gather_nodes(C.byref(ar.c), C.byref(run.sel), C.byref(out), L.stream()))
    sel = ids[5:70]
    rows = np.concatenate([np.arange(roll.node_ptr[t], roll.node_ptr[t + 1]) for t in sel])
    obs32 = roll.obs.astype(np.float32)
    st = np.repeat(roll.stage[sel], roll.num_nodes[sel])
    want = np.zeros((len(rows), ldx), dtype=np.float32)
    want[:, :lead] = obs32[rows, :lead]
    want[:, lead:lead + tail] = obs32[rows, D - tail:]
    want[:, lead + tail:lead + tail + 8] = consts
    want[np.arange(len(rows)), lead + tail + 8 + st] = 1.0
    assert np.array_equal(x.cpu().numpy(), want)
    assert np.array_equal(nda.cpu().numpy(), rows)
    assert np.array_equal(ndg.cpu().numpy(), np.repeat(np.arange(len(sel)), roll.num_nodes[sel]))
    assert np.array_equal(body.cpu().numpy(), roll.body_ind[rows])
    assert np.array_equal(root.cpu().numpy(), np.concatenate([[0], np.cumsum(roll.num_nodes[sel])[:-1]]))


@pytest.mark.parametrize('n', [1, 255, 256, 257, 5000])
def test_sort_by_index_is_stable_and_tables_cover_rows(L, n):
    rng = np.random.default_rng(n)
    live = np.array([0, 1, 2, 3, 4, 6, 7, 8, 9, 31, 32, 33, 34])
    body = live[rng.integers(0, len(live), size=n)].astype(np.int32)
    mi, tr, cr = 40, 128, 256
    b = dev(body)
    i32 = lambda k: torch.zeros(k, dtype=torch.int32, device='cuda')
    srow, spos, grp, gcp, counts = i32(n), i32(n), i32(mi + 1), i32(mi + 1), i32(4)
    tiles, chunks = i32(4 * (n // tr + 1 + mi)), i32(4 * (n // cr + 1 + mi))
    scratch = i32(((n + 255) // 256 + 2) * mi)
    L.check(L.lib.sard_sort_by_index(L.ptr(b), n, mi, tr, cr, L.ptr(srow), L.ptr(spos), L.ptr(grp), L.ptr(tiles), L.ptr(chunks),
                                     L.ptr(gcp), L.ptr(counts), L.ptr(scratch), L.stream()))
    want = np.argsort(body, kind='stable')
    assert np.array_equal(srow.cpu().numpy(), want)
    assert np.array_equal(spos.cpu().numpy()[want], np.arange(n))
    cnt = np.bincount(body, minlength=mi)
    assert np.array_equal(grp.cpu().numpy(), np.concatenate([[0], np.cumsum(cnt)]))
    nt, nc = counts.cpu().numpy()[:2]
    t = tiles.cpu().numpy().reshape(-1, 4)[:nt]
    assert nt == sum((c + tr - 1) // tr for c in cnt)
    covered = np.concatenate([np.arange(a, e) for a, e, _, _ in t]) if nt else np.zeros(0)
    assert np.array_equal(covered, np.arange(n))
    for a, e, u, _ in t:
        assert e - a <= tr and np.all(body[want[a:e]] == u)
    ch = chunks.cpu().numpy().reshape(-1, 4)[:nc]
    g = gcp.cpu().numpy()
    for u in range(mi):
        rows = [np.arange(a, e) for a, e, uu, _ in ch[g[u]:g[u + 1]]]
        assert all(uu == u for _, _, uu, _ in ch[g[u]:g[u + 1]])
        assert sum(len(r) for r in rows) == cnt[u]


# ------------------------------------------------------------------------------------------ RunningNorm
def test_running_norm_matches_oracle_including_constant_columns(L):
    from oracle import sard_oracle as O
    rng = np.random.default_rng(1)
    D = 38
    sd = {'n.n': torch.tensor(0), 'n.mean': torch.zeros(D, dtype=torch.float64), 'n.var': torch.zeros(D, dtype=torch.float64),
          'n.std': torch.zeros(D, dtype=torch.float64)}
    mean, var, std, inv = (torch.zeros(D, device='cuda') for _ in range(4))
    count = torch.zeros((), dtype=torch.long, device='cuda')
    for it, n in enumerate([700, 129, 3000]):
        x = rng.normal(size=(n, D)).astype(np.float32) * rng.uniform(0.1, 3, size=D).astype(np.float32) + 0.5
        x[:, 5] = 1.0 / 3.0            # constant column (the subgroup embedding)
        x[:, 6] = 0.0                  # all-zero column (sim obs in the design stages)
        x[:, 7] = (rng.random(n) < 0.3)   # one-hot style column
        ldx = 40
        xg = torch.zeros(n, ldx, device='cuda'); xg[:, :D] = dev(x)
        part = torch.zeros(2 * D * ((n + 127) // 128), device='cuda')
        rec = torch.zeros(1 + 2 * D, dtype=torch.float64, device='cuda')
        L.check(L.lib.sard_norm_moments(L.ptr(xg), ldx, n, D, L.ptr(part), L.ptr(rec), L.stream()))
        L.check(L.lib.sard_norm_update(L.ptr(rec), 1, D, L.ptr(mean), L.ptr(var), L.ptr(std), L.ptr(count), L.ptr(inv), L.stream()))
        L.check(L.lib.sard_norm_apply(L.ptr(xg), ldx, n, D, L.ptr(mean), L.ptr(inv), L.ptr(count), 5.0, L.stream()))
        want = O.running_norm(sd, 'n.', torch.from_numpy(x.astype(np.float64)), training=True)
        assert int(count.item()) == int(sd['n.n'])
        assert_close(mean.cpu(), sd['n.mean'], rtol=1e-5, atol=1e-6, what=f'mean it{it}')
        assert_close(var.cpu(), sd['n.var'], rtol=1e-4, atol=1e-7, what=f'var it{it}')
        assert_close(xg[:, :D].cpu(), want, rtol=1e-4, atol=2e-5, what=f'normalised it{it}')
        assert float(xg[:, 5].abs().max()) < 1e-3 and float(xg[:, 6].abs().max()) == 0.0


def test_norm_records_merge_like_one_big_batch(L):
    """Two ranks' records merged by sard_norm_update == one batch of the concatenation."""
    rng = np.random.default_rng(2)
    D, n1, n2 = 17, 300, 555
    x = rng.normal(size=(n1 + n2, D)).astype(np.float32)
    recs = torch.zeros(2, 1 + 2 * D, dtype=torch.float64, device='cuda')
    for k, (a, b) in enumerate([(0, n1), (n1, n1 + n2)]):
        xg = dev(x[a:b])
        part = torch.zeros(2 * D * ((b - a + 127) // 128), device='cuda')
        L.check(L.lib.sard_norm_moments(L.ptr(xg), D, b - a, D, L.ptr(part), recs[k].data_ptr(), L.stream()))
    mean, var, std, inv = (torch.zeros(D, device='cuda') for _ in range(4))
    count = torch.zeros((), dtype=torch.long, device='cuda')
    L.check(L.lib.sard_norm_update(L.ptr(recs), 2, D, L.ptr(mean), L.ptr(var), L.ptr(std), L.ptr(count), L.ptr(inv), L.stream()))
    x64 = x.astype(np.float64)
    assert int(count.item()) == n1 + n2
    assert_close(mean.cpu(), x64.mean(0), rtol=1e-5, atol=1e-6, what='mean')
    assert_close(var.cpu(), x64.var(0), rtol=1e-4, atol=1e-7, what='var')


# ------------------------------------------------------------------------------------------ aggregate
def test_csr_aggregate_forward_and_adjoint(L):
    from sard_b200.packed import Ordering, PackedRollout
    roll = make_roll(T=150)
    ar = PackedRollout(roll, 'cuda')
    order = Ordering(ar, np.random.default_rng(3).permutation(150))
    run = order.run(10, 140)
    sel = order.ids_np[10:140]
    F = 64
    n = run.n
    ndg = dev(np.repeat(np.arange(len(sel)), roll.num_nodes[sel]), torch.int32)
    nda = dev(np.concatenate([np.arange(roll.node_ptr[t], roll.node_ptr[t + 1]) for t in sel]), torch.int32)
    x = torch.randn(n, 2 * F, device='cuda')
    base = torch.randn(n, F, device='cuda')
    y = torch.randn(n, F, device='cuda')
    out = torch.zeros(n, F, device='cuda')
    # global edge list of the run
    start = np.concatenate([[0], np.cumsum(roll.num_nodes[sel])[:-1]])
    src = np.concatenate([roll.edges[0, roll.edge_ptr[t]:roll.edge_ptr[t + 1]] + s for t, s in zip(sel, start)])
    dst = np.concatenate([roll.edges[1, roll.edge_ptr[t]:roll.edge_ptr[t + 1]] + s for t, s in zip(sel, start)])
    src_t, dst_t = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    L.check(L.lib.sard_csr_aggregate(x.data_ptr() + 4 * F, 2 * F, L.ptr(out), F, None, 0, None, 0, 0, n, F, L.ptr(ndg), L.ptr(nda),
                                     run.sel.cum, run.sel.node_base, L.ptr(ar.in_ptr), L.ptr(ar.in_nbr), L.stream()))
    want = torch.zeros(n, F, dtype=torch.float64, device='cuda').index_add_(0, dst_t, x[:, F:].double()[src_t])
    assert_close(out.cpu(), want.cpu(), rtol=1e-5, atol=1e-5, what='aggregate')
    # adjoint with base and relu'
    L.check(L.lib.sard_csr_aggregate(L.ptr(x), 2 * F, L.ptr(out), F, L.ptr(base), F, L.ptr(y), F, 1, n, F, L.ptr(ndg), L.ptr(nda),
                                     run.sel.cum, run.sel.node_base, L.ptr(ar.out_ptr), L.ptr(ar.out_nbr), L.stream()))
    want = (base.double() + torch.zeros(n, F, dtype=torch.float64, device='cuda').index_add_(0, src_t, x[:, :F].double()[dst_t])) * (y > 0)
    assert_close(out.cpu(), want.cpu(), rtol=1e-5, atol=1e-5, what='aggregate adjoint')


@pytest.mark.parametrize('n_legs', [4, 7])
def test_fused_graphconv_gemm_matches_aggregate_then_linear(L, n_legs):
    """SardGemm.agg_*: forward Y = relu([sum_nbr x | x] [Wl|Wr]^T + b) and the adjoint [A^T dz | dz] [Wl; Wr] * act',
    against fp64 torch built from the edge list.  7 legs give the root 7 neighbours: the loader keeps 4 neighbour
    rows in registers and walks longer lists from memory."""
    from sard_b200 import synthetic
    from sard_b200.packed import Ordering, PackedRollout
    shape = synthetic.EnvShape('wide', n_legs=n_legs, max_body_depth=5)
    roll = synthetic.generate_rollout(shape, 180, seed=11, min_exec=5, max_exec=30)
    ar = PackedRollout(roll, 'cuda')
    order = Ordering(ar, np.random.default_rng(4).permutation(180))
    run = order.run(7, 171)
    sel = order.ids_np[7:171]
    H = 64
    n = run.n
    assert L.lib.sard_gemm_agg_supported(H, H) == 1
    ndg = dev(np.repeat(np.arange(len(sel)), roll.num_nodes[sel]), torch.int32)
    nda = dev(np.concatenate([np.arange(roll.node_ptr[t], roll.node_ptr[t + 1]) for t in sel]), torch.int32)
    start = np.concatenate([[0], np.cumsum(roll.num_nodes[sel])[:-1]])
    src = np.concatenate([roll.edges[0, roll.edge_ptr[t]:roll.edge_ptr[t + 1]] + s for t, s in zip(sel, start)])
    dst = np.concatenate([roll.edges[1, roll.edge_ptr[t]:roll.edge_ptr[t + 1]] + s for t, s in zip(sel, start)])
    src_t, dst_t = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    assert np.bincount(dst).max() >= min(n_legs, 5)
    g = torch.Generator().manual_seed(n_legs)
    xa = torch.zeros(n, 2 * H, device='cuda')
    xa[:, H:] = torch.randn(n, H, generator=g).cuda()
    Wl, Wr, b = (torch.randn(H, H, generator=g) / 8).cuda(), (torch.randn(H, H, generator=g) / 8).cuda(), torch.randn(H, generator=g).cuda()
    Y = torch.zeros(n, H, device='cuda')
    gm = L.SardGemm()
    gm.A, gm.lda = xa.data_ptr() + 4 * H, 2 * H
    gm.B0, gm.ldb0, gm.B1, gm.ldb1, gm.b_split, gm.bias = L.ptr(Wl), H, L.ptr(Wr), H, H, L.ptr(b)
    gm.Y, gm.ldy, gm.M, gm.N, gm.R, gm.act = L.ptr(Y), H, n, H, 2 * H, L.ACT['relu']
    gm.agg_cols, gm.agg_ptr, gm.agg_nbr, gm.agg_node, gm.agg_graph = H, L.ptr(ar.in_ptr), L.ptr(ar.in_nbr), L.ptr(nda), L.ptr(ndg)
    gm.agg_cum, gm.agg_node_base, gm.agg_out, gm.ld_agg_out = run.sel.cum, run.sel.node_base, L.ptr(xa), 2 * H
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    x64 = xa[:, H:].double()
    agg = torch.zeros(n, H, dtype=torch.float64, device='cuda').index_add_(0, dst_t, x64[src_t])
    want = torch.relu(agg @ Wl.double().t() + x64 @ Wr.double().t() + b.double())
    assert_close(xa[:, :H].cpu(), agg.cpu(), rtol=1e-6, atol=1e-6, what='fused aggregate side product')
    assert_close(Y.cpu(), want.cpu(), rtol=1e-4, atol=2e-5, what='fused graphconv forward')
    # adjoint: dx = (A^T dz Wl + dz Wr) * relu'(yprev)
    dz = torch.randn(n, H, generator=g).cuda()
    yprev = torch.randn(n, H, generator=g).cuda()
    dx = torch.zeros(n, H, device='cuda')
    gb = L.SardGemm()
    gb.A, gb.lda, gb.B0, gb.ldb0, gb.B1, gb.ldb1, gb.b_split, gb.b_ksplit = L.ptr(dz), H, L.ptr(Wl), H, L.ptr(Wr), H, H, H
    gb.Y, gb.ldy, gb.M, gb.N, gb.R, gb.dsplit = L.ptr(dx), H, n, H, 2 * H, H
    gb.Dact, gb.ldd, gb.dact0, gb.dact1 = L.ptr(yprev), H, L.ACT['relu'], L.ACT['relu']
    gb.agg_cols, gb.agg_ptr, gb.agg_nbr, gb.agg_node, gb.agg_graph = H, L.ptr(ar.out_ptr), L.ptr(ar.out_nbr), L.ptr(nda), L.ptr(ndg)
    gb.agg_cum, gb.agg_node_base = run.sel.cum, run.sel.node_base
    L.check(L.lib.sard_linear_bwd_data(C.byref(gb), L.stream()))
    dz64 = dz.double()
    aggT = torch.zeros(n, H, dtype=torch.float64, device='cuda').index_add_(0, src_t, dz64[dst_t])
    want = (aggT @ Wl.double() + dz64 @ Wr.double()) * (yprev > 0)
    assert_close(dx.cpu(), want.cpu(), rtol=1e-4, atol=2e-5, what='fused graphconv adjoint')


@pytest.mark.parametrize('M,N,R', [(2048, 256, 512), (2048, 64, 512), (1000, 512, 256)])
def test_linear_split_k_matches_dense(L, M, N, R):
    """SardGemm.splitk_ws: few rows, long reduction (critic MLP): forward with tanh and backward-data with act' and
    scattered rows must match fp64, and must be bit-identical from run to run (slice-ordered sum)."""
    g = torch.Generator().manual_seed(M + N + R)
    A = torch.randn(M, R, generator=g).cuda(); W = (torch.randn(N, R, generator=g) / np.sqrt(R)).cuda(); b = torch.randn(N, generator=g).cuda()
    ws = torch.zeros(1024 * M, device='cuda')
    Y = torch.zeros(M, N, device='cuda')
    gm = L.SardGemm()
    gm.A, gm.lda, gm.B0, gm.ldb0, gm.b_split, gm.bias = L.ptr(A), R, L.ptr(W), R, R, L.ptr(b)
    gm.Y, gm.ldy, gm.M, gm.N, gm.R, gm.act = L.ptr(Y), N, M, N, R, L.ACT['tanh']
    gm.splitk_ws, gm.splitk_floats = L.ptr(ws), ws.numel()
    n0 = L.lib.sard_launch_count()
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    launched = L.lib.sard_launch_count() - n0
    want = torch.tanh(A.double() @ W.double().t() + b.double())
    assert_close(Y.cpu(), want.cpu(), rtol=1e-4, atol=2e-5, what='split-K forward')
    if (M, N, R) != (1000, 512, 256):
        assert launched == 2, 'expected the split-K pair (partial tiles + finish)'
    Y1 = Y.clone(); Y.zero_()
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    assert torch.equal(Y, Y1)
    # backward-data: dA[yrows] = (dY W) * tanh'(saved)
    Nb, Rb = R, N                                   # dY [M, N] times W [N, R] -> [M, R]
    dY = torch.randn(M, Rb, generator=g).cuda()
    saved = torch.tanh(torch.randn(M, Nb, generator=g)).cuda()
    yrows = torch.randperm(M, generator=g).int().cuda()
    dA = torch.zeros(M, Nb, device='cuda')
    gb = L.SardGemm()
    gb.A, gb.lda, gb.B0, gb.ldb0, gb.b_split = L.ptr(dY), Rb, L.ptr(W), R, Nb
    gb.Y, gb.ldy, gb.y_rows, gb.M, gb.N, gb.R, gb.dsplit = L.ptr(dA), Nb, L.ptr(yrows), M, Nb, Rb, Nb
    gb.Dact, gb.ldd, gb.dact0, gb.dact1 = L.ptr(saved), Nb, L.ACT['tanh'], L.ACT['tanh']
    gb.splitk_ws, gb.splitk_floats = L.ptr(ws), ws.numel()
    L.check(L.lib.sard_linear_bwd_data(C.byref(gb), L.stream()))
    want = torch.zeros(M, Nb, dtype=torch.float64, device='cuda')
    want[yrows.long()] = (dY.double() @ W.double())
    want = want * (1 - saved.double() ** 2)
    assert_close(dA.cpu(), want.cpu(), rtol=1e-4, atol=2e-5, what='split-K backward-data')


# ------------------------------------------------------------------------------------------ linear layers
@pytest.mark.parametrize('M,N,R,act', [(1000, 64, 71, 'none'), (517, 64, 128, 'relu'), (300, 128, 256, 'tanh'),
                                       (77, 1, 448, 'none'), (2048, 64, 1410, 'relu'), (130, 6, 128, 'none'),
                                       (64, 512, 64, 'tanh'), (3, 3, 5, 'sigmoid')])
def test_linear_forward_dense(L, M, N, R, act):
    g = torch.Generator().manual_seed(M + N + R)
    A = torch.randn(M, R, generator=g); W = torch.randn(N, R, generator=g) / np.sqrt(R); b = torch.randn(N, generator=g)
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    Y = torch.zeros(M, N, device='cuda')
    gm = L.SardGemm()
    gm.A, gm.lda, gm.B0, gm.ldb0, gm.b_split, gm.bias = L.ptr(Ad), R, L.ptr(Wd), R, R, L.ptr(bd)
    gm.Y, gm.ldy, gm.M, gm.N, gm.R, gm.act = L.ptr(Y), N, M, N, R, L.ACT[act]
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    f = {'none': lambda v: v, 'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid}[act]
    want = f(A.double() @ W.double().t() + b.double())
    assert_close(Y.cpu(), want, rtol=1e-4, atol=2e-5, what='linear fwd')


def test_linear_forward_split_weights_gather_scatter_and_sin(L):
    g = torch.Generator().manual_seed(0)
    M, H = 333, 64
    src = torch.randn(500, 2 * H, generator=g)
    rows = torch.randperm(500, generator=g)[:M].int()
    yrows = torch.randperm(M, generator=g).int()
    Wl, Wr, b = torch.randn(H, H, generator=g) / 8, torch.randn(H, H, generator=g) / 8, torch.randn(H, generator=g)
    d = lambda t: t.cuda().contiguous()
    srcd, rowsd, yrowsd, Wld, Wrd, bd = d(src), d(rows), d(yrows), d(Wl), d(Wr), d(b)
    Y = torch.zeros(M, 80, device='cuda'); Ypre = torch.zeros(M, 80, device='cuda')
    gm = L.SardGemm()
    gm.A, gm.lda, gm.a_rows = L.ptr(srcd), 2 * H, L.ptr(rowsd)
    gm.B0, gm.ldb0, gm.B1, gm.ldb1, gm.b_split, gm.bias = L.ptr(Wld), H, L.ptr(Wrd), H, H, L.ptr(bd)
    gm.Y, gm.ldy, gm.y_rows, gm.Ypre, gm.post = Y.data_ptr() + 4 * 8, 80, L.ptr(yrowsd), Ypre.data_ptr() + 4 * 8, L.POST_SIN
    gm.M, gm.N, gm.R = M, H, 2 * H
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    a = src[rows.long()].double()
    pre = a[:, :H] @ Wl.double().t() + a[:, H:] @ Wr.double().t() + b.double()
    want = torch.zeros(M, H, dtype=torch.float64); want[yrows.long()] = torch.sin(pre * np.pi / 2)
    wpre = torch.zeros(M, H, dtype=torch.float64); wpre[yrows.long()] = pre
    assert_close(Y[:, 8:8 + H].cpu(), want, rtol=1e-4, atol=2e-5, what='sin out')
    assert_close(Ypre[:, 8:8 + H].cpu(), wpre, rtol=1e-4, atol=2e-5, what='pre')
    assert float(Y[:, :8].abs().max()) == 0 and float(Y[:, 72:].abs().max()) == 0


def _sorted_tables(L, body, mi):
    n = len(body)
    b = dev(body, torch.int32)
    i32 = lambda k: torch.zeros(k, dtype=torch.int32, device='cuda')
    t = dict(srow=i32(n), spos=i32(n), grp=i32(mi + 1), gcp=i32(mi + 1), counts=i32(4),
             tiles=i32(4 * (n // 128 + 1 + mi)), chunks=i32(4 * (n // 256 + 1 + mi)), scratch=i32(((n + 255) // 256 + 2) * mi))
    L.check(L.lib.sard_sort_by_index(L.ptr(b), n, mi, 128, 256, L.ptr(t['srow']), L.ptr(t['spos']), L.ptr(t['grp']), L.ptr(t['tiles']),
                                     L.ptr(t['chunks']), L.ptr(t['gcp']), L.ptr(t['counts']), L.ptr(t['scratch']), L.stream()))
    t['max_tiles'] = n // 128 + 1 + mi
    t['max_chunks'] = n // 256 + 1 + mi
    return t


def test_index_linear_grouped_forward_backward_vs_oracle(L):
    """IndexLinear (jsmlp.py:32-40) as a grouped GEMM over rows sorted by body index, all three passes."""
    from oracle import sard_oracle as O
    rng = np.random.default_rng(11)
    n, K, N, mi = 1500, 96, 48, 40
    live = np.array([0, 1, 2, 3, 4, 6, 7, 8, 9, 31, 32, 33, 34])
    body = live[rng.integers(0, len(live), size=n)]
    x = torch.from_numpy(rng.normal(size=(n, K)).astype(np.float32))
    W = torch.from_numpy((rng.normal(size=(mi, N, K)) / np.sqrt(K)).astype(np.float32))
    bb = torch.from_numpy(rng.normal(size=(mi, N)).astype(np.float32))
    dY = torch.from_numpy(rng.normal(size=(n, N)).astype(np.float32))
    t = _sorted_tables(L, body, mi)
    srow = t['srow'].long()
    xs = x.cuda()[srow].contiguous()                   # rows in sorted order
    Wd, bd = W.cuda(), bb.cuda()
    Y = torch.zeros(n, N, device='cuda')
    gm = L.SardGemm()
    gm.A, gm.lda, gm.B0, gm.ldb0, gm.b_split = L.ptr(xs), K, L.ptr(Wd), K, K
    gm.b_group_stride, gm.bias, gm.bias_group_stride = N * K, L.ptr(bd), N
    gm.Y, gm.ldy, gm.y_rows, gm.M, gm.N, gm.R, gm.act = L.ptr(Y), N, L.ptr(t['srow']), n, N, K, L.ACT['tanh']
    gm.tiles, gm.num_tiles, gm.max_tiles = L.ptr(t['tiles']), L.ptr(t['counts']), t['max_tiles']
    L.check(L.lib.sard_linear_fwd(C.byref(gm), L.stream()))
    x64 = x.double().requires_grad_(True); W64 = W.double().requires_grad_(True); b64 = bb.double().requires_grad_(True)
    want = torch.tanh(O.index_linear(W64, b64, x64, torch.from_numpy(body)))
    assert_close(Y.cpu(), want.detach(), rtol=1e-4, atol=2e-5, what='index_linear fwd')
    want.backward(dY.double())
    # backward-data: dX (sorted space) = (dZ W[u]) with dZ = dY * tanh'(Y)
    dZ = (dY.cuda() * (1 - Y * Y)).contiguous()
    dX = torch.zeros(n, K, device='cuda')
    gd = L.SardGemm()
    gd.A, gd.lda, gd.a_rows, gd.B0, gd.ldb0, gd.b_split, gd.b_group_stride = L.ptr(dZ), N, L.ptr(t['srow']), L.ptr(Wd), K, K, N * K
    gd.Y, gd.ldy, gd.M, gd.N, gd.R, gd.dsplit = L.ptr(dX), K, n, K, N, K
    gd.tiles, gd.num_tiles, gd.max_tiles = L.ptr(t['tiles']), L.ptr(t['counts']), t['max_tiles']
    L.check(L.lib.sard_linear_bwd_data(C.byref(gd), L.stream()))
    got_dx = torch.zeros(n, K, device='cuda'); got_dx[srow] = dX
    assert_close(got_dx.cpu(), x64.grad, rtol=1e-3, atol=1e-4, what='index_linear dX')
    # backward-weight into compact slots
    slots = np.full(mi, -1, dtype=np.int32); slots[live] = np.arange(len(live))
    slot_d = dev(slots)
    gW = torch.zeros(len(live), N, K, device='cuda'); gb = torch.zeros(len(live), N, device='cuda')
    part = torch.zeros(t['max_chunks'] * N * (K + 4), device='cuda')
    w = L.SardWGrad()
    w.P, w.ldp, w.p_rows, w.Q, w.ldq, w.M, w.N1, w.N2 = L.ptr(dZ), N, L.ptr(t['srow']), L.ptr(xs), K, n, N, K
    w.chunks, w.num_chunks, w.max_chunks, w.chunk_rows = L.ptr(t['chunks']), t['counts'].data_ptr() + 4, t['max_chunks'], 256
    w.grp_chunk_ptr, w.n_groups, w.slot_of_group = L.ptr(t['gcp']), mi, L.ptr(slot_d)
    w.dW0, w.ldw0, w.w_split, w.w_slot_stride, w.db, w.b_slot_stride, w.partials = L.ptr(gW), K, K, N * K, L.ptr(gb), N, L.ptr(part)
    L.check(L.lib.sard_linear_bwd_weight(C.byref(w), L.stream()))
    assert_close(gW.cpu(), W64.grad[live], rtol=1e-3, atol=1e-4, what='index_linear dW')
    assert_close(gb.cpu(), b64.grad[live], rtol=1e-3, atol=1e-4, what='index_linear db')
    dead = np.setdiff1d(np.arange(mi), live)
    assert float(W64.grad[dead].abs().max()) == 0.0      # rows the compact buffer never stores are exactly zero


@pytest.mark.parametrize('M,N1,N2', [(1000, 64, 128), (5000, 64, 71), (2048, 64, 1410), (300, 1, 448), (257, 16, 38)])
def test_weight_gradient_dense_split(L, M, N1, N2):
    g = torch.Generator().manual_seed(M)
    dZ = torch.randn(M, N1, generator=g); A = torch.randn(M, N2, generator=g)
    dZd, Ad = dZ.cuda(), A.cuda()
    split = N2 // 2
    dW0 = torch.ones(N1, split, device='cuda'); dW1 = torch.ones(N1, N2 - split, device='cuda'); db = torch.ones(N1, device='cuda')
    part = torch.zeros(((M + 255) // 256) * N1 * (N2 + 4), device='cuda')
    w = L.SardWGrad()
    w.P, w.ldp, w.Q, w.ldq, w.M, w.N1, w.N2, w.chunk_rows = L.ptr(dZd), N1, L.ptr(Ad), N2, M, N1, N2, 256
    w.dW0, w.ldw0, w.dW1, w.ldw1, w.w_split, w.db, w.partials = L.ptr(dW0), split, L.ptr(dW1), N2 - split, split, L.ptr(db), L.ptr(part)
    L.check(L.lib.sard_linear_bwd_weight(C.byref(w), L.stream()))
    want = dZ.double().t() @ A.double() + 1.0            # accumulates into the existing gradient
    assert_close(torch.cat([dW0, dW1], 1).cpu(), want, rtol=1e-3, atol=1e-3, what='dW')
    assert_close(db.cpu(), dZ.double().sum(0) + 1.0, rtol=1e-3, atol=1e-3, what='db')


def test_backward_data_with_split_activation_derivative(L):
    g = torch.Generator().manual_seed(4)
    M, N, R, ds = 700, 448, 1, 256          # the value head: dcat = dv * W, tanh' then relu'
    dv = torch.randn(M, R, generator=g); W = torch.randn(R, N, generator=g); Yp = torch.randn(M, N, generator=g)
    Y = torch.zeros(M, N, device='cuda')
    dvd, Wd, Ypd = dv.cuda(), W.cuda(), Yp.cuda()
    gd = L.SardGemm()
    gd.A, gd.lda, gd.B0, gd.ldb0, gd.b_split = L.ptr(dvd), R, L.ptr(Wd), N, N
    gd.Y, gd.ldy, gd.M, gd.N, gd.R = L.ptr(Y), N, M, N, R
    gd.Dact, gd.ldd, gd.dact0, gd.dact1, gd.dsplit = L.ptr(Ypd), N, L.ACT['tanh'], L.ACT['relu'], ds
    L.check(L.lib.sard_linear_bwd_data(C.byref(gd), L.stream()))
    want = dv.double() @ W.double()
    want[:, :ds] *= (1 - Yp[:, :ds].double() ** 2)
    want[:, ds:] *= (Yp[:, ds:] > 0)
    assert_close(Y.cpu(), want, rtol=1e-4, atol=1e-5, what='bwd data')


# ------------------------------------------------------------------------------------------ GAE / Adam
@pytest.mark.parametrize('T', [1, 7, 2048, 2049, 50000])
def test_gae_matches_oracle_within_1e6(L, T):
    from oracle import sard_oracle as O
    from sard_b200.rl_core import estimate_advantages
    if T < 2:
        pytest.skip('std of one element is undefined (reference gives nan)')
    rng = np.random.default_rng(T)
    r = torch.from_numpy(rng.normal(0.01, 0.1, size=T).astype(np.float32))
    v = torch.from_numpy(rng.normal(0, 1, size=(T, 1)).astype(np.float32))
    m = torch.from_numpy((rng.random(T) > 0.01).astype(np.float32))
    adv, ret = estimate_advantages(r.cuda(), m.cuda(), v.cuda(), 0.995, 0.95)
    wadv, wret = O.estimate_advantages(r.double(), m.double(), v.double(), 0.995, 0.95)
    assert_close(adv.cpu(), wadv, rtol=0, atol=1e-6, what='advantages')
    assert_close(ret.cpu(), wret, rtol=0, atol=2e-6, what='returns')
    a64, r64 = estimate_advantages(r.double().cuda(), m.double().cuda(), v.double().cuda(), 0.995, 0.95)
    assert_close(a64.cpu(), wadv, rtol=0, atol=1e-10, what='advantages f64')
    assert_close(r64.cpu(), wret, rtol=0, atol=1e-10, what='returns f64')


def test_clip_and_adam_match_torch(L):
    """sqnorm + prepare + apply == clip_grad_norm_ + torch.optim.Adam over several steps, incl. a gated step."""
    from sard_b200.nets import AdamState
    g = torch.Generator().manual_seed(8)
    n = 70000
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref], lr=5e-5)
    p = p0.cuda().clone()
    st = AdamState(5e-5)
    st.ensure(p)
    grad = torch.zeros(16 + n, device='cuda')
    segs = [(p.data_ptr() + 4 * o, st.m.data_ptr() + 4 * o, st.v.data_ptr() + 4 * o, grad.data_ptr() + 4 * (16 + o), min(32768, n - o), 0)
            for o in range(0, n, 32768)]
    st.set_segments(segs, p.device)
    sq_s, sq_o = torch.zeros(1024, device='cuda'), torch.zeros(1, device='cuda')
    log = torch.zeros(8, device='cuda')
    for it in range(5):
        gk = torch.randn(n, generator=g) * (3.0 if it == 0 else 0.01)
        kl = 0.5 if it == 2 else 0.01                      # step 2 is gated off (approx_kl > 0.1)
        grad.zero_(); grad[16:] = gk.cuda(); grad[0] = -1.0; grad[1] = kl * 4
        L.check(L.lib.sard_sqnorm(grad.data_ptr() + 64, n, L.ptr(sq_s), L.ptr(sq_o), L.stream()))
        L.check(L.lib.sard_adam_prepare(L.ptr(grad), 0.25, 1.0, 0.0, L.ptr(sq_o), 40.0, 1, 1, 1, 0.9, 0.999, L.ptr(st.ctl_f), L.ptr(st.ctl_i),
                                        L.ptr(log), L.stream()))
        L.check(L.lib.sard_adam_apply(L.ptr(st.segs), st.n_seg, st.max_seg, 5e-5, 0.9, 0.999, 1e-8, 1, L.ptr(st.ctl_f), L.ptr(st.ctl_i), L.stream()))
        lg = log.cpu().numpy()
        assert_close(lg[4], float(gk.double().norm()), rtol=1e-5, atol=0, what='grad norm')
        if it != 2:
            ref.grad = gk.double().clone()
            if it == 0:                                   # the reference clips only on its first policy step
                torch.nn.utils.clip_grad_norm_([ref], 40)
            opt.step()
            assert lg[3] == 1.0
        else:
            assert lg[3] == 0.0
        assert_close(p.cpu(), ref.detach(), rtol=2e-6, atol=5e-7, what=f'params after step {it}')
    assert int(st.ctl_i[8].item()) == 4 and int(st.ctl_i[1].item()) == 0


@pytest.mark.parametrize('sticky', [False, True])
def test_adam_update_single_launch_equals_prepare_plus_apply(L, sticky):
    """sard_adam_update without a gradient norm runs gate + bias corrections + Adam in ONE launch (adam_fused_kernel);
    bit-identical to sard_adam_prepare + sard_adam_apply over several steps with three groups, absent groups, a gated
    step and ragged / unaligned segments, and it leaves statistics and gradients zeroed."""
    from sard_b200 import _lib
    from sard_b200.nets import AdamState
    g = torch.Generator().manual_seed(21)
    sizes = [(4096, 0), (4096, 0), (1234, 1), (4096, 2), (3, 2), (4095, 1)]
    n = sum(s for s, _ in sizes) + 8
    p0 = torch.randn(n, generator=g)
    states = []
    for _ in range(2):
        p = p0.cuda().clone()
        st = AdamState(3e-4)
        st.ensure(p)
        grad = torch.zeros(16 + n, device='cuda')
        segs, o = [], 0                                   # the first three segments are 16-byte aligned, the rest are not
        for sz, grp in sizes:
            segs.append((p.data_ptr() + 4 * o, st.m.data_ptr() + 4 * o, st.v.data_ptr() + 4 * o, grad.data_ptr() + 4 * (16 + o), sz, grp))
            o += sz
        st.set_segments(segs, p.device)
        states.append((p, st, grad, torch.zeros(8, device='cuda')))
    masks = [0b111, 0b001, 0b011, 0b101, 0b111, 0b001]
    for it, mask in enumerate(masks):
        gk = torch.randn(n, generator=g).cuda() * 0.1
        kl = 0.9 if it == 3 else 0.02                     # step 3 is gated off
        m = mask | (_lib.ADAM_STICKY if sticky else 0)
        for which, (p, st, grad, log) in enumerate(states):
            grad.zero_()
            o = 0
            for sz, grp in sizes:                          # only the active groups receive gradients
                if (mask >> grp) & 1:
                    grad[16 + o:16 + o + sz] = gk[o:o + sz]
                o += sz
            grad[0] = -0.3; grad[1] = kl * 8; grad[2] = 0.7; grad[4] = 1.1; grad[5] = 2.0
            if which == 0:
                L.lib.sard_set_adam_single_launch(1)
                L.check(L.lib.sard_adam_update(L.ptr(grad), 0.125, 1.0, 0.5, None, 0, None, 40.0, 1, 1, m, 0.9, 0.999,
                                               L.ptr(st.ctl_f), L.ptr(st.ctl_i), L.ptr(log), L.ptr(st.segs), st.n_seg, st.max_seg,
                                               3e-4, 1e-8, L.stream()))
                L.lib.sard_set_adam_single_launch(0)
            else:
                L.check(L.lib.sard_adam_prepare(L.ptr(grad), 0.125, 1.0, 0.5, None, 40.0, 1, 1, m, 0.9, 0.999, L.ptr(st.ctl_f),
                                                L.ptr(st.ctl_i), L.ptr(log), L.stream()))
                L.check(L.lib.sard_adam_apply(L.ptr(st.segs), st.n_seg, st.max_seg, 3e-4, 0.9, 0.999, 1e-8, m, L.ptr(st.ctl_f),
                                              L.ptr(st.ctl_i), L.stream()))
        (pa, sa, ga, la), (pb, sb_, gb, lb) = states
        if not torch.equal(pa, pb):
            bad = (pa != pb).nonzero().flatten()
            raise AssertionError(f'parameters differ after step {it}: {bad.numel()} entries, first {bad[:4].tolist()} last {bad[-4:].tolist()}, '
                                 f'max diff {float((pa - pb).abs().max())}; ctl_i {sa.ctl_i[:16].tolist()} vs {sb_.ctl_i[:16].tolist()}; '
                                 f'ctl_f {sa.ctl_f[8:11].tolist()} {sa.ctl_f[16:19].tolist()} vs {sb_.ctl_f[8:11].tolist()} {sb_.ctl_f[16:19].tolist()}')
        assert torch.equal(sa.m, sb_.m) and torch.equal(sa.v, sb_.v)
        assert torch.equal(la[:6], lb[:6]), (la, lb)
        ia, ib = sa.ctl_i.cpu(), sb_.ctl_i.cpu()
        assert torch.equal(ia[8:16], ib[8:16]) and ia[0] == ib[0] and ia[2] == ib[2] and int(ia[4]) == 0
        assert float(ga.abs().sum()) == 0.0               # statistics and gradients left zeroed

"""GPU probe of the fused tcgen05 JSMLP kernels against fp64 torch (and timing against the SIMT path).

    python tools/tc_probe.py [fwd] [bwd] [wgrad]      (default: everything that is built)
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sard_b200 import _lib as L

torch.cuda.set_device(0)
dev = 'cuda'
LIVE = [0, 1, 2, 3, 4, 6, 7, 8, 9, 31, 32, 33, 34]
MAXI = 256


def make_case(n, in_dim, out_dim, seed=0, post=0):
    g = torch.Generator().manual_seed(seed)
    body = torch.tensor(LIVE)[torch.randint(0, len(LIVE), (n,), generator=g)].int()
    # skew like real rollouts: many roots / legs, few deep nodes
    body[: n // 3] = torch.tensor(LIVE[:5])[torch.randint(0, 5, (n // 3,), generator=g)].int()
    W = [torch.zeros(MAXI, 128, in_dim), torch.zeros(MAXI, 128, 128), torch.zeros(MAXI, out_dim, 128)]
    b = [torch.zeros(MAXI, 128), torch.zeros(MAXI, 128), torch.zeros(MAXI, out_dim)]
    for u in LIVE:
        W[0][u] = torch.randn(128, in_dim, generator=g) / in_dim ** 0.5
        W[1][u] = torch.randn(128, 128, generator=g) / 128 ** 0.5
        W[2][u] = torch.randn(out_dim, 128, generator=g) / 128 ** 0.5
        for j in range(3):
            b[j][u] = torch.randn(b[j].shape[1], generator=g) * 0.1
    x = torch.randn(n, in_dim, generator=g)
    slot = torch.full((MAXI,), -1, dtype=torch.int32)
    slot[LIVE] = torch.arange(len(LIVE), dtype=torch.int32)
    return dict(n=n, in_dim=in_dim, out_dim=out_dim, post=post, body=body.to(dev), W=[w.to(dev) for w in W],
                b=[v.to(dev) for v in b], x=x.to(dev), slot=slot.to(dev))


def sort_rows(c, tile_rows=128, chunk_rows=256):
    n = c['n']
    i32 = lambda k: torch.zeros(k, dtype=torch.int32, device=dev)
    max_tiles = n // tile_rows + len(LIVE)
    max_chunks = n // chunk_rows + len(LIVE)
    c.update(srow=i32(n), spos=i32(n), grp_ptr=i32(MAXI + 1), tiles=i32(4 * max_tiles), chunks=i32(4 * max_chunks),
             grp_chunk_ptr=i32(MAXI + 1), counts=i32(4), max_tiles=max_tiles, max_chunks=max_chunks)
    scratch = i32(((n + 255) // 256 + 2) * MAXI)
    L.check(L.lib.sard_sort_by_index(L.ptr(c['body']), n, MAXI, tile_rows, chunk_rows, L.ptr(c['srow']), L.ptr(c['spos']),
                                     L.ptr(c['grp_ptr']), L.ptr(c['tiles']), L.ptr(c['chunks']), L.ptr(c['grp_chunk_ptr']),
                                     L.ptr(c['counts']), L.ptr(scratch), L.stream()))
    torch.cuda.synchronize()
    c['xs'] = c['x'][c['srow'].long()].contiguous()           # sorted rows
    c['body_s'] = c['body'][c['srow'].long()].long()


def reference(c):
    xs, ind = c['xs'].double(), c['body_s']
    W, b = [w.double() for w in c['W']], [v.double() for v in c['b']]
    h1 = torch.tanh(torch.einsum('nk,nok->no', xs, W[0][ind]) + b[0][ind])
    h2 = torch.tanh(torch.einsum('nk,nok->no', h1, W[1][ind]) + b[1][ind])
    pre = torch.einsum('nk,nok->no', h2, W[2][ind]) + b[2][ind]
    out = torch.sin(pre * np.pi / 2) if c['post'] else pre
    return h1, h2, pre, out


def tc_args(c, raw=True):
    n, in_dim, out_dim = c['n'], c['in_dim'], c['out_dim']
    a = L.SardJsmlpTc()
    a.n, a.in_dim, a.hid, a.out_dim, a.act, a.post, a.max_index, a.n_slots = n, in_dim, 128, out_dim, 2, c['post'], MAXI, len(LIVE)
    a.xs, a.ld_xs = L.ptr(c['xs']), in_dim
    for j in range(3):
        a.W[j] = L.ptr(c['W'][j]); a.b[j] = L.ptr(c['b'][j])
    a.slot_of_index = L.ptr(c['slot'])
    a.tiles, a.num_tiles, a.max_tiles = L.ptr(c['tiles']), L.ptr(c['counts']), c['max_tiles']
    a.srow = L.ptr(c['srow'])
    c['H'] = [torch.zeros(n, 128, device=dev) for _ in range(2)]
    c['Hhl'] = [torch.zeros(n, 256, device=dev) for _ in range(2)]
    ldo = (out_dim + 3) // 4 * 4
    c['out'] = torch.zeros(n, ldo, device=dev); c['pre'] = torch.zeros(n, ldo, device=dev)
    for j in range(2):
        a.H[j] = L.ptr(c['H'][j]) if raw else None
        a.Hhl[j] = L.ptr(c['Hhl'][j])
    a.ldh, a.out, a.pre, a.ldo = 128, L.ptr(c['out']), L.ptr(c['pre']), ldo
    need = L.lib.sard_jsmlp_tc_workspace(n, in_dim, out_dim, len(LIVE))
    c['ws'] = torch.zeros(need + 64, device=dev)
    off = (-c['ws'].data_ptr() // 4) % 64
    a.ws, a.ws_floats = c['ws'].data_ptr() + 4 * off, need
    return a


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def probe_fwd(n, in_dim, out_dim, post):
    c = make_case(n, in_dim, out_dim, seed=n % 97, post=post)
    sort_rows(c)
    a = tc_args(c)
    st = L.stream()
    L.check(L.lib.sard_jsmlp_tc_presplit(C.byref(a), st))
    L.check(L.lib.sard_jsmlp_tc_split_input(C.byref(a), st))
    L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st))
    torch.cuda.synchronize()
    h1, h2, pre, out = reference(c)
    inv = c['srow'].long()                      # out rows are scattered to their original row srow[p]
    d = out_dim
    e = dict(h1=rel(c['H'][0], h1), h2=rel(c['H'][1], h2), pre=rel(c['pre'][inv][:, :d], pre), out=rel(c['out'][inv][:, :d], out),
             h1hl=rel(c['Hhl'][0][:, :128].double() + c['Hhl'][0][:, 128:].double(), h1),
             h2hl=rel(c['Hhl'][1][:, :128].double() + c['Hhl'][1][:, 128:].double(), h2))
    us_raw = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st)))
    a = tc_args(c, raw=False)                   # the production configuration: pair copies only (TMA stores)
    L.check(L.lib.sard_jsmlp_tc_presplit(C.byref(a), st))
    L.check(L.lib.sard_jsmlp_tc_split_input(C.byref(a), st))
    L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st))
    torch.cuda.synchronize()
    e['h2hl_noraw'] = rel(c['Hhl'][1][:, :128].double() + c['Hhl'][1][:, 128:].double(), h2)
    e['out_noraw'] = rel(c['out'][inv][:, :d], out)
    us = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st)))
    us_pre = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_presplit(C.byref(a), st)))
    us_split = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_split_input(C.byref(a), st)))
    flops = 2.0 * n * (in_dim * 128 + 128 * 128 + 128 * out_dim)
    ok = all(v < 2e-5 for v in e.values())
    print(f'fwd n={n} in={in_dim} out={out_dim} post={post}: {"OK " if ok else "BAD"} rel err {e}  '
          f'{us:.1f} us ({flops / us / 1e6:.1f} TFLOP/s; with raw H copies {us_raw:.1f} us), presplit {us_pre:.1f} us, split_input {us_split:.1f} us, '
          f'tiles={int(c["counts"][0])}', flush=True)
    buf = (C.c_longlong * 32)()
    L.lib.sard_tc_debug_read(buf, 32)
    t = list(buf)
    print('   block0 cycles: pdl_wait %d | L1 issue done +%d | d1 ready +%d | epi1 done +%d | L2 issue done +%d | d2 ready +%d | epi2 done +%d | d3 ready +%d | end +%d'
          % (t[1] - t[0], t[8] - t[1], t[2] - t[1], t[3] - t[1], t[9] - t[1], t[4] - t[1], t[5] - t[1], t[6] - t[1], t[7] - t[1]), flush=True)
    return ok


def probe_simt_fwd(n, in_dim, out_dim):
    """the three grouped sard_linear_fwd launches of the SIMT path on the same case, for the timing comparison"""
    c = make_case(n, in_dim, out_dim, seed=n % 97)
    sort_rows(c)
    H = [torch.zeros(n, 128, device=dev), torch.zeros(n, 128, device=dev), torch.zeros(n, 4 * ((out_dim + 3) // 4), device=dev)]
    gs = []
    dims = [(in_dim, 128), (128, 128), (128, out_dim)]
    for j, (i, o) in enumerate(dims):
        g = L.SardGemm()
        A = c['xs'] if j == 0 else H[j - 1]
        g.A, g.lda = L.ptr(A), i
        g.B0, g.ldb0, g.b_split, g.b_group_stride = L.ptr(c['W'][j]), i, i, o * i
        g.bias, g.bias_group_stride = L.ptr(c['b'][j]), o
        g.M, g.N, g.R = n, o, i
        g.tiles, g.num_tiles, g.max_tiles = L.ptr(c['tiles']), L.ptr(c['counts']), c['max_tiles']
        g.act = 2 if j < 2 else 0
        g.Y, g.ldy = L.ptr(H[j]), H[j].shape[1]
        gs.append(g)
    c['keep'] = H
    st = L.stream()

    def run():
        for g in gs:
            L.check(L.lib.sard_linear_fwd(C.byref(g), st))
    us = timeit(run)
    print(f'simt fwd n={n}: {us:.1f} us (3 launches)', flush=True)


def probe_bwd(n, out_dim, seed=0, chunk_rows=256, time_it=True):
    """fused backward-data + weight gradients in the padded row space against fp64 autograd"""
    in_dim = 256
    c = make_case(n, in_dim, out_dim, seed=seed)
    nl = len(LIVE)
    cap = (n + 127) // 128 * 128 + nl * 128
    i32 = lambda k: torch.zeros(k, dtype=torch.int32, device=dev)
    max_tiles, max_chunks = cap // 128, n // chunk_rows + nl
    srow, spos, grp_ptr, tiles, chunks, gcp, counts = i32(cap), i32(n), i32(MAXI + 1), i32(4 * max_tiles), i32(4 * max_chunks), i32(MAXI + 1), i32(4)
    scratch = i32(((n + 255) // 256 + 2) * MAXI)
    st = L.stream()
    L.check(L.lib.sard_sort_by_index_padded(L.ptr(c['body']), n, MAXI, 128, chunk_rows, cap, L.ptr(srow), L.ptr(spos), L.ptr(grp_ptr),
                                            L.ptr(tiles), L.ptr(chunks), L.ptr(gcp), L.ptr(counts), L.ptr(scratch), st))
    torch.cuda.synchronize()
    used = int(counts[3])
    assert used <= cap and used % 128 == 0, (used, cap)
    sr = srow.long()
    real = sr >= 0
    assert int(real.sum()) == n and bool((sr[spos.long()] == torch.arange(n, device=dev)).all())
    a = L.SardJsmlpTc()
    a.n, a.in_dim, a.hid, a.out_dim, a.act, a.post, a.max_index, a.n_slots = cap, in_dim, 128, out_dim, 2, 0, MAXI, nl
    for j in range(3):
        a.W[j] = L.ptr(c['W'][j]); a.b[j] = L.ptr(c['b'][j])
    a.slot_of_index = L.ptr(c['slot'])
    a.tiles, a.num_tiles, a.max_tiles = L.ptr(tiles), L.ptr(counts), max_tiles
    a.srow, a.padded = L.ptr(srow), 1
    Hhl = [torch.zeros(cap, 256, device=dev) for _ in range(2)]
    ldo = (out_dim + 3) // 4 * 4
    out = torch.zeros(n, ldo, device=dev)
    for j in range(2):
        a.Hhl[j] = L.ptr(Hhl[j])
    a.ldh, a.out, a.ldo = 128, L.ptr(out), ldo
    need = L.lib.sard_jsmlp_tc_workspace(cap, in_dim, out_dim, nl)
    ws = torch.zeros(need + 64, device=dev)
    off = (-ws.data_ptr() // 4) % 64
    a.ws, a.ws_floats = ws.data_ptr() + 4 * off, need
    xs_hl = L.lib.sard_jsmlp_tc_xs_hl(C.byref(a))
    nd_graph = i32(n)
    L.check(L.lib.sard_jsmlp_tc_presplit(C.byref(a), st))
    L.check(L.lib.sard_concat_sorted_split(L.ptr(c['x']), in_dim, in_dim, L.ptr(c['x']), in_dim, 0, L.ptr(srow), L.ptr(nd_graph), cap, xs_hl, st))
    L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st))
    torch.cuda.synchronize()
    # fp64 autograd reference in the original row order
    x = c['x'].double().requires_grad_(True)
    ind = c['body'].long()
    W = [w.double().requires_grad_(True) for w in c['W']]
    bb = [v.double().requires_grad_(True) for v in c['b']]
    h1 = torch.tanh(torch.einsum('nk,nok->no', x, W[0][ind]) + bb[0][ind])
    h2 = torch.tanh(torch.einsum('nk,nok->no', h1, W[1][ind]) + bb[1][ind])
    o = torch.einsum('nk,nok->no', h2, W[2][ind]) + bb[2][ind]
    g = torch.Generator().manual_seed(seed + 5)
    dout = torch.zeros(n, ldo)
    dout[:, :out_dim] = torch.randn(n, out_dim, generator=g)
    dout = dout.to(dev)
    (o * dout[:, :out_dim].double()).sum().backward()
    e = dict(out=rel(out[:, :out_dim], o.detach()))
    b = L.SardJsmlpTcBwd()
    dOs, dZ, dxs = torch.zeros(cap, 64, device=dev), [torch.zeros(cap, 256, device=dev) for _ in range(2)], torch.zeros(cap, 256, device=dev)
    part = torch.zeros(L.lib.sard_jsmlp_tc_partials_floats(max_chunks), device=dev)
    gW = [torch.zeros(nl, 128, 256, device=dev), torch.zeros(nl, 128, 128, device=dev), torch.zeros(nl, out_dim, 128, device=dev)]
    gb = [torch.zeros(nl, 128, device=dev), torch.zeros(nl, 128, device=dev), torch.zeros(nl, out_dim, device=dev)]
    b.dout, b.ld_dout, b.dOs_hl, b.dxs, b.ld_dxs = L.ptr(dout), ldo, L.ptr(dOs), L.ptr(dxs), 256
    for j in range(2):
        b.dZhl[j] = L.ptr(dZ[j])
    b.chunks, b.num_chunks, b.max_chunks, b.grp_chunk_ptr, b.partials = L.ptr(chunks), counts.data_ptr() + 4, max_chunks, L.ptr(gcp), L.ptr(part)
    for j in range(3):
        b.gW[j] = L.ptr(gW[j]); b.gb[j] = L.ptr(gb[j])
    L.check(L.lib.sard_jsmlp_tc_backward(C.byref(a), C.byref(b), st))
    torch.cuda.synchronize()
    e['dx'] = rel(dxs[spos.long()], x.grad)
    L.check(L.lib.sard_jsmlp_tc_wgrad(C.byref(a), C.byref(b), st))
    torch.cuda.synchronize()
    if os.environ.get('PROBE_DEBUG'):
        P = part.view(-1, 128, 480)
        buf = (C.c_float * 64)()
        L.lib.sard_jw_debug_read(buf, 64)
        print('jw dbg', [round(v, 4) for v in buf])
        print('dZ1 row0', dZ[0][0, :8].tolist(), 'dOs row0', dOs[0, :4].tolist())
        ch = chunks.view(-1, 4)[: int(counts[1])].cpu()
        print('partials absmax', float(P.abs().max()), 'gW absmax', [float(g.abs().max()) for g in gW], 'chunks', ch[:3].tolist(), 'gcp', gcp[:12].tolist())
        r0, r1 = int(ch[0, 0]), int(ch[0, 1])
        z1 = (dZ[0][r0:r1, :128].double() + dZ[0][r0:r1, 128:].double())
        xsv = torch.zeros(cap, 512, device=dev)
        C.memmove  # noqa
        xs_t = torch.empty(0)
        xs_full = c['x'][sr.clamp_min(0)] * real[:, None]
        want1 = z1.T @ xs_full[r0:r1].double()
        print('chunk0 dW1 partial: got absmax', float(P[0, :, :256].abs().max()), 'want absmax', float(want1.abs().max()),
              'err', float((P[0, :, :256].double() - want1).abs().max()))
        print('got[0,:8]', P[0, 0, :8].tolist()); print('want[0,:8]', want1[0, :8].tolist())
        print('db1 got', P[0, :4, 416].tolist(), 'want', z1.sum(0)[:4].tolist())
        print('db3 got', P[0, 0, 448:452].tolist())
    live = torch.tensor(LIVE, device=dev)
    for j in range(3):
        e[f'dW{j + 1}'] = rel(gW[j], W[j].grad[live])
        e[f'db{j + 1}'] = rel(gb[j], bb[j].grad[live])
    ok = all(v < 3e-5 for v in e.values())
    msg = f'bwd n={n} out={out_dim}: {"OK " if ok else "BAD"} rel err ' + ' '.join(f'{k}={v:.2e}' for k, v in e.items())
    if time_it:
        us_b = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_backward(C.byref(a), C.byref(b), st)))
        us_w = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_wgrad(C.byref(a), C.byref(b), st)))
        us_f = timeit(lambda: L.check(L.lib.sard_jsmlp_tc_forward(C.byref(a), st)))
        msg += f'  fwd {us_f:.1f} us, bwd-data {us_b:.1f} us, wgrad+reduce {us_w:.1f} us (tiles {int(counts[0])}, chunks {int(counts[1])}, chunk_rows {chunk_rows})'
    print(msg, flush=True)
    return ok


def probe_umma():
    L.lib.sard_umma_probe.restype = C.c_int
    L.lib.sard_umma_probe.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]
    g = torch.Generator().manual_seed(1)
    for N in (32, 128, 256):
        A = torch.randint(-4, 5, (8, 128), generator=g).float().to(dev)
        B = torch.randint(-4, 5, (8, N), generator=g).float().to(dev)
        want = A.T @ B
        for a_mn, b_mn in ((0, 0), (1, 0), (0, 1), (1, 1)):
            for lbo, sbo in ((1024, 512), (512, 1024), (1024, 1024)):
                D = torch.full((128, N), -77.0, device=dev)
                L.check(L.lib.sard_umma_probe(L.ptr(A), L.ptr(B), L.ptr(D), N, a_mn, b_mn, lbo, sbo, L.stream()))
                torch.cuda.synchronize()
                err = float((D - want).abs().max())
                print(f'umma N={N} a_mn={a_mn} b_mn={b_mn} lbo={lbo} sbo={sbo}: maxerr {err} absmax {float(D.abs().max())}', flush=True)


if __name__ == '__main__':
    what = set(sys.argv[1:]) or {'fwd'}
    ok = True
    if 'umma' in what:
        probe_umma()
    if 'bwd' in what:
        cases = [(300, 1)] if os.environ.get('PROBE_DEBUG') else [(300, 1), (5000, 6), (18432, 1), (18432, 3)]
        for n, o in cases:
            ok &= probe_bwd(n, o, seed=n % 89)
        if not os.environ.get('PROBE_DEBUG'):
            for cr in (128, 512, 1024):
                ok &= probe_bwd(18432, 1, seed=3, chunk_rows=cr)
    if 'fwd' in what:
        for n, i, o, post in [(300, 256, 1, 0), (5000, 256, 6, 1), (18432, 256, 1, 0), (18432, 256, 3, 0), (26000, 256, 6, 1)]:
            ok &= probe_fwd(n, i, o, post)
        probe_simt_fwd(18432, 256, 1)
    print('PROBE', 'PASS' if ok else 'FAIL')
    sys.exit(0 if ok else 1)

"""Micro-benchmark of sard_linear_fwd / bwd_data / bwd_weight at the bench workload's shapes, both engines."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sard_b200 import _lib as L
torch.cuda.set_device(0)
shapes = [('x tiny', 128, 64, 8, True), ('x onetile', 128, 128, 256, True), ('x 18k r8', 18432, 128, 8, True), ('x 18k n64 r8', 18432, 64, 8, True), ('x 2k n128 r64', 2048, 128, 64, True),
          ('il1 fwd', 18432, 128, 256, True), ('il2 fwd', 18432, 128, 128, True), ('gconv fwd', 18432, 64, 128, True),
          ('in_fc fwd', 18432, 64, 72, True), ('mlp1 fwd', 2048, 512, 64, True), ('mlp2 fwd', 2048, 256, 512, True),
          ('il1 bwd_data', 18432, 256, 128, False), ('gconv bwd_data', 18432, 128, 64, False)]
iters = int(os.environ.get('ITERS', '30'))
only = os.environ.get('ONLY')
for name, M, N, R, nt in shapes:
    if only and only not in name:
        continue
    A = torch.randn(M, R, device='cuda'); W = torch.randn(N, R, device='cuda') / R ** 0.5 if nt else torch.randn(R, N, device='cuda') / R ** 0.5
    b = torch.randn(N, device='cuda'); Y = torch.zeros(M, N, device='cuda')
    g = L.SardGemm()
    g.A, g.lda, g.B0 = L.ptr(A), R, L.ptr(W)
    if nt:
        g.ldb0, g.b_split, g.bias, g.act = R, R, L.ptr(b), 2
    else:
        g.ldb0, g.b_split, g.dsplit = N, N, N
    g.Y, g.ldy, g.M, g.N, g.R = L.ptr(Y), N, M, N, R
    fn = L.lib.sard_linear_fwd if nt else L.lib.sard_linear_bwd_data
    ref = (torch.tanh(A.double() @ W.double().t() + b.double()) if nt else A.double() @ W.double())
    for eng in [int(e) for e in os.environ.get('ENGINES', '0,1,2').split(',')]:
        L.lib.sard_set_gemm_engine(eng)
        for _ in range(3):
            L.check(fn(C.byref(g), L.stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            L.check(fn(C.byref(g), L.stream()))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        err = float((Y.double() - ref).abs().max())
        print(f'{name:16s} M={M} N={N} R={R} engine={eng}: {us:7.1f} us  {2*M*N*R/us/1e6:7.2f} TFLOP/s  max|err|={err:.2e}')

# ---- weight gradient (ungrouped) ----
for name, M, N1, N2 in [('gconv wgrad', 18432, 64, 128), ('il1 wgrad', 18432, 128, 256), ('in_fc wgrad', 18432, 64, 82)]:
    if only and only not in name:
        continue
    dZ = torch.randn(M, N1, device='cuda'); X = torch.randn(M, N2, device='cuda')
    dW = torch.zeros(N1, N2, device='cuda'); db = torch.zeros(N1, device='cuda')
    rows = L.lib.sard_wgrad_chunk_rows(M, N1, N2)
    part = torch.zeros(((M + rows - 1) // rows) * N1 * (N2 + 4), device='cuda')
    w = L.SardWGrad()
    w.P, w.ldp, w.Q, w.ldq, w.M, w.N1, w.N2, w.chunk_rows = L.ptr(dZ), N1, L.ptr(X), N2, M, N1, N2, 0
    w.dW0, w.ldw0, w.w_split, w.db, w.partials = L.ptr(dW), N2, N2, L.ptr(db), L.ptr(part)
    ref = dZ.double().t() @ X.double()
    for eng in [int(e) for e in os.environ.get('ENGINES', '0,1,2').split(',')]:
        L.lib.sard_set_gemm_engine(eng)
        for _ in range(3):
            L.check(L.lib.sard_linear_bwd_weight(C.byref(w), L.stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dW.zero_()
        e0.record()
        for _ in range(iters):
            L.check(L.lib.sard_linear_bwd_weight(C.byref(w), L.stream()))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        err = float((dW.double() / iters - ref).abs().max())
        print(f'{name:16s} M={M} N1={N1} N2={N2} engine={eng}: {us:7.1f} us  {2*M*N1*N2/us/1e6:7.2f} TFLOP/s  max|err|={err:.2e}')

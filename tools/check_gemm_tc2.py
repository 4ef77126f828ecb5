"""2-CTA (cta_group::2) tcgen05 TF32 GEMM vs fp64 reference and vs the 1-CTA kernel (time).  Usage: python tools/check_gemm_tc2.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib                      # noqa: E402
from ast_b200._lib import ptr                  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
rng = np.random.default_rng(0)
bad = 0
for (tb, M, N, K) in [(1, 256, 256, 32), (1, 256, 256, 64), (1, 512, 512, 256), (1, 300, 200, 120), (1, 1000, 1024, 1536), (1, 5120, 1024, 1536),
                      (0, 256, 256, 64), (0, 130, 260, 72), (0, 4000, 1536, 1024), (0, 1500, 1152, 512)]:
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    want = A.astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
    dA, dB, db = (torch.as_tensor(x, device=dev) for x in (A, B, bias))
    dC = torch.full((M, N), 7.0, device=dev)
    rc = lib.ast_gemm(-2, 0, tb, M, N, K, 1.0, ptr(dA), K, ptr(dB), B.shape[1], 0.0, ptr(dC), N, ptr(db), st())
    torch.cuda.synchronize()
    if rc != 0:
        print(f"[FAIL] tb={tb} {M}x{N}x{K}: rc={rc} {lib.ast_last_error().decode()}"); bad += 1; continue
    got = dC.cpu().numpy()
    err = np.abs(got - want).max() / np.sqrt(K)
    ok = err < 6e-3
    bad += (not ok)
    print(f"[{'ok' if ok else 'FAIL'}] 2-CTA gemm tb={tb} {M}x{N}x{K}: max err / sqrt(K) = {err:.2e}", flush=True)
    if not ok:
        d = np.abs(got - want)
        print("     row-block errs", [float(d[r:r + 32].max()) for r in range(0, min(M, 512), 32)])
        print("     col-block errs", [float(d[:, c:c + 32].max()) for c in range(0, min(N, 512), 32)])
# M-major A (weight-gradient form) and split-K
for (tb, M, N, K, which) in [(0, 256, 256, 64, -2), (0, 512, 256, 1000, -2), (1, 256, 128, 64, -2), (0, 1024, 1536, 4000, -3), (0, 1024, 256, 5120, -3),
                             (0, 1100, 520, 2000, -3), (1, 512, 512, 3000, -3)]:
    A = rng.standard_normal((K, M)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    want = A.T.astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    dA, dB = (torch.as_tensor(x, device=dev) for x in (A, B))
    dC = torch.full((M, N), 7.0, device=dev)
    rc = lib.ast_gemm(which, 1, tb, M, N, K, 1.0, ptr(dA), M, ptr(dB), B.shape[1], 0.0, ptr(dC), N, None, st())
    torch.cuda.synchronize()
    if rc != 0:
        print(f"[FAIL] ta=1 tb={tb} {M}x{N}x{K}: rc={rc} {lib.ast_last_error().decode()}"); bad += 1; continue
    err = np.abs(dC.cpu().numpy() - want).max() / np.sqrt(K)
    ok = err < 6e-3
    bad += (not ok)
    print(f"[{'ok' if ok else 'FAIL'}] 2-CTA gemm ta=1 tb={tb} {M}x{N}x{K} {'split-K' if which == -3 else ''}: max err / sqrt(K) = {err:.2e}", flush=True)
# grouped: n same-shape problems in one launch
for (n, ta, tb, M, N, K, sk) in [(4, 1, 0, 1024, 256, 5120, -1), (2, 1, 0, 1024, 256, 3000, -1), (3, 0, 1, 512, 256, 640, 0), (4, 1, 0, 512, 256, 2048, 3)]:
    A = rng.standard_normal((n, K, M) if ta else (n, M, K)).astype(np.float32)
    B = rng.standard_normal((n, N, K) if tb else (n, K, N)).astype(np.float32)
    want = np.stack([(A[g].T if ta else A[g]).astype(np.float64) @ (B[g].T if tb else B[g]).astype(np.float64) for g in range(n)])
    dA, dB = torch.as_tensor(A, device=dev), torch.as_tensor(B, device=dev)
    dC = torch.full((n, M, N), 7.0, device=dev)
    rc = lib.ast_gemm_grouped(n, ta, tb, M, N, K, ptr(dA), A[0].size, A.shape[2], ptr(dB), B[0].size, B.shape[2], ptr(dC), M * N, N, sk, st())
    torch.cuda.synchronize()
    if rc != 0:
        print(f"[FAIL] grouped n={n}: rc={rc} {lib.ast_last_error().decode()}"); bad += 1; continue
    err = np.abs(dC.cpu().numpy() - want).max() / np.sqrt(K)
    ok = err < 6e-3
    bad += (not ok)
    print(f"[{'ok' if ok else 'FAIL'}] grouped 2-CTA gemm n={n} ta={ta} tb={tb} {M}x{N}x{K} split_k={sk}: max err / sqrt(K) = {err:.2e}", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
A4 = torch.randn(4, 5120, 1024, device=dev); B4 = torch.randn(4, 5120, 256, device=dev); C4 = torch.zeros(4, 1024, 256, device=dev)
for n in (4, 2):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.ast_gemm_grouped(n, 1, 0, 1024, 256, 5120, ptr(A4), 5120 * 1024, 1024, ptr(B4), 5120 * 256, 256, ptr(C4), 1024 * 256, 256, -1, st())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts[3:]))
    print(f"  time grouped TN split-K {n} x 1024x256x5120: {ms*1e3:.1f} us  {n*2.0*1024*256*5120/ms/1e9:.1f} TFLOP/s ({ms*1e3/n:.1f} us per problem)", flush=True)
for (M, N, K) in [(1024, 1536, 5120), (1024, 256, 5120), (512, 1152, 15744)]:
    A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev); Cc = torch.zeros(M, N, device=dev)
    for w in (-3, 2):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.ast_gemm(w, 1, 0, M, N, K, 1.0, ptr(A), M, ptr(B), N, 0.0, ptr(Cc), N, None, st())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts[3:]))
        print(f"  time TN split-K {M}x{N}x{K} {'2-CTA' if w == -3 else '1-CTA'}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
for (tb, M, N, K) in [(1, 5120, 1024, 1536), (0, 5120, 1536, 1024), (0, 15744, 1152, 512), (1, 8192, 4096, 4096)]:
    A = torch.randn(M, K, device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev)
    Cc = torch.zeros(M, N, device=dev)
    for w in (-2, 1):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.ast_gemm(w, 0, tb, M, N, K, 1.0, ptr(A), K, ptr(B), B.shape[1], 0.0, ptr(Cc), N, None, st())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts[3:]))
        print(f"  time tb={tb} {M}x{N}x{K} {'2-CTA' if w == -2 else '1-CTA'}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
print("FAILED" if bad else "ALL OK", bad)

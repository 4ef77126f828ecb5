"""tcgen05 TF32 GEMM vs fp64 reference (TF32 tolerance) for every operand-major combination, tails, split-K."""
import ctypes as C
import os
import sys
import time

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
cases = [(0, 1, 256, 256, 64, 1), (0, 1, 128, 128, 32, 1), (0, 1, 300, 200, 120, 1), (0, 1, 4000, 1024, 1536, 1),
         (0, 0, 256, 256, 64, 1), (0, 0, 130, 260, 72, 1), (0, 0, 4000, 1536, 1024, 1),
         (1, 0, 256, 256, 64, 1), (1, 0, 116, 96, 1000, 1), (1, 0, 1024, 1536, 4000, 2), (1, 0, 1024, 256, 4000, 2), (1, 0, 1104, 512, 480, 2),
         (1, 1, 256, 128, 64, 1), (0, 1, 1000, 512, 1152, 1)]
for (ta, tb, M, N, K, which) in cases:
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    want = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
    dA, dB, db = (torch.as_tensor(x, device=dev) for x in (A, B, bias))
    dC = torch.full((M, N), 7.0, device=dev)
    rc = lib.ast_gemm(which, ta, tb, M, N, K, 1.0, ptr(dA), A.shape[1], ptr(dB), B.shape[1], 0.0, ptr(dC), N, ptr(db), st())
    torch.cuda.synchronize()
    if rc != 0:
        print(f"[FAIL] ta={ta} tb={tb} {M}x{N}x{K} which={which}: rc={rc} {lib.ast_last_error().decode()}"); bad += 1; continue
    got = dC.cpu().numpy()
    err = np.abs(got - want).max() / (np.sqrt(K) * 1.0)
    ok = err < 6e-3
    bad += (not ok)
    print(f"[{'ok' if ok else 'FAIL'}] tc gemm ta={ta} tb={tb} {M}x{N}x{K} which={which}: max err / sqrt(K) = {err:.2e}", flush=True)
    if not ok:
        d = np.abs(got - want)
        i, j = np.unravel_index(d.argmax(), d.shape)
        print("     worst at", i, j, got[i, j], want[i, j], "row-block errs", [float(d[r:r + 32].max()) for r in range(0, min(M, 256), 32)],
              "col-block errs", [float(d[:, c:c + 32].max()) for c in range(0, min(N, 256), 32)])
# overlapping-rows operand (CNN_1 implicit GEMM): A rows start every 256 floats, K = 1152
M, N, K, lda = 500, 512, 1152, 256
buf = rng.standard_normal(M * lda + K).astype(np.float32)
W = rng.standard_normal((N, K)).astype(np.float32)
Av = np.lib.stride_tricks.as_strided(buf, (M, K), (lda * 4, 4))
want = Av.astype(np.float64) @ W.T.astype(np.float64)
dbuf, dW, dC = torch.as_tensor(buf, device=dev), torch.as_tensor(W, device=dev), torch.zeros(M, N, device=dev)
rc = lib.ast_gemm(1, 0, 1, M, N, K, 1.0, ptr(dbuf), lda, ptr(dW), K, 0.0, ptr(dC), N, None, st())
torch.cuda.synchronize()
if rc != 0:
    print("[info] overlapping-rows tensor map rejected:", lib.ast_last_error().decode())
else:
    err = np.abs(dC.cpu().numpy() - want).max() / np.sqrt(K)
    print(f"[{'ok' if err < 6e-3 else 'FAIL'}] tc gemm overlapping rows (implicit conv): err {err:.2e}"); bad += err >= 6e-3
# timing
for (ta, tb, M, N, K, which) in [(0, 1, 5120, 1024, 1536, 1), (0, 0, 5120, 1536, 1024, 1), (1, 0, 1024, 1536, 5120, 2), (1, 0, 1024, 256, 5120, 2),
                                 (0, 1, 15744, 512, 1152, 1)]:
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev)
    Cc = torch.zeros(M, N, device=dev)
    for w in (which, 0):
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.ast_gemm(w, ta, tb, M, N, K, 1.0, ptr(A), A.shape[1], ptr(B), B.shape[1], 0.0, ptr(Cc), N, None, st())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = np.mean(ts[2:])
        print(f"  time ta={ta} tb={tb} {M}x{N}x{K} which={w}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
print("FAILED" if bad else "ALL OK", bad)

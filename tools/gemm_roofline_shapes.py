"""Launch the bench.py roofline kernels a few times each (for ncu captures): the grouped 2-CTA layer-0 projection and the 1-CTA CNN_1
data-gradient (transposed-convolution, even rows) GEMM.  Usage: python tools/gemm_roofline_shapes.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib                      # noqa: E402
from ast_b200._lib import check, ptr           # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
G, M, N, K = 2, 5120, 1024, 1536
A = torch.randn(G, M, K, device=dev); W = torch.randn(G, N, K, device=dev); Cc = torch.empty(G, M, N, device=dev)
for it in range(5):
    flush.zero_()
    check(lib.ast_gemm_grouped(G, 0, 1, M, N, K, ptr(A), M * K, K, ptr(W), N * K, K, ptr(Cc), M * N, N, 0, st), "grouped")
M2, N2, K2 = 15744, 128, 2560
A2 = torch.randn(M2 + 8, 512, device=dev); W2 = torch.randn(K2, N2, device=dev); C2 = torch.empty(M2, 2 * N2, device=dev)
for it in range(5):
    flush.zero_()
    check(lib.ast_gemm(1, 0, 0, M2, N2, K2, 1.0, ptr(A2), 512, ptr(W2), N2, 0.0, ptr(C2), 2 * N2, None, st), "conv dx")   # overlapping rows: lda 512 < K
torch.cuda.synchronize()
print("ok")

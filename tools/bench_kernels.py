"""Micro-benchmarks of the latency-bound kernels (CUDA events, warm): persistent LSTM recurrence per step."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib
from ast_b200._lib import ptr
dev = torch.device("cuda", 0); lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def timeit(fn, n=5):
    ts = []
    for i in range(n + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
for (T, B, h) in [(250, 16, 256), (250, 32, 256), (160, 32, 256), (750, 1, 256)]:
    G = torch.randn(T, B, 4 * h, device=dev) * 0.5; W = torch.randn(4 * h, h, device=dev) / 16
    Hs = torch.zeros(T + 1, B, h, device=dev); Cs = torch.zeros(T + 1, B, h, device=dev); out = torch.zeros(T, B, h, device=dev)
    dh = torch.randn(B, h, device=dev); dc = torch.randn(B, h, device=dev)
    for exact in (1, 0):
        G0 = G.clone()
        f = timeit(lambda: (G0.copy_(G), lib.ast_lstm_seq(0, ptr(G0), ptr(W), ptr(Hs), ptr(Cs), ptr(out), T, B, h, None, None, exact, st())))
        cp = timeit(lambda: G0.copy_(G))
        b = timeit(lambda: lib.ast_lstm_seq(1, ptr(G0), ptr(W), ptr(Hs), ptr(Cs), ptr(out), T, B, h, ptr(dh), ptr(dc), exact, st()))
        print(f"lstm_seq T={T} B={B} h={h} exact={exact}: fwd {1e3*(f-cp)/T:.2f} us/step ({f-cp:.3f} ms)  bwd {1e3*b/T:.2f} us/step ({b:.3f} ms)", flush=True)

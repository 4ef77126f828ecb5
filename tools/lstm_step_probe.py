"""Cycle breakdown of one step of the tcgen05 LSTM forward recurrence (lstm_seq_tc.cu): clock64() stamps taken by CTA 0
(issuer thread: h landed, MMAs issued+committed; epilogue thread 0: accumulator ready, gates exchanged, cell done, h sent,
bookkeeping done) for steps 8..23 of a 64-step launch.  Usage: python tools/lstm_step_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib                     # noqa: E402
from ast_b200._lib import check, ptr          # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
T, B, h = 64, 16, 256
rng = np.random.default_rng(0)
G = torch.as_tensor(rng.standard_normal((T, B, 4 * h)).astype(np.float32), device=dev)
W = torch.as_tensor((rng.standard_normal((4 * h, h)) / 16).astype(np.float32), device=dev)
Hs = torch.zeros(T + 1, B, h, device=dev); Cs = torch.zeros(T + 1, B, h, device=dev); out = torch.zeros(T, B, h, device=dev)
prof = torch.zeros(512, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for it in range(3):
    Gc = G.clone()
    check(lib.ast_lstm_probe(ptr(prof)))
    check(lib.ast_lstm_seq(0, ptr(Gc), ptr(W), ptr(Hs), ptr(Cs), ptr(out), T, B, h, None, None, 0, st))
    torch.cuda.synchronize()
check(lib.ast_lstm_probe(None))
p = prof.cpu().numpy()[:128].reshape(16, 8)[:, :7]          # 32-bit clock stamps
names = ["h landed (issuer)", "MMAs issued + commit", "accumulator ready (epilogue)", "gates exchanged (tcgen05.ld, smem, bar)",
         "cell math done", "h sent (st.async)", "bookkeeping done"]
step = ((p[1:, 0] - p[:-1, 0]) & 0xFFFFFFFF).astype(np.float64)
print(f"step period: median {np.median(step):.0f} cycles = {np.median(step) / 1.965e3:.2f} us @1.965 GHz")
rel = ((p - p[:, :1]) & 0xFFFFFFFF).astype(np.float64)
for k in range(7):
    print(f"  {names[k]:42s} +{np.median(rel[:, k]):7.0f} cycles after 'h landed'")
print("  next 'h landed' (other CTAs' sends + mbarrier)   +%7.0f" % np.median(step))

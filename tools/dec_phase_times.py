"""Per-phase timing of the persistent decoder-sequence kernels (option dec_prof): CTA 0 stores %globaltimer after every
grid barrier; this prints the median nanoseconds of each phase of a decoder step, forward and backward.
Usage: python tools/dec_phase_times.py [--B 32 --T 640 --L 24]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg          # noqa: E402
from ast_b200.seq2seq import SpeechEncoderDecoder        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--T", type=int, default=640)
ap.add_argument("--L", type=int, default=24)
ap.add_argument("--ratio", type=float, default=1.0, help="teacher-forcing ratio (1.0: no sampled steps, uniform phases)")
ap.add_argument("--fine", type=int, default=-1, help="CTA whose thread 0 also stamps clock64 inside the phases of step 6 (forward)")
ap.add_argument("--sync", action="store_true", help="grid barrier after every phase (option dec_sync)")
ap.add_argument("--v1", action="store_true", help="first-generation decoder forward kernel")
ap.add_argument("--cg", action="store_true", help="use cooperative_groups grid.sync instead of the counter barrier")
a = ap.parse_args()
rng = np.random.default_rng(0)
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(dropout=(0.3, 0.3, 0.0)), feat_dim=40)
m.init_params(seed=0)
e = m._engine
e.set_option("exact", 0); e.set_option("tc_gemm", 1); e.set_option("dec_prof", 1 if a.fine < 0 else 2 + a.fine)
e.set_option("dec_sync", 1 if a.sync else 0)
e.set_option("dec_fast_barrier", 0 if a.cg else 1)
e.set_option("dec_v2", 0 if a.v1 else 1)
X = torch.as_tensor(rng.standard_normal((a.B, a.T, 40)).astype(np.float32), device=e.device)
y = rng.integers(4, 1098, (a.B, a.L)).astype(np.int32); y[:, 0] = 1; y[:, -1] = 2
y = torch.as_tensor(y, device=e.device)
bits = torch.as_tensor((rng.random(a.L - 1) < a.ratio).astype(np.uint8), device=e.device)
for it in range(3):
    e.forward_loss(X, y, use_true=bits, noise_sigma=0.25)
    e.backward()
torch.cuda.synchronize()
raw = e.debug_fetch("dec_prof").view(torch.int64).cpu().numpy().reshape(2, 4096)
S = a.L - 1
for name, row in (("forward", raw[0]), ("backward", raw[1])):
    n = int(row[0]); ts = row[1:n + 1].astype(np.float64)
    d = np.diff(ts)
    if row[4090] > 0 and not a.v1:
        print(f"{name}: kernel entry -> end of setup (TMEM/smem weight fill, prologue) {1e-3 * (row[1] - row[4090]):.1f} us")
    print(f"{name}: {n} stamps, total {1e-3 * (ts[-1] - ts[0]):.1f} us, {1e-3 * (ts[-1] - ts[0]) / S:.2f} us/step")
    per = (n - 1) // S
    head = (n - 1) - per * S
    ph = d[head:].reshape(S, per)
    print("  phases/step:", per, " leading stamps:", head, d[:head])
    print("  median ns per phase:", np.median(ph, axis=0).round(0).tolist())
    print("  min    ns per phase:", ph.min(axis=0).round(0).tolist())

if a.fine >= 0:
    LBL = (["L%d: start->step-old operands staged", "L%d: MMA on them", "L%d: poll + stage the fresh operand", "L%d: MMA on it", "L%d: exchange (send, wait)", "L%d: cell epilogue + stores"] * 3,
           ["attn: prefetch", "attn: poll h2", "attn: online-softmax pass", "attn: warp merge + send", "attn: exchange wait", "attn: cv / alpha stores"],
           ["ht: h2 part staged + MMA", "ht: poll cv", "ht: stage + MMA", "ht: exchange", "ht: tanh + stores"])
    for name, row in (("forward", raw[0]), ("backward", raw[1])):
        f = row[3000:3064].astype(np.float64); f = f[f > 0]
        if len(f) < 2:
            continue
        print(f"{name}: clock64 stamps of CTA {a.fine}, step 6 ({len(f)} stamps), ns at 1.965 GHz between consecutive stamps:")
        print("  ", (np.diff(f) / 1.965).round(0).astype(int).tolist())

"""One warm-up + N training steps on a fixed synthetic batch, for ncu (launch list / --set full captures).
Usage: python tools/profile_step.py [--B 32 --T 640 --L 24 --steps 1 --precision f32|tf32]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg          # noqa: E402
from ast_b200.seq2seq import SpeechEncoderDecoder        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--T", type=int, default=640)
ap.add_argument("--L", type=int, default=24)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--precision", default="f32")
ap.add_argument("--dropout", type=float, default=0.3)
ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable), e.g. --opt enc_chunk=16")
a = ap.parse_args()

rng = np.random.default_rng(0)
cfg = es_en_20h_model_cfg(dropout=(a.dropout, a.dropout, 0.0))
m = SpeechEncoderDecoder(0, cfg, feat_dim=40)
m.init_params(seed=0)
e = m._engine
e.set_option("exact", 0 if a.precision == "tf32" else 1)
e.set_option("tc_gemm", 1 if a.precision == "tf32" else 0)
for kv in a.opt:
    k, v = kv.split("=")
    e.set_option(k, float(v))
X = torch.as_tensor(rng.standard_normal((a.B, a.T, 40)).astype(np.float32), device=e.device)
y = rng.integers(4, 1098, (a.B, a.L)).astype(np.int32); y[:, 0] = 1; y[:, -1] = 2
y = torch.as_tensor(y, device=e.device)
bits = torch.as_tensor((rng.random(a.L - 1) < 0.8).astype(np.uint8), device=e.device)
mm, v, vh = (torch.zeros_like(e.params) for _ in range(3))
for it in range(a.warmup + a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = e.forward_loss(X, y, use_true=bits, noise_sigma=0.25)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    e.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    e.opt_step(mm, v, vh, it + 1, 1e-3, 1e-4, 2.0)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if "stage_timing=1" in a.opt:
        print("   stages:", ", ".join(f"{n} {ms:.3f}" for n, ms in e.stage_times()), flush=True)
    print(f"step {it}: loss {float(loss):.4f} fwd {1e3*(t1-t0):.2f} ms bwd {1e3*(t2-t1):.2f} ms opt {1e3*(t3-t2):.2f} ms", flush=True)

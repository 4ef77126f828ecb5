"""Timeline of the persistent encoder wavefront: %globaltimer stamps written by block 0 of every layer's recurrence kernel at each
chunk boundary (forward and backward).  Usage: python tools/enc_timeline.py [--T 640 --pchunk 8 --opt k=v]"""
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
ap.add_argument("--pchunk", type=int, default=8)      # the engine default is 4; 8 keeps the periods comparable with earlier profiles
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
rng = np.random.default_rng(0)
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(dropout=(0.3, 0.3, 0.0)), feat_dim=40)
m.init_params(seed=0)
e = m._engine
e.set_option("exact", 0); e.set_option("tc_gemm", 1); e.set_option("enc_persist", 3); e.set_option("enc_pchunk", a.pchunk); e.set_option("enc_ts", 1)
for kv in a.opt:
    k, v = kv.split("=")
    e.set_option(k, float(v))
X = torch.as_tensor(rng.standard_normal((a.B, a.T, 40)).astype(np.float32), device=e.device)
y = rng.integers(4, 1098, (a.B, a.L)).astype(np.int32); y[:, 0] = 1; y[:, -1] = 2
y = torch.as_tensor(y, device=e.device)
bits = torch.as_tensor((rng.random(a.L - 1) < 0.8).astype(np.uint8), device=e.device)
for it in range(3):
    float(e.forward_loss(X, y, use_true=bits, noise_sigma=0.25)); e.backward(); torch.cuda.synchronize()
ts = e.debug_fetch("enc_ts").cpu().numpy().view(np.uint64).reshape(2, 4, 256)
Tp = e.enc_len(a.T); nq = (Tp + a.pchunk - 1) // a.pchunk
for p, name in enumerate(("forward", "backward")):
    t = ts[p, :3, :nq].astype(np.int64)
    t0 = t.min()
    print(f"{name}: T'={Tp}, {nq} chunks of {a.pchunk}; microseconds since the first stamp of the pass")
    for l in range(3):
        us = (t[l] - t0) / 1e3
        d = np.diff(us)
        print(f"  layer {l}: first {us[0]:7.1f} last {us[-1]:7.1f}  chunk period median {np.median(d):6.2f} (= {np.median(d) / a.pchunk:.2f} us/step) max {d.max():6.1f}")
    order = (2, 1, 0) if p else (0, 1, 2)
    for x, yv in zip(order[:-1], order[1:]):
        lag = (t[yv] - t[x]) / 1e3
        print(f"  lag layer {x} -> {yv}: median {np.median(lag):6.1f} us, first {lag[0]:6.1f}, last {lag[-1]:6.1f}")

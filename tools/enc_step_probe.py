"""Cycle breakdown of one step of the tcgen05 encoder recurrences INSIDE a real training step (persistent wavefront running: three
layers' clusters, gated GEMMs and the side stream beside them): clock64 stamps of CTA 0 of the (layer 0, forward direction) launch
for its 9th..24th step, forward and backward kernel.  Usage: python tools/enc_step_probe.py [--T 640]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib
from ast_b200._lib import check, ptr
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder
ap = argparse.ArgumentParser(); ap.add_argument("--T", type=int, default=640); ap.add_argument("--L", type=int, default=24)
a = ap.parse_args()
rng = np.random.default_rng(0)
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(dropout=(0.3, 0.3, 0.0)), feat_dim=40); m.init_params(seed=0)
e = m._engine; lib = _lib.load()
e.set_option("exact", 0); e.set_option("tc_gemm", 1)
X = torch.as_tensor(rng.standard_normal((32, a.T, 40)).astype(np.float32), device=e.device)
y = rng.integers(4, 1098, (32, a.L)).astype(np.int32); y[:, 0] = 1; y[:, -1] = 2
y = torch.as_tensor(y, device=e.device)
prof = torch.zeros(512, dtype=torch.int64, device=e.device)
for it in range(4):
    if it == 3:
        check(lib.ast_lstm_probe(ptr(prof)))
    e.forward_loss(X, y, noise_sigma=0.25); e.backward()
torch.cuda.synchronize()
check(lib.ast_lstm_probe(None))
print("persistent wavefront active:", bool(e.get_option("enc_persist_active") == 1))
p = prof.cpu().numpy().astype(np.int64)
rel = lambda x, x0: ((x - x0) & 0xFFFFFFFF).astype(np.float64)          # 32-bit clock stamps
f = p[:128].reshape(16, 8)[:, :7]
names = ["h landed (issuer)", "MMAs issued + commit", "accumulator ready (epilogue)", "gates exchanged", "cell math done", "h sent", "bookkeeping done"]
st = rel(f[1:, 0], f[:-1, 0])
print(f"forward step period: median {np.median(st):.0f} cycles = {np.median(st) / 1.965e3:.2f} us; per step:", " ".join(f"{x:.0f}" for x in st))
for k in range(7):
    print(f"  {names[k]:34s} +{np.median(rel(f[:, k], f[:, 0])):7.0f}")
b = p[128:384].reshape(16, 16); iss = p[384:416].reshape(16, 2)
names = ["partial dh of all CTAs landed", "dG computed, operand in smem", "proxy fence + arrive done", "dG stored / bookkeeping done",
         "MMAs retired (epilogue sees commit)", "accumulator in registers", "partial dh staged (fence + syncwarp)", "bulk send issued"]
st = rel(b[1:, 0], b[:-1, 0])
print(f"backward step period: median {np.median(st):.0f} cycles = {np.median(st) / 1.965e3:.2f} us; per step:", " ".join(f"{x:.0f}" for x in st))
print(f"  {'':40s} {'warp 0':>8s} {'warp 7':>8s}")
for k in range(8):
    print(f"  {names[k]:40s} +{np.median(rel(b[:, k], b[:, 0])):7.0f} +{np.median(rel(b[:, 8 + k], b[:, 0])):7.0f}")
print(f"  {'operand ready (issuer)':40s} +{np.median(rel(iss[:, 0], b[:, 0])):7.0f}")
print(f"  {'MMAs issued + commit (issuer)':40s} +{np.median(rel(iss[:, 1], b[:, 0])):7.0f}")

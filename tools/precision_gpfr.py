"""Which contraction sets the worst gradient error of the TF32 training mode on the asr_gpfr-shaped case (B9 x T420 x D13,
V59)?  Prints the worst per-tensor max-rel errors against the fp64 oracle with one GEMM call-site class at a time moved back
to the fp32 SIMT kernel (tc_mask), and with the decoder / recurrences in exact mode."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.engine import Engine
from oracle import ast_oracle as O
cfg = O.default_model_cfg(vocab=59)
P = O.init_params(cfg, 13, seed=21)
X, y, _ = O.synth_batch(9, 420, 13, 59, 10, 30, seed=22, Tmin=300)
om = O.OracleModel(cfg, P, dtype=np.float64)
loss = float(om.forward_loss(X, y)); g = om.backward()
e = Engine(cfg, 13, 0)
for k in e.info: e.view(k).copy_(torch.as_tensor(P[k], device=e.device))
e.weights_changed()
def run(tag):
    got = float(e.forward_loss(X, y)); e.backward()
    errs = {k: np.abs(e.view(k, grad=True).cpu().numpy().astype(np.float64) - g[k]).max() / (np.abs(g[k]).max() + 1e-30) for k in e.info}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"{tag:34s} loss err {abs(got-loss)/abs(loss):.1e}; worst:", ", ".join(f"{k} {v:.1e}" for k, v in top), flush=True)
e.set_option("exact", 1); run("exact")
e.set_option("exact", 0); e.set_option("tc_gemm", 0); run("tf32 recurrences, SIMT GEMMs")
e.set_option("tc_gemm", 1); run("full TF32 training mode")
names = ["conv0", "conv1", "enc_proj", "dec_wgrad", "enc_dx", "enc_wgrad", "conv1_wgrad", "conv1_dx", "conv0_wgrad", "dec_pre"]
for site in range(2, 10):
    e.set_option("tc_mask", 0x3 | (1 << site)); run(f"{names[site]} back on fp32 SIMT")
e.set_option("tc_mask", 0x3)
e.set_option("dec_v2", 0); run("first-generation decoder kernels"); e.set_option("dec_v2", 1)
e.set_option("conv3x", 0); run("conv1 fwd on SIMT (no 3xTF32)"); e.set_option("conv3x", 1)

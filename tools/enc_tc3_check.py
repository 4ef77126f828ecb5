"""Decode-time encoder: exact (fp32 SIMT) vs 3xTF32 tensor-core projections / convolutions, both against the fp64 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ast_oracle as O
from ast_b200.engine import Engine
cfg = O.default_model_cfg(vocab=300); D = 40
P = O.init_params(cfg, D, seed=61)
rng = np.random.default_rng(62)
X = rng.standard_normal((4, 230, D)).astype(np.float32)
om = O.OracleModel(cfg, P, dtype=np.float64)
om.train = False
om.encode(X)
want = np.asarray(om.enc_states, dtype=np.float64) if hasattr(om, "enc_states") else None
e = Engine(cfg, D, 0)
for k in e.info: e.view(k).copy_(torch.as_tensor(P[k], device=e.device))
e.weights_changed()
outs = {}
for mode in (0, 1):
    e.set_option("enc_tc3", mode)
    e.encode(X, train=False)
    outs[mode] = e.enc_states().cpu().numpy().astype(np.float64)
d = np.abs(outs[0] - outs[1]).max()
print(f"max |exact - tc3| = {d:.3e}, max |enc| = {np.abs(outs[0]).max():.3f}")
if want is not None:
    for mode in (0, 1):
        print(f"mode {mode}: max |enc - oracle fp64| = {np.abs(outs[mode] - want).max():.3e}")

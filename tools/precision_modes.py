"""Gradient / loss error of each precision mode against the float64 oracle at C1 size."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.engine import Engine
from oracle import ast_oracle as O
cfg = O.default_model_cfg(vocab=1098)
P = O.init_params(cfg, 40, seed=0)
X, y, lens = O.synth_batch(16, 1000, 40, 1098, 20, 40, seed=1, Tmin=921)
om = O.OracleModel(cfg, P, dtype=np.float64)
loss = float(om.forward_loss(X, y)); g = om.backward()
e = Engine(cfg, 40, 0)
for k in e.info: e.view(k).copy_(torch.as_tensor(P[k], device=e.device))
e.weights_changed()
for exact, tc in ((1, 0), (0, 0), (0, 1)):
    e.set_option("exact", exact); e.set_option("tc_gemm", tc)
    got = float(e.forward_loss(X, y)); e.backward()
    errs = {}
    for k in e.info:
        a = e.view(k, grad=True).cpu().numpy().astype(np.float64)
        errs[k] = (np.abs(a - g[k]).max() / (np.abs(g[k]).max() + 1e-30), np.linalg.norm(a - g[k]) / (np.linalg.norm(g[k]) + 1e-30))
    top = sorted(errs.items(), key=lambda kv: -kv[1][0])[:6]
    print(f"exact={exact} tc_gemm={tc}: loss rel err {abs(got-loss)/abs(loss):.2e}; worst max-rel grads:", ", ".join(f"{k} {v[0]:.1e} (l2 {v[1]:.1e})" for k, v in top), flush=True)
names = ["conv0", "conv1", "enc_proj", "dec_wgrad", "enc_dx", "enc_wgrad", "conv1_wgrad", "conv1_dx", "conv0_wgrad"]
e.set_option("exact", 0); e.set_option("tc_gemm", 1)
for site in range(9):
    e.set_option("tc_mask", 0x1FF & ~(1 << site))      # ONLY this site on tensor cores
    got = float(e.forward_loss(X, y)); e.backward()
    errs = {k: np.abs(e.view(k, grad=True).cpu().numpy().astype(np.float64) - g[k]).max() / (np.abs(g[k]).max() + 1e-30) for k in e.info}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"only {names[site]:12s} on tcgen05: loss err {abs(got-loss)/abs(loss):.1e}; worst:", ", ".join(f"{k} {v:.1e}" for k, v in top), flush=True)

"""Rebuilds the inputs of a golden case (tests/golden/*.npz) — parameters are a pure function of the seed."""
import os

import numpy as np

from oracle import ast_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"tiny_d13_tf": 23, "tiny_d40_ss": 29}


def tiny_cfg(V):
    return O.default_model_cfg(vocab=V, hidden=128, embed=16, attn=128, layers=3,
                               cnn=((8, (9, 13), (2, 13), (4, 0)), (16, (9, 1), (2, 1), (4, 0))))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = tiny_cfg(CASES[name])
    D, seed = int(z["D"]), int(z["seed"])
    P = O.init_params(cfg, D, seed=seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 7)
    for k in P:
        if k.endswith(("gamma", "beta", "/b")):
            P[k] = P[k] + 0.1 * rng.standard_normal(P[k].shape)
    chk = sum(float(np.abs(v).sum()) for k, v in sorted(P.items()) if v.dtype.kind == "f")
    assert abs(chk - float(z["param_checksum"])) <= 1e-9 * abs(chk), "numpy RNG stream changed: regenerate the goldens"
    return cfg, D, P, z


def decode_params(P, z, eos_boost):
    """float32 parameters for decoding; the beam golden raises the EOS bias so hypotheses finish."""
    P32 = {k: (v.astype(np.float32) if v.dtype.kind == "f" else v) for k, v in P.items()}
    if eos_boost:
        P32["out/b"] = P32["out/b"].copy()
        P32["out/b"][O.EOS_ID] += float(z["out_b_eos_boost"])
    return P32


def beam_hyps(z):
    out, i = [], 0
    for n in z["beam_hyp_lens"]:
        out.append([int(t) for t in z["beam_hyps"][i:i + n]])
        i += n
    return out

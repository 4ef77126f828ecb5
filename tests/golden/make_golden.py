"""Generates tests/golden/*.npz from the float64 oracle (the reference ships no golden vectors;
these are this repo's pins).  Run once:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ast_oracle as O  # noqa: E402


def tiny_cfg(V=23):
    return O.default_model_cfg(vocab=V, hidden=128, embed=16, attn=128, layers=3,
                               cnn=((8, (9, 13), (2, 13), (4, 0)), (16, (9, 1), (2, 1), (4, 0))))


def case(name, cfg, D, B, T, Lmin, Lmax, seed, bits_ratio=None):
    V = cfg["rnn_config"]["dec_vocab_size"]
    P = O.init_params(cfg, D, seed=seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 7)
    for k in P:
        if k.endswith(("gamma", "beta", "/b")):
            P[k] = P[k] + 0.1 * rng.standard_normal(P[k].shape)
    X, y, lens = O.synth_batch(B, T, D, V, Lmin, Lmax, seed=seed + 1, Tmin=max(1, T - 40))
    L = y.shape[1]
    bits = None
    if bits_ratio is not None:
        bits = np.asarray([True if not (0 < i < L - 2) else bool(rng.random() < bits_ratio) for i in range(L - 1)])
    m = O.OracleModel(cfg, P, dtype=np.float64)
    loss = m.forward_loss(X.astype(np.float64), y, tf_bits=None if bits is None else list(bits))
    g = m.backward()
    out = {"X": X, "y": y, "loss": np.float64(loss), "enc_states": m.enc_states, "step_losses": np.asarray(m.step_losses),
           "step_argmax": np.stack(m.step_argmax).astype(np.int32), "bits": np.ones(L - 1, bool) if bits is None else bits,
           "D": np.int64(D)}
    # parameters are re-derived from (cfg, D, seed, the 0.1*N(0,1) bias/BN perturbation) by tests/golden_util.py;
    # gradients are pinned by per-tensor L2 norm and sum, plus every tensor under 4096 elements in full.
    out["seed"] = np.int64(seed)
    out["param_checksum"] = np.float64(sum(float(np.abs(v).sum()) for k, v in sorted(P.items()) if v.dtype.kind == "f"))
    for k, v in g.items():
        out["gnorm:" + k] = np.float64(np.sqrt((v ** 2).sum()))
        out["gsum:" + k] = np.float64(v.sum())
        if v.size <= 4096:
            out["grad:" + k] = v
    # greedy + beam with the float32 oracle (decode parity is defined at fp32)
    P32 = {k: (v.astype(np.float32) if v.dtype.kind == "f" else v) for k, v in P.items()}
    out["greedy"] = O.OracleModel(cfg, P32, dtype=np.float32).predict(X, O.GO_ID, O.EOS_ID, 12)   # no EOS boost: 12 steps
    P32["out/b"] = P32["out/b"].copy(); P32["out/b"][O.EOS_ID] += 1.5                           # beam: EOS reachable
    m32 = O.OracleModel(cfg, P32, dtype=np.float32)
    nb = m32.decode_beam(X[:1, :lens[0]], 12, 4, 3)
    out["beam_len0"] = np.int64(lens[0])
    out["beam_scores"] = np.asarray([float(e["score"]) for e in nb], dtype=np.float32)
    out["beam_hyp_lens"] = np.asarray([len(e["hyp"]) for e in nb])
    out["beam_hyps"] = np.concatenate([np.asarray(e["hyp"]) for e in nb])
    out["out_b_eos_boost"] = np.float32(1.5)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", float(loss), "greedy", out["greedy"].shape, "beam", [len(e["hyp"]) for e in nb])


if __name__ == "__main__":
    case("tiny_d13_tf", tiny_cfg(), 13, 3, 45, 4, 7, 101)
    case("tiny_d40_ss", tiny_cfg(29), 40, 5, 61, 4, 8, 202, bits_ratio=0.5)

"""Readers of tests/golden/ref_*.npz (generated from the reference's own source by oracle/gen_ref_golden.py)."""
import os

import numpy as np

from oracle import ast_oracle as O
from oracle import ref_golden_common as C
from oracle import synth_corpus as SC

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def _params(z, cfg):
    D, seed = int(z["D"]), int(z["seed"])
    P = C.golden_params(cfg, D, seed)
    chk = C.param_checksum(P)
    assert abs(chk - float(z["param_checksum"])) <= 1e-9 * abs(chk), "numpy RNG stream changed: regenerate the goldens"
    return D, P


def load_model_case(name, dropout=(0.0, 0.0, 0.0), eos_boost=False):
    z = load(name)
    cfg = C.model_cfg(dropout)
    D, P = _params(z, cfg)
    if eos_boost:
        P["out/b"] = P["out/b"].copy()
        P["out/b"][O.EOS_ID] += float(z["eos_boost"])
    return cfg, D, P, z


def load_full_case(name="ref_full_es_en_20h", eos_boost=False):
    """The shipped es_en_20h geometry (model_cfg.json as the reference ships it, stored in the fixture) at V=1098, D=40."""
    import json
    z = load(name)
    cfg = json.loads(str(z["model_cfg_json"]))
    cfg["rnn_config"]["dec_vocab_size"] = int(z["V"])
    D, P = _params(z, cfg)
    if eos_boost:
        P["out/b"] = P["out/b"].copy()
        P["out/b"][O.EOS_ID] += float(z["eos_boost"])
    return cfg, D, P, z


def epoch_case_model(z):
    cfg = C.model_cfg()
    D, P = _params(z, cfg)
    return cfg, D, P


def rebuild_epoch_corpus(z, root):
    """The experiment directory the generator handed to the reference's NN(cfg_path), rebuilt from the same arguments,
    plus the seq2seq_0.model checkpoint it resumed from."""
    cfg, D, P = epoch_case_model(z)
    mc = SC.small_model_cfg(hidden=128, embed=16, attn=128, c0=8, c1=16)
    exp = SC.write_experiment(root, mc, feat_dim=D, vocab_words=C.VOCAB_WORDS,
                              **C.epoch_corpus_kwargs(int(z["seed"]), bool(z["globalphone"]), [str(s) for s in z["freeze"]]))
    with open(os.path.join(exp, "seq2seq_0.model"), "wb") as f:
        np.savez_compressed(f, **P)
    return exp


def tensor_errors(z, prefix, tensors):
    """Per tensor: (max-norm relative error, L2 relative error) of `tensors[k]` against the stored float32 values
    (all of them, or every SAMPLE-th element of the flattened tensor)."""
    out = {}
    for key in z.files:
        if not key.startswith(prefix + "/"):
            continue
        k = key[len(prefix) + 1:]
        want = z[key].astype(np.float64)
        got = np.asarray(tensors[k], dtype=np.float64)
        if want.shape != got.shape:
            got = got.ravel()[::int(z["sample"]) if "sample" in z.files else C.SAMPLE]
        assert got.shape == want.shape, (k, got.shape, want.shape)
        out[k] = (float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30)),
                  float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-30)))
    assert out, f"no tensors stored under {prefix}/"
    return out


def assert_tensors_match(z, prefix, tensors, rel):
    """float64 agreement of EVERY element through the stored whole-tensor checksums [sum, l2, <t, probe>] (tolerance `rel`
    of the tensor's l2 norm), plus element-wise agreement with the stored float32 values at float32 resolution."""
    errs = tensor_errors(z, prefix, tensors)
    for k, (emax, el2) in errs.items():
        assert emax <= 3e-7 and el2 <= 3e-7, (prefix, k, emax, el2)
        want = z[f"{prefix}_chk/{k}"]
        got = C.checksum(tensors[k])
        scale = max(abs(want[1]), 1e-300) * np.sqrt(np.asarray(tensors[k]).size)
        assert abs(got[1] - want[1]) <= rel * max(abs(want[1]), 1e-300), (prefix, k, "l2", got[1], want[1])
        assert abs(got[0] - want[0]) <= rel * scale and abs(got[2] - want[2]) <= rel * scale, (prefix, k, got, want)


def beam_from_fixture(z, N, K):
    p = f"beam_N{N}K{K}/"
    hyps, i = [], 0
    for n in z[p + "beam_hyp_lens"]:
        hyps.append([int(t) for t in z[p + "beam_hyps"][i:i + n]])
        i += n
    return hyps, z[p + "beam_scores"], z[p + "beam_attn_last"]

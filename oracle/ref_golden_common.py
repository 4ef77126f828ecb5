"""TEST INFRASTRUCTURE ONLY - what the golden generator (oracle/gen_ref_golden.py, runs where /root/reference exists) and
the tests that read tests/golden/ref_*.npz (run everywhere) must agree on: the toy model geometry, how parameters are derived
from a seed, the float64 checksum direction, and the arguments the synthetic corpora are built from."""
import numpy as np

from oracle import ast_oracle as O

VOCAB_WORDS = 40
V = VOCAB_WORDS + 4
DROP_SEED = 0x5EED0042
SAMPLE = 7          # strided tensors keep every 7th element (+ float64 checksums of the whole tensor)
EOS_BOOST = 1.5


def model_cfg(dropout=(0.0, 0.0, 0.0)):
    """Smallest geometry the CUDA path accepts (H % 128, E/A % 16, channels % 4) with the shipped structure."""
    return O.default_model_cfg(vocab=V, hidden=128, embed=16, attn=128, layers=3,
                               cnn=((8, (9, 13), (2, 13), (4, 0)), (16, (9, 1), (2, 1), (4, 0))), dropout=dropout)


def golden_params(cfg, D, seed):
    """Parameters = pure function of (cfg, D, seed); biases / BN affine are perturbed so no gradient path is trivially 0."""
    P = O.init_params(cfg, D, seed=seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 7)
    for k in sorted(P):
        if k.endswith(("gamma", "beta", "/b")):
            P[k] = P[k] + 0.1 * rng.standard_normal(P[k].shape)
    P["out/W"] = P["out/W"] * 3.0               # spread the logits: greedy / beam hypotheses are not all alike
    return P


def param_checksum(P):
    return sum(float(np.abs(v).sum()) for k, v in sorted(P.items()) if v.dtype.kind == "f")


def probe(n):
    """Fixed pseudo-random direction for a float64 checksum of a tensor stored in float32."""
    return np.cos(0.37 * np.arange(n, dtype=np.float64) + 0.11)


def checksum(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.sqrt((a * a).sum()), float(a.ravel() @ probe(a.size))])


def epoch_corpus_kwargs(seed, globalphone, freeze):
    return dict(seed=seed, globalphone=bool(globalphone), batch_size=4, buckets_num=4, buckets_width=16, max_pred=8,
                teach_ratio=0.8, speech_noise=0.0, zero_input=0.1, freeze=tuple(freeze))

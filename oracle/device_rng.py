"""TEST INFRASTRUCTURE ONLY - numpy replica of the library's counter-based dropout RNG (ast_b200/csrc/common.cuh:143-160
`hash_u32 / rng_u32 / rng_uniform / dropout_scale`) and of the (stream, index) each dropout site uses, so that the masks
the CUDA path applies in a training step can be handed to the oracle (`OracleModel.dropout_masks`) or to the reference
running on the Chainer stand-in (`chainer.functions.dropout_hook`).  Chainer's own RNG stream is not reproducible
(SURVEY 0.10), so parity WITH dropout is defined as "same masks in, same numbers out".

Sites (model.cu:405,491-492; dec_seq2.cu:182,334,480; dec_seq.cu:41,86; decoder.cu:57):
  step seed        cur_seed = seed + 0x9E3779B97F4A7C15 * k for the k-th training-mode encode since the seed was set
  encoder layer    stream 1 + 2*l + d (d = 0 fwd, 1 rev), index (i*B + b)*h + j, i = step in processing order (seq2seq.py:198)
  decoder layer    stream 16 + l, index (s*B + b)*H + j                                                   (seq2seq.py:198)
  embedding        stream 32, index (s*B + b)*E + j                                                       (seq2seq.py:365)
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)
GOLDEN = 0x9E3779B97F4A7C15


def _hash_u32(x):
    x = x & _M32
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x = x ^ (x >> np.uint64(15))
    x = (x * np.uint64(0x846CA68B)) & _M32
    x = x ^ (x >> np.uint64(16))
    return x


def rng_u32(seed, stream, idx):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    idx = np.asarray(idx, dtype=np.uint64) & _M32
    a = _hash_u32(np.uint64((seed & 0xFFFFFFFF) ^ ((int(stream) * 0x9E3779B9) & 0xFFFFFFFF)))
    b = _hash_u32((np.uint64(seed >> 32) + ((idx * np.uint64(0x85EBCA6B)) & _M32) + a) & _M32)
    return _hash_u32(a ^ b ^ idx)


def rng_uniform(seed, stream, idx):
    return ((rng_u32(seed, stream, idx) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def dropout_scale(seed, stream, idx, ratio):
    """Scaled keep mask of F.dropout: 0 with probability ratio, else 1/(1-ratio) (float32 arithmetic like the kernel)."""
    if ratio <= 0:
        return np.ones(np.shape(idx), dtype=np.float32)
    u = rng_uniform(seed, stream, idx)
    keep = np.float32(1.0) / (np.float32(1.0) - np.float32(ratio))
    return np.where(u < np.float32(ratio), np.float32(0), keep).astype(np.float32)


def step_seed(seed, k=1):
    return (int(seed) + GOLDEN * int(k)) & 0xFFFFFFFFFFFFFFFF


def training_masks(seed, k, B, Tp, S, h, H, E, nl, drop_rnn, drop_embed):
    """All scaled keep masks of one training step -> dict keyed like OracleModel.dropout_masks:
    ("L{l}_enc"|"L{l}_rev_enc", i) -> (B,h); ("L{l}_dec", s) -> (B,H); ("embed", s) -> (B,E)."""
    cs = step_seed(seed, k)
    out = {}
    if drop_rnn > 0:
        for l in range(nl):
            for d, stack in enumerate(("enc", "rev_enc")):
                m = dropout_scale(cs, 1 + 2 * l + d, np.arange(Tp * B * h, dtype=np.uint64), drop_rnn).reshape(Tp, B, h)
                for i in range(Tp):
                    out[(f"L{l}_{stack}", i)] = m[i]
            m = dropout_scale(cs, 16 + l, np.arange(S * B * H, dtype=np.uint64), drop_rnn).reshape(S, B, H)
            for s in range(S):
                out[(f"L{l}_dec", s)] = m[s]
    if drop_embed > 0:
        m = dropout_scale(cs, 32, np.arange(S * B * E, dtype=np.uint64), drop_embed).reshape(S, B, E)
        for s in range(S):
            out[("embed", s)] = m[s]
    return out


def reference_call_order(Tp, S, nl, drop_rnn, drop_embed):
    """Keys in the order the reference calls F.dropout with ratio > 0 during forward_loss (seq2seq.py:211-225: per
    timestep the forward stack's layers then the reverse stack's; :365,375: per decode step the embedding, then the
    decoder layers; dropout.out = 0 is never > 0)."""
    keys = []
    if drop_rnn > 0:
        for i in range(Tp):
            keys += [(f"L{l}_enc", i) for l in range(nl)]
            keys += [(f"L{l}_rev_enc", i) for l in range(nl)]
    for s in range(S):
        if drop_embed > 0:
            keys.append(("embed", s))
        if drop_rnn > 0:
            keys += [(f"L{l}_dec", s) for l in range(nl)]
    return keys

"""Pins for the CPU oracle (the reference ships no tests or golden vectors):
(1) float64 hand backward vs central finite differences, (2) vs a torch-autograd re-expression,
(3) fp32 vs fp64 agreement, (4) hand-computed known answers, (5) a scripted beam search,
(6) frozen golden outputs under tests/golden/."""
import math
import random

import numpy as np
import pytest

from oracle import ast_oracle as O
from golden_util import CASES, beam_hyps, decode_params, load_case
from torch_ref import torch_forward_loss


def _tiny(V=11):
    return O.default_model_cfg(vocab=V, hidden=8, embed=5, attn=8, layers=3,
                               cnn=((6, (9, 13), (2, 13), (4, 0)), (10, (9, 1), (2, 1), (4, 0))))


def _perturbed(cfg, D, seed):
    P = O.init_params(cfg, D, seed=seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 1)
    for k in P:
        if k.endswith(("gamma", "beta", "/b")):
            P[k] = P[k] + 0.1 * rng.standard_normal(P[k].shape)
    return P


@pytest.mark.parametrize("D", [13, 40])
def test_backward_matches_torch_autograd(D):
    cfg = _tiny()
    P = _perturbed(cfg, D, 1)
    X, y, _ = O.synth_batch(3, 23, D, 11, 4, 7, seed=3, Tmin=15)
    bits = [True, False, True, False, True, True, True][:y.shape[1] - 1]
    m = O.OracleModel(cfg, P, dtype=np.float64)
    loss = m.forward_loss(X.astype(np.float64), y, tf_bits=bits)
    g = m.backward()
    tl, tg, tenc = torch_forward_loss(cfg, P, X.astype(np.float64), y, tf_bits=bits)
    assert abs(loss - tl) < 1e-10
    assert np.abs(tenc - m.enc_states).max() < 1e-12
    assert set(tg) == set(g)
    for k in tg:
        assert np.abs(tg[k] - g[k]).max() <= 1e-9 * (np.abs(tg[k]).max() + 1e-12), k


def test_backward_matches_finite_differences():
    cfg = _tiny(7)
    D = 13
    P = _perturbed(cfg, D, 5)
    X, y, _ = O.synth_batch(2, 12, D, 7, 4, 4, seed=6)
    X = X.astype(np.float64)

    def loss_of(Pm):
        m = O.OracleModel(cfg, Pm, dtype=np.float64)
        return float(m.forward_loss(X, y))
    m = O.OracleModel(cfg, P, dtype=np.float64)
    m.forward_loss(X, y)
    g = m.backward()
    rng = np.random.default_rng(0)
    eps = 1e-6
    for k in sorted(g):
        idx = tuple(rng.integers(0, s) for s in P[k].shape)
        Pp = {a: b.copy() for a, b in P.items()}; Pm_ = {a: b.copy() for a, b in P.items()}
        Pp[k][idx] += eps; Pm_[k][idx] -= eps
        fd = (loss_of(Pp) - loss_of(Pm_)) / (2 * eps)
        assert abs(fd - g[k][idx]) <= 1e-5 * max(1.0, abs(fd)), (k, idx, fd, g[k][idx])


def test_fp32_agrees_with_fp64():
    cfg = O.default_model_cfg(vocab=50, hidden=64, embed=16, attn=64)
    P = O.init_params(cfg, 40, seed=2, dtype=np.float64)
    X, y, _ = O.synth_batch(4, 120, 40, 50, 5, 8, seed=3, Tmin=100)
    l64 = float(O.OracleModel(cfg, P, np.float64).forward_loss(X, y))
    l32 = float(O.OracleModel(cfg, P, np.float32).forward_loss(X, y))
    assert abs(l32 - l64) <= 1e-5 * abs(l64)


# ---- hand-computed known answers --------------------------------------------------------------------
def test_kat_lstm_cell_interleaved_gates_and_forget_bias():
    # one unit, gates laid out (a, i, f, o) at indices 4j+k; forget bias 1 lives at index 4j+2
    gates = np.array([[0.5, -1.0, 1.0, 2.0]])
    c_prev = np.array([[0.3]])
    c, h, (a, i, f, o) = O.lstm_cell(c_prev, gates)
    sig = lambda v: 1 / (1 + math.exp(-v))
    ca = math.tanh(0.5) * sig(-1.0) + sig(1.0) * 0.3
    assert np.allclose(c, ca) and np.allclose(h, sig(2.0) * math.tanh(ca))
    b = O.init_params(_tiny(), 13)["L0_enc/upward/b"]
    assert (b[2::4] == 1).all() and b.sum() == len(b) // 4


def test_kat_reverse_frame_order():
    assert O.reverse_frame_order(6) == [0, 5, 4, 3, 2, 1]          # X[-i], seq2seq.py:219
    assert O.reverse_frame_order(1) == [0]


def test_kat_cross_entropy_pad_row_counts_in_denominator():
    z = np.log(np.array([[0.1, 0.2, 0.7], [0.3, 0.3, 0.4]]))
    w = np.array([0.0, 1.0, 1.0])
    loss, dz = O.softmax_cross_entropy(z, np.array([2, 0]), w)
    assert np.isclose(loss, -math.log(0.7) / 2)                      # PAD row contributes 0, divisor stays B = 2
    assert np.allclose(dz[1], 0) and np.allclose(dz[0], (np.array([0.1, 0.2, 0.7]) - [0, 0, 1]) / 2)


def test_kat_attention_step():
    cfg = _tiny()
    P = {k: np.zeros_like(v) for k, v in O.init_params(cfg, 13, dtype=np.float64).items()}
    P["attn_Wa/W"] = np.eye(8)
    m = O.OracleModel(cfg, P, np.float64)
    enc = np.zeros((1, 3, 8)); enc[0, 0, 0] = 1.0; enc[0, 1, 0] = 2.0; enc[0, 2, 1] = 5.0
    m.enc_states = enc
    h = np.zeros((1, 8)); h[0, 0] = 1.0
    cv, alpha, q = m.compute_context_vector(h)
    e = np.exp([1.0, 2.0, 0.0]); a = e / e.sum()
    assert np.allclose(alpha[0], a) and np.allclose(cv[0, 0], a[0] + 2 * a[1]) and np.allclose(cv[0, 1], 5 * a[2])


def test_kat_amsgrad_two_steps_with_weight_decay_and_clip():
    p = {"w": np.array([1.0, -2.0])}
    opt = O.OracleAMSGrad(p, lr=0.1, l2=0.5, grad_clip=2.0)
    g1 = np.array([3.0, 4.0])
    gg = g1 + 0.5 * np.array([1.0, -2.0])                   # WeightDecay first
    n = math.sqrt((gg ** 2).sum()); gg = gg * (2.0 / n)      # then clip to norm 2
    m1 = 0.1 * gg; v1 = 0.001 * gg * gg
    a1 = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = np.array([1.0, -2.0]) - a1 * m1 / (np.sqrt(v1) + 1e-8)
    opt.update(p, {"w": g1.copy()})
    assert np.allclose(p["w"], want) and np.isclose(opt.last_norm, n)
    # second step with a tiny gradient: vhat keeps the larger v (AMSGrad)
    p0 = p["w"].copy()
    g2 = np.array([1e-3, 1e-3])
    gg2 = g2 + 0.5 * p0
    m2 = m1 + 0.1 * (gg2 - m1); v2 = v1 + 0.001 * (gg2 * gg2 - v1)
    a2 = 0.1 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    opt.update(p, {"w": g2.copy()})
    assert np.allclose(p["w"], p0 - a2 * m2 / (np.sqrt(np.maximum(v1, v2)) + 1e-8))


def test_kat_teacher_forcing_draw_order():
    class R:
        def __init__(self): self.n = 0
        def random(self):
            self.n += 1
            return [0.9, 0.1, 0.5][self.n - 1]
    r = R()
    bits = O.teacher_forcing_bits(6, 0.8, rng=r)              # steps 0..4; draws only for i = 1,2,3
    assert bits == [True, False, True, True, True] and r.n == 3
    assert O.teacher_forcing_bits(3, 0.0) == [True, True]      # L-1 = 2 steps, no draws


class _ScriptedModel(O.OracleModel):
    """Beam KAT: logits come from a table keyed by the last token (3-token toy vocabulary + specials)."""

    def __init__(self, table):
        self.table, self.train, self.A = table, False, 2
        self.dtype = np.dtype(np.float32)

    def encode(self, X, noise=None):
        pass

    def get_encoder_states(self):
        return {"c": [], "h": []}

    def get_decoder_states(self):
        return {"c": [], "h": []}

    def set_decoder_states(self, st):
        pass

    def decode_step(self, word, ht, step_key=None):
        z = np.log(np.asarray(self.table[int(word[0])], dtype=np.float32))[None]
        return z, ht, np.zeros((1, 1, 1), np.float32)


def test_kat_beam_carry_over_tie_and_stable_sort():
    # vocab ids: 0 PAD 1 GO 2 EOS 3 a 4 b ; after GO: a .5, b .3, EOS .2 ; after a: EOS .6 a .2 b .2 (tie) ; after b: EOS 1
    table = {1: [1e-9, 1e-9, .2, .5, .3], 3: [1e-9, 1e-9, .6, .2, .2], 4: [1e-9, 1e-9, 1 - 3e-9, 1e-9, 1e-9]}
    m = _ScriptedModel(table)
    nb = m.decode_beam(np.zeros((1, 4, 1), np.float32), stop_limit=5, N=3, K=2)
    hyps = [e["hyp"] for e in nb]
    assert hyps[0] == [1, 3, 2] and hyps[1] == [1, 4, 2]       # .5*.6 = .30 == .3*1 -> tie keeps insertion order
    assert all(h[-1] == 2 for h in hyps)
    assert np.isclose(float(nb[0]["score"]), math.log(.5) + math.log(.6), atol=1e-6)
    # top-K tie (a vs b after a): larger id first
    lp = O.log_softmax(np.log(np.asarray([table[3]], dtype=np.float32)))[0]
    assert list(np.argsort(lp, kind="stable")[-3:][::-1]) == [2, 4, 3]


def test_rerank_length_normalisation():
    beam = {"u": [([1, 5, 6, 7, 2], -4.0, []), ([1, 5, 2], -3.0, [])]}
    assert O.get_best_hyps(beam, 1.0)["u"] == [1, 5, 6, 7, 2]    # -4/3 > -3/1
    assert O.get_best_hyps(beam, 0.0)["u"] == [1, 5, 2]


def test_cmvn_matches_definition():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((200, 5)) * [1, 2, 3, 4, 5] + [5, 4, 3, 2, 1]).astype(np.float32)
    y = O.apply_cmvn(x, x.astype(np.float64).sum(0), (x.astype(np.float64) ** 2).sum(0), len(x))
    assert np.abs(y.mean(0)).max() < 1e-4 and np.abs(y.std(0) - 1).max() < 1e-4


def test_bucket_rule_and_labels():
    assert O.bucket_index(79, 80, 20) == 0 and O.bucket_index(80, 80, 20) == 1 and O.bucket_index(5000, 80, 20) == 19
    assert list(O.make_labels([7, 8, 9, 10], 5)) == [1, 7, 8, 9, 2]
    p = O.pad_sequence([np.ones((2, 3)), np.ones((4, 3))])
    assert p.shape == (2, 4, 3) and p[0, 2:].sum() == 0


# ---- golden fixtures ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    cfg, D, P, z = load_case(name)
    m = O.OracleModel(cfg, P, dtype=np.float64)
    loss = m.forward_loss(z["X"].astype(np.float64), z["y"], tf_bits=list(z["bits"]))
    g = m.backward()
    assert abs(loss - float(z["loss"])) <= 1e-10 * abs(loss)
    assert np.abs(m.enc_states - z["enc_states"]).max() < 1e-11
    assert (np.stack(m.step_argmax) == z["step_argmax"]).all()
    for k in g:
        assert np.isclose(np.sqrt((g[k] ** 2).sum()), float(z["gnorm:" + k]), rtol=1e-9), k
        if "grad:" + k in z.files:
            assert np.abs(g[k] - z["grad:" + k]).max() <= 1e-10 * (np.abs(g[k]).max() + 1e-12), k
    m32 = O.OracleModel(cfg, decode_params(P, z, False), dtype=np.float32)
    assert (m32.predict(z["X"], O.GO_ID, O.EOS_ID, 12) == z["greedy"]).all()
    mb = O.OracleModel(cfg, decode_params(P, z, True), dtype=np.float32)
    nb = mb.decode_beam(z["X"][:1, :int(z["beam_len0"])], 12, 4, 3)
    assert [e["hyp"] for e in nb] == beam_hyps(z)
    assert np.allclose([float(e["score"]) for e in nb], z["beam_scores"], rtol=1e-5)

"""CPU oracle for the 0xSameer/ast speech encoder-decoder hot path.

TEST INFRASTRUCTURE ONLY.  This module is the parity checker for the CUDA path in
``ast_b200``; it may be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by nothing else.  The
product path never routes through it.

PARITY PINNED TO THE REFERENCE'S OWN SOURCE (control flow and data path), op semantics restated.  The reference ships no
tests, golden vectors or known-answer files and its arithmetic lives in Chainer/CuPy, which are
neither vendored in /root/reference nor installable here.  oracle/gen_ref_golden.py therefore imports the reference's
seq2seq.py / nn.py / dataloader.py / config.py / eval.py UNMODIFIED on top of a float64 stand-in for the ~25 Chainer/CuPy names
they use (oracle/_ref_shim) and records what the reference computes: tests/golden/ref_*.npz (loss, per-step losses, logits,
encoder states, every gradient, two optimizer steps, BN running statistics, greedy tokens, beam hypotheses, a bucketed epoch).
tests/test_ref_golden.py checks this oracle against those fixtures to 1e-12 (float64); tests/test_gpu_ref_parity.py checks the
CUDA path against the same fixtures.  What stays unpinned is the semantics of the Chainer ops themselves (SURVEY.md Appendix
A: BN running-variance convention, hook order, beam tie order), restated in the stand-in because the library is absent; those
are additionally covered by finite differences in float64, a torch-autograd re-expression of the same graph and hand-computed
known answers (tests/test_oracle.py, tests/golden/).

Every function cites the reference file:line it follows.  The execution shape is
deliberately the one Chainer would run (per-timestep LSTM link calls with separate
``upward``/``lateral`` GEMMs, im2col convolution, per-step attention and
softmax-cross-entropy) because the same code is timed as the "reference CPU path".
"""
from __future__ import annotations

import math
import random as _pyrandom
from typing import Dict, List, Optional, Sequence

import numpy as np

# dataloader.py:26-36
PAD_ID, GO_ID, EOS_ID, UNK_ID = 0, 1, 2, 3
N_SPECIAL = 4

BN_EPS = 2e-5      # chainer.links.BatchNormalization default (Appendix A.2)
BN_DECAY = 0.9


# --------------------------------------------------------------------------------------
# configuration helpers
# --------------------------------------------------------------------------------------
def default_model_cfg(vocab: int = 1098, hidden: int = 512, embed: int = 128,
                      attn: int = 512, layers: int = 3,
                      cnn=((128, (9, 13), (2, 13), (4, 0)), (512, (9, 1), (2, 1), (4, 0))),
                      dropout=(0.0, 0.0, 0.0)) -> dict:
    """Same nested dict as experiments/es_en_20h/model_cfg.json + config.py:25 injection."""
    return {
        "dropout": {"embed": dropout[0], "rnn": dropout[1], "out": dropout[2]},
        "rnn_config": {"bi_rnn": True, "enc_layers": layers, "dec_layers": layers,
                       "hidden_units": hidden, "embedding_units": embed,
                       "attn_units": attn, "n_attn": 1, "feed_attn": True, "ln": False,
                       "dec_vocab_size": vocab},
        "cnn_config": {"bn": True, "cnn_layers": [
            {"in_channels": None, "out_channels": oc, "ksize": list(k), "stride": list(s),
             "pad": list(p)} for (oc, k, s, p) in cnn]},
    }


def conv_out_len(n: int, k: int, s: int, p: int) -> int:
    """Chainer get_conv_outsize, cover_all=False, dilate=1 (Appendix A.1)."""
    return (n + 2 * p - k) // s + 1


def cnn_shapes(cfg: dict, T: int, D: int):
    """[(Cin, Cout, kh, kw, sh, sw, ph, pw, Hin, Win, Hout, Wout), ...] per CNN layer."""
    out = []
    cin, h, w = 1, T, D
    for l in cfg["cnn_config"]["cnn_layers"]:
        kh, kw = l["ksize"]; sh, sw = l["stride"]; ph, pw = l["pad"]
        ho, wo = conv_out_len(h, kh, sh, ph), conv_out_len(w, kw, sw, pw)
        out.append((cin, l["out_channels"], kh, kw, sh, sw, ph, pw, h, w, ho, wo))
        cin, h, w = l["out_channels"], ho, wo
    return out


def rnn_in_dim(cfg: dict, D: int) -> int:
    sh = cnn_shapes(cfg, 64, D)
    return sh[-1][1] * sh[-1][11]


def param_shapes(cfg: dict, D: int) -> Dict[str, tuple]:
    """Key set / shapes of chainer.serializers.save_npz for seq2seq.py (Appendix A.9).

    Order is the canonical flat-buffer order used by ast_b200 as well.
    """
    r = cfg["rnn_config"]
    H, E, A, V = r["hidden_units"], r["embedding_units"], r["attn_units"], r["dec_vocab_size"]
    h = H // 2 if r["bi_rnn"] else H
    shapes: Dict[str, tuple] = {}
    cin = 1
    for i, l in enumerate(cfg["cnn_config"]["cnn_layers"]):
        oc = l["out_channels"]
        shapes[f"CNN_{i}/W"] = (oc, cin, l["ksize"][0], l["ksize"][1])
        shapes[f"CNN_{i}_bn/gamma"] = (oc,)
        shapes[f"CNN_{i}_bn/beta"] = (oc,)
        cin = oc
    rin = rnn_in_dim(cfg, D)
    for stack in ("enc", "rev_enc"):
        for l in range(r["enc_layers"]):
            ind = rin if l == 0 else h
            shapes[f"L{l}_{stack}/upward/W"] = (4 * h, ind)
            shapes[f"L{l}_{stack}/upward/b"] = (4 * h,)
            shapes[f"L{l}_{stack}/lateral/W"] = (4 * h, h)
    shapes["attn_Wa/W"] = (H, H)
    shapes["attn_Wa/b"] = (H,)
    shapes["context/W"] = (A, 2 * H)
    shapes["context/b"] = (A,)
    shapes["embed_dec/W"] = (V, E)
    for l in range(r["dec_layers"]):
        ind = E + A if l == 0 else H
        shapes[f"L{l}_dec/upward/W"] = (4 * H, ind)
        shapes[f"L{l}_dec/upward/b"] = (4 * H,)
        shapes[f"L{l}_dec/lateral/W"] = (4 * H, H)
    shapes["out/W"] = (V, A)
    shapes["out/b"] = (V,)
    return shapes


def persistent_shapes(cfg: dict) -> Dict[str, tuple]:
    """BatchNormalization persistents saved next to the params (Appendix A.2 / A.9)."""
    out = {}
    for i, l in enumerate(cfg["cnn_config"]["cnn_layers"]):
        oc = l["out_channels"]
        out[f"CNN_{i}_bn/avg_mean"] = (oc,)
        out[f"CNN_{i}_bn/avg_var"] = (oc,)
        out[f"CNN_{i}_bn/N"] = ()
    return out


def init_params(cfg: dict, D: int, seed: int = 0, dtype=np.float32) -> Dict[str, np.ndarray]:
    """Chainer-default initialisation (Appendix A.1-A.4) from a numpy Generator.

    Conv: HeNormal N(0, sqrt(2/fan_in)); Linear / LSTM blocks: LeCunNormal N(0, sqrt(1/fan_in));
    EmbedID: N(0,1); biases 0 except LSTM forget-gate bias 1 at interleaved index 4j+2;
    BN gamma 1, beta 0, avg_mean 0, avg_var 1, N 0.
    """
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, shp in param_shapes(cfg, D).items():
        if name.startswith("CNN_") and name.endswith("/W"):
            fan_in = shp[1] * shp[2] * shp[3]
            p[name] = rng.normal(0.0, math.sqrt(2.0 / fan_in), shp)
        elif name.endswith("/gamma"):
            p[name] = np.ones(shp)
        elif name.endswith("/beta"):
            p[name] = np.zeros(shp)
        elif name == "embed_dec/W":
            p[name] = rng.normal(0.0, 1.0, shp)
        elif name.endswith("/W"):
            p[name] = rng.normal(0.0, math.sqrt(1.0 / shp[1]), shp)
        elif name.endswith("upward/b"):
            b = np.zeros(shp)
            b[2::4] = 1.0                      # forget gate, interleaved layout (A.3)
            p[name] = b
        else:
            p[name] = np.zeros(shp)
    for name, shp in persistent_shapes(cfg).items():
        if name.endswith("avg_var"):
            p[name] = np.ones(shp)
        elif name.endswith("/N"):
            p[name] = np.zeros(shp, dtype=np.int64)
        else:
            p[name] = np.zeros(shp)
    return {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in p.items()}


def reverse_frame_order(T: int) -> List[int]:
    """Frame consumed by the reverse stack at step i: ``X[-i]`` (seq2seq.py:219,224)."""
    return [(-i) % T for i in range(T)]


def teacher_forcing_bits(L: int, teach_ratio: float, rng=_pyrandom) -> List[bool]:
    """The host random draws of seq2seq.py:431-436, one per decode step i in [0, L-2].

    bits[i] True -> the true token y[i] is fed at step i.  Steps 0 and >= L-2 are always
    True and consume no random number; steps 1..L-3 draw ``random.random() < teach_ratio``.
    """
    bits = []
    for i in range(L - 1):
        if 0 < i < L - 2:
            bits.append(rng.random() < teach_ratio)
        else:
            bits.append(True)
    return bits


# --------------------------------------------------------------------------------------
# small numerics
# --------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _im2col(x, kh, kw, sh, sw, ph, pw):
    """x (B,C,H,W) -> cols (B*Ho*Wo, C*kh*kw), row order (b,ho,wo), col order (c,kh,kw)."""
    B, C, H, W = x.shape
    ho, wo = conv_out_len(H, kh, sh, ph), conv_out_len(W, kw, sw, pw)
    xp = np.zeros((B, C, H + 2 * ph, W + 2 * pw), dtype=x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    cols = np.empty((B, ho, wo, C, kh, kw), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, :, :, i, j] = xp[:, :, i:i + sh * ho:sh, j:j + sw * wo:sw].transpose(0, 2, 3, 1)
    return cols.reshape(B * ho * wo, C * kh * kw), ho, wo


def _col2im(dcols, xshape, kh, kw, sh, sw, ph, pw):
    B, C, H, W = xshape
    ho, wo = conv_out_len(H, kh, sh, ph), conv_out_len(W, kw, sw, pw)
    d = dcols.reshape(B, ho, wo, C, kh, kw)
    dxp = np.zeros((B, C, H + 2 * ph, W + 2 * pw), dtype=dcols.dtype)
    for i in range(kh):
        for j in range(kw):
            dxp[:, :, i:i + sh * ho:sh, j:j + sw * wo:sw] += d[:, :, :, :, i, j].transpose(0, 3, 1, 2)
    return dxp[:, :, ph:ph + H, pw:pw + W]


def lstm_cell(c_prev, gates):
    """chainer.functions.lstm with the interleaved gate layout (Appendix A.3).

    gates (B,4h) viewed (B,h,4): k=0 a(tanh), 1 i, 2 f, 3 o (sigmoid).
    Returns c, h and the activated gates (a,i,f,o) for backward.
    """
    B = gates.shape[0]
    g = gates.reshape(B, -1, 4)
    a = np.tanh(g[:, :, 0]); i = _sigmoid(g[:, :, 1]); f = _sigmoid(g[:, :, 2]); o = _sigmoid(g[:, :, 3])
    c = a * i + f * c_prev
    h = o * np.tanh(c)
    return c, h, (a, i, f, o)


def lstm_cell_bwd(dh, dc_next, c_prev, c, act):
    """Backward of lstm_cell. Returns dgates (B,4h interleaved), dc_prev."""
    a, i, f, o = act
    tc = np.tanh(c)
    dc = dc_next + dh * o * (1.0 - tc * tc)
    dg = np.empty(a.shape + (4,), dtype=a.dtype)
    dg[:, :, 0] = dc * i * (1.0 - a * a)
    dg[:, :, 1] = dc * a * i * (1.0 - i)
    dg[:, :, 2] = dc * c_prev * f * (1.0 - f)
    dg[:, :, 3] = dh * tc * o * (1.0 - o)
    return dg.reshape(a.shape[0], -1), dc * f


def log_softmax(z):
    """chainer F.log_softmax: x - logsumexp(x) with logsumexp = max + log(sum(exp(x - max)))."""
    m = z.max(axis=-1, keepdims=True)
    return z - (m + np.log(np.exp(z - m).sum(axis=-1, keepdims=True)))


def softmax_cross_entropy(z, t, class_weight):
    """F.softmax_cross_entropy(x, t, class_weight=w), normalize=True, reduce='mean' (A.7).

    Divisor is the number of rows whose label != -1, i.e. B here: PAD rows are counted
    in the denominator and contribute 0 through w[PAD]=0.  Returns (loss, dz).
    """
    B = z.shape[0]
    lp = log_softmax(z)
    w = class_weight[t]
    loss = -(w * lp[np.arange(B), t]).sum() / max(B, 1)
    dz = np.exp(lp)
    dz[np.arange(B), t] -= 1.0
    dz *= (w / max(B, 1))[:, None]
    return loss, dz


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class OracleModel:
    """numpy restatement of seq2seq.py::SpeechEncoderDecoder (live model)."""

    def __init__(self, cfg: dict, params: Dict[str, np.ndarray], dtype=np.float32):
        self.cfg = cfg
        self.dtype = np.dtype(dtype)
        self.p = {k: (np.array(v, dtype=dtype) if np.asarray(v).dtype.kind == "f" else np.array(v))
                  for k, v in params.items()}
        r = cfg["rnn_config"]
        self.H, self.E, self.A, self.V = (r["hidden_units"], r["embedding_units"],
                                          r["attn_units"], r["dec_vocab_size"])
        self.nl = r["enc_layers"]
        self.h = self.H // 2
        self.n_cnn = len(cfg["cnn_config"]["cnn_layers"])
        # seq2seq.py:152-156
        self.mask_pad_id = np.ones(self.V, dtype=dtype)
        self.mask_pad_id[PAD_ID] = 0
        self.train = True
        self.grads: Dict[str, np.ndarray] = {}

    # ---- CNN front-end: seq2seq.py:158-180 --------------------------------------------
    def forward_cnn(self, X):
        p = self.p
        h = X[:, None, :, :]                                   # expand_dims + swapaxes :160-161
        self._cnn_cache = []
        for i, l in enumerate(self.cfg["cnn_config"]["cnn_layers"]):
            kh, kw = l["ksize"]; sh, sw = l["stride"]; ph, pw = l["pad"]
            W = p[f"CNN_{i}/W"]
            cols, ho, wo = _im2col(h, kh, kw, sh, sw, ph, pw)
            y = cols @ W.reshape(W.shape[0], -1).T             # (B*ho*wo, Cout), no bias (:54)
            B = h.shape[0]
            gamma, beta = p[f"CNN_{i}_bn/gamma"], p[f"CNN_{i}_bn/beta"]
            if self.train:
                m = y.shape[0]
                mu = y.mean(axis=0)
                var = y.var(axis=0)                            # biased (A.2)
                inv = 1.0 / np.sqrt(var + self.dtype.type(BN_EPS))
                xhat = (y - mu) * inv
                adj = m / max(m - 1.0, 1.0)
                p[f"CNN_{i}_bn/avg_mean"] = (BN_DECAY * p[f"CNN_{i}_bn/avg_mean"]
                                             + (1 - BN_DECAY) * mu).astype(self.dtype)
                p[f"CNN_{i}_bn/avg_var"] = (BN_DECAY * p[f"CNN_{i}_bn/avg_var"]
                                            + (1 - BN_DECAY) * var * adj).astype(self.dtype)
                p[f"CNN_{i}_bn/N"] = p[f"CNN_{i}_bn/N"] + 1
            else:
                inv = 1.0 / np.sqrt(p[f"CNN_{i}_bn/avg_var"] + self.dtype.type(BN_EPS))
                xhat = (y - p[f"CNN_{i}_bn/avg_mean"]) * inv
            z = gamma * xhat + beta
            a = np.maximum(z, 0)                               # relu :171
            self._cnn_cache.append((h.shape, cols, xhat, inv, a > 0, (kh, kw, sh, sw, ph, pw)))
            h = a.reshape(B, ho, wo, -1).transpose(0, 3, 1, 2)  # (B,C,T',F')
        B, C, Tp, Fp = h.shape
        self._cnn_out_shape = h.shape
        # swapaxes(1,2) -> (B,T',C,F'); reshape -> (B,T',C*F'); rollaxis(1) -> (T',B,C*F')  :177-179
        return np.ascontiguousarray(h.transpose(2, 0, 1, 3).reshape(Tp, B, C * Fp))

    def backward_cnn(self, d_rnn_in):
        p, g = self.p, self.grads
        B, C, Tp, Fp = self._cnn_out_shape
        dh = d_rnn_in.reshape(Tp, B, C, Fp).transpose(1, 2, 0, 3)       # (B,C,T',F')
        for i in reversed(range(self.n_cnn)):
            xshape, cols, xhat, inv, mask, (kh, kw, sh, sw, ph, pw) = self._cnn_cache[i]
            dy = dh.transpose(0, 2, 3, 1).reshape(-1, dh.shape[1]) * mask
            gamma = p[f"CNN_{i}_bn/gamma"]
            m = dy.shape[0]
            dbeta = dy.sum(axis=0)
            dgamma = (dy * xhat).sum(axis=0)
            g[f"CNN_{i}_bn/beta"] = dbeta
            g[f"CNN_{i}_bn/gamma"] = dgamma
            dconv = (gamma * inv) * (dy - dbeta / m - xhat * (dgamma / m))
            W = p[f"CNN_{i}/W"]
            g[f"CNN_{i}/W"] = (dconv.T @ cols).reshape(W.shape)
            if i > 0:
                dcols = dconv @ W.reshape(W.shape[0], -1)
                dh = _col2im(dcols, xshape, kh, kw, sh, sw, ph, pw)

    # ---- encoder LSTM stacks: seq2seq.py:182-242 --------------------------------------
    def _lstm_link(self, name, x, state):
        """One L.LSTM.__call__ (A.3): separate upward / lateral GEMMs, like Chainer."""
        p = self.p
        gates = x @ p[f"{name}/upward/W"].T + p[f"{name}/upward/b"]
        if state["h"] is not None:
            gates = gates + state["h"] @ p[f"{name}/lateral/W"].T
        if state["c"] is None:
            state["c"] = np.zeros((x.shape[0], gates.shape[1] // 4), dtype=self.dtype)
        c_prev, h_prev = state["c"], state["h"]
        c, h, act = lstm_cell(c_prev, gates)
        state["c"], state["h"] = c, h
        return h, (x, h_prev, c_prev, c, act)

    def reset_rnn_state(self):
        names = ([f"L{l}_enc" for l in range(self.nl)] + [f"L{l}_rev_enc" for l in range(self.nl)]
                 + [f"L{l}_dec" for l in range(self.nl)])
        self.state = {n: {"h": None, "c": None} for n in names}

    def _dropout(self, x, ratio, key):
        """F.dropout (A.5). Masks are injected via self.dropout_masks[key] (already scaled)
        because Chainer's RNG stream is not reproducible; absent mask == ratio 0."""
        m = getattr(self, "dropout_masks", None)
        if not self.train or ratio <= 0 or m is None or key not in m:
            return x, None
        return x * m[key], m[key]

    def forward_rnn_encode(self, Xr):
        """seq2seq.py:205-242.  Xr (T',B,R).  Sets self.enc_states (B,T',H)."""
        self.reset_rnn_state()
        Tp, B, _ = Xr.shape
        rr = self.cfg["dropout"]["rnn"]
        order = reverse_frame_order(Tp)
        self._enc_cache = {"fwd": [], "rev": []}
        h_fwd = np.empty((Tp, B, self.h), dtype=self.dtype)
        h_rev = np.empty((Tp, B, self.h), dtype=self.dtype)
        for i in range(Tp):
            for tag, stack, src, dst in (("fwd", "enc", i, h_fwd), ("rev", "rev_enc", order[i], h_rev)):
                hs = Xr[src]
                step = []
                for l in range(self.nl):
                    name = f"L{l}_{stack}"
                    hs, cache = self._lstm_link(name, hs, self.state[name])
                    hs, dm = self._dropout(hs, rr, (name, i))
                    step.append((cache, dm))
                dst[i] = hs
                self._enc_cache[tag].append(step)
        h_rev = h_rev[::-1]                                                # flipud :231
        self.enc_states = np.ascontiguousarray(
            np.concatenate((h_fwd, h_rev), axis=2).transpose(1, 0, 2))     # :232,242
        self._enc_T = Tp

    def backward_rnn_encode(self, d_enc_states, dh_fin, dc_fin):
        """BPTT through both stacks. d_enc_states (B,T',H); dh_fin/dc_fin: per-layer (B,H)
        gradients w.r.t. the decoder's initial state (init_decoder_state :318-334).
        Returns d_rnn_in (T',B,R)."""
        p, g = self.p, self.grads
        Tp, hh = self._enc_T, self.h
        order = reverse_frame_order(Tp)
        d_te = d_enc_states.transpose(1, 0, 2)                             # (T',B,H)
        d_rnn_in = None
        for tag, stack, sl in (("fwd", "enc", slice(0, hh)), ("rev", "rev_enc", slice(hh, 2 * hh))):
            names = [f"L{l}_{stack}" for l in range(self.nl)]
            for n in names:
                for k in ("upward/W", "upward/b", "lateral/W"):
                    g[f"{n}/{k}"] = np.zeros_like(p[f"{n}/{k}"])
            dh_rec = [dh_fin[l][:, sl].copy() for l in range(self.nl)]
            dc_rec = [dc_fin[l][:, sl].copy() for l in range(self.nl)]
            for i in reversed(range(Tp)):
                d_out = d_te[i, :, sl] if tag == "fwd" else d_te[Tp - 1 - i, :, sl]
                step = self._enc_cache[tag][i]
                for l in reversed(range(self.nl)):
                    (x, h_prev, c_prev, c, act), dm = step[l]
                    if dm is not None:
                        d_out = d_out * dm
                    dh = d_out + dh_rec[l]
                    dgate, dc_rec[l] = lstm_cell_bwd(dh, dc_rec[l], c_prev, c, act)
                    n = names[l]
                    g[f"{n}/upward/W"] += dgate.T @ x
                    g[f"{n}/upward/b"] += dgate.sum(axis=0)
                    if h_prev is not None:
                        g[f"{n}/lateral/W"] += dgate.T @ h_prev
                        dh_rec[l] = dgate @ p[f"{n}/lateral/W"]
                    else:
                        dh_rec[l] = np.zeros_like(dh)
                    d_out = dgate @ p[f"{n}/upward/W"]
                if d_rnn_in is None:
                    d_rnn_in = np.zeros((Tp,) + d_out.shape, dtype=self.dtype)
                d_rnn_in[i if tag == "fwd" else order[i]] += d_out
        return d_rnn_in

    # ---- encode: seq2seq.py:293-315 ---------------------------------------------------
    def encode(self, X, noise=None):
        """noise: optional multiplicative tensor standing in for np.random.normal(1,sigma)
        (:300-305); None == add_noise 0."""
        X = np.asarray(X, dtype=self.dtype)
        if noise is not None and self.train:
            X = X * noise.astype(self.dtype)
        self.forward_rnn_encode(self.forward_cnn(X))

    # ---- decoder ----------------------------------------------------------------------
    def get_encoder_states(self):
        """seq2seq.py:529-547 — pre-dropout link states, fwd||rev."""
        out = {"c": [], "h": []}
        for l in range(self.nl):
            f, r = self.state[f"L{l}_enc"], self.state[f"L{l}_rev_enc"]
            out["h"].append(np.concatenate((f["h"], r["h"]), axis=1))
            out["c"].append(np.concatenate((f["c"], r["c"]), axis=1))
        return out

    def init_decoder_state(self):
        """seq2seq.py:318-334 (set_state(c, h) argument order)."""
        self.set_decoder_states(self.get_encoder_states())

    def get_decoder_states(self):
        return {"c": [self.state[f"L{l}_dec"]["c"] for l in range(self.nl)],
                "h": [self.state[f"L{l}_dec"]["h"] for l in range(self.nl)]}

    def set_decoder_states(self, st):
        for l in range(self.nl):
            self.state[f"L{l}_dec"] = {"c": st["c"][l], "h": st["h"][l]}

    def compute_context_vector(self, dec_h):
        """seq2seq.py:336-358: Luong 'general' attention, NO length mask (:344-347)."""
        p = self.p
        q = dec_h @ p["attn_Wa/W"].T + p["attn_Wa/b"]                 # :341
        s = np.einsum("bth,bh->bt", self.enc_states, q)               # batch_matmul :342
        s = s - s.max(axis=1, keepdims=True)
        e = np.exp(s)
        alpha = e / e.sum(axis=1, keepdims=True)                      # softmax over T' :351
        cv = np.einsum("bth,bt->bh", self.enc_states, alpha)          # :355
        return cv, alpha, q

    def decode_step(self, word, ht, step_key=None):
        """seq2seq.py:361-396.  Returns logits (B,V), ht (B,A), alphas (B,T',1)."""
        p = self.p
        d = self.cfg["dropout"]
        emb = p["embed_dec/W"][word]                                   # :365
        emb, dm_e = self._dropout(emb, d["embed"], ("embed", step_key))
        x = np.concatenate((emb, ht), axis=1)                          # input feeding :372
        caches = []
        hs = x
        for l in range(self.nl):
            name = f"L{l}_dec"
            hs, cache = self._lstm_link(name, hs, self.state[name])
            hs, dm = self._dropout(hs, d["rnn"], (name, step_key))
            caches.append((cache, dm))
        cv, alpha, q = self.compute_context_vector(hs)                 # :379
        cvh = np.concatenate((cv, hs), axis=1)                         # :386
        ht_new = np.tanh(cvh @ p["context/W"].T + p["context/b"])      # :390
        logits = ht_new @ p["out/W"].T + p["out/b"]                    # :394 (dropout out = 0)
        self._last_dec = (word, dm_e, caches, hs, alpha, q, cvh, ht_new)
        return logits, ht_new, alpha[:, :, None]

    # ---- loss: seq2seq.py:399-473 -----------------------------------------------------
    def forward_loss(self, X, y, tf_bits: Optional[Sequence[bool]] = None, noise=None, feedback=None):
        """Sum over decode steps of the batch-mean, PAD-weighted CE.  ``tf_bits`` are the
        scheduled-sampling draws (see teacher_forcing_bits); None == teach_ratio 1.
        ``feedback`` ((L-1, B) ints, test hook): tokens fed back at sampled steps INSTEAD of this model's own argmax -
        lets a gradient comparison against an implementation in lower precision (whose argmax may flip on a near-tie
        and then follows a different, equally valid, trajectory) be made along the same token path."""
        X = np.asarray(X, dtype=self.dtype)
        y = np.asarray(y)
        B, L = y.shape
        self.encode(X, noise)
        self.init_decoder_state()
        self._dec_init = self.get_decoder_states()
        yT = y.T
        ht = np.zeros((B, self.A), dtype=self.dtype)                   # :420
        loss = self.dtype.type(0)
        self._dec_cache = []
        decoder_input = None
        self.step_losses = []
        self.step_argmax = []
        for i in range(L - 1):                                         # zip(y, y[1:]) :423
            if tf_bits is None or tf_bits[i] or i == 0 or i >= L - 2:
                decoder_input = yT[i]
            logits, ht, _ = self.decode_step(decoder_input.astype(np.int64), ht, step_key=i)
            decoder_input = logits.argmax(axis=1)                      # :448 (ties -> lowest idx)
            li, dz = softmax_cross_entropy(logits, yT[i + 1].astype(np.int64), self.mask_pad_id)
            loss = loss + li
            self.step_losses.append(float(li))
            self.step_argmax.append(decoder_input.copy())
            if feedback is not None:
                decoder_input = np.asarray(feedback[i]).astype(np.int64)
            self._dec_cache.append(self._last_dec + (dz,))
        self.loss = loss
        return loss

    def backward(self):
        """Reverse-mode through forward_loss (nn.py:181).  Fills self.grads (cleargrads
        semantics: every grad is overwritten)."""
        p = self.p
        g = self.grads = {}
        for k in ("attn_Wa/W", "attn_Wa/b", "context/W", "context/b", "embed_dec/W", "out/W", "out/b"):
            g[k] = np.zeros_like(p[k])
        for l in range(self.nl):
            for k in ("upward/W", "upward/b", "lateral/W"):
                g[f"L{l}_dec/{k}"] = np.zeros_like(p[f"L{l}_dec/{k}"])
        enc = self.enc_states
        B = enc.shape[0]
        H, E = self.H, self.E
        d_enc = np.zeros_like(enc)
        dh_rec = [np.zeros((B, H), dtype=self.dtype) for _ in range(self.nl)]
        dc_rec = [np.zeros((B, H), dtype=self.dtype) for _ in range(self.nl)]
        dht_feed = np.zeros((B, self.A), dtype=self.dtype)
        for (word, dm_e, caches, hs, alpha, q, cvh, ht_new, dz) in reversed(self._dec_cache):
            g["out/W"] += dz.T @ ht_new
            g["out/b"] += dz.sum(axis=0)
            dht = dz @ p["out/W"] + dht_feed
            du = dht * (1.0 - ht_new * ht_new)
            g["context/W"] += du.T @ cvh
            g["context/b"] += du.sum(axis=0)
            dcvh = du @ p["context/W"]
            dcv, dh_top = dcvh[:, :H], dcvh[:, H:].copy()
            dalpha = np.einsum("bth,bh->bt", enc, dcv)
            d_enc += alpha[:, :, None] * dcv[:, None, :]
            ds = alpha * (dalpha - (alpha * dalpha).sum(axis=1, keepdims=True))
            dq = np.einsum("bt,bth->bh", ds, enc)
            d_enc += ds[:, :, None] * q[:, None, :]
            g["attn_Wa/W"] += dq.T @ hs
            g["attn_Wa/b"] += dq.sum(axis=0)
            dh_top += dq @ p["attn_Wa/W"]
            d_out = dh_top
            for l in reversed(range(self.nl)):
                (x, h_prev, c_prev, c, act), dm = caches[l]
                if dm is not None:
                    d_out = d_out * dm
                dh = d_out + dh_rec[l]
                dgate, dc_rec[l] = lstm_cell_bwd(dh, dc_rec[l], c_prev, c, act)
                n = f"L{l}_dec"
                g[f"{n}/upward/W"] += dgate.T @ x
                g[f"{n}/upward/b"] += dgate.sum(axis=0)
                g[f"{n}/lateral/W"] += dgate.T @ h_prev
                dh_rec[l] = dgate @ p[f"{n}/lateral/W"]
                d_out = dgate @ p[f"{n}/upward/W"]
            de = d_out[:, :E]
            if dm_e is not None:
                de = de * dm_e
            np.add.at(g["embed_dec/W"], word, de)                      # EmbedID scatter-add (A.4)
            dht_feed = d_out[:, E:]
        d_rnn_in = self.backward_rnn_encode(d_enc, dh_rec, dc_rec)
        self.backward_cnn(d_rnn_in)
        return g

    # ---- greedy decode: seq2seq.py:475-527 --------------------------------------------
    def predict(self, X, start_token=GO_ID, end_token=EOS_ID, stop_limit=175):
        was = self.train
        self.train = False
        try:
            X = np.asarray(X, dtype=self.dtype)
            B = X.shape[0]
            self.encode(X)
            self.init_decoder_state()
            done = np.zeros(B, dtype=bool)
            ht = np.zeros((B, self.A), dtype=self.dtype)
            cur = np.full((B,), start_token, dtype=np.int64)
            preds = []
            npred = 0
            while npred < stop_limit:
                logits, ht, _ = self.decode_step(cur, ht)
                cur = logits.argmax(axis=1)
                preds.append(cur.astype(np.int32))
                done[cur == end_token] = True
                if done.all():
                    break
                npred += 1
            return np.stack(preds, axis=0).T                            # (B, n_steps), not cut at EOS
        finally:
            self.train = was

    # ---- beam search: nn.py:235-322 ---------------------------------------------------
    def decode_beam(self, X, stop_limit, N, K):
        """Batch-size-1 beam search with float32 score accumulation (nn.py:289), finished
        hypotheses carried over (:317-318) and a stable descending sort (:320).
        top-K order: descending log-prob; exact ties -> larger token id first (what a stable
        ascending argsort reversed yields; numpy's introsort leaves ties unspecified)."""
        was = self.train
        self.train = False
        try:
            X = np.asarray(X, dtype=self.dtype)
            assert X.shape[0] == 1
            self.encode(X)
            n_best = [{"hyp": [GO_ID], "score": 0, "dec_state": self.get_encoder_states(),
                       "attn_v": np.zeros((1, self.A), dtype=self.dtype), "attn_history": []}]
            for _ in range(stop_limit):
                if all(e["hyp"][-1] == EOS_ID for e in n_best):
                    break
                cur = []
                for e in n_best:
                    if e["hyp"][-1] == EOS_ID:
                        cur.append(e)
                        continue
                    self.set_decoder_states(e["dec_state"])
                    logits, ht, alphas = self.decode_step(np.array([e["hyp"][-1]], dtype=np.int64),
                                                          e["attn_v"])
                    lp = log_softmax(logits.astype(np.float32))[0].astype(np.float32)
                    top = np.argsort(lp, kind="stable")[-K:][::-1]
                    st = self.get_decoder_states()
                    for pi in top:
                        cur.append({"hyp": e["hyp"] + [int(pi)],
                                    "score": np.float32(np.float32(e["score"]) + lp[pi]),
                                    "dec_state": st, "attn_v": ht,
                                    "attn_history": e["attn_history"] + [alphas[0, :, 0].copy()]})
                n_best = sorted(cur, reverse=True, key=lambda t: t["score"])[:N]
            return n_best
        finally:
            self.train = was


# beam.py:30-42
def rerank_hypothesis(beam_hyps, weight):
    return sorted([(i[0], i[1] / math.pow(len(i[0]) - 2, weight), len(i[0])) for i in beam_hyps],
                  reverse=True, key=lambda t: t[1])


def get_best_hyps(utts_beam, W):
    return {u: list(rerank_hypothesis(utts_beam[u], W)[0][0]) for u in utts_beam}


# --------------------------------------------------------------------------------------
# optimizer: nn.py:81-119 + optimizer.update() (Appendix A.10)
# --------------------------------------------------------------------------------------
class OracleAMSGrad:
    """WeightDecay -> GradientClipping (global L2 norm) -> AMSGrad-Adam, in hook order."""

    def __init__(self, params: Dict[str, np.ndarray], lr=1e-3, l2=1e-4, grad_clip=2.0,
                 beta1=0.9, beta2=0.999, eps=1e-8, freeze: Sequence[str] = ()):
        self.lr, self.l2, self.clip = lr, l2, grad_clip
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.t = 0
        self.freeze = tuple(freeze)
        self.m = {k: np.zeros_like(v) for k, v in params.items() if v.dtype.kind == "f"}
        self.v = {k: np.zeros_like(v) for k, v in self.m.items()}
        self.vhat = {k: np.zeros_like(v) for k, v in self.m.items()}
        self.last_norm = None

    def update(self, params: Dict[str, np.ndarray], grads: Dict[str, np.ndarray]):
        names = [k for k in grads]
        dt = params[names[0]].dtype.type
        if self.l2 > 0:
            for k in names:
                grads[k] = grads[k] + dt(self.l2) * params[k]
        sq = 0.0
        for k in names:
            sq += float((grads[k].astype(np.float64) ** 2).sum())
        norm = math.sqrt(sq)
        self.last_norm = norm
        rate = self.clip / norm if norm > 0 else 1.0
        if rate < 1:
            for k in names:
                grads[k] = grads[k] * dt(rate)
        self.t += 1
        fix1 = 1.0 - self.b1 ** self.t
        fix2 = 1.0 - self.b2 ** self.t
        alpha_t = self.lr * math.sqrt(fix2) / fix1
        for k in names:
            if any(k.startswith(f + "/") for f in self.freeze):
                continue
            gk = grads[k]
            self.m[k] += dt(1 - self.b1) * (gk - self.m[k])
            self.v[k] += dt(1 - self.b2) * (gk * gk - self.v[k])
            np.maximum(self.vhat[k], self.v[k], out=self.vhat[k])
            params[k] -= dt(alpha_t) * self.m[k] / (np.sqrt(self.vhat[k]) + dt(self.eps))


# --------------------------------------------------------------------------------------
# data-side restatements
# --------------------------------------------------------------------------------------
def bucket_index(n_frames: int, width_b: int, num_b: int) -> int:
    """preprocessing/prep_buckets.py:52."""
    return min(n_frames // width_b, num_b - 1)


def pad_sequence(xs: Sequence[np.ndarray], padding=0) -> np.ndarray:
    """F.pad_sequence: pad to the longest in the list (A.8; dataloader.py:156-161)."""
    L = max(len(x) for x in xs)
    out = np.full((len(xs), L) + xs[0].shape[1:], padding, dtype=xs[0].dtype)
    for i, x in enumerate(xs):
        out[i, :len(x)] = x
    return out


def make_labels(ids: Sequence[int], max_pred: int) -> np.ndarray:
    """dataloader.py:151."""
    return np.asarray([GO_ID] + list(ids)[:max_pred - 2] + [EOS_ID], dtype=np.int32)


def apply_cmvn(feats: np.ndarray, stats_sum: np.ndarray, stats_sumsq: np.ndarray, count: float,
               norm_vars: bool = True) -> np.ndarray:
    """Kaldi ``apply-cmvn --norm-vars=true`` arithmetic (linking_files/apply_cmvn.sh:11;
    parity unpinned — Kaldi is external).  Stats are accumulated in double; per dim
    mean = sum/count, var = sumsq/count - mean^2 (floored at 1e-20), scale = 1/sqrt(var),
    offset = -mean*scale; y = x*scale + offset applied in the feature precision."""
    mean = np.asarray(stats_sum, dtype=np.float64) / count
    if norm_vars:
        var = np.maximum(np.asarray(stats_sumsq, dtype=np.float64) / count - mean * mean, 1e-20)
        scale = 1.0 / np.sqrt(var)
    else:
        scale = np.ones_like(mean)
    offset = -mean * scale
    return (feats * scale.astype(feats.dtype) + offset.astype(feats.dtype)).astype(feats.dtype)


def pack_cmvn_batch(utts: Sequence[np.ndarray], spk_sum, spk_sumsq, spk_count, max_sp: int,
                    drop_masks=None) -> np.ndarray:
    """CMVN per utterance-speaker, truncate to max_sp (dataloader.py:103), optional frame
    zeroing (:83-93, mask injected), zero-pad to the batch max (:156)."""
    outs = []
    for i, u in enumerate(utts):
        x = apply_cmvn(u[:max_sp].astype(np.float32), spk_sum[i], spk_sumsq[i], spk_count[i])
        if drop_masks is not None:
            x = x * drop_masks[i][:len(x), None].astype(np.float32)
        outs.append(x)
    return pad_sequence(outs, 0)


# --------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------
def synth_batch(B: int, T: int, D: int, V: int, Lmin: int, Lmax: int, seed: int,
                Tmin: Optional[int] = None):
    """C1-style synthetic batch: lengths U[Tmin,T] with one utterance at exactly T, X~N(0,1)
    zero beyond each length, labels GO + U[4,V) + EOS padded with PAD."""
    rng = np.random.default_rng(seed)
    Tmin = T if Tmin is None else Tmin
    lens = rng.integers(Tmin, T + 1, size=B)
    lens[rng.integers(0, B)] = T
    X = np.zeros((B, T, D), dtype=np.float32)
    for b in range(B):
        X[b, :lens[b]] = rng.standard_normal((lens[b], D)).astype(np.float32)
    ys = []
    for b in range(B):
        n = int(rng.integers(Lmin, Lmax + 1))
        ys.append(make_labels(rng.integers(N_SPECIAL, V, size=n - 2).tolist(), 10 ** 9))
    y = pad_sequence(ys, PAD_ID)
    return X, y, lens

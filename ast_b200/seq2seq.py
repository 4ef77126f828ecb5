"""Drop-in for the reference's live model, seq2seq.py::SpeechEncoderDecoder.

Same constructor, method names, argument meaning and return shapes as /root/reference/seq2seq.py;
everything below the method surface is the sm_100a CUDA path (libast_b200.so).  Links are reachable
as attributes and through ``model["name"]``, expose ``.W.data`` / ``.lateral.W.data`` etc. as views
of the flat device buffer, and re-binding a link (copy_params.py:26-43) transfers its parameters.
"""
import math
import random

import numpy as np
import torch

from .engine import Engine
from .symbols import SYMBOLS  # noqa: F401  (re-exported like `from dataloader import SYMBOLS`, seq2seq.py:20)

# chainer.config.train stand-in, toggled by ast_b200.nn.using_config (nn.py:174,216)
class _Config:
    train = True


config = _Config()


def draw_use_true(L, teach_ratio):
    """Scheduled-sampling bits for L-1 decode steps with the reference's exact random.random() call
    order (seq2seq.py:431-436): one draw for each 1 <= i <= L-3, steps 0 and >= L-2 always teacher-forced."""
    return [True if not (0 < i < L - 2) else (random.random() < teach_ratio) for i in range(L - 1)]


class Variable:
    """Minimal chainer.Variable stand-in: ``.data`` is a torch CUDA tensor."""

    def __init__(self, data, backward_fn=None):
        self.data = data
        self.array = data
        self._backward_fn = backward_fn

    @property
    def shape(self):
        return tuple(self.data.shape)

    def backward(self):
        if self._backward_fn is None:
            raise RuntimeError("this Variable is not a loss")
        self._backward_fn()

    def __float__(self):
        return float(self.data)

    def __len__(self):
        return self.data.shape[0]

    def __array__(self, dtype=None):
        a = self.data.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


class Param:
    """chainer.Parameter stand-in over a view of the flat parameter / gradient buffers."""

    def __init__(self, engine, key):
        self._e, self._key = engine, key

    @property
    def data(self):
        return self._e.view(self._key)

    @data.setter
    def data(self, value):
        v = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value, dtype=np.float32))
        self._e.view(self._key).copy_(v)
        self._e.weights_changed()

    array = data

    @property
    def grad(self):
        return self._e.view(self._key, grad=True)

    @property
    def shape(self):
        return self._e.info[self._key][2]


class Link:
    """chainer.Link stand-in: named params (+ persistents), optional child links."""

    def __init__(self, engine, prefix, params=(), persistents=(), children=()):
        object.__setattr__(self, "_e", engine)
        object.__setattr__(self, "_prefix", prefix)
        object.__setattr__(self, "_params", tuple(params))
        object.__setattr__(self, "_persistents", tuple(persistents))
        object.__setattr__(self, "_children", tuple(children))
        object.__setattr__(self, "update_enabled", True)
        for p in params:
            object.__setattr__(self, p, Param(engine, f"{prefix}/{p}"))
        for c in children:
            object.__setattr__(self, c.name, c)

    @property
    def name(self):
        return self._prefix.split("/")[-1]

    def param_keys(self):
        keys = [f"{self._prefix}/{p}" for p in self._params]
        for c in self._children:
            keys += c.param_keys()
        return keys

    def persistent_keys(self):
        return [f"{self._prefix}/{p}" for p in self._persistents]

    def disable_update(self):          # nn.py:116
        self.update_enabled = False
        for c in self._children:
            c.disable_update()

    def enable_update(self):
        self.update_enabled = True
        for c in self._children:
            c.enable_update()

    def copy_from(self, other):
        """Transfer parameters (and BN persistents) from another link of the same structure."""
        mine, theirs = self.param_keys(), other.param_keys()
        if len(mine) != len(theirs):
            raise ValueError(f"cannot bind {other._prefix} onto {self._prefix}: different structure")
        for a, b in zip(mine, theirs):
            src = other._e.view(b)
            dst = self._e.view(a)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"shape mismatch binding {b}{tuple(src.shape)} onto {a}{tuple(dst.shape)}")
            dst.copy_(src)
        for a, b in zip(self.persistent_keys(), other.persistent_keys()):
            if a.endswith("/N"):
                i = 0 if "CNN_0" in a else 1
                self._e.bn_N[i] = other._e.bn_N[0 if "CNN_0" in b else 1]
            else:
                self._e.bn_view(a).copy_(other._e.bn_view(b))
        self._e.weights_changed()


class BNLink(Link):
    @property
    def avg_mean(self):
        return self._e.bn_view(f"{self._prefix}/avg_mean")

    @property
    def avg_var(self):
        return self._e.bn_view(f"{self._prefix}/avg_var")

    @property
    def N(self):
        return self._e.bn_N[0 if "CNN_0" in self._prefix else 1]


class LSTMLink(Link):
    """L.LSTM stand-in: ``upward`` (W, b), ``lateral`` (W); h / c are owned by the engine."""

    def __init__(self, engine, prefix):
        up = Link(engine, f"{prefix}/upward", params=("W", "b"))
        lat = Link(engine, f"{prefix}/lateral", params=("W",))
        super().__init__(engine, prefix, children=(up, lat))


class SpeechEncoderDecoder:
    def __init__(self, gpuid, cfg, feat_dim=None):
        """seq2seq.py:23-33.  ``feat_dim``: the reference sizes its first layers lazily at the first
        call (in_channels: null, L.LSTM(None, ...)); here the feature dimension may be given up
        front, else the engine is built at the first encode()."""
        object.__setattr__(self, "_links", {})
        self.gpuid = gpuid
        self.cfg = cfg
        self._engine = None
        self._feat_dim = None
        self.enc_states = None
        self.loss = 0
        r = cfg["rnn_config"]
        self.bi_rnn = r["bi_rnn"]
        self.cnn_bn = cfg["cnn_config"]["bn"]
        self.rnn_ln = r["ln"]
        self.rnn_linear_proj = bool(r.get("linear_proj", False))
        self.n_attn = r.get("n_attn", 1)
        nl = r["enc_layers"]
        self.cnns = [f"CNN_{i}" for i in range(len(cfg["cnn_config"]["cnn_layers"]))]
        self.rnn_enc = [f"L{i}_enc" for i in range(nl)]
        self.rnn_rev_enc = [f"L{i}_rev_enc" for i in range(nl)] if self.bi_rnn else []
        self.rnn_dec = [f"L{i}_dec" for i in range(r["dec_layers"])]
        self._pending_state = {}
        self._pending_frozen = set()
        self._provisional_dim = False
        self._seed = None
        if feat_dim is not None:
            self._build(feat_dim)

    @property
    def link_names(self):
        """Names of the child links (what `l in model.__dict__` tests in nn.py:114), known before the lazy build."""
        return ([n for c in self.cnns for n in (c, c + "_bn")] + self.rnn_enc + self.rnn_rev_enc
                + ["attn_Wa", "context", "embed_dec"] + self.rnn_dec + ["out"])

    def disable_update(self, name):
        """`model[name].disable_update()` (nn.py:116) that also works while the model is still unshaped."""
        if name not in self.link_names:
            raise KeyError(name)
        if self._engine is None:
            self._pending_frozen.add(name)
        else:
            self._links[name].disable_update()

    # ---- construction ---------------------------------------------------------------------------
    def _build(self, feat_dim, provisional=False):
        dev = self.gpuid if self.gpuid is not None and self.gpuid >= 0 else 0
        old = self._engine
        frozen = {n for n, l in self._links.items() if not l.update_enabled} | set(self._pending_frozen)
        e = Engine(self.cfg, feat_dim, dev)
        self._engine, self._feat_dim, self._provisional_dim = e, feat_dim, provisional
        links = {}
        for c in self.cnns:
            links[c] = Link(e, c, params=("W",))
            links[c + "_bn"] = BNLink(e, c + "_bn", params=("gamma", "beta"), persistents=("avg_mean", "avg_var", "N"))
        for n in self.rnn_enc + self.rnn_rev_enc + self.rnn_dec:
            links[n] = LSTMLink(e, n)
        links["attn_Wa"] = Link(e, "attn_Wa", params=("W", "b"))
        links["context"] = Link(e, "context", params=("W", "b"))
        links["embed_dec"] = Link(e, "embed_dec", params=("W",))
        links["out"] = Link(e, "out", params=("W", "b"))
        for k, v in links.items():
            self._links[k] = v
            object.__setattr__(self, k, v)
        self.mask_pad_id = torch.ones(e.V, dtype=torch.float32, device=e.device)   # seq2seq.py:152-156
        self.mask_pad_id[0] = 0
        if old is not None:
            # re-shaped for another feature dimension with the same number of CNN frequency positions: no parameter shape
            # depends on D itself (CNN_0/W is (C0,1,kh,kw); the RNN input width is C1*F'), so everything carries over
            assert old.nfloats == e.nfloats, "parameter layout changed with the feature dimension"
            e.params.copy_(old.params)
            e.bn_state.copy_(old.bn_state)
            e.bn_N = list(old.bn_N)
            e.weights_changed()
        else:
            self.init_params(seed=self._seed)
        for n in frozen:
            links[n].disable_update()
        self._pending_frozen = set()

    def _require(self, X=None):
        if self._engine is None:
            if X is None:
                raise RuntimeError("model parameters are shaped at the first encode() (lazy in_channels); "
                                   "pass feat_dim= or call encode first")
            self._build(int(X.shape[-1]))
        elif X is not None and self._provisional_dim and int(X.shape[-1]) != self._feat_dim:
            # shaped from a checkpoint, which fixes only the number of frequency positions F' (serializers.load_npz)
            D = int(X.shape[-1])
            l0 = self.cfg["cnn_config"]["cnn_layers"][0]
            fp = (D + 2 * l0["pad"][1] - l0["ksize"][1]) // l0["stride"][1] + 1
            if fp * self.cfg["cnn_config"]["cnn_layers"][-1]["out_channels"] != self._engine.info["L0_enc/upward/W"][2][1]:
                raise ValueError(f"feature dimension {D} does not fit the loaded parameters "
                                 f"(L0_enc input width {self._engine.info['L0_enc/upward/W'][2][1]})")
            self._build(D)
        if X is not None:
            self._provisional_dim = False
        return self._engine

    def init_params(self, seed=None):
        """Chainer-default initialisers (Appendix A.1-A.4), drawn on the host."""
        e = self._engine
        rng = np.random.default_rng(seed)
        for key, (_, _, shp) in e.info.items():
            if key.startswith("CNN_") and key.endswith("/W"):
                a = rng.normal(0.0, math.sqrt(2.0 / (shp[1] * shp[2] * shp[3])), shp)
            elif key.endswith("/gamma"):
                a = np.ones(shp)
            elif key == "embed_dec/W":
                a = rng.normal(0.0, 1.0, shp)
            elif key.endswith("/W"):
                a = rng.normal(0.0, math.sqrt(1.0 / shp[1]), shp)
            elif key.endswith("upward/b"):
                a = np.zeros(shp)
                a[2::4] = 1.0
            else:
                a = np.zeros(shp)
            e.view(key).copy_(torch.as_tensor(a, dtype=torch.float32))
        e.weights_changed()

    # ---- chainer.Chain protocol -----------------------------------------------------------------
    def __getitem__(self, name):
        return self._links[name]

    def __setattr__(self, name, value):
        links = self.__dict__.get("_links")
        if links is not None and name in links and isinstance(value, Link):
            links[name].copy_from(value)          # re-binding transfers the parameters
            return
        object.__setattr__(self, name, value)

    def to_gpu(self, gpuid=None):
        return self

    def namedparams(self):
        e = self._require()
        for key in e.info:
            yield "/" + key, Param(e, key)

    def params(self):
        for _, p in self.namedparams():
            yield p

    def cleargrads(self):
        """nn.py:180.  Gradients are fully overwritten by the next backward; nothing to do eagerly."""
        return None

    def load_state(self, arrays: dict):
        """name -> numpy array for params and BN persistents (serializers.load_npz)."""
        e = self._require()
        for key, (_, _, shp) in e.info.items():
            a = np.asarray(arrays[key], dtype=np.float32)
            if tuple(a.shape) != tuple(shp):
                raise ValueError(f"{key}: checkpoint shape {a.shape} != model shape {shp}")
            e.view(key).copy_(torch.as_tensor(a))
        for i, c in enumerate(self.cnns):
            for p in ("avg_mean", "avg_var"):
                k = f"{c}_bn/{p}"
                if k in arrays:
                    e.bn_view(k).copy_(torch.as_tensor(np.asarray(arrays[k], dtype=np.float32)))
            if f"{c}_bn/N" in arrays:
                e.bn_N[i] = int(arrays[f"{c}_bn/N"])
        e.weights_changed()

    def state_arrays(self):
        e = self._require()
        out = {k: e.view(k).detach().cpu().numpy().copy() for k in e.info}
        for i, c in enumerate(self.cnns):
            for p in ("avg_mean", "avg_var"):
                out[f"{c}_bn/{p}"] = e.bn_view(f"{c}_bn/{p}").detach().cpu().numpy().copy()
            out[f"{c}_bn/N"] = np.asarray(e.bn_N[i])
        return out

    # ---- hot path: same names as seq2seq.py -----------------------------------------------------------
    def reset_rnn_state(self):
        self.loss = 0

    def encode(self, X, add_noise=0):
        """seq2seq.py:293-315.  Sets self.enc_states (B,T',H)."""
        X = X.data if isinstance(X, Variable) else X
        e = self._require(X)
        train = bool(config.train)
        e.encode(X, train=train, noise_sigma=float(add_noise) if (add_noise and train) else 0.0)
        self.enc_states = Variable(e.enc_states())

    forward_enc = encode          # legacy name (enc_dec.py:517)

    def init_decoder_state(self):
        self._engine.init_decoder_state()

    def get_encoder_states(self):
        """seq2seq.py:529-547 -> {"c": [L x (B,H)], "h": [...]}"""
        st = self._engine.get_encoder_states()
        return {"c": [Variable(st[l, 0]) for l in range(st.shape[0])], "h": [Variable(st[l, 1]) for l in range(st.shape[0])]}

    def get_decoder_states(self):
        st = self._engine.get_decoder_states()
        return {"c": [Variable(st[l, 0]) for l in range(st.shape[0])], "h": [Variable(st[l, 1]) for l in range(st.shape[0])]}

    def set_decoder_states(self, rnn_states):
        def t(v):
            return v.data if isinstance(v, Variable) else torch.as_tensor(np.asarray(v))
        nl = len(rnn_states["c"])
        st = torch.stack([torch.stack((t(rnn_states["c"][l]).to(self._engine.device), t(rnn_states["h"][l]).to(self._engine.device)))
                          for l in range(nl)])
        self._engine.set_decoder_states(st)

    def decode_step(self, word, ht):
        """seq2seq.py:361-396 -> (logits (B,V), ht (B,A), alphas (B,T',1)).  Eval-mode step on the
        decoder state held by the engine (training goes through forward_loss)."""
        word = word.data if isinstance(word, Variable) else word
        ht = ht.data if isinstance(ht, Variable) else ht
        logits, ht_out, alphas = self._engine.decode_step(word, ht)
        return Variable(logits), Variable(ht_out), Variable(alphas.unsqueeze(2))

    forward_dec = decode_step

    def compute_context_vector(self, dec_h, attn_Wa=None):
        """seq2seq.py:336-358 -> (cv (B,H), alphas (B,T',1)) over the encoder states of the last encode(): q = attn_Wa(dec_h),
        scores = enc_states . q, softmax over T' with NO length mask (:344-347), cv = sum_t alpha_t enc_states[:, t].
        `attn_Wa`: an attention link (`model.attn_Wa`, also of another model) or None for this model's own."""
        dec_h = dec_h.data if isinstance(dec_h, Variable) else dec_h
        W = b = None
        if attn_Wa is not None and attn_Wa is not self._links.get("attn_Wa"):
            W, b = attn_Wa.W.data, attn_Wa.b.data
        cv, alphas = self._engine.attention(dec_h, W, b)
        return Variable(cv), Variable(alphas.unsqueeze(2))

    attention = compute_context_vector

    def forward_loss(self, X, y, teach_ratio, random_out=0, add_noise=0, use_true=None):
        """seq2seq.py:399-473.  Scheduled-sampling bits are drawn here with the reference's exact
        ``random.random()`` call order (:431-436) unless ``use_true`` is given."""
        X = X.data if isinstance(X, Variable) else X
        y = y.data if isinstance(y, Variable) else y
        if random_out and random_out > 0:
            raise NotImplementedError("random_out > 0 is not supported (0 in every shipped config)")
        e = self._require(X)
        L = int(y.shape[1])
        if use_true is None:
            use_true = draw_use_true(L, teach_ratio)
        loss = e.forward_loss(X, y, use_true=use_true, noise_sigma=float(add_noise) if add_noise else 0.0)
        self.loss = Variable(loss.reshape(()), backward_fn=e.backward)
        return self.loss

    loss_fn = forward_loss

    def predict(self, X, start_token, end_token, stop_limit):
        """seq2seq.py:475-527 -> (B, n_steps) int32, not truncated at EOS."""
        X = X.data if isinstance(X, Variable) else X
        e = self._require(X)
        return e.predict(X, int(start_token), int(end_token), int(stop_limit))

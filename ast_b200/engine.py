"""Engine: owns the device buffers (flat params / grads / BN state / workspace) and drives the C ABI.

torch is used only as the device allocator, stream provider and (in dist.py) NCCL host; every
arithmetic kernel is in libast_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AstConfig, check, ptr


def config_from_dict(cfg: dict, feat_dim: int) -> AstConfig:
    """model_cfg.json dict (config.py:15-29) -> ast_config."""
    r, cn, dr = cfg["rnn_config"], cfg["cnn_config"], cfg["dropout"]
    layers = cn["cnn_layers"]
    if len(layers) != 2:
        raise ValueError("the B200 path implements the shipped 2-layer CNN front-end")
    if not cn.get("bn", True):
        raise ValueError("cnn_config.bn=false is not supported (every shipped config uses BN)")
    if not r.get("bi_rnn", True):
        raise ValueError("rnn_config.bi_rnn=false is not supported (every shipped config is bidirectional)")
    if r.get("ln", False) or r.get("linear_proj", False) or r.get("n_attn", 1) != 1 or not r.get("feed_attn", True):
        raise ValueError("ln / linear_proj / n_attn>1 / feed_attn=false are outside the hot-path scope (SURVEY 8f-4)")
    for l in layers:
        if l.get("dilate", 1) != 1:
            raise ValueError("dilated convolutions are not supported")
    c = AstConfig()
    c.feat_dim = feat_dim
    for i, l in enumerate(layers):
        c.cnn_cout[i] = l["out_channels"]
        c.cnn_kh[i], c.cnn_kw[i] = l["ksize"]
        c.cnn_sh[i], c.cnn_sw[i] = l["stride"]
        c.cnn_ph[i], c.cnn_pw[i] = l["pad"]
    c.enc_layers, c.dec_layers = r["enc_layers"], r["dec_layers"]
    c.hidden_units, c.embedding_units, c.attn_units = r["hidden_units"], r["embedding_units"], r["attn_units"]
    c.vocab = r["dec_vocab_size"]
    c.drop_embed, c.drop_rnn, c.drop_out = dr["embed"], dr["rnn"], dr["out"]
    return c


class AsyncScalar:
    """Device scalar -> host without stalling the stream: `push(t)` enqueues a D2H copy into a pinned slot and records an
    event right behind it; `pop()` waits for THAT event only.  (`float(t)` one step late is not enough: `.item()` copies
    on the current stream, i.e. behind everything enqueued since - including the whole next training step.)"""

    def __init__(self, device, depth=4):
        self.device = device
        self.buf = torch.zeros(depth, dtype=torch.float32).pin_memory()
        self.ev = [torch.cuda.Event() for _ in range(depth)]
        self.head = self.tail = 0
        self.depth = depth

    def __len__(self):
        return self.head - self.tail

    def push(self, t):
        assert len(self) < self.depth, "AsyncScalar ring full: pop() before pushing more"
        i = self.head % self.depth
        self.buf[i:i + 1].copy_(t.reshape(1), non_blocking=True)
        self.ev[i].record(torch.cuda.current_stream(self.device))
        self.head += 1

    def pop(self):
        assert len(self) > 0
        i = self.tail % self.depth
        self.ev[i].synchronize()
        self.tail += 1
        return float(self.buf[i])


class Engine:
    def __init__(self, cfg: dict, feat_dim: int, device: int = 0):
        if not torch.cuda.is_available():
            raise _lib.AstError("ast_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.feat_dim = feat_dim
        self.cfg = cfg
        self._c = config_from_dict(cfg, feat_dim)
        h = C.c_void_p()
        check(self.lib.ast_create(C.byref(self._c), device, C.byref(h)), "ast_create")
        self.h = h
        self.nfloats = int(self.lib.ast_param_floats(h))
        with torch.cuda.device(self.device):
            self.params = torch.zeros(self.nfloats, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros(self.nfloats, dtype=torch.float32, device=self.device)
            self.bn_state = torch.zeros(int(self.lib.ast_bn_state_floats(h)), dtype=torch.float32, device=self.device)
        self.bn_N = [0, 0]
        check(self.lib.ast_bind_params(h, ptr(self.params), ptr(self.grads), ptr(self.bn_state)), "ast_bind_params")
        self.info = {}
        name = C.create_string_buffer(128)
        off, nd, shp = C.c_longlong(), C.c_int(), (C.c_int * 4)()
        for i in range(self.lib.ast_param_count(h)):
            check(self.lib.ast_param_info(h, i, name, 128, C.byref(off), C.byref(nd), shp), "ast_param_info")
            self.info[name.value.decode()] = (i, int(off.value), tuple(shp[k] for k in range(nd.value)))
        C0, C1 = self._c.cnn_cout[0], self._c.cnn_cout[1]
        self._bn_slices = {"CNN_0_bn/avg_mean": (0, C0), "CNN_0_bn/avg_var": (C0, 2 * C0),
                           "CNN_1_bn/avg_mean": (2 * C0, 2 * C0 + C1), "CNN_1_bn/avg_var": (2 * C0 + C1, 2 * C0 + 2 * C1)}
        for k in ("CNN_0_bn/avg_var", "CNN_1_bn/avg_var"):
            self.bn_view(k).fill_(1.0)
        self.ws = None
        self.ws_shape = (0, 0, 0, 0, 0)
        self.H, self.A, self.V, self.E = self._c.hidden_units, self._c.attn_units, self._c.vocab, self._c.embedding_units
        self.NL = self._c.enc_layers
        self.B = self.T = self.Tp = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ast_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------------
    def view(self, name, grad=False):
        _, off, shp = self.info[name]
        n = int(np.prod(shp))
        return (self.grads if grad else self.params)[off:off + n].view(*shp)

    def bn_view(self, name):
        a, b = self._bn_slices[name]
        return self.bn_state[a:b]

    def weights_changed(self):
        check(self.lib.ast_weights_changed(self.h))

    def set_option(self, key, value):
        check(self.lib.ast_set_option(self.h, key.encode(), float(value)), f"set_option({key})")

    def get_option(self, key):
        return self.lib.ast_get_option(self.h, key.encode())

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- workspace ----------------------------------------------------------------------------
    def ensure_workspace(self, B=1, T=16, L=2, N=1, steps=1):
        cur = self.ws_shape
        if all(a <= b for a, b in zip((B, T, L, N, steps), cur)) and self.ws is not None:
            return
        need = tuple(max(a, b) for a, b in zip((B, T, L, N, steps), cur))
        torch.cuda.synchronize(self.device)
        nbytes = int(self.lib.ast_workspace_bytes(self.h, *need))
        self.ws = None
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        check(self.lib.ast_bind_workspace(self.h, ptr(self.ws), nbytes, *need), "ast_bind_workspace")
        self.ws_shape = need

    # ---- hot path -----------------------------------------------------------------------------
    def _as_f32(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=self.device)

    def _as_i32(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32), device=self.device)

    def enc_len(self, T):
        return int(self.lib.ast_enc_len(self.h, T))

    def encode(self, X, train=False, noise=None, noise_sigma=0.0):
        X = self._as_f32(X)
        B, T, D = X.shape
        assert D == self.feat_dim, f"feature dim {D} != model feature dim {self.feat_dim}"
        self.ensure_workspace(B=B, T=T, N=max(B, 16))
        nz = self._as_f32(noise) if noise is not None else None
        check(self.lib.ast_encode(self.h, ptr(X), B, T, int(train), ptr(nz), float(noise_sigma), self.stream()), "ast_encode")
        if train:
            self.bn_N = [n + 1 for n in self.bn_N]
        self.B, self.T, self.Tp = B, T, self.enc_len(T)
        self._keep = (X, nz)

    def enc_states(self):
        out = torch.empty(self.B, self.Tp, self.H, dtype=torch.float32, device=self.device)
        check(self.lib.ast_get_enc_states(self.h, ptr(out), self.stream()), "ast_get_enc_states")
        return out

    def forward_loss(self, X, y, use_true=None, noise=None, noise_sigma=0.0):
        X = self._as_f32(X)
        y = self._as_i32(y)
        B, T, D = X.shape
        L = y.shape[1]
        assert D == self.feat_dim and y.shape[0] == B
        self.ensure_workspace(B=B, T=T, L=L)
        ut = None
        if use_true is not None:
            if isinstance(use_true, torch.Tensor):
                ut = use_true.to(device=self.device, dtype=torch.uint8).contiguous()
            else:
                ut = torch.as_tensor(np.asarray(use_true, dtype=np.uint8), device=self.device)
            assert ut.numel() == L - 1
        nz = self._as_f32(noise) if noise is not None else None
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        check(self.lib.ast_forward_loss(self.h, ptr(X), ptr(y), B, T, L, ptr(ut), ptr(nz), float(noise_sigma), ptr(loss),
                                        self.stream()), "ast_forward_loss")
        self.bn_N = [n + 1 for n in self.bn_N]
        self.B, self.T, self.Tp, self.L = B, T, self.enc_len(T), L
        self._keep = (X, y, ut, nz)
        return loss

    def backward(self):
        check(self.lib.ast_backward(self.h, self.stream()), "ast_backward")

    def grad_buckets(self):
        """[(offset, count)] in floats: contiguous ranges of ``grads`` in the order backward completes them."""
        out = []
        for i in range(self.lib.ast_grad_bucket_count(self.h)):
            off, cnt = C.c_longlong(), C.c_longlong()
            check(self.lib.ast_grad_bucket_range(self.h, i, C.byref(off), C.byref(cnt)), "ast_grad_bucket_range")
            out.append((off.value, cnt.value))
        return out

    def grad_bucket_wait(self, bucket, stream):
        """Make ``stream`` (torch.cuda.Stream) wait until bucket ``bucket`` of the last enqueued backward is final."""
        check(self.lib.ast_grad_bucket_wait(self.h, int(bucket), C.c_void_p(stream.cuda_stream)), "ast_grad_bucket_wait")

    def step_argmax(self):
        out = torch.empty(self.L - 1, self.B, dtype=torch.int32, device=self.device)
        check(self.lib.ast_get_step_argmax(self.h, ptr(out), self.stream()))
        return out

    def opt_step(self, m, v, vhat, t, lr, l2, clip, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, frozen=()):
        idx = [self.info[n][0] for n in frozen]
        arr = (C.c_int * max(len(idx), 1))(*idx)
        self.ensure_workspace()
        check(self.lib.ast_opt_step(self.h, ptr(m), ptr(v), ptr(vhat), int(t), lr, l2, clip, beta1, beta2, eps, grad_scale,
                                    arr, len(idx), self.stream()), "ast_opt_step")

    def opt_step_sgd(self, lr, l2, clip, grad_scale=1.0, frozen=()):
        """optimizers.SGD(lr).update() behind the WeightDecay / GradientClipping (/ GradientNoise) hooks (nn.py:91-110)."""
        idx = [self.info[n][0] for n in frozen]
        arr = (C.c_int * max(len(idx), 1))(*idx)
        self.ensure_workspace()
        check(self.lib.ast_opt_step_sgd(self.h, lr, l2, clip, grad_scale, arr, len(idx), self.stream()), "ast_opt_step_sgd")

    def scale_grads(self, weight):
        """grads *= weight, then re-arm the bucket events so a gated all-reduce sees the scaled values."""
        check(self.lib.ast_scale_grads(self.h, float(weight), self.stream()), "ast_scale_grads")
        self.mark_grads_final()

    def mark_grads_final(self):
        check(self.lib.ast_grad_buckets_mark(self.h, self.stream()), "ast_grad_buckets_mark")

    def last_grad_norm(self):
        return float(self.lib.ast_last_grad_norm(self.h, self.stream()))

    # ---- decoder protocol ---------------------------------------------------------------------------
    def init_decoder_state(self, Bd=None):
        Bd = self.B if Bd is None else Bd
        self.ensure_workspace(B=Bd, N=Bd)
        check(self.lib.ast_init_decoder_state(self.h, Bd, self.stream()), "ast_init_decoder_state")
        self.Bd = Bd

    def get_encoder_states(self):
        out = torch.empty(self.NL, 2, self.B, self.H, dtype=torch.float32, device=self.device)
        check(self.lib.ast_get_encoder_states(self.h, ptr(out), self.stream()), "ast_get_encoder_states")
        return out

    def get_decoder_states(self):
        out = torch.empty(self.NL, 2, self.Bd, self.H, dtype=torch.float32, device=self.device)
        check(self.lib.ast_get_decoder_states(self.h, ptr(out), self.Bd, self.stream()), "ast_get_decoder_states")
        return out

    def set_decoder_states(self, states):
        states = self._as_f32(states)
        Bd = states.shape[2]
        self.ensure_workspace(B=Bd, N=Bd)
        check(self.lib.ast_set_decoder_states(self.h, ptr(states), Bd, self.stream()), "ast_set_decoder_states")
        self.Bd = Bd
        self._keep_states = states

    def decode_step(self, word, ht):
        word = self._as_i32(word).reshape(-1)
        ht = self._as_f32(ht)
        Bd = word.shape[0]
        logits = torch.empty(Bd, self.V, dtype=torch.float32, device=self.device)
        ht_out = torch.empty(Bd, self.A, dtype=torch.float32, device=self.device)
        alphas = torch.empty(Bd, self.Tp, dtype=torch.float32, device=self.device)
        check(self.lib.ast_decode_step(self.h, ptr(word), ptr(ht), Bd, ptr(logits), ptr(ht_out), ptr(alphas), self.stream()),
              "ast_decode_step")
        self._keep_step = (word, ht)
        return logits, ht_out, alphas

    def attention(self, dec_h, W=None, b=None):
        """compute_context_vector (seq2seq.py:336-358): dec_h (Bd,H) -> (cv (Bd,H), alphas (Bd,T'))."""
        dec_h = self._as_f32(dec_h)
        Bd = dec_h.shape[0]
        assert dec_h.shape[1] == self.H
        self.ensure_workspace(B=Bd, N=Bd)
        if W is not None:
            W, b = self._as_f32(W), self._as_f32(b)
            assert tuple(W.shape) == (self.H, self.H) and tuple(b.shape) == (self.H,)
        cv = torch.empty(Bd, self.H, dtype=torch.float32, device=self.device)
        alphas = torch.empty(Bd, self.Tp, dtype=torch.float32, device=self.device)
        check(self.lib.ast_attention(self.h, ptr(dec_h), Bd, ptr(W), ptr(b), ptr(cv), ptr(alphas), self.stream()), "ast_attention")
        self._keep_attn = (dec_h, W, b)
        return cv, alphas

    def softmax_ce(self, logits, targets):
        """F.softmax_cross_entropy(class_weight=mask_pad_id) + argmax for one step (seq2seq.py:448,468): logits (B,V), targets (B,)
        -> (row_loss (B,) already divided by B and PAD-weighted, argmax (B,) int32).  The kernel works in place: a copy is passed."""
        z = self._as_f32(logits).clone()
        t = self._as_i32(targets).reshape(-1)
        B, Vv = z.shape
        row_loss = torch.empty(B, dtype=torch.float32, device=self.device)
        am = torch.empty(B, dtype=torch.int32, device=self.device)
        check(self.lib.ast_softmax_ce(ptr(z), Vv, ptr(t), B, Vv, ptr(row_loss), ptr(am), self.stream()), "ast_softmax_ce")
        return row_loss, am

    def predict(self, X, start_token, end_token, stop_limit):
        X = self._as_f32(X)
        B, T, _ = X.shape
        self.ensure_workspace(B=B, T=T, N=B, steps=stop_limit)
        preds = torch.zeros(stop_limit, B, dtype=torch.int32, device=self.device)
        n = C.c_int(0)
        check(self.lib.ast_predict(self.h, ptr(X), B, T, start_token, end_token, stop_limit, ptr(preds), C.byref(n),
                                   self.stream()), "ast_predict")
        self.B, self.T, self.Tp = B, T, self.enc_len(T)
        return preds[:n.value].t().contiguous()

    def beam_search(self, X, stop_limit, N, K, go=1, eos=2):
        X = self._as_f32(X)
        assert X.shape[0] == 1, "decode_beam is batch-size-1 (beam.py:111)"
        T = X.shape[1]
        self.ensure_workspace(B=1, T=T, N=N, steps=stop_limit)
        Tp = self.enc_len(T)
        dev = self.device
        hist_parent = torch.zeros(stop_limit, N, dtype=torch.int32, device=dev)
        hist_tok = torch.zeros(stop_limit, N, dtype=torch.int32, device=dev)
        scores = torch.zeros(N, dtype=torch.float32, device=dev)
        alpha_hist = torch.zeros(stop_limit, N, Tp, dtype=torch.float32, device=dev)
        states = torch.zeros(self.NL, 2, N, self.H, dtype=torch.float32, device=dev)
        attn_v = torch.zeros(N, self.A, dtype=torch.float32, device=dev)
        ns, nh = C.c_int(0), C.c_int(0)
        check(self.lib.ast_beam_search(self.h, ptr(X), T, stop_limit, N, K, go, eos, C.byref(ns), C.byref(nh), ptr(hist_parent),
                                       ptr(hist_tok), ptr(scores), ptr(alpha_hist), ptr(states), ptr(attn_v), self.stream()),
              "ast_beam_search")
        self.B, self.T, self.Tp = 1, T, Tp
        return dict(n_steps=ns.value, n_hyps=nh.value, hist_parent=hist_parent, hist_tok=hist_tok, scores=scores,
                    alpha_hist=alpha_hist, states=states, attn_v=attn_v)

    def beam_search_batch(self, Xs, stop_limit, N, K, go=1, eos=2):
        """decode_beam for up to 32 utterances in lock-step (ast_beam_search_batch): Xs = list of (1, T_g, D) / (T_g, D) arrays or
        tensors of ANY lengths -> list of per-utterance result dicts in the same order, each with the fields of `beam_search`
        (so nn.beam_result_to_entries applies).  Utterances are sorted by length internally so equal lengths share an encoder
        batch; every utterance's hypotheses, scores and attention history equal those of `beam_search` on it alone."""
        G = len(Xs)
        assert 1 <= G <= 32, "1..32 utterances per call"
        dev = self.device
        if all(not isinstance(x, torch.Tensor) for x in Xs):
            # host features: one pinned staging buffer, ONE host-to-device copy for the whole group (32 pageable copies + a
            # device concatenation were a millisecond of the call)
            arrs = [np.asarray(x, dtype=np.float32).reshape(-1, np.shape(x)[-1]) for x in Xs]
            assert all(a.shape[1] == self.feat_dim for a in arrs)
            order = sorted(range(G), key=lambda i: arrs[i].shape[0])
            lens = [int(arrs[i].shape[0]) for i in order]
            stage = self._pinned("beam_X", (sum(lens), self.feat_dim), torch.float32)
            np.concatenate([arrs[i] for i in order], axis=0, out=stage.numpy())
            X = stage.to(dev, non_blocking=True)
        else:
            feats = []
            for x in Xs:
                t = self._as_f32(x)
                t = t.reshape(-1, t.shape[-1])
                assert t.shape[1] == self.feat_dim
                feats.append(t)
            order = sorted(range(G), key=lambda i: feats[i].shape[0])
            lens = [int(feats[i].shape[0]) for i in order]
            X = torch.cat([feats[i] for i in order], dim=0).contiguous()
        run = max(sum(1 for _ in grp) for _, grp in __import__("itertools").groupby(lens))
        Tmax = max(lens)
        self.ensure_workspace(B=min(run, 32), T=Tmax, N=G * N, steps=stop_limit)
        Tp_ld = self.enc_len(Tmax)
        hist_parent = torch.zeros(G, stop_limit, N, dtype=torch.int32, device=dev)
        hist_tok = torch.zeros(G, stop_limit, N, dtype=torch.int32, device=dev)
        scores = torch.zeros(G, N, dtype=torch.float32, device=dev)
        alpha_hist = torch.zeros(G, stop_limit, N, Tp_ld, dtype=torch.float32, device=dev)
        states = torch.zeros(self.NL, 2, G * N, self.H, dtype=torch.float32, device=dev)
        attn_v = torch.zeros(G * N, self.A, dtype=torch.float32, device=dev)
        IntG = C.c_int * G
        c_lens, ns, nh, tp = IntG(*lens), IntG(), IntG(), IntG()
        check(self.lib.ast_beam_search_batch(self.h, ptr(X), c_lens, G, int(stop_limit), int(N), int(K), go, eos, ns, nh, tp,
                                             ptr(hist_parent), ptr(hist_tok), ptr(scores), ptr(alpha_hist), Tp_ld, ptr(states),
                                             ptr(attn_v), self.stream()), "ast_beam_search_batch")
        self._keep = (X,)
        # the group's buffers, shared by the per-utterance result dicts: nn.beam_result_to_entries copies them to the host once per
        # group (pinned) and backtracks all utterances together
        shared = dict(engine=self, hist_parent=hist_parent, hist_tok=hist_tok, scores=scores, alpha_hist=alpha_hist,
                      n_steps=[int(v) for v in ns], n_hyps=[int(v) for v in nh], tp=[int(v) for v in tp], host=None)
        out = [None] * G
        for j, i in enumerate(order):
            out[i] = dict(n_steps=ns[j], n_hyps=nh[j], hist_parent=hist_parent[j], hist_tok=hist_tok[j], scores=scores[j],
                          alpha_hist=alpha_hist[j][:, :, :tp[j]], states=states[:, :, j * N:(j + 1) * N, :],
                          attn_v=attn_v[j * N:(j + 1) * N], _group=shared, _slot=j)
        return out

    def _pinned(self, key, shape, dtype):
        """Cached page-locked staging tensor (cudaHostAlloc costs milliseconds: one buffer per use, grown on demand)."""
        cache = self.__dict__.setdefault("_pin_cache", {})
        n = int(np.prod(shape))
        buf = cache.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
            cache[key] = buf
        return buf[:n].view(*shape)

    def fetch_host(self, tensors):
        """Device tensors -> numpy VIEWS of cached pinned buffers: all copies asynchronous, one synchronisation.  The views are
        valid until the next fetch_host (``self.fetch_gen`` counts them: a holder compares the value it saw)."""
        outs = []
        for i, t in enumerate(tensors):
            h = self._pinned(f"fetch{i}", tuple(t.shape), t.dtype)
            h.copy_(t, non_blocking=True)
            outs.append(h)
        torch.cuda.current_stream(self.device).synchronize()
        self.fetch_gen = getattr(self, "fetch_gen", 0) + 1
        return [h.numpy() for h in outs]

    def stage_times(self):
        """[(stage, ms)] of the last forward_loss + backward (needs set_option('stage_timing', 1)); synchronises."""
        ms = (C.c_float * 16)()
        names = C.create_string_buffer(16 * 48)
        n = self.lib.ast_stage_times(self.h, ms, names, 48, 16)
        if n < 0:
            raise RuntimeError(self.lib.ast_last_error().decode())
        return [(names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode(), float(ms[i])) for i in range(n)]

    def debug_fetch(self, name):
        n = C.c_longlong(0)
        check(self.lib.ast_debug_fetch(self.h, name.encode(), None, 0, C.byref(n), self.stream()), "ast_debug_fetch")
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        check(self.lib.ast_debug_fetch(self.h, name.encode(), ptr(out), n.value, C.byref(n), self.stream()), "ast_debug_fetch")
        return out

"""Drop-in for the reference runtime nn.py::NN (model build + resume, optimizer + hooks, train_epoch,
greedy predict, beam search) over the CUDA engine.  Same attributes and method names as
/root/reference/nn.py so train.py / beam.py / copy_params.py keep working.
"""
import contextlib
import os
import random

import numpy as np
import torch

from . import serializers
from .config import Config
from .dataloader import FisherDataLoader, GlobalPhoneDataLoader
from .seq2seq import SpeechEncoderDecoder, Variable, config as _train_config
from .symbols import SYMBOLS

_ADAM = 0
_SGD = 1


@contextlib.contextmanager
def using_config(name, value):
    """chainer.using_config('train', flag) stand-in (nn.py:174,216)."""
    assert name == "train"
    old = _train_config.train
    _train_config.train = value
    try:
        yield
    finally:
        _train_config.train = old


# ---- optimizer + hooks (nn.py:81-119; Appendix A.10) --------------------------------------------------
class WeightDecay:
    name = "WeightDecay"

    def __init__(self, rate):
        self.rate = rate


class GradientClipping:
    name = "GradientClipping"

    def __init__(self, threshold):
        self.threshold = threshold


class GradientNoise:
    name = "GradientNoise"

    def __init__(self, eta):
        self.eta = eta


class _GradientMethod:
    """Shared hook handling of the two update rules (chainer.optimizer.GradientMethod): hooks run before the update in the
    order they were added (WeightDecay, GradientClipping, GradientNoise: nn.py:98-110); ``t`` counts updates."""

    def __init__(self):
        self.t = 0
        self.l2 = 0.0
        self.clip = 0.0
        self.noise_eta = 0.0
        self.grad_scale = 1.0
        self.target = None
        self.pre_update = None        # e.g. the data-parallel all-reduce

    def setup(self, model):
        self.target = model
        return self

    def add_hook(self, hook):
        if isinstance(hook, WeightDecay):
            self.l2 = hook.rate
        elif isinstance(hook, GradientClipping):
            self.clip = float(hook.threshold)
        elif isinstance(hook, GradientNoise):
            self.noise_eta = float(hook.eta)
        else:
            raise TypeError(f"unknown hook {hook!r}")

    def frozen(self):
        out = []
        for name, link in self.target._links.items():
            if not link.update_enabled:
                out += link.param_keys()
        return out

    def _pre(self, e):
        if self.pre_update is not None:
            self.pre_update()
        # chainer.optimizer_hooks.GradientNoise: N(0, eta / (1 + t)^0.55) per element, t = updates made so far
        sigma = (self.noise_eta / (1.0 + self.t) ** 0.55) ** 0.5 if self.noise_eta > 0 else 0.0
        if sigma != getattr(e, "_grad_noise_sigma", 0.0):          # nothing to tell the library on the usual (noise-free) path
            e.set_option("grad_noise_sigma", sigma)
            e._grad_noise_sigma = sigma
        self.t += 1


class Adam(_GradientMethod):
    """optimizers.Adam(alpha, beta1, beta2, eps, amsgrad=True) as one fused multi-tensor kernel over the
    flat parameter buffer: WeightDecay -> GradientClipping (global L2 norm, computed on device, no host
    sync) -> GradientNoise -> AMSGrad.  ``grad_scale`` is 1/world_size under data parallelism."""

    def __init__(self, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, amsgrad=True):
        if not amsgrad:
            raise NotImplementedError("the reference always sets amsgrad=True (nn.py:89)")
        super().__init__()
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self._m = self._v = self._vhat = None

    def _state(self, e):
        if self._m is None:
            self._m = torch.zeros_like(e.params)
            self._v = torch.zeros_like(e.params)
            self._vhat = torch.zeros_like(e.params)

    def update(self):
        e = self.target._require()
        self._state(e)
        self._pre(e)
        e.opt_step(self._m, self._v, self._vhat, self.t, self.alpha, self.l2, self.clip, self.beta1, self.beta2, self.eps,
                   self.grad_scale, self.frozen())


class SGD(_GradientMethod):
    """optimizers.SGD(lr) (nn.py:91-93, optimizer.type = 1): p -= lr * g behind the same hooks, one fused pass."""

    def __init__(self, lr=0.01):
        super().__init__()
        self.lr = lr

    def update(self):
        e = self.target._require()
        self._pre(e)
        e.opt_step_sgd(self.lr, self.l2, self.clip, self.grad_scale, self.frozen())


def _backtrack(hp, hk, ns):
    """Parent-pointer walk for a whole group at once.  hp, hk: (G, S, N) int arrays, ns: (G,) steps made per utterance ->
    toks, slots (G, S, N): final hypothesis j of utterance g emitted toks[g, s, j] at step s (-1: none: finished earlier) from
    beam slot slots[g, s, j]."""
    G, S, N = hp.shape
    slot = np.tile(np.arange(N), (G, 1))
    toks = np.full((G, S, N), -1, dtype=np.int64)
    slots = np.zeros((G, S, N), dtype=np.int64)
    gi = np.arange(G)[:, None]
    ns = np.asarray(ns)
    for s in range(int(ns.max()) - 1, -1, -1):
        act = (s < ns)[:, None]
        toks[:, s] = np.where(act, hk[gi, s, slot], -1)
        slots[:, s] = slot
        slot = np.where(act, hp[gi, s, slot], slot)
    return toks, slots


def _entries(toks, hist, sc, ns, nh, r, go_id):
    """One utterance's hypothesis dicts.  ``hist[s, j]`` is the attention vector hypothesis j's ancestor had at step s (already
    gathered along the parent chain); ``attn_history`` is a (tokens, T') array - a view when the hypothesis emitted a token at
    every step - whose indexing, iteration and len() behave like the reference's list of per-step vectors (one list of 175 small
    arrays per hypothesis was most of the conversion time).  The device views of the decoder state are cut with a handful of
    split/unbind calls per utterance, not seven slices per hypothesis."""
    st = r["states"]                                                        # (NL, 2, N, H)
    per_layer = [[part.split(1, 0) for part in layer.unbind(0)] for layer in st.unbind(0)]      # [l][c|h][j] -> (1, H)
    av = r["attn_v"].split(1, 0)
    out = []
    for j in range(nh):
        valid = toks[:, j] >= 0
        full = bool(valid.all())
        out.append({"hyp": [go_id] + (toks[:, j] if full else toks[valid, j]).tolist(), "score": np.float32(sc[j]) if ns > 0 else 0,
                    "dec_state": {"c": [Variable(pl[0][j]) for pl in per_layer],
                                  "h": [Variable(pl[1][j]) for pl in per_layer]},
                    "attn_v": Variable(av[j]),
                    "attn_history": hist[:, j] if full else hist[valid, j]})
    return out


def beam_result_to_entries(r, go_id=SYMBOLS.GO_ID, model=None):
    """Device beam-search buffers -> the reference's list of hypothesis dicts (nn.py:286-294).  Results of a lock-step group
    (Engine.beam_search_batch) are copied to the host (pinned, one synchronisation), backtracked together and their attention
    histories gathered along the parent chains in ONE pass when the first of them is converted."""
    grp = r.get("_group")
    if grp is not None:
        if grp["host"] is None:
            hp, hk, sc, ah = grp["engine"].fetch_host([grp["hist_parent"], grp["hist_tok"], grp["scores"], grp["alpha_hist"]])
            toks, slots = _backtrack(hp, hk, grp["n_steps"])
            G_, S_, N_, Tp_ = ah.shape                                       # row gather: hist[g, s, j] = ah[g, s, slots[g, s, j]]
            rows = (np.arange(G_)[:, None, None] * S_ + np.arange(S_)[None, :, None]) * N_ + slots
            hist = np.take(ah.reshape(G_ * S_ * N_, Tp_), rows.ravel(), axis=0).reshape(G_, S_, N_, Tp_)      # owns its memory
            grp["host"] = (toks, hist, sc.copy())
        toks, hist, sc = grp["host"]
        g = r["_slot"]
        ns, nh = grp["n_steps"][g], grp["n_hyps"][g]
        return _entries(toks[g, :ns], hist[g, :ns, :, :grp["tp"][g]], sc[g], ns, nh, r, go_id)
    ns, nh = r["n_steps"], r["n_hyps"]
    hp = r["hist_parent"][:ns].cpu().numpy()
    hk = r["hist_tok"][:ns].cpu().numpy()
    ah = r["alpha_hist"][:ns].cpu().numpy()
    sc = r["scores"].cpu().numpy()
    if ns == 0:
        return _entries(np.zeros((0, max(nh, 1)), np.int64), ah, sc, ns, nh, r, go_id)
    toks, slots = _backtrack(hp[None], hk[None], [ns])
    S_, N_, Tp_ = ah.shape
    rows = np.arange(S_)[:, None] * N_ + slots[0]
    return _entries(toks[0], np.take(ah.reshape(S_ * N_, Tp_), rows.ravel(), axis=0).reshape(S_, N_, Tp_), sc, ns, nh, r, go_id)


class NN:
    def __init__(self, cfg_path, feat_dim=None, data_loader=None, cfg=None):
        """nn.py:43-79.  ``feat_dim`` / ``data_loader`` / ``cfg`` are optional injection points (synthetic
        benchmarks, tests); with only ``cfg_path`` this behaves like the reference.

        Launched under torchrun (WORLD_SIZE > 1) the same object trains data-parallel (new work, SURVEY 8e / BASELINE config
        4; the reference is single-GPU): one process per GPU, the device is LOCAL_RANK (train_cfg's ``gpuid`` is ignored),
        rank 0's parameters are broadcast, every global batch of batch_size * world utterances is split r::world, and the
        flat gradient is sum-all-reduced over NCCL in buckets overlapped with backward (dist.GradAllReduce)."""
        from . import dist as adist
        self.cfg = cfg if cfg is not None else Config(cfg_path)
        self.model_dir = self.cfg.model["model_dir"]
        self.rank, local_rank, self.world = adist.env_rank()
        if self.world > 1:
            adist.init_process_group()
            self.cfg.train["gpuid"] = local_rank
        self.gpuid = self.cfg.train["gpuid"]
        random.seed(self.cfg.train["seed"])
        # data-parallel batch plans come from their own generator: identical on every rank in every epoch, whatever each
        # rank's global `random` stream has consumed for scheduled sampling meanwhile
        self._plan_rng = random.Random(f"{self.cfg.train['seed']}/plan") if self.world > 1 else None
        if data_loader is not None:
            self.data_loader = data_loader
        elif self.cfg.train["data"].get("dataloader") == "globalphone":
            self.data_loader = GlobalPhoneDataLoader(self.cfg.train["data"], self.model_dir, self.gpuid)
        else:
            self.data_loader = FisherDataLoader(self.cfg.train["data"], self.model_dir, self.gpuid)
        if feat_dim is None:
            feat_dim = getattr(self.data_loader, "feat_dim", None)
        self._feat_dim = feat_dim
        self.get_model()
        self.init_optimizer(self.cfg.train["optimizer"])
        self._allreduce = None
        self.train_log = os.path.join(self.model_dir, "train.log")
        self.dev_log = os.path.join(self.model_dir, "dev.log")

    def init_optimizer(self, opt_cfg):
        """nn.py:81-119"""
        if opt_cfg["type"] == _ADAM:
            self.optimizer = Adam(alpha=opt_cfg["lr"], beta1=0.9, beta2=0.999, eps=1e-08, amsgrad=True)
        else:
            self.optimizer = SGD(lr=opt_cfg["lr"])
        self.optimizer.setup(self.model)
        if opt_cfg["l2"] > 0:
            self.optimizer.add_hook(WeightDecay(opt_cfg["l2"]))
        self.optimizer.add_hook(GradientClipping(threshold=opt_cfg["grad_clip"]))
        if opt_cfg["grad_noise_eta"] > 0:
            self.optimizer.add_hook(GradientNoise(eta=opt_cfg["grad_noise_eta"]))
        for l in opt_cfg["freeze"]:
            if l in self.model.link_names:          # known before the lazy build (`l in self.model.__dict__`, nn.py:114)
                print("freezing: {0:s}".format(l))
                self.model.disable_update(l)
            else:
                print("layer {0:s} not in model".format(l))

    def get_model(self):
        """nn.py:122-155: build, then resume from the newest seq2seq_<N>.model in the model dir."""
        self.model_fname = os.path.join(self.model_dir, "seq2seq.model")
        self.model = SpeechEncoderDecoder(self.gpuid, self.cfg.model, feat_dim=self._feat_dim)
        self.model.to_gpu(self.gpuid)
        self.max_epoch = 0
        model_fil = self.model_fname
        d = os.path.dirname(model_fil)
        stem = os.path.basename(model_fil).replace(".model", "")
        model_files = [f for f in os.listdir(d) if stem in f] if os.path.isdir(d) else []
        if len(model_files) > 0:
            max_model_fil = max(model_files, key=lambda s: int(s.split("_")[-1].split(".")[0]))
            max_model_fil = os.path.join(d, max_model_fil)
            print("model found = \n{0:s}".format(max_model_fil))
            serializers.load_npz(max_model_fil, self.model)
            self.max_epoch = int(max_model_fil.split("_")[-1].split(".")[0])
        else:
            print("model not found")

    def _dp_setup(self, engine):
        """First data-parallel step: same start on every rank, then the bucketed all-reduce hook (dist.GradAllReduce)."""
        from . import dist as adist
        adist.broadcast_params_(engine)
        self._allreduce = adist.GradAllReduce(engine, self.optimizer, self.world)

    def train_epoch(self, set_key, max_batches=None):
        """nn.py:158-200.  The per-batch loss read-back is one step late (pinned, asynchronous) so the
        training loop never blocks on the device; the returned average is the same quantity."""
        from .engine import AsyncScalar
        from . import dist as adist
        total_loss, n_batches = 0.0, 0
        batch_size = self.cfg.train["batch_size"]
        random_out = self.cfg.train["extras"]["random_out"]
        add_noise = self.cfg.train["extras"]["speech_noise"]
        teach_ratio = self.cfg.train["extras"]["teach_ratio"]
        reader = None            # needs the device, i.e. the (possibly lazily built) engine: created after the first forward
        sizes = []
        kw = dict(rank=self.rank, world=self.world, plan_rng=self._plan_rng) if self.world > 1 else {}
        for batch in self.data_loader.get_batch(batch_size, set_key, train=True, labels=True, **kw):
            weight = 1.0
            with using_config("train", True):
                if batch["X"] is None:
                    # empty shard of a data-parallel tail batch: contribute exact zeros to the all-reduce
                    e = self.model._require()
                    if self._allreduce is None:
                        self._dp_setup(e)
                    e.scale_grads(0.0)
                    loss = None
                else:
                    loss = self.model.forward_loss(X=batch["X"], y=batch["y"], teach_ratio=teach_ratio, random_out=random_out,
                                                   add_noise=add_noise)
                    if self.world > 1 and self._allreduce is None:
                        self._dp_setup(self.model._engine)
                    self.model.cleargrads()
                    loss.backward()
                    if self.world > 1:
                        weight = adist.shard_weight(len(batch["y"]), batch["global_batch"], self.world)
                        if weight != 1.0:
                            self.model._engine.scale_grads(weight)
                self.optimizer.update()
            n_batches += 1
            if loss is not None:
                if reader is None:
                    reader = AsyncScalar(self.model._engine.device)
                reader.push(loss.data)             # D2H into pinned memory + event; read back one step late (nn.py:189)
                sizes.append(len(batch["y"]))
                if len(reader) > 1:
                    total_loss += reader.pop() / sizes.pop(0)
            if max_batches is not None and n_batches >= max_batches:
                break
        while reader is not None and len(reader):
            total_loss += reader.pop() / sizes.pop(0)
        avg = total_loss / max(n_batches, 1)
        if self.world > 1:       # the epoch loss every rank logs is the mean over ranks of the per-replica averages
            avg = adist.sum_over_ranks(avg, self.model._engine.device) / self.world
        return avg

    def predict(self, set_key):
        """nn.py:202-233"""
        batch_size = self.cfg.train["batch_size"]
        stop_limit = self.cfg.train["data"]["max_pred"]
        preds = []
        for batch in self.data_loader.get_batch(batch_size, set_key, train=False, labels=False):
            with using_config("train", False):
                p = self.model.predict(batch["X"], SYMBOLS.GO_ID, SYMBOLS.EOS_ID, stop_limit)
                preds.extend(zip(batch["utts"], p.tolist()))
        return preds

    def init_hyp(self):
        """nn.py:235-243"""
        e = self.model._engine
        return {"hyp": [SYMBOLS.GO_ID], "score": 0, "dec_state": self.model.get_encoder_states(),
                "attn_v": Variable(torch.zeros(1, e.A, device=e.device)), "attn_history": []}

    def decode_beam_step(self, decode_entry, beam_width):
        """nn.py:245-297, one hypothesis through the public decode_step protocol (host-driven; the
        production path is decode_beam below, which keeps the whole search on the device)."""
        with using_config("train", False):
            self.model.set_decoder_states(decode_entry["dec_state"])
            e = self.model._engine
            word = torch.full((1,), int(decode_entry["hyp"][-1]), dtype=torch.int32, device=e.device)
            pred_out, ht, alphas = self.model.decode_step(word, decode_entry["attn_v"])
            z = pred_out.data[0]
            m = z.max()
            lp = (z - (m + torch.log(torch.exp(z - m).sum()))).cpu().numpy()
            top = np.argsort(lp, kind="stable")[-beam_width:]
            st = self.model.get_decoder_states()
            out = []
            for pi in top[::-1]:
                out.append({"hyp": decode_entry["hyp"] + [int(pi)],
                            "score": np.float32(np.float32(decode_entry["score"]) + lp[pi]),
                            "dec_state": st, "attn_v": ht,
                            "attn_history": list(decode_entry["attn_history"]) + [alphas.data[0, :, 0].cpu().numpy()]})
            return out

    def decode_beam_batch(self, Xs, stop_limit, N, K):
        """Throughput form of decode_beam: up to 32 utterances (any lengths) searched in lock-step on the device - one pass over
        the decoder weights per step serves every in-flight search.  Returns one n_best list per utterance, each identical to
        decode_beam(X) on that utterance alone (nn.py:299-322)."""
        Xs = [x.data if isinstance(x, Variable) else x for x in Xs]
        with using_config("train", False):
            e = self.model._require(Xs[0])
            res = e.beam_search_batch(Xs, int(stop_limit), int(N), int(K), SYMBOLS.GO_ID, SYMBOLS.EOS_ID)
            self.model.enc_states = None
            return [beam_result_to_entries(r, SYMBOLS.GO_ID) if r["n_steps"] > 0 else [self.init_hyp()] for r in res]

    def decode_beam(self, X, stop_limit, N, K):
        """nn.py:299-322 with the search loop, top-K, pruning and state gather on the device."""
        X = X.data if isinstance(X, Variable) else X
        with using_config("train", False):
            e = self.model._require(X)
            r = e.beam_search(X, int(stop_limit), int(N), int(K), SYMBOLS.GO_ID, SYMBOLS.EOS_ID)
            self.model.enc_states = None
            if r["n_steps"] == 0:
                return [self.init_hyp()]
            return beam_result_to_entries(r, SYMBOLS.GO_ID)

"""TEST INFRASTRUCTURE ONLY - generates tests/golden/ref_*.npz by executing the reference's OWN source.

    python oracle/gen_ref_golden.py            # needs /root/reference (this container only); writes tests/golden/ref_*.npz

`/root/reference/{seq2seq,nn,dataloader,config,eval}.py` are imported UNMODIFIED on top of the Chainer/CuPy stand-in in
oracle/_ref_shim (float64, torch-CPU conv / matmul / autograd).  Every number stored here therefore went through the
reference's control flow: `X[-i]` reverse order (seq2seq.py:219), dropout on every layer output (:198), set_state(c, h)
(:329), scheduled-sampling draws (:431-436), PAD-weighted CE summed over steps (:468-470), hook order + AMSGrad
(nn.py:85-118), the bucketed batch plan and frame zeroing (dataloader.py:83-164), greedy predict (:475-527) and the beam
bookkeeping (nn.py:235-322).  What stays unpinned: the Chainer op semantics themselves, restated in the
stand-in because the library is absent.

The fixtures travel to the GPU box; /root/reference does not.  tests/test_ref_golden.py checks oracle == fixture (CPU) and
tests/test_gpu_parity.py checks CUDA path == fixture (GPU).
"""
import json
import os
import random
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("AST_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")


def _setup_path():
    assert os.path.isdir(REF), f"{REF} not found: the reference only exists in the build container"
    for p in (REPO, REF, os.path.join(HERE, "_ref_shim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


_setup_path()

import chainer                               # noqa: E402  (the stand-in)
import chainer.functions as F                # noqa: E402
from chainer import serializers              # noqa: E402
import nn as ref_nn                          # noqa: E402  (/root/reference/nn.py, unmodified)
from dataloader import SYMBOLS               # noqa: E402  (/root/reference/dataloader.py)

from oracle import ast_oracle as O           # noqa: E402
from oracle import device_rng as R           # noqa: E402
from oracle import synth_corpus as SC        # noqa: E402

assert ref_nn.__file__.startswith(REF) and "_ref_shim" in chainer.__file__

from oracle.ref_golden_common import (DROP_SEED, EOS_BOOST, SAMPLE, V, VOCAB_WORDS, checksum, epoch_corpus_kwargs,   # noqa: E402
                                      golden_params, model_cfg, param_checksum)


def write_checkpoint(exp, P, epoch=0, eos_boost=0.0):
    P = dict(P)
    if eos_boost:
        P["out/b"] = P["out/b"].copy()
        P["out/b"][SYMBOLS.EOS_ID] += eos_boost
    with open(os.path.join(exp, f"seq2seq_{epoch}.model"), "wb") as f:
        np.savez_compressed(f, **P)


def model_params(model):
    return {k[1:]: p.data.copy() for k, p in model.namedparams()}


def model_grads(model):
    return {k[1:]: p.grad.copy() for k, p in model.namedparams()}


def bn_state(model):
    out = {}
    for c in model.cnns:
        bn = model[c + "_bn"]
        out[f"{c}_bn/avg_mean"] = np.asarray(bn.avg_mean).copy()
        out[f"{c}_bn/avg_var"] = np.asarray(bn.avg_var).copy()
        out[f"{c}_bn/N"] = np.asarray(bn.N)
    return out


def pack_tensors(out, prefix, tensors, stride=1):
    """Store each tensor in float32 (every `stride`-th element of the flattened tensor: fixture size) plus float64
    [sum, l2, <t, probe>] over the WHOLE tensor, so float64 agreement of every element can still be asserted."""
    for k, a in tensors.items():
        a = np.asarray(a, dtype=np.float64)
        out[f"{prefix}/{k}"] = a.astype(np.float32) if stride == 1 else a.ravel()[::stride].astype(np.float32)
        out[f"{prefix}_chk/{k}"] = checksum(a)


class Recorder:
    """Wraps functions of the STAND-IN (never of the reference) to expose per-step values the reference does not return."""

    def __init__(self):
        self.logits, self.step_losses, self.noise = [], [], []
        self._ce, self._normal = F.softmax_cross_entropy, np.random.normal

    def __enter__(self):
        def ce(x, t, **kw):
            l = self._ce(x, t, **kw)
            self.logits.append(np.asarray(x.data).copy())
            self.step_losses.append(float(l.data))
            return l

        def normal(*a, **kw):
            z = self._normal(*a, **kw)
            self.noise.append(np.asarray(z).copy())
            return z
        F.softmax_cross_entropy = ce
        np.random.normal = normal
        return self

    def __exit__(self, *a):
        F.softmax_cross_entropy = self._ce
        np.random.normal = self._normal
        return False


def synth_xy(rng, B, T, D, L, lens=None, ylens=None, V=V):
    lens = lens if lens is not None else [T] + [int(v) for v in rng.integers(max(T - 15, 9), T, size=B - 1)]
    X = np.zeros((B, T, D), dtype=np.float32)
    for b, n in enumerate(lens):
        X[b, :n] = rng.standard_normal((n, D)).astype(np.float32)
    ylens = ylens if ylens is not None else [L] + [int(v) for v in rng.integers(3, L + 1, size=B - 1)]
    y = np.zeros((B, L), dtype=np.int32)
    for b, n in enumerate(ylens):
        y[b, :n] = [SYMBOLS.GO_ID] + [int(v) for v in rng.integers(4, V, size=n - 2)] + [SYMBOLS.EOS_ID]   # noqa
    return X, y


def fresh_nn(root, cfg, D, P, dropout=(0.0, 0.0, 0.0), eos_boost=0.0, mc=None, vocab_words=VOCAB_WORDS, **kw):
    """A new experiment directory + checkpoint -> the reference's NN(cfg_path) (resumes from seq2seq_0.model, nn.py:142-152)."""
    if os.path.isdir(root):
        shutil.rmtree(root)
    os.makedirs(root)
    if mc is None:
        mc = SC.small_model_cfg(hidden=128, embed=16, attn=128, c0=8, c1=16, dropout=dropout)
    else:
        mc = json.loads(json.dumps(mc))
        mc["dropout"] = {"embed": dropout[0], "rnn": dropout[1], "out": dropout[2]}
    assert mc["rnn_config"] == {k: v for k, v in cfg["rnn_config"].items() if k != "dec_vocab_size"}
    exp = SC.write_experiment(root, mc, feat_dim=D, vocab_words=vocab_words, **kw)
    write_checkpoint(exp, P, 0, eos_boost)
    n = ref_nn.NN(exp)
    assert n.max_epoch == 0 and n.model.cfg["rnn_config"]["dec_vocab_size"] == vocab_words + 4
    return n


def beam_to_arrays(n_best):
    hyps = [list(map(int, e["hyp"])) for e in n_best]
    return {"beam_hyps": np.asarray([t for h in hyps for t in h], dtype=np.int32),
            "beam_hyp_lens": np.asarray([len(h) for h in hyps], dtype=np.int32),
            "beam_scores": np.asarray([float(e["score"]) for e in n_best], dtype=np.float64),
            "beam_attn_last": np.stack([np.asarray(e["attn_history"][-1], dtype=np.float64) for e in n_best])
            if n_best and n_best[0]["attn_history"] else np.zeros((0,))}


def gen_model_case(name, D, seed, B=3, T=57, L=9, eos_boost=EOS_BOOST):
    """Model-level calls on the reference's SpeechEncoderDecoder (through NN for construction / resume)."""
    cfg = model_cfg()
    P = golden_params(cfg, D, seed)
    rng = np.random.default_rng(seed + 1)
    X, y = synth_xy(rng, B, T, D, L)
    out = {"D": D, "seed": seed, "V": V, "X": X, "y": y, "param_checksum": param_checksum(P)}
    tmp = tempfile.mkdtemp(prefix="ast_ref_")
    try:
        # (i) teacher forcing 1.0, no dropout / noise: loss, per-step loss + logits, enc_states, every gradient, BN stats
        n = fresh_nn(os.path.join(tmp, "a"), cfg, D, P)
        m = n.model
        with Recorder() as rec, chainer.using_config("train", True):
            loss = m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=1.0)
            m.cleargrads()
            loss.backward()
        out["tf_loss"] = float(loss.data)
        out["tf_step_losses"] = np.asarray(rec.step_losses)
        out["tf_logits"] = np.stack(rec.logits)                              # (L-1, B, V) float64
        out["tf_enc_states"] = np.asarray(m.enc_states.data).copy()          # (B, T', H) float64
        pack_tensors(out, "tf_grad", model_grads(m))
        for k, v in bn_state(m).items():
            out["tf_bn/" + k] = v
        # (ii) one optimizer step on those gradients (hooks + AMSGrad, nn.py:85-118,182), then a second full step
        n.optimizer.update()
        pack_tensors(out, "tf_param_after1", model_params(m), stride=SAMPLE)
        out["tf_grad_norm1"] = n.optimizer.last_grad_norm
        with chainer.using_config("train", True):
            loss2 = m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=1.0)
            m.cleargrads()
            loss2.backward()
            n.optimizer.update()
        out["tf_loss2"] = float(loss2.data)
        pack_tensors(out, "tf_param_after2", model_params(m), stride=SAMPLE)

        # (iii) eval mode on a fresh model after exactly ONE training-mode forward (so the BN running statistics are
        # non-trivial): batched greedy predict, then beam search on each utterance cut to its own length
        n = fresh_nn(os.path.join(tmp, "b"), cfg, D, P, eos_boost=eos_boost)
        m = n.model
        with chainer.using_config("train", True):
            m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=1.0)
        with chainer.using_config("train", False):
            pred = m.predict(chainer.Variable(X), SYMBOLS.GO_ID, SYMBOLS.EOS_ID, 12)
        out["greedy"] = np.asarray(pred, dtype=np.int32)
        out["eos_boost"] = eos_boost
        chainer.config.dtype = np.float32                                    # hypotheses "identical at fp32" (north_star)
        try:
            n32 = fresh_nn(os.path.join(tmp, "b32"), cfg, D, P, eos_boost=eos_boost)
            with chainer.using_config("train", True):
                n32.model.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=1.0)
            with chainer.using_config("train", False):
                out["greedy_f32"] = np.asarray(n32.model.predict(chainer.Variable(X), SYMBOLS.GO_ID, SYMBOLS.EOS_ID, 12), dtype=np.int32)
            for (N_, K_, stop) in ((4, 3, 12), (10, 10, 12), (1, 1, 6), (3, 5, 10)):
                nb = n32.decode_beam(chainer.Variable(X[0:1]), stop_limit=stop, N=N_, K=K_)
                for k, v in beam_to_arrays(nb).items():
                    out[f"beam_N{N_}K{K_}/{k}"] = v
        finally:
            chainer.config.dtype = np.float64

        # (iv) scheduled sampling: teach_ratio 0.5, draws from Python's `random` in the reference's order
        n = fresh_nn(os.path.join(tmp, "c"), cfg, D, P)
        m = n.model
        random.seed(4242)
        with Recorder() as rec, chainer.using_config("train", True):
            loss = m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=0.5)
            m.cleargrads()
            loss.backward()
        random.seed(4242)
        out["ss_bits"] = np.asarray([True if not (0 < i < L - 2) else (random.random() < 0.5) for i in range(L - 1)])
        out["ss_seed"] = 4242
        out["ss_loss"] = float(loss.data)
        out["ss_step_losses"] = np.asarray(rec.step_losses)
        out["ss_argmax"] = np.stack([lg.argmax(axis=1) for lg in rec.logits]).astype(np.int32)
        pack_tensors(out, "ss_grad", model_grads(m), stride=SAMPLE)

        # (v) the benchmarked training configuration: dropout .3/.3, speech_noise .25, teach_ratio .8.  Masks = the CUDA
        # library's counter RNG at seed DROP_SEED, first training step (oracle/device_rng.py), injected in F.dropout call order.
        n = fresh_nn(os.path.join(tmp, "d"), cfg, D, P, dropout=(0.3, 0.3, 0.0))
        m = n.model
        Tp = O.cnn_shapes(cfg, T, D)[-1][10]
        masks = R.training_masks(DROP_SEED, 1, B, Tp, L - 1, 64, 128, 16, 3, 0.3, 0.3)
        order = R.reference_call_order(Tp, L - 1, 3, 0.3, 0.3)
        calls = []

        def hook(shape, ratio):
            k = order[len(calls)]
            calls.append(k)
            assert tuple(masks[k].shape) == tuple(shape) and abs(ratio - 0.3) < 1e-12, (k, shape, ratio)
            return masks[k].astype(np.float64)
        F.dropout_hook = hook
        random.seed(778)
        np.random.seed(31337)
        try:
            with Recorder() as rec, chainer.using_config("train", True):
                loss = m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=0.8, add_noise=0.25)
                m.cleargrads()
                loss.backward()
        finally:
            F.dropout_hook = None
        assert len(calls) == len(order) and len(rec.noise) == 1
        random.seed(778)
        out["do_bits"] = np.asarray([True if not (0 < i < L - 2) else (random.random() < 0.8) for i in range(L - 1)])
        out["do_noise"] = rec.noise[0].astype(np.float32)                    # the reference casts to float32 (seq2seq.py:302)
        out["do_seed"] = DROP_SEED
        out["do_loss"] = float(loss.data)
        out["do_step_losses"] = np.asarray(rec.step_losses)
        out["do_enc_states"] = np.asarray(m.enc_states.data).copy()
        pack_tensors(out, "do_grad", model_grads(m), stride=SAMPLE)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: tf_loss {out['tf_loss']:.6f} ss_loss {out['ss_loss']:.6f} do_loss {out['do_loss']:.6f} "
          f"greedy {out['greedy'].shape} -> {path} ({os.path.getsize(path) / 1024:.0f} KB)")


def gen_epoch_case(name, D, seed, globalphone=False, freeze=()):
    """Runtime-level: the reference's NN.train_epoch + NN.predict on a tiny on-disk corpus (bucketed plan, frame zeroing,
    scheduled sampling, hooks + AMSGrad over several steps, greedy eval).  The GPU test runs ast_b200.nn.NN on a corpus
    rebuilt from the same arguments."""
    cfg = model_cfg()
    P = golden_params(cfg, D, seed)
    kw = epoch_corpus_kwargs(seed, globalphone, freeze)
    out = {"D": D, "seed": seed, "V": V, "param_checksum": param_checksum(P), "globalphone": globalphone,
           "freeze": np.asarray(list(freeze), dtype="U32"), "np_seed": 2024}
    tmp = tempfile.mkdtemp(prefix="ast_ref_")
    try:
        np.random.seed(2024)                                      # frame zeroing draws from the global numpy RNG (dataloader.py:88)
        n = fresh_nn(os.path.join(tmp, "e"), cfg, D, P, **kw)
        losses, plan = [], []
        orig = n.model.forward_loss

        def logged(**kwargs):
            l = orig(**kwargs)
            losses.append(float(l.data))
            plan.append((np.asarray(kwargs["X"].data).copy(), np.asarray(kwargs["y"].data).copy()))
            return l
        n.model.forward_loss = logged                             # instance attribute; the reference's source is untouched
        avg = n.train_epoch("fisher_train")
        del n.model.__dict__["forward_loss"]
        out["epoch_avg_loss"] = avg
        out["batch_losses"] = np.asarray(losses)
        out["n_batches"] = len(losses)
        for i, (Xb, yb) in enumerate(plan):
            out[f"batch{i}/X_absum"] = np.abs(Xb).sum(axis=(1, 2))            # identifies utterance order + zeroed frames
            out[f"batch{i}/y"] = yb
        pack_tensors(out, "param_after", model_params(n.model), stride=SAMPLE)
        for k, v in bn_state(n.model).items():
            out["bn_after/" + k] = v
        preds = n.predict("fisher_dev")
        out["pred_utts"] = np.asarray([u for u, _ in preds], dtype="U64")
        out["pred_lens"] = np.asarray([len(p) for _, p in preds], dtype=np.int32)
        out["pred_tokens"] = np.asarray([t for _, p in preds for t in p], dtype=np.int32)
        hyps = n.data_loader.get_hyps(preds)
        out["pred_text"] = np.asarray([" ".join(hyps[u]) for u, _ in preds], dtype="U256")
        # save through the reference's own call (train.py:75) and record the key set
        ck = os.path.join(tmp, "saved.model")
        serializers.save_npz(ck, n.model)
        with np.load(ck) as z:
            out["npz_keys"] = np.asarray(sorted(z.files), dtype="U64")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {out['n_batches']} batches, avg loss {avg:.6f} -> {path} ({os.path.getsize(path) / 1024:.0f} KB)")


def gen_full_case(name, seed=505, B=4, T=120, L=10, stride=97, eos_boost=1.4):
    """The SHIPPED geometry (the reference's own experiments/es_en_20h/model_cfg.json: H=512, E=128, A=512, CNN 128/512,
    dropout .3/.3) with the BPE vocabulary size of data/fisher/fisher.vocab (1098), D=40, and the shipped training extras
    (speech_noise .25, teach_ratio .8): the configuration bench.py measures, on a batch small enough for a fixture.  This
    is the geometry the tcgen05 recurrences and the TMEM-resident decoder kernels are specialised for."""
    with open(os.path.join(REF, "experiments", "es_en_20h", "model_cfg.json")) as f:
        mc = json.load(f)
    D, Vf = 40, 1098
    cfg = json.loads(json.dumps(mc))
    cfg["rnn_config"]["dec_vocab_size"] = Vf
    drop = (mc["dropout"]["embed"], mc["dropout"]["rnn"], mc["dropout"]["out"])
    assert drop == (0.3, 0.3, 0)
    P = golden_params(cfg, D, seed)
    rng = np.random.default_rng(seed + 1)
    X, y = synth_xy(rng, B, T, D, L, V=Vf)
    out = {"D": D, "seed": seed, "V": Vf, "X": X, "y": y, "param_checksum": param_checksum(P), "sample": stride,
           "model_cfg_json": json.dumps(mc, sort_keys=True)}
    tmp = tempfile.mkdtemp(prefix="ast_ref_")
    try:
        n = fresh_nn(os.path.join(tmp, "f"), cfg, D, P, dropout=drop, mc=mc, vocab_words=Vf - 4)
        m = n.model
        Tp = O.cnn_shapes(cfg, T, D)[-1][10]
        masks = R.training_masks(DROP_SEED, 1, B, Tp, L - 1, 256, 512, 128, 3, 0.3, 0.3)
        order = R.reference_call_order(Tp, L - 1, 3, 0.3, 0.3)
        calls = []

        def hook(shape, ratio):
            k = order[len(calls)]
            calls.append(k)
            assert tuple(masks[k].shape) == tuple(shape), (k, shape)
            return masks[k].astype(np.float64)
        F.dropout_hook = hook
        random.seed(779)
        np.random.seed(31338)
        try:
            with Recorder() as rec, chainer.using_config("train", True):
                loss = m.forward_loss(X=chainer.Variable(X), y=chainer.Variable(y), teach_ratio=0.8, add_noise=0.25)
                m.cleargrads()
                loss.backward()
        finally:
            F.dropout_hook = None
        assert len(calls) == len(order) and len(rec.noise) == 1
        random.seed(779)
        out["do_bits"] = np.asarray([True if not (0 < i < L - 2) else (random.random() < 0.8) for i in range(L - 1)])
        out["do_noise"] = rec.noise[0].astype(np.float32)
        out["do_seed"] = DROP_SEED
        out["do_loss"] = float(loss.data)
        out["do_step_losses"] = np.asarray(rec.step_losses)
        out["do_enc_states"] = np.asarray(m.enc_states.data).astype(np.float32)
        out["do_logits"] = np.stack(rec.logits).astype(np.float32)
        pack_tensors(out, "do_grad", model_grads(m), stride=stride)
        n.optimizer.update()
        out["do_grad_norm"] = n.optimizer.last_grad_norm
        pack_tensors(out, "do_param_after1", model_params(m), stride=stride)
        for k, v in bn_state(m).items():
            out["do_bn/" + k] = v
        # decoding at float32 after that one training step: batched greedy + beam-10 that runs 40 steps
        chainer.config.dtype = np.float32
        try:
            n32 = fresh_nn(os.path.join(tmp, "g"), cfg, D, P, dropout=drop, mc=mc, vocab_words=Vf - 4, eos_boost=eos_boost)
            with chainer.using_config("train", False):
                out["greedy_f32"] = np.asarray(n32.model.predict(chainer.Variable(X), SYMBOLS.GO_ID, SYMBOLS.EOS_ID, 20), dtype=np.int32)
                nb = n32.decode_beam(chainer.Variable(X[0:1]), stop_limit=40, N=10, K=10)
            for k, v in beam_to_arrays(nb).items():
                out[f"beam_N10K10/{k}"] = v
            out["eos_boost"] = eos_boost
        finally:
            chainer.config.dtype = np.float64
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: do_loss {out['do_loss']:.6f} greedy {out['greedy_f32'].shape} beam lens {out['beam_N10K10/beam_hyp_lens'].tolist()} "
          f"-> {path} ({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_full_case("ref_full_es_en_20h")
    gen_model_case("ref_model_d13", D=13, seed=101)
    gen_model_case("ref_model_d40", D=40, seed=202)
    gen_epoch_case("ref_epoch_fisher_d13", D=13, seed=303)
    gen_epoch_case("ref_epoch_gp_d40_freeze", D=40, seed=404, globalphone=True, freeze=("L1_enc", "L0_rev_enc", "L2_rev_enc", "embed_dec"))

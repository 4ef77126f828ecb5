"""Oracle pinned to the reference's OWN source.

tests/golden/ref_*.npz were produced by oracle/gen_ref_golden.py, which imports /root/reference/{seq2seq,nn,dataloader,
config,eval}.py UNMODIFIED over the Chainer/CuPy stand-in of oracle/_ref_shim (float64) and records what the reference
computes.  Here the numpy oracle (and the repo's host-side loader / scheduled-sampling / beam logic) must reproduce those
numbers: loss, per-step losses, logits, encoder states, every gradient, two optimizer steps, BN running statistics, greedy
tokens, beam hypotheses and scores, and a whole bucketed training epoch + dev prediction through the on-disk formats.
The CUDA path is compared with the same fixtures in tests/test_gpu_parity.py.
"""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import ast_oracle as O
from oracle import device_rng as R
from oracle import synth_corpus as SC
import ref_golden_util as G

MODEL_CASES = ["ref_model_d13", "ref_model_d40"]
EPOCH_CASES = ["ref_epoch_fisher_d13", "ref_epoch_gp_d40_freeze"]


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_reference_teacher_forced_step(name):
    cfg, D, P, z = G.load_model_case(name)
    om = O.OracleModel(cfg, P, dtype=np.float64)
    loss = float(om.forward_loss(z["X"], z["y"]))
    assert abs(loss - float(z["tf_loss"])) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(om.step_losses, z["tf_step_losses"], rtol=1e-11, atol=0)
    np.testing.assert_allclose(om.enc_states, z["tf_enc_states"], rtol=0, atol=1e-12)
    logits = np.stack([c[-2] @ om.p["out/W"].T + om.p["out/b"] for c in om._dec_cache])
    np.testing.assert_allclose(logits, z["tf_logits"], rtol=0, atol=1e-11)
    g = om.backward()
    G.assert_tensors_match(z, "tf_grad", g, rel=1e-9)
    for k in ("CNN_0_bn/avg_mean", "CNN_0_bn/avg_var", "CNN_1_bn/avg_mean", "CNN_1_bn/avg_var"):
        np.testing.assert_allclose(om.p[k], z["tf_bn/" + k], rtol=1e-12, atol=1e-14)
        assert int(om.p[k.rsplit("/", 1)[0] + "/N"]) == int(z["tf_bn/" + k.rsplit("/", 1)[0] + "/N"]) == 1


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_reference_two_optimizer_steps(name):
    """nn.py:85-118,182: WeightDecay -> GradientClipping -> AMSGrad through chainer.optimizers.Adam(amsgrad=True)."""
    cfg, D, P, z = G.load_model_case(name)
    om = O.OracleModel(cfg, P, dtype=np.float64)
    opt = O.OracleAMSGrad(om.p, lr=1e-3, l2=1e-4, grad_clip=2.0)
    om.forward_loss(z["X"], z["y"])
    opt.update(om.p, om.backward())
    assert abs(opt.last_norm - float(z["tf_grad_norm1"])) <= 1e-10 * opt.last_norm
    G.assert_tensors_match(z, "tf_param_after1", {k: v for k, v in om.p.items() if k in om.grads}, rel=1e-10)
    loss2 = float(om.forward_loss(z["X"], z["y"]))
    assert abs(loss2 - float(z["tf_loss2"])) <= 1e-10 * abs(loss2)
    opt.update(om.p, om.backward())
    G.assert_tensors_match(z, "tf_param_after2", {k: v for k, v in om.p.items() if k in om.grads}, rel=1e-10)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_reference_scheduled_sampling(name):
    """seq2seq.py:431-436: one random.random() draw for each 1 <= i <= L-3; a sampled step feeds the argmax back."""
    from ast_b200.seq2seq import draw_use_true
    cfg, D, P, z = G.load_model_case(name)
    L = z["y"].shape[1]
    random.seed(int(z["ss_seed"]))
    bits = draw_use_true(L, 0.5)                                   # the product's host code, same draw order
    assert bits == [bool(b) for b in z["ss_bits"]] and not all(bits)
    om = O.OracleModel(cfg, P, dtype=np.float64)
    loss = float(om.forward_loss(z["X"], z["y"], tf_bits=bits))
    assert abs(loss - float(z["ss_loss"])) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(om.step_losses, z["ss_step_losses"], rtol=1e-11)
    assert (np.stack(om.step_argmax) == z["ss_argmax"]).all()
    G.assert_tensors_match(z, "ss_grad", om.backward(), rel=1e-9)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_reference_with_dropout_noise_and_sampling(name):
    """The benchmarked training configuration (dropout .3/.3, speech_noise .25, teach_ratio .8): the reference ran with the
    CUDA library's counter-RNG masks injected in its F.dropout call order (seq2seq.py:198,365) and its own
    np.random.normal noise (:300) recorded; the oracle gets the same masks by key."""
    cfg, D, P, z = G.load_model_case(name, dropout=(0.3, 0.3, 0.0))
    B, L = z["y"].shape
    Tp = O.cnn_shapes(cfg, z["X"].shape[1], D)[-1][10]
    om = O.OracleModel(cfg, P, dtype=np.float64)
    om.dropout_masks = {k: v.astype(np.float64) for k, v in
                        R.training_masks(int(z["do_seed"]), 1, B, Tp, L - 1, 64, 128, 16, 3, 0.3, 0.3).items()}
    bits = [bool(b) for b in z["do_bits"]]
    assert not all(bits)
    loss = float(om.forward_loss(z["X"], z["y"], tf_bits=bits, noise=z["do_noise"]))
    assert abs(loss - float(z["do_loss"])) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(om.step_losses, z["do_step_losses"], rtol=1e-11)
    np.testing.assert_allclose(om.enc_states, z["do_enc_states"], rtol=0, atol=1e-12)
    G.assert_tensors_match(z, "do_grad", om.backward(), rel=1e-9)
    # the masks really are Bernoulli(0.7)/0.7 and really changed the result
    m = np.concatenate([v.ravel() for v in om.dropout_masks.values()])
    assert set(np.unique(m.astype(np.float32)).tolist()) == {0.0, float(np.float32(1) / (np.float32(1) - np.float32(0.3)))}
    assert abs((m > 0).mean() - 0.7) < 0.01 and abs(loss - float(z["tf_loss"])) > 1e-3


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_reference_greedy_and_beam(name):
    """seq2seq.py:475-527 and nn.py:235-322 after exactly one training-mode forward (BN running stats non-trivial)."""
    cfg, D, P, z = G.load_model_case(name, eos_boost=True)
    for dt, key in ((np.float64, "greedy"), (np.float32, "greedy_f32")):
        om = O.OracleModel(cfg, P, dtype=dt)
        om.forward_loss(z["X"], z["y"])
        pred = om.predict(z["X"], O.GO_ID, O.EOS_ID, 12)
        assert pred.shape == z[key].shape and (pred == z[key]).all()
    assert (z["greedy"] == O.EOS_ID).any() and z["greedy"].shape[1] > 1        # a row keeps decoding after its EOS
    om = O.OracleModel(cfg, P, dtype=np.float32)
    om.forward_loss(z["X"], z["y"])
    for (N, K, stop) in ((4, 3, 12), (10, 10, 12), (1, 1, 6), (3, 5, 10)):
        nb = om.decode_beam(z["X"][0:1], stop, N, K)
        hyps, scores, attn = G.beam_from_fixture(z, N, K)
        assert [list(map(int, e["hyp"])) for e in nb] == hyps, (N, K)
        np.testing.assert_allclose([float(e["score"]) for e in nb], scores, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(np.stack([e["attn_history"][-1] for e in nb]), attn, rtol=0, atol=2e-6)
    lens = z["beam_N10K10/beam_hyp_lens"]
    assert lens.min() < lens.max()                                            # finished hypotheses were carried along


def test_oracle_matches_reference_at_the_shipped_geometry():
    """experiments/es_en_20h/model_cfg.json as shipped (H=512, dropout .3/.3), V=1098, D=40, speech_noise .25,
    teach_ratio .8 - the configuration bench.py measures - then decoding at float32: greedy and a 40-step beam-10."""
    cfg, D, P, z = G.load_full_case()
    assert (cfg["dropout"]["embed"], cfg["dropout"]["rnn"], cfg["rnn_config"]["hidden_units"]) == (0.3, 0.3, 512)
    B, L = z["y"].shape
    Tp = O.cnn_shapes(cfg, z["X"].shape[1], D)[-1][10]
    om = O.OracleModel(cfg, P, dtype=np.float64)
    om.dropout_masks = {k: v.astype(np.float64) for k, v in
                        R.training_masks(int(z["do_seed"]), 1, B, Tp, L - 1, 256, 512, 128, 3, 0.3, 0.3).items()}
    loss = float(om.forward_loss(z["X"], z["y"], tf_bits=[bool(b) for b in z["do_bits"]], noise=z["do_noise"]))
    assert abs(loss - float(z["do_loss"])) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(om.step_losses, z["do_step_losses"], rtol=1e-11)
    np.testing.assert_allclose(om.enc_states, z["do_enc_states"], rtol=0, atol=2e-7)
    g = om.backward()
    G.assert_tensors_match(z, "do_grad", g, rel=1e-9)
    opt = O.OracleAMSGrad(om.p)
    opt.update(om.p, g)
    assert abs(opt.last_norm - float(z["do_grad_norm"])) <= 1e-10 * opt.last_norm
    G.assert_tensors_match(z, "do_param_after1", {k: v for k, v in om.p.items() if k in g}, rel=1e-10)
    cfg, D, P, z = G.load_full_case(eos_boost=True)
    om = O.OracleModel(cfg, P, dtype=np.float32)
    pred = om.predict(z["X"], O.GO_ID, O.EOS_ID, 20)
    assert pred.shape == z["greedy_f32"].shape and (pred == z["greedy_f32"]).all()
    nb = om.decode_beam(z["X"][0:1], 40, 10, 10)
    hyps, scores, attn = G.beam_from_fixture(z, 10, 10)
    assert [list(map(int, e["hyp"])) for e in nb] == hyps
    assert max(len(h) for h in hyps) == 41 and min(len(h) for h in hyps) == 2      # 40-step hypotheses next to finished ones
    np.testing.assert_allclose([float(e["score"]) for e in nb], scores, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("name", EPOCH_CASES)
def test_host_loader_and_oracle_reproduce_reference_training_epoch(name, tmp_path):
    """train.py:56 -> nn.py:158-200 on an on-disk corpus: the repo's loaders (bucket plan, per-speaker sub-directories / one
    pickle, frame zeroing from numpy's global RNG, labels with UNK and max_pred truncation), its scheduled-sampling draw
    and the oracle's step + optimizer (with a frozen-link list) must land on the reference's parameters after the epoch,
    then nn.py:202-233 + dataloader.py:167-183 on the dev set."""
    from ast_b200.config import Config
    from ast_b200.dataloader import FisherDataLoader, GlobalPhoneDataLoader
    from ast_b200.seq2seq import draw_use_true
    z = G.load(name)
    cfg, D, P = G.epoch_case_model(z)
    exp = G.rebuild_epoch_corpus(z, str(tmp_path))
    c = Config(exp)
    assert c.model["rnn_config"]["dec_vocab_size"] == int(z["V"])
    np.random.seed(int(z["np_seed"]))
    random.seed(c.train["seed"])                                              # nn.py:54
    cls = GlobalPhoneDataLoader if bool(z["globalphone"]) else FisherDataLoader
    loader = cls(c.train["data"], exp, c.train["gpuid"])
    assert loader.feat_dim == D
    freeze = [str(s) for s in z["freeze"]]
    om = O.OracleModel(cfg, P, dtype=np.float64)
    opt = O.OracleAMSGrad(om.p, lr=c.train["optimizer"]["lr"], l2=c.train["optimizer"]["l2"],
                          grad_clip=c.train["optimizer"]["grad_clip"], freeze=freeze)
    losses = []
    for i, (utts, feats, keep, y, max_sp) in enumerate(loader.host_batches(c.train["batch_size"], "fisher_train", True, True)):
        X = O.pad_sequence([f[:max_sp] * k[:, None] for f, k in zip(feats, keep)])
        np.testing.assert_allclose(np.abs(X).sum(axis=(1, 2)), z[f"batch{i}/X_absum"], rtol=1e-6)
        assert (y == z[f"batch{i}/y"]).all()
        bits = draw_use_true(y.shape[1], c.train["extras"]["teach_ratio"])
        losses.append(float(om.forward_loss(X, y, tf_bits=bits)))
        opt.update(om.p, om.backward())
    assert len(losses) == int(z["n_batches"])
    np.testing.assert_allclose(losses, z["batch_losses"], rtol=1e-9)
    sizes = [z[f"batch{i}/y"].shape[0] for i in range(len(losses))]
    avg = sum(l / b for l, b in zip(losses, sizes)) / len(losses)             # nn.py:189-192: loss / len(batch['y'])
    assert abs(avg - float(z["epoch_avg_loss"])) <= 1e-9 * abs(avg)
    G.assert_tensors_match(z, "param_after", {k: v for k, v in om.p.items() if k in om.grads}, rel=1e-9)
    if freeze:
        P0 = G.epoch_case_model(z)[2]
        for k in om.grads:
            frozen = any(k.startswith(f + "/") for f in freeze)
            assert frozen == bool((om.p[k] == P0[k]).all()), k
    for k in ("CNN_0_bn/avg_mean", "CNN_0_bn/avg_var", "CNN_1_bn/avg_mean", "CNN_1_bn/avg_var"):
        np.testing.assert_allclose(om.p[k], z["bn_after/" + k], rtol=1e-10, atol=1e-13)
    # dev-set greedy prediction + ids -> text
    preds = []
    for utts, feats, keep, y, max_sp in loader.host_batches(c.train["batch_size"], "fisher_dev", False, False):
        assert keep is None and y is None
        p = om.predict(O.pad_sequence([f[:max_sp] for f in feats]), O.GO_ID, O.EOS_ID, c.train["data"]["max_pred"])
        preds.extend(zip(utts, p.tolist()))
    assert [u for u, _ in preds] == [str(u) for u in z["pred_utts"]]
    assert [t for _, p in preds for t in p] == z["pred_tokens"].tolist()
    hyps = loader.get_hyps(preds)
    assert [" ".join(hyps[u]) for u, _ in preds] == [str(t) for t in z["pred_text"]]
    # serializer key set (train.py:75 -> chainer.serializers.save_npz)
    want = sorted(list(O.param_shapes(cfg, D)) + list(O.persistent_shapes(cfg)))
    assert [str(k) for k in z["npz_keys"]] == want


def test_device_rng_replica_known_answers():
    """oracle/device_rng.py restates csrc/common.cuh:143-160; these values were computed by hand-tracing the C code
    (uint32 wrap-around) and are checked against the device in tests/test_gpu_parity.py."""
    def h(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF; x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF; x ^= x >> 16
        return x
    seed, stream = 0x123456789ABCDEF0, 17
    for idx in (0, 1, 255, 2 ** 31 + 5):
        a = h((seed & 0xFFFFFFFF) ^ ((stream * 0x9E3779B9) & 0xFFFFFFFF))
        b = h(((seed >> 32) + ((idx * 0x85EBCA6B) & 0xFFFFFFFF) + a) & 0xFFFFFFFF)
        want = h(a ^ b ^ idx)
        assert int(R.rng_u32(seed, stream, np.array([idx], dtype=np.uint64))[0]) == want
    u = R.rng_uniform(seed, stream, np.arange(200000, dtype=np.uint64))
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    m = R.dropout_scale(seed, stream, np.arange(200000, dtype=np.uint64), 0.3)
    assert abs((m == 0).mean() - 0.3) < 5e-3
    order = R.reference_call_order(2, 2, 3, 0.3, 0.3)
    assert order[:6] == [("L0_enc", 0), ("L1_enc", 0), ("L2_enc", 0), ("L0_rev_enc", 0), ("L1_rev_enc", 0), ("L2_rev_enc", 0)]
    assert order[12:16] == [("embed", 0), ("L0_dec", 0), ("L1_dec", 0), ("L2_dec", 0)]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference only exists in the build container")
def test_fixtures_regenerate_from_the_unmodified_reference(tmp_path):
    """Re-runs the generator (reference source + stand-in) into a scratch directory and compares with the committed
    fixtures: they are what the reference computes today, not hand-edited numbers."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, os; sys.argv=['gen']; import runpy; "
            f"sys.path.insert(0, {root!r}); "
            "import oracle.gen_ref_golden as g; "
            f"g.OUT = {str(tmp_path)!r}; "
            "g.gen_model_case('ref_model_d13', D=13, seed=101); "
            "g.gen_epoch_case('ref_epoch_fisher_d13', D=13, seed=303)")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("ref_model_d13", "ref_epoch_fisher_d13"):
        a, b = np.load(os.path.join(str(tmp_path), name + ".npz")), G.load(name)
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            if a[k].dtype.kind in "fc":
                np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-12, err_msg=k)
            else:
                assert (a[k] == b[k]).all(), k

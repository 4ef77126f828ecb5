"""CPU-side tests: the C-ABI library loads and exports every symbol include/ast_b200.h declares, the
host logic (bucketing, batch plan, scheduled-sampling draw order, rerank, config, parameter key set)
mirrors the reference, and nothing in the product path imports the oracle."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ast_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ast_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    from ast_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.ast_abi_version() == 1


def test_create_reports_errors_without_gpu_compute(lib):
    from ast_b200._lib import AstConfig
    c = AstConfig()
    h = ctypes.c_void_p()
    assert lib.ast_create(ctypes.byref(c), 0, ctypes.byref(h)) != 0      # all-zero config is rejected before any CUDA call
    assert b"enc_layers" in lib.ast_last_error()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ast_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"


def test_bucketing_and_batch_plan_follow_the_reference():
    from ast_b200.dataloader import create_buckets, plan_batches
    info = {f"u{i}": {"sp": n} for i, n in enumerate([10, 79, 80, 159, 160, 1599, 1600, 5000, 81, 82])}
    b = create_buckets(info, 20, 80, "sp", 1, "haha")
    assert b["buckets"][0] == ["u0", "u1"] and b["buckets"][1] == ["u2", "u3", "u8", "u9"] and b["buckets"][2] == ["u4"]
    assert b["buckets"][19] == ["u5", "u6", "u7"]                       # min(len // 80, 19)
    # same draws as dataloader.py:125-135: shuffle each bucket in order, then shuffle the batch list
    random.seed("seed-ast-20h")
    plan = plan_batches({"buckets": [list(x) for x in b["buckets"]], "width_b": 80, "num_b": 20}, 2)
    random.seed("seed-ast-20h")
    want = []
    for bi, bucket in enumerate([list(x) for x in b["buckets"]]):
        random.shuffle(bucket)
        for i in range(0, len(bucket), 2):
            want.append((bucket[i:i + 2], (bi + 1) * 80))
    random.shuffle(want)
    assert plan == want
    assert sorted(u for utts, _ in plan for u in utts) == sorted(info)


def test_train_scale_subsamples_with_seed():
    from ast_b200.dataloader import create_buckets
    info = {f"u{i}": {"sp": 100} for i in range(10)}
    b = create_buckets(info, 20, 80, "sp", 2, "haha")
    random.seed("haha")
    assert b["buckets"][1] == random.sample([f"u{i}" for i in range(10)], 5)


def test_scheduled_sampling_bits_use_reference_draw_order():
    # forward_loss draws one random.random() for each 1 <= i <= L-3 (seq2seq.py:431-436)
    L, ratio = 9, 0.8
    random.seed(7)
    want = [True if not (0 < i < L - 2) else (random.random() < ratio) for i in range(L - 1)]
    from oracle.ast_oracle import teacher_forcing_bits
    random.seed(7)
    assert teacher_forcing_bits(L, ratio) == want
    random.seed(7)
    [random.random() for _ in range(L - 3)]
    nxt = random.random()
    random.seed(7)
    teacher_forcing_bits(L, ratio)
    assert random.random() == nxt                                       # exactly L-3 draws consumed


def test_rerank_matches_reference_formula():
    from ast_b200.beam import get_best_hyps, rerank_hypothesis
    hyps = [([1, 4, 5, 6, 2], -6.0, []), ([1, 4, 2], -2.5, [])]
    rr = rerank_hypothesis(hyps, 0.5)
    assert rr[0][0] == [1, 4, 2] and np.isclose(rr[0][1], -2.5 / 1 ** 0.5) and np.isclose(rr[1][1], -6.0 / 3 ** 0.5)
    assert get_best_hyps({"u": hyps}, 2.0)["u"] == [1, 4, 5, 6, 2]


def test_config_reads_experiment_dir(tmp_path):
    import json, pickle
    from ast_b200.config import Config
    vocab = {"bpe_w": {"w2i": {i: i for i in range(37)}, "i2w": {}}}
    pickle.dump(vocab, open(tmp_path / "v.vocab", "wb"))
    json.dump({"rnn_config": {}, "cnn_config": {}, "dropout": {}}, open(tmp_path / "model_cfg.json", "w"))
    json.dump({"data": {"vocab_path": str(tmp_path / "v.vocab"), "dec_key": "bpe_w"}}, open(tmp_path / "train_cfg.json", "w"))
    c = Config(str(tmp_path))
    assert c.model["rnn_config"]["dec_vocab_size"] == 37 and c.model["model_dir"] == str(tmp_path)


def test_param_key_set_is_chainers():
    from oracle.ast_oracle import default_model_cfg, param_shapes, persistent_shapes
    keys = set(param_shapes(default_model_cfg(), 40)) | set(persistent_shapes(default_model_cfg()))
    want = {"CNN_0/W", "CNN_1/W", "attn_Wa/W", "attn_Wa/b", "context/W", "context/b", "embed_dec/W", "out/W", "out/b"}
    for i in (0, 1):
        want |= {f"CNN_{i}_bn/{p}" for p in ("gamma", "beta", "avg_mean", "avg_var", "N")}
    for l in range(3):
        for s in ("enc", "rev_enc", "dec"):
            want |= {f"L{l}_{s}/upward/W", f"L{l}_{s}/upward/b", f"L{l}_{s}/lateral/W"}
    assert keys == want
    sh = param_shapes(default_model_cfg(), 40)
    assert sh["L0_enc/upward/W"] == (1024, 1536) and sh["L0_dec/upward/W"] == (2048, 640) and sh["context/W"] == (512, 1024)
    assert param_shapes(default_model_cfg(), 13)["L0_enc/upward/W"] == (1024, 512)
    assert sum(int(np.prod(s)) for s in sh.values()) == 14430410


def test_cmvn_coefficients():
    from ast_b200.dataloader import cmvn_scale_offset
    from oracle.ast_oracle import apply_cmvn
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((50, 4)) * 3 + 2).astype(np.float32)
    s, o = cmvn_scale_offset(x.astype(np.float64).sum(0), (x.astype(np.float64) ** 2).sum(0), 50)
    assert np.allclose(x * s + o, apply_cmvn(x, x.astype(np.float64).sum(0), (x.astype(np.float64) ** 2).sum(0), 50), atol=1e-6)


def test_corpus_bleu_known_answers(tmp_path):
    """eval.py:29-38 (nltk corpus_bleu + smoothing method2), hand-computed."""
    import math
    from ast_b200.eval import Eval, corpus_bleu, brevity_penalty, closest_ref_length, modified_precision
    # Papineni's clipping example: 'the' x7 against two references -> 2/7
    hyp = "the the the the the the the".split()
    refs = ["the cat is on the mat".split(), "there is a cat on the mat".split()]
    assert modified_precision(refs, hyp, 1) == (2, 7)
    assert modified_precision(refs, hyp, 2) == (0, 6)
    assert closest_ref_length(refs, 7) == 7 and closest_ref_length([[0] * 5, [0] * 9], 7) == 5   # tie -> shorter
    assert brevity_penalty(6, 7) == 1.0 and brevity_penalty(7, 0) == 0.0
    assert abs(brevity_penalty(9, 6) - math.exp(1 - 9 / 6)) < 1e-15
    # perfect match of one 5-token sentence: counts 5,4,3,2 -> smoothed p = 6/6, 5/5, ... = 1 -> BLEU 1
    s = "a b c d e".split()
    assert abs(corpus_bleu([[s]], [s]) - 1.0) < 1e-12
    # corpus of two sentences, one reference each
    r1, h1 = "the quick brown fox jumps".split(), "the quick brown dog jumps".split()
    r2, h2 = "hello world again".split(), "hello world".split()
    # unigrams: 4/5 + 2/2 = 6/7 ; bigrams: 2/4 + 1/1 = 3/5 ; trigrams: 1/3 + 0/max(1,0) = 1/4 ; 4-grams: 0/2 + 0/1 = 0/3
    want_all = math.exp(1 - 8 / 7) * math.exp(0.25 * (math.log(7 / 8) + math.log(4 / 6) + math.log(2 / 5) + math.log(1 / 4)))
    want_new = math.exp(1 - 8 / 7) * math.exp(0.25 * (math.log(6 / 7) + math.log(4 / 6) + math.log(2 / 5) + math.log(1 / 4)))
    assert abs(corpus_bleu([[r1], [r2]], [h1, h2]) - want_all) < 1e-12
    assert abs(corpus_bleu([[r1], [r2]], [h1, h2], smooth_unigram=False) - want_new) < 1e-12
    assert corpus_bleu([[r1]], ["x y z".split()]) == 0.0            # no unigram match
    # Eval reads the reference's directory layout (data/fisher/refs/<set>/eval.ids, ref.en0..)
    (tmp_path / "eval.ids").write_text("u1\nu2\n")
    (tmp_path / "ref.en0").write_text(" ".join(r1) + "\n" + " ".join(r2) + "\n")
    (tmp_path / "ref.en1").write_text("a fast brown fox leaps\nhello there world\n")
    ev = Eval(str(tmp_path), 2)
    assert len(ev.refs) == 2 and len(ev.refs[0]) == 2
    b = ev.calc_bleu({"u1": h1, "u2": h2})
    assert 0.0 < b < 1.0
    ev.write_to_file({"u1": h1, "u2": h2}, str(tmp_path / "out.txt"))
    assert (tmp_path / "out.txt").read_text() == "the quick brown dog jumps\nhello world\n"


def test_grad_allreduce_rejects_buckets_that_do_not_tile_the_buffer():
    """GradAllReduce (data-parallel hook): the gradient buckets an engine reports must tile the flat buffer exactly - a gap
    would leave gradients unreduced, an overlap would reduce them twice."""
    import torch
    from ast_b200 import dist as adist

    class _Opt:
        grad_scale = 1.0
        pre_update = None

    class _Eng:
        grads = torch.zeros(1000)

        def __init__(self, buckets):
            self._b = buckets

        def grad_buckets(self):
            return self._b

    ok = adist.GradAllReduce(_Eng([(600, 400), (100, 500), (0, 100)]), _Opt(), world=1)
    assert ok.buckets[0] == (600, 400) and not ok.overlap
    for bad in ([(600, 400), (0, 100)], [(500, 500), (100, 500), (0, 100)], [(0, 999)]):
        with pytest.raises(AssertionError):
            adist.GradAllReduce(_Eng(bad), _Opt(), world=1)


def test_declared_c_abi_covers_the_data_parallel_and_grouped_entry_points(lib):
    """The round-1 additions are part of the C ABI (include/ast_b200.h) and of the ctypes table."""
    import re
    hdr = open(os.path.join(ROOT, "include", "ast_b200.h")).read()
    for sym in ("ast_grad_bucket_count", "ast_grad_bucket_range", "ast_grad_bucket_wait", "ast_gemm_grouped"):
        assert re.search(r"\b%s\s*\(" % sym, hdr), sym
        assert hasattr(lib, sym), sym

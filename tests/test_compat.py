"""The drop-in boundary at driver level (SURVEY 8b "runtime level"): with `compat/` on PYTHONPATH the
reference's own train.py / beam.py / copy_params.py resolve every name they import, and - on a GPU, where the reference
checkout is available (AST_REFERENCE_DIR or /root/reference; it never travels with the repo) - run UNMODIFIED end to end on
a synthetic experiment directory."""
import ast
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "compat")
REF = os.environ.get("AST_REFERENCE_DIR", "/root/reference")
DRIVERS = ("train.py", "beam.py", "copy_params.py")
have_ref = all(os.path.isfile(os.path.join(REF, d)) for d in DRIVERS)


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([COMPAT, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    return env


def test_compat_modules_resolve_to_ast_b200():
    code = ("import nn, eval, seq2seq, dataloader, config, cupy\n"
            "from chainer import serializers, cuda, Function, utils, Variable\n"
            "import chainer, ast_b200.nn, ast_b200.eval, ast_b200.serializers, ast_b200.seq2seq, ast_b200.dataloader\n"
            "assert nn.NN is ast_b200.nn.NN and eval.Eval is ast_b200.eval.Eval\n"
            "assert serializers.save_npz is ast_b200.serializers.save_npz and serializers.load_npz is ast_b200.serializers.load_npz\n"
            "assert seq2seq.SpeechEncoderDecoder is ast_b200.seq2seq.SpeechEncoderDecoder\n"
            "assert dataloader.FisherDataLoader is ast_b200.dataloader.FisherDataLoader and dataloader.SYMBOLS.EOS_ID == 2\n"
            "assert cuda.cupy is cupy and callable(cupy.all) and callable(chainer.using_config)\n"
            "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], env=_env(), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


# attributes / methods the drivers use on the objects they get back (train.py:34-75, beam.py:80-144, copy_params.py:15-68)
NN_ATTRS = ["cfg", "model_fname", "max_epoch", "train_log", "dev_log", "gpuid", "data_loader", "model", "optimizer",
            "train_epoch", "predict", "decode_beam"]


@pytest.mark.skipif(not have_ref, reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("driver", DRIVERS)
def test_every_import_of_the_unmodified_drivers_resolves_under_compat(driver):
    """Static check (no GPU here): every `import X` / `from X import a, b` of the driver resolves with compat/ first on the
    path, and every attribute the driver reads off `nn` / `metrics` exists on the drop-in classes."""
    with open(os.path.join(REF, driver)) as f:
        tree = ast.parse(f.read())
    lines = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            lines += [f"import {a.name}" for a in node.names]
        elif isinstance(node, ast.ImportFrom):
            lines.append(f"from {node.module} import {', '.join(a.name for a in node.names)}")
    assert any(l.startswith("from nn import") for l in lines)
    used = sorted({n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)
                   and n.value.id in ("nn", "nn_1", "nn_2")})
    assert set(used) <= set(NN_ATTRS), used
    code = "\n".join(lines) + (
        "\nimport inspect\nsrc = inspect.getsource(NN)\n"
        f"missing = [a for a in {used!r} if not hasattr(NN, a) and ('self.' + a) not in src]\n"
        "assert not missing, missing\n"
        "assert all(hasattr(Eval, m) for m in ('calc_bleu', 'write_to_file'))\nprint('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], env=_env(), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (lines, r.stderr[-2000:])


def _write_exp(tmp, name, seed, **kw):
    from oracle import synth_corpus as SC
    root = os.path.join(tmp, name)
    mc = SC.small_model_cfg(hidden=128, embed=16, attn=128, c0=8, c1=16, dropout=(0.3, 0.3, 0.0))
    return SC.write_experiment(root, mc, feat_dim=40, vocab_words=40, seed=seed, batch_size=4, buckets_num=4,
                               buckets_width=16, max_pred=8, teach_ratio=0.8, speech_noise=0.25, zero_input=0.1, **kw)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="needs the reference checkout next to a GPU (AST_REFERENCE_DIR); it never travels with the repo")
def test_unmodified_reference_drivers_run_on_the_cuda_path(tmp_path):
    """`python <ref>/train.py -m <exp> -e 2`, `python <ref>/beam.py -m <exp> -n 5 -k 5 -s fisher_dev -w 1.0` and
    `python <ref>/copy_params.py` with PYTHONPATH=compat: logs, checkpoints, beam pickle and hypothesis file appear in the
    formats the reference writes; a second train.py invocation resumes from the newest checkpoint (nn.py:142-152)."""
    import pickle
    tmp = str(tmp_path)
    exp = _write_exp(tmp, "st", 11)
    env = _env()

    def run(script, *args, cwd=None):
        r = subprocess.run([sys.executable, os.path.join(REF, script), *args], env=env, capture_output=True, text=True,
                           cwd=cwd or tmp, timeout=900)
        assert r.returncode == 0, (script, r.stdout[-1500:], r.stderr[-3000:])
        return r.stdout

    out = run("train.py", "-m", exp, "-e", "2")
    assert "BLEU" in out and "Saving model" in out
    with open(os.path.join(exp, "train.log")) as f:
        tl = [l.strip().split(", ") for l in f]
    assert [int(a) for a, _ in tl] == [1, 2] and all(np.isfinite(float(b)) and float(b) > 0 for _, b in tl)
    assert float(tl[1][1]) < float(tl[0][1]) + 0.5
    with open(os.path.join(exp, "dev.log")) as f:
        assert [int(l.split(",")[0]) for l in f] == [1, 2]
    assert os.path.isfile(os.path.join(exp, "seq2seq_1.model")) and os.path.isfile(os.path.join(exp, "seq2seq_2.model"))
    with np.load(os.path.join(exp, "seq2seq_2.model")) as z:
        assert "L0_enc/upward/W" in z.files and "CNN_1_bn/avg_var" in z.files and int(z["CNN_0_bn/N"]) == 6
    out = run("train.py", "-m", exp, "-e", "1")                      # resumes: epoch 3
    assert "epoch: 3" in out and os.path.isfile(os.path.join(exp, "seq2seq_3.model"))
    out = run("beam.py", "-m", exp, "-n", "5", "-k", "5", "-s", "fisher_dev", "-w", "1.0")
    assert "BLEU" in out
    with open(os.path.join(exp, "fisher_dev_beam_N-5_K-5.p"), "rb") as f:
        beam = pickle.load(f)
    assert len(beam) == 3
    for u, hyps in beam.items():
        assert 1 <= len(hyps) <= 5 and all(h[0][0] == 1 for h in hyps)
        assert all(hyps[i][1] >= hyps[i + 1][1] for i in range(len(hyps) - 1))
        assert all(len(h[2]) == len(h[0]) - 1 for h in hyps)            # one attention vector per emitted token
    with open(os.path.join(exp, "fisher_dev_beam_N-5_K-5_W-1.00.en")) as f:
        assert len(f.read().splitlines()) == 3
    # copy_params.py: hard-coded sibling paths (copy_params.py:12-13), run from a working directory that makes them resolve
    work = os.path.join(tmp, "work")
    os.makedirs(work)
    base = os.path.join(tmp, "safe-copy-ast", "experiments")
    os.makedirs(base)
    asr = _write_exp(base, "asr_sw_GOLD_root", 21)
    st = _write_exp(base, "pretrain_sw_GOLD_root", 22)
    os.symlink(asr, os.path.join(base, "asr_sw_GOLD"))
    os.symlink(st, os.path.join(base, "pretrain_sw_GOLD"))
    run("train.py", "-m", asr, "-e", "1")
    out = run("copy_params.py", cwd=work)
    assert out.count("True") >= 3 and "Finished saving model" in out
    with np.load(os.path.join(asr, "seq2seq_1.model")) as a, np.load(os.path.join(st, "seq2seq_0.model")) as b:
        for k in ("CNN_0/W", "CNN_1_bn/gamma", "CNN_1_bn/avg_mean", "L2_rev_enc/lateral/W", "L0_enc/upward/b"):
            assert np.array_equal(a[k], b[k]), k
        assert not np.array_equal(a["out/W"], b["out/W"])            # the decoder is NOT copied (copy_params.py:58)

"""TEST INFRASTRUCTURE ONLY - writes a tiny synthetic experiment directory in the on-disk formats the reference's
drivers read (train.py:34-44, nn.py:44-65, config.py:17-29, dataloader.py:50-72,95-108,186-211, eval.py:14-26):

  <root>/exp/model_cfg.json, train_cfg.json
  <root>/data/corpus.vocab   pickle {dec_key: {"w2i": {bytes: id}, "i2w": [bytes]}}
  <root>/data/corpus.map     pickle {set: {utt: {dec_key: [bytes words]}}}
  <root>/data/corpus.info    pickle {set: {utt: {"sp": n_frames, ...}}}
  <root>/speech/<set>/<utt>.npy            (Fisher layout)   or   <root>/data/speech.pkl {set: {utt: (T,D)}} (GlobalPhone)
  <root>/data/refs/<set>/eval.ids, ref.en0..ref.en{n_evals-1}

Used by the golden generator (through the UNMODIFIED reference NN) and by the tests of the repo's loaders / compat layer.
"""
import json
import os
import pickle

import numpy as np

START_VOCAB = [b"_PAD", b"_GO", b"_EOS", b"_UNK"]


def small_model_cfg(hidden=32, embed=8, attn=32, c0=4, c1=8, kw=13, dropout=(0.0, 0.0, 0.0), layers=3):
    """The shipped model_cfg.json structure (experiments/es_en_20h/model_cfg.json) at toy widths."""
    return {
        "dropout": {"embed": dropout[0], "rnn": dropout[1], "out": dropout[2]},
        "rnn_config": {"bi_rnn": True, "enc_layers": layers, "dec_layers": layers, "hidden_units": hidden,
                       "embedding_units": embed, "attn_units": attn, "n_attn": 1, "feed_attn": True, "ln": False},
        "cnn_config": {"bn": True, "cnn_layers": [
            {"in_channels": None, "out_channels": c0, "ksize": [9, kw], "stride": [2, kw], "pad": [4, 0]},
            {"in_channels": None, "out_channels": c1, "ksize": [9, 1], "stride": [2, 1], "pad": [4, 0]}]},
    }


def write_experiment(root, model_cfg, feat_dim=13, vocab_words=40, sets=None, seed=0, dec_key="bpe_w", globalphone=False,
                     batch_size=4, buckets_num=4, buckets_width=16, max_pred=12, teach_ratio=1.0, speech_noise=0.0,
                     zero_input=0.0, freeze=(), n_evals=2, lr=1e-3, train_set="fisher_train", dev_set="fisher_dev",
                     seed_str="seed-synth"):
    """sets: {set_key: [n_frames per utterance]}.  Returns the experiment dir (the `-m` argument of train.py / beam.py)."""
    rng = np.random.default_rng(seed)
    if sets is None:
        sets = {train_set: [20, 23, 31, 34, 37, 45, 52, 61, 47, 29], dev_set: [22, 35, 50]}
    exp = os.path.join(root, "exp")
    data = os.path.join(root, "data")
    os.makedirs(exp, exist_ok=True)
    os.makedirs(data, exist_ok=True)
    words = [("w%d" % i).encode() if i % 3 else ("w%d@@" % i).encode() for i in range(vocab_words)]
    i2w = START_VOCAB + words
    vocab = {dec_key: {"w2i": {w: i for i, w in enumerate(i2w)}, "i2w": i2w}}
    cmap, info, speech = {}, {}, {}
    for set_key, lens in sets.items():
        cmap[set_key], info[set_key], speech[set_key] = {}, {}, {}
        for j, n in enumerate(lens):
            utt = "spk%d_%s_%03d" % (j % 3, set_key[-3:], j)
            nw = int(rng.integers(2, max(3, max_pred - 3)))
            toks = [words[int(k)] for k in rng.integers(0, vocab_words, size=nw)]
            if j == 0:
                toks[0] = b"oov-word"                       # exercises the UNK path (dataloader.py:146)
            cmap[set_key][utt] = {dec_key: toks}
            info[set_key][utt] = {"sp": int(n), dec_key: nw}
            speech[set_key][utt] = rng.standard_normal((int(n), feat_dim)).astype(np.float32)
    with open(os.path.join(data, "corpus.vocab"), "wb") as f:
        pickle.dump(vocab, f)
    with open(os.path.join(data, "corpus.map"), "wb") as f:
        pickle.dump(cmap, f)
    with open(os.path.join(data, "corpus.info"), "wb") as f:
        pickle.dump(info, f)
    if globalphone:
        speech_path = os.path.join(data, "speech.pkl")
        with open(speech_path, "wb") as f:
            pickle.dump(speech, f)
    else:
        speech_path = os.path.join(root, "speech")
        for set_key in speech:
            os.makedirs(os.path.join(speech_path, set_key), exist_ok=True)
            for k, (utt, a) in enumerate(speech[set_key].items()):
                d = os.path.join(speech_path, set_key)
                if k % 2:                                   # the per-speaker sub-directory fallback (dataloader.py:100-102)
                    d = os.path.join(d, utt.split("_", 1)[0])
                    os.makedirs(d, exist_ok=True)
                np.save(os.path.join(d, utt + ".npy"), a)
    refs = os.path.join(data, "refs")
    for set_key in sets:
        d = os.path.join(refs, set_key)
        os.makedirs(d, exist_ok=True)
        utts = list(cmap[set_key])
        with open(os.path.join(d, "eval.ids"), "w", encoding="utf-8") as f:
            f.write("\n".join(utts) + "\n")
        for r in range(n_evals):
            with open(os.path.join(d, "ref.en%d" % r), "w", encoding="utf-8") as f:
                for u in utts:
                    t = " ".join(w.decode() for w in cmap[set_key][u][dec_key]).replace("@@ ", "")
                    f.write(t + (" extra" if r else "") + "\n")
    train_cfg = {
        "seed": seed_str, "gpuid": 0, "iters_save": 1, "train_set": train_set, "dev_set": dev_set,
        "extras": {"random_out": 0, "speech_noise": speech_noise, "teach_ratio": teach_ratio},
        "data": {"enc_key": "sp", "dec_key": dec_key, "speech_path": speech_path,
                 "map_path": os.path.join(data, "corpus.map"), "vocab_path": os.path.join(data, "corpus.vocab"),
                 "max_pred": max_pred, "info_path": os.path.join(data, "corpus.info"), "refs_path": refs,
                 "n_evals": n_evals, "buckets_num": buckets_num, "buckets_width": buckets_width, "train_scale": 1,
                 "zero_input": zero_input},
        "optimizer": {"type": 0, "lr": lr, "l2": 1e-4, "grad_clip": 2, "grad_noise_eta": 0, "freeze": list(freeze)},
        "batch_size": batch_size,
    }
    if globalphone:
        train_cfg["data"]["dataloader"] = "globalphone"
    with open(os.path.join(exp, "model_cfg.json"), "w") as f:
        json.dump(model_cfg, f, indent=1)
    with open(os.path.join(exp, "train_cfg.json"), "w") as f:
        json.dump(train_cfg, f, indent=1)
    return exp

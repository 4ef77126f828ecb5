"""Legacy surface: enc_dec.py::SpeechEncoderDecoder(m_cfg, gpuid) with flat config keys, forward()'s
return conventions, per-step teacher forcing and add_weight_noise, over the same CUDA engine."""
import random

import numpy as np
import pytest

from oracle import ast_oracle as O

LEGACY_CFG = {
    "enc_key": "sp", "dec_key": "bpe_w", "vocab_path": "unused", "sp_dim": 40, "max_en_pred": 9,
    "hidden_units": 64, "enc_layers": 3, "dec_layers": 3, "embedding_units": 16, "attn_units": 128, "bi_rnn": True, "ln": False,
    "bn": True, "rnn_dropout": 0.0, "out_dropout": 0.0,
    "cnn_layers": [{"in_channels": 1, "out_channels": 8, "ksize": [9, 13], "stride": [2, 13], "pad": [4, 0]},
                   {"in_channels": 8, "out_channels": 16, "ksize": [9, 1], "stride": [2, 1], "pad": [4, 0]}],
}


def test_legacy_flat_config_maps_onto_the_live_geometry():
    from ast_b200.enc_dec import nested_config
    c = nested_config(dict(LEGACY_CFG, rnn_dropout=0.2, embed_dropout=True), 44)
    # enc_dec.py:80-106: with bi_rnn each direction has the full hidden_units -> live hidden_units = 2h
    assert c["rnn_config"]["hidden_units"] == 128 and c["rnn_config"]["dec_vocab_size"] == 44
    assert c["dropout"] == {"embed": 0.2, "rnn": 0.2, "out": 0.0}
    assert nested_config(dict(LEGACY_CFG, rnn_dropout=0.2), 44)["dropout"]["embed"] == 0.0      # enc_dec.py:295: key must be present
    shapes = O.param_shapes(c, 40)
    assert shapes["attn_Wa/W"] == (128, 128) and shapes["context/W"] == (128, 256) and shapes["L0_dec/upward/W"] == (512, 16 + 128)
    assert shapes["L0_enc/lateral/W"] == (256, 64)
    for bad in ({"cnn_pool": [[2, 1], [2, 1]]}, {"leaky_relu": True}, {"rnn_relu": True}, {"ln": True}, {"enc_key": "es_w"},
                {"bi_rnn": False}, {"random_out": True}):
        with pytest.raises(ValueError):
            nested_config(dict(LEGACY_CFG, **bad), 44)


@pytest.mark.gpu
def test_legacy_forward_conventions_match_oracle():
    torch = pytest.importorskip("torch")
    from ast_b200.enc_dec import SpeechEncoderDecoder, nested_config
    from ast_b200.nn import using_config
    V = 44
    cfg = nested_config(LEGACY_CFG, V)
    P = O.init_params(cfg, 40, seed=9)
    P["out/W"] = P["out/W"] * 3.0
    X, y, _ = O.synth_batch(3, 70, 40, V, 5, 8, seed=10, Tmin=55)
    m = SpeechEncoderDecoder(LEGACY_CFG, 0, vocab_size=V)            # legacy argument order (enc_dec.py:14)
    assert m.v_size_en == V and m.m_cfg is LEGACY_CFG
    m.load_state(P)
    L = y.shape[1]
    # train mode: ([], loss), teacher forcing drawn at EVERY step (enc_dec.py:344)
    random.seed(5)
    with using_config("train", True):
        out, loss = m.forward(X, add_noise=0, teacher_ratio=0.5, y=y)
    random.seed(5)
    bits = [random.random() < 0.5 for _ in range(L - 1)]
    assert not all(bits[1:])
    bits[0] = True
    om = O.OracleModel(cfg, P, dtype=np.float64)
    # the oracle's tf_bits force steps 0 and >= L-2 (live rule); emulate "every step" with explicit feedback of its own argmax
    om2 = O.OracleModel(cfg, P, dtype=np.float64)
    want = float(_legacy_loss(om2, X, y, bits))
    assert out == [] and abs(float(loss.data) - want) <= 1e-4 * abs(want)
    loss.backward()
    # eval mode: (pred (B, n), loss along the greedy path)
    m.load_state(P)
    with using_config("train", False):
        pred, eloss = m.forward(X, y=y)
    om = O.OracleModel(cfg, P, dtype=np.float32)
    om.train = False
    om.encode(X.astype(np.float32)); om.init_decoder_state()
    ht = np.zeros((3, 128), np.float32); word = y[:, 0].astype(np.int64); want_pred = []; want_loss = 0.0
    done = np.zeros(3, bool)
    for n in range(L - 1):
        logits, ht, _ = om.decode_step(word, ht)
        li, _ = O.softmax_cross_entropy(logits.astype(np.float64), y[:, n + 1].astype(np.int64), om.mask_pad_id)
        want_loss += float(li); word = logits.argmax(1); want_pred.append(word.copy())
        done |= word == O.EOS_ID
        if done.all():
            break
    assert (pred.cpu().numpy() == np.stack(want_pred).T).all()
    assert abs(float(eloss.data) - want_loss) <= 1e-4 * abs(want_loss)
    with using_config("train", False):
        pred2, l2 = m.forward(X)                                     # no labels: (pred, 0), up to max_en_pred steps
    assert l2 == 0 and pred2.shape[0] == 3 and pred2.shape[1] <= LEGACY_CFG["max_en_pred"]
    # add_weight_noise (enc_dec.py:587-624): LSTM W / upward b and the decoder embedding move, nothing else does
    before = {k: m._engine.view(k).clone() for k in m._engine.info}
    m.add_weight_noise(0.0, 0.01)
    for k in m._engine.info:
        moved = not torch.equal(before[k], m._engine.view(k))
        should = ("_enc/" in k or "_dec/" in k) and (k.endswith("/W") or k.endswith("upward/b")) or k == "embed_dec/W"
        assert moved == bool(should), k
        if moved:
            d = (m._engine.view(k) - before[k]).float()
            assert abs(float(d.std()) - 0.01) < 2e-3 or d.numel() < 600


def _legacy_loss(om, X, y, bits):
    """decode_batch (enc_dec.py:328-370) on the oracle: use the label where bits[i], else the previous argmax."""
    X = np.asarray(X, dtype=om.dtype)
    B, L = y.shape
    om.encode(X); om.init_decoder_state()
    ht = np.zeros((B, om.A), dtype=om.dtype)
    loss, inp = 0.0, y[:, 0].astype(np.int64)
    for i in range(L - 1):
        if bits[i]:
            inp = y[:, i].astype(np.int64)
        logits, ht, _ = om.decode_step(inp, ht, step_key=i)
        inp = logits.argmax(axis=1)
        li, _ = O.softmax_cross_entropy(logits, y[:, i + 1].astype(np.int64), om.mask_pad_id)
        loss += float(li)
    return loss

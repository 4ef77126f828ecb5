"""Legacy surface: the reference's older model file enc_dec.py::SpeechEncoderDecoder, used by nmt_run.py (SURVEY 8a-legacy,
8f-4).  Same constructor order, flat config keys, method names and return conventions as /root/reference/enc_dec.py; underneath
it is the same CUDA engine as the live model (ast_b200/seq2seq.py).

How the legacy geometry maps onto the live one (enc_dec.py:61-106): with ``bi_rnn`` every direction has the FULL ``hidden_units``
h, so enc_states are 2h wide, the decoder LSTMs have 2h units, attn_Wa is (2h, 2h) and context is (4h -> attn_units) - exactly the
live model with ``hidden_units = 2h`` (seq2seq.py:73, 103-116).  Differences that stay visible at this surface:
  * constructor ``(m_cfg, gpuid)`` (enc_dec.py:14), vocabulary size read from ``m_cfg["vocab_path"]`` (:34-39);
  * explicit ``sp_dim`` instead of lazy shaping (:110);
  * teacher forcing drawn at EVERY step, the first included (:344; the first step's input is the label either way, :339);
  * dropout only when its ratio is > 0 (:268, :317 - the same arithmetic), embedding dropout uses ``rnn_dropout`` and only if
    the key ``embed_dropout`` is present (:295-297);
  * ``forward()`` returns ``([], loss)`` in train mode and ``(pred.T, loss)`` in eval mode (:569-584);
  * ``add_weight_noise(mu, sigma)`` perturbs every LSTM W / upward b and the decoder embedding (:587-624).
Options of the legacy file that no shipped experiment uses are rejected loudly: text encoders (``enc_key != 'sp'``),
``cnn_pool``, ``leaky_relu``, ``rnn_relu``, ``ln``, ``random_out``, unidirectional encoders.
"""
import pickle
import random

import numpy as np
import torch

from . import seq2seq as live
from .seq2seq import Variable, config as _train_config
from .symbols import SYMBOLS

GO_ID, EOS_ID = SYMBOLS.GO_ID, SYMBOLS.EOS_ID


def nested_config(m_cfg: dict, vocab_size: int) -> dict:
    """Flat legacy keys (enc_dec.py, nmt_run.py's model.cfg) -> the nested dict the live model takes (config.py:15-29)."""
    for key in ("cnn_pool", "leaky_relu", "rnn_relu"):
        if m_cfg.get(key):
            raise ValueError(f"legacy option {key!r} is outside the hot-path scope (no shipped experiment uses it)")
    if m_cfg.get("enc_key", "sp") != "sp":
        raise ValueError("text encoders (enc_key != 'sp', enc_dec.py:162-164) are outside the hot-path scope")
    if not m_cfg.get("bi_rnn", True):
        raise ValueError("unidirectional legacy encoders are not supported (every shipped config is bidirectional)")
    if m_cfg.get("ln", False):
        raise ValueError("layer normalisation (enc_dec.py:56-57) is outside the hot-path scope")
    if m_cfg.get("random_out", False):
        raise ValueError("random_out (enc_dec.py:361-368) is not supported")
    rd = float(m_cfg.get("rnn_dropout", 0.0))
    return {
        "dropout": {"embed": rd if "embed_dropout" in m_cfg else 0.0, "rnn": rd, "out": float(m_cfg.get("out_dropout", 0.0))},
        "rnn_config": {"bi_rnn": True, "enc_layers": m_cfg["enc_layers"], "dec_layers": m_cfg["dec_layers"],
                       "hidden_units": 2 * m_cfg["hidden_units"], "embedding_units": m_cfg["embedding_units"],
                       "attn_units": m_cfg["attn_units"], "n_attn": 1, "feed_attn": True, "ln": False,
                       "dec_vocab_size": int(vocab_size)},
        "cnn_config": {"bn": bool(m_cfg.get("bn", True)),
                       "cnn_layers": [{k: l[k] for k in ("in_channels", "out_channels", "ksize", "stride", "pad") if k in l}
                                      for l in m_cfg["cnn_layers"]]},
    }


class SpeechEncoderDecoder(live.SpeechEncoderDecoder):
    def __init__(self, m_cfg, gpuid, vocab_size=None):
        """enc_dec.py:14-24.  ``vocab_size`` is an injection point for tests; otherwise the vocabulary pickle is read."""
        self.m_cfg = m_cfg
        if vocab_size is None:
            with open(m_cfg["vocab_path"], "rb") as f:
                vocab_size = len(pickle.load(f)[m_cfg["dec_key"]]["w2i"])           # enc_dec.py:34-39
        self.v_size_en, self.v_size_es = int(vocab_size), 0
        super().__init__(gpuid, nested_config(m_cfg, vocab_size), feat_dim=m_cfg["sp_dim"])

    # ---- names of the legacy file -----------------------------------------------------------------------------------
    def reset_state(self):                       # enc_dec.py:206
        self.reset_rnn_state()

    def set_decoder_state(self):                 # enc_dec.py:220
        self.init_decoder_state()

    def forward_enc(self, X, l=None):            # enc_dec.py:517
        self.encode(X)

    def compute_context_vector(self, dec_h, attn_Wa=None):     # enc_dec.py:238 takes no link argument
        return super().compute_context_vector(dec_h, attn_Wa)

    def decode(self, word, ht, get_alphas=False):              # enc_dec.py:291
        logits, ht_out, alphas = self.decode_step(word, ht)
        return (logits, ht_out, alphas) if get_alphas else (logits, ht_out)

    @staticmethod
    def draw_use_label(n_steps, teacher_ratio):
        """enc_dec.py:344: one random.random() draw for EVERY decode step, the first included."""
        return [random.random() < teacher_ratio for _ in range(n_steps)]

    def decode_batch(self, decoder_batch, teacher_ratio, X=None, add_noise=0):
        """enc_dec.py:328-370 (time-major labels (L, B)).  The engine runs the encoder and the decoder of a training step as
        one forward pass, so the features of the last forward_enc() are used unless ``X`` is given."""
        yb = decoder_batch.data if isinstance(decoder_batch, Variable) else decoder_batch
        y = yb.t().contiguous() if isinstance(yb, torch.Tensor) else np.ascontiguousarray(np.asarray(yb).T)
        bits = self.draw_use_label(int(y.shape[1]) - 1, teacher_ratio)
        bits[0] = True                           # the first input is decoder_batch[0] whatever the draw says (:339)
        X = self._last_X if X is None else X
        return self.forward_loss(X, y, teach_ratio=teacher_ratio, add_noise=add_noise, use_true=bits)

    def predict_batch(self, batch_size, pred_limit, y=None, display=False, X=None):
        """enc_dec.py:372-429 -> (pred (B, n_steps), loss).  With labels the loss is accumulated along the GREEDY path in eval
        mode (stop_limit = len(y) - 1) through the public decode_step protocol and the fused CE kernel."""
        e = self._engine
        X = self._last_X if X is None else X
        if y is None:
            return self.predict(X, GO_ID, EOS_ID, int(pred_limit)), 0
        yb = y.data if isinstance(y, Variable) else y
        yb = yb if isinstance(yb, torch.Tensor) else torch.as_tensor(np.asarray(yb))
        yb = yb.to(device=e.device, dtype=torch.int32)                  # (L, B) time-major
        self.encode(X)
        self.init_decoder_state()
        B = int(yb.shape[1])
        ht = torch.zeros(B, e.A, device=e.device)
        word = yb[0].contiguous()
        seen = torch.zeros(B, dtype=torch.bool, device=e.device)
        preds, loss = [], torch.zeros((), device=e.device)
        for n in range(int(yb.shape[0]) - 1):
            logits, ht_v, _ = self.decode_step(word, ht)
            ht = ht_v.data
            row_loss, word = e.softmax_ce(logits.data, yb[n + 1].contiguous())
            loss = loss + row_loss.sum()
            preds.append(word)
            seen |= word == EOS_ID
            if bool(seen.all()):
                break
        return torch.stack(preds, dim=0).t().contiguous(), Variable(loss)

    def forward(self, X, add_noise=0, teacher_ratio=0, y=None):
        """enc_dec.py:537-584: ``([], loss)`` in train mode, ``(pred.T, loss)`` in eval mode; ``y`` is (B, L) batch-major."""
        X = X.data if isinstance(X, Variable) else X
        self._last_X = X
        if _train_config.train:
            yb = y.data if isinstance(y, Variable) else y
            L = int(yb.shape[1])
            bits = self.draw_use_label(L - 1, teacher_ratio)
            bits[0] = True
            self.loss = self.forward_loss(X, yb, teach_ratio=teacher_ratio, add_noise=add_noise, use_true=bits)
            return [], self.loss
        yt = None
        if y is not None:
            yb = y.data if isinstance(y, Variable) else y
            yt = yb.t() if isinstance(yb, torch.Tensor) else np.asarray(yb).T
        return self.predict_batch(batch_size=int(X.shape[0]), pred_limit=self.m_cfg["max_en_pred"], y=yt, X=X)

    def add_weight_noise(self, mu, sigma):
        """enc_dec.py:587-624: W += N(mu, sigma) on every LSTM upward / lateral W, upward b, and the decoder embedding.
        (The reference draws with cupy's generator; here torch's CUDA generator - noise streams are backend-specific.)"""
        e = self._require()
        for name in self.rnn_enc + self.rnn_rev_enc + self.rnn_dec:
            for key in (f"{name}/upward/W", f"{name}/lateral/W", f"{name}/upward/b"):
                v = e.view(key)
                v.add_(torch.empty_like(v).normal_(float(mu), float(sigma)))
        v = e.view("embed_dec/W")
        v.add_(torch.empty_like(v).normal_(float(mu), float(sigma)))
        e.weights_changed()

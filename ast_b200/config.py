"""Experiment configuration (config.py:15-29): two JSON files per experiment directory plus the
decoder vocabulary size read from the vocab pickle.  Consumed unchanged by the new path."""
import json
import os
import pickle


class Config:
    def __init__(self, cfg_path: str, vocab_size: int = None) -> None:
        with open(os.path.join(cfg_path, "model_cfg.json"), "r") as model_f:
            self.model = json.load(model_f)
        with open(os.path.join(cfg_path, "train_cfg.json"), "r") as train_f:
            self.train = json.load(train_f)
        if vocab_size is None:
            with open(self.train["data"]["vocab_path"], "rb") as f:
                vocab = pickle.load(f)
            vocab_size = len(vocab[self.train["data"]["dec_key"]]["w2i"])
        self.model["rnn_config"]["dec_vocab_size"] = vocab_size
        print("vocab size {0:s} = {1:d}".format(self.train["data"]["dec_key"], vocab_size))
        self.model["model_dir"] = cfg_path


def es_en_20h_model_cfg(vocab=1098, dropout=(0.3, 0.3, 0.0)):
    """The shipped experiments/es_en_20h/model_cfg.json as a dict (+ injected dec_vocab_size)."""
    return {
        "dropout": {"embed": dropout[0], "rnn": dropout[1], "out": dropout[2]},
        "rnn_config": {"bi_rnn": True, "enc_layers": 3, "dec_layers": 3, "hidden_units": 512, "embedding_units": 128,
                       "attn_units": 512, "n_attn": 1, "feed_attn": True, "ln": False, "dec_vocab_size": vocab},
        "cnn_config": {"bn": True, "cnn_layers": [
            {"in_channels": None, "out_channels": 128, "ksize": [9, 13], "stride": [2, 13], "pad": [4, 0]},
            {"in_channels": None, "out_channels": 512, "ksize": [9, 1], "stride": [2, 1], "pad": [4, 0]}]},
    }

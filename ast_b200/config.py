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

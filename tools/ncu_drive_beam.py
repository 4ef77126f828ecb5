"""Driver for ncu: one lock-step beam search (32 utterances x beam 10, T = 1000, 6 steps)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder, config
config.train = False
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(), feat_dim=40); m.init_params(seed=0); e = m._engine
rng = np.random.default_rng(7)
utts = [rng.standard_normal((1, 1000, 40), dtype=np.float32) for _ in range(32)]
e.beam_search_batch(utts, 6, 10, 10); torch.cuda.synchronize()
print("searched")

"""Beam-10 throughput of the lock-step group search through the public path (host features in, hypothesis lists out), with the
time split into the device call and the host conversion.  Usage: python tools/beam_bench.py [T] [stop] [G] [n_utts]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder, config
from ast_b200.nn import beam_result_to_entries
config.train = False
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
stop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
G = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n = int(sys.argv[4]) if len(sys.argv) > 4 else 96
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(), feat_dim=40); m.init_params(seed=0); e = m._engine
rng = np.random.default_rng(7)
utts = [rng.standard_normal((1, T, 40), dtype=np.float32) for _ in range(n)]
[beam_result_to_entries(r) for r in e.beam_search_batch(utts[:G], stop, 10, 10)]; torch.cuda.synchronize()      # warm: workspace, kernels, pinned staging
t_dev = t_host = 0.0
t0 = time.perf_counter()
for i in range(0, n, G):
    a = time.perf_counter()
    res = e.beam_search_batch(utts[i:i + G], stop, 10, 10)
    torch.cuda.synchronize()
    b = time.perf_counter()
    ent = [beam_result_to_entries(r) for r in res]
    c = time.perf_counter()
    t_dev += b - a; t_host += c - b
dt = time.perf_counter() - t0
import ast_b200.nn as NNm
res = e.beam_search_batch(utts[:G], stop, 10, 10); torch.cuda.synchronize()
grp = res[0]["_group"]
a = time.perf_counter(); hp, hk, sc, ah = e.fetch_host([grp["hist_parent"], grp["hist_tok"], grp["scores"], grp["alpha_hist"]])
b = time.perf_counter(); toks, slots = NNm._backtrack(hp, hk, grp["n_steps"])
c = time.perf_counter(); ent = [beam_result_to_entries(r) for r in res]
d = time.perf_counter()
print(f"   conversion parts: fetch {1e3 * (b - a):.1f} ms, backtrack {1e3 * (c - b):.1f} ms, everything (incl. a second fetch + backtrack) {1e3 * (d - c):.1f} ms")
print(f"T={T} stop={stop} G={G}: {n / dt:.1f} utts/s; per group of {G}: search call {1e3 * t_dev * G / n:.1f} ms, conversion {1e3 * t_host * G / n:.1f} ms")

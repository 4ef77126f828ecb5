import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder
from ast_b200.beam import BeamPool
from ast_b200.nn import beam_result_to_entries
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(), feat_dim=40); m.init_params(seed=0)
e = m._engine; e.set_option("exact", 1)
rng = np.random.default_rng(7)
for T, stop in ((1000, 40), (1000, 175), (3000, 40)):
    utts = [rng.standard_normal((1, T, 40), dtype=np.float32) for _ in range(16)]
    e.beam_search(utts[0], stop, 10, 10); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for x in utts[:4]: beam_result_to_entries(e.beam_search(x, stop, 10, 10))
    torch.cuda.synchronize(); seq = 4 / (time.perf_counter() - t0)
    print(f"T={T} stop={stop}: sequential {seq:.1f} utts/s", flush=True)
    for n in (2, 4, 8):
        pool = BeamPool(e, n=n)
        pool.decode(utts[:n], stop, 10, 10, convert=beam_result_to_entries)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pool.decode(utts, stop, 10, 10, convert=beam_result_to_entries)
        torch.cuda.synchronize(); r = len(utts) / (time.perf_counter() - t0)
        print(f"   {n} in flight: {r:.1f} utts/s ({r/seq:.2f}x)", flush=True)
        del pool

"""Beam throughput of the decode-step variants (AST_BEAM_TC=0/1/3) on one GPU: sequential, lock-step 32."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder, config
from ast_b200.nn import beam_result_to_entries
config.train = False
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(), feat_dim=40); m.init_params(seed=0); e = m._engine
print("beam_tc =", e.get_option("beam_tc"))
rng = np.random.default_rng(7)
for T, stop in ((1000, 40), (1000, 175), (3000, 40)):
    utts = [rng.standard_normal((1, T, 40), dtype=np.float32) for _ in range(64)]
    e.beam_search(utts[0], stop, 10, 10); torch.cuda.synchronize(); t0 = time.perf_counter()
    for x in utts[:4]: beam_result_to_entries(e.beam_search(x, stop, 10, 10))
    torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    e.beam_search_batch(utts[:32], stop, 10, 10); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in (0, 32): [beam_result_to_entries(r) for r in e.beam_search_batch(utts[i:i + 32], stop, 10, 10)]
    torch.cuda.synchronize(); t2 = time.perf_counter() - t0
    t0 = time.perf_counter(); e.encode(np.concatenate(utts[:32], 0)); torch.cuda.synchronize(); t3 = time.perf_counter() - t0
    print(f"T{T} x {stop} steps: sequential {4 / t1:.1f} utts/s ({1e6 * t1 / 4 / stop:.0f} us/step incl. encoder); lock-step 32: {64 / t2:.1f} utts/s "
          f"({1e6 * t2 / 2 / stop:.0f} us per lock-step step incl. encoders); encode B=32: {1e3 * t3:.1f} ms", flush=True)

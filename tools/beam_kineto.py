"""Kernel list (CUPTI, via torch.profiler) of one lock-step beam search: where a decode step's time goes."""
import collections, os, sys, time
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg
from ast_b200.seq2seq import SpeechEncoderDecoder, config
config.train = False
G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
stop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(), feat_dim=40); m.init_params(seed=0); e = m._engine
rng = np.random.default_rng(7)
utts = [rng.standard_normal((1, T, 40), dtype=np.float32) for _ in range(G)]
run = (lambda: e.beam_search_batch(utts, stop, 10, 10)) if G > 1 else (lambda: e.beam_search(utts[0], stop, 10, 10))
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
ev = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for x in ev:
    k = x.name.split("(")[0][:80]
    c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += x.time_range.end - x.time_range.start
tot = sum(c[1] for c in agg.values())
print(f"beam_tc={e.get_option('beam_tc')} G={G} T={T} stop={stop}: wall {1e3 * wall:.1f} ms unprofiled; {len(ev)} device records, sum of durations {tot / 1e3:.1f} ms")
print(f"{'kernel':80s} {'count':>6s} {'total us':>10s} {'avg us':>9s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k:80s} {n:6d} {us:10.1f} {us / n:9.2f}")

"""Kernel list of REAL training steps (persistent encoder wavefront + cooperative decoder running as in production) through
torch.profiler (CUPTI concurrent-kernel activity records: kernels are not serialised, unlike under ncu).  Durations of the
kernels that wait for one another include their waiting.  Usage: python tools/kineto_step.py [--steps 2]"""
import argparse
import collections
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.config import es_en_20h_model_cfg          # noqa: E402
from ast_b200.seq2seq import SpeechEncoderDecoder        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--T", type=int, default=640)
ap.add_argument("--L", type=int, default=24)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--timeline", action="store_true", help="also print every device record of the last step in start order (us since its first record, duration, stream, name)")
a = ap.parse_args()
rng = np.random.default_rng(0)
m = SpeechEncoderDecoder(0, es_en_20h_model_cfg(dropout=(0.3, 0.3, 0.0)), feat_dim=40)
m.init_params(seed=0)
e = m._engine
e.set_option("exact", 0); e.set_option("tc_gemm", 1)
X = torch.as_tensor(rng.standard_normal((a.B, a.T, 40)).astype(np.float32), device=e.device)
y = rng.integers(4, 1098, (a.B, a.L)).astype(np.int32); y[:, 0] = 1; y[:, -1] = 2
y = torch.as_tensor(y, device=e.device)
bits = torch.as_tensor((rng.random(a.L - 1) < 0.8).astype(np.uint8), device=e.device)
mm, v, vh = (torch.zeros_like(e.params) for _ in range(3))


def step(it):
    loss = e.forward_loss(X, y, use_true=bits, noise_sigma=0.25)
    e.backward()
    e.opt_step(mm, v, vh, it + 1, 1e-3, 1e-4, 2.0)
    return loss


for it in range(3):
    step(it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(a.steps):
        step(3 + it)
    torch.cuda.synchronize()
ev = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA and "mem" not in x.name.lower()[:6]]
agg = collections.OrderedDict()
t_lo = min(x.time_range.start for x in ev); t_hi = max(x.time_range.end for x in ev)
for x in ev:
    k = x.name.split("(")[0][:70]
    c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += x.time_range.end - x.time_range.start
tot = sum(c[1] for c in agg.values())
print(f"{len(ev)} kernel records over {a.steps} steps, wall span {(t_hi - t_lo) / a.steps:.0f} us per step, sum of kernel durations {tot / a.steps:.0f} us per step "
      f"(kernels overlap: persistent wavefront, side streams)")
print(f"{'kernel':70s} {'count':>6s} {'total us':>10s} {'avg us':>9s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {n:6d} {us:10.1f} {us / n:9.2f}")

if a.timeline:
    allev = sorted((x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA), key=lambda x: x.time_range.start)
    # the last step starts at the last mul_noise / im2col0 kernel
    starts = [x.time_range.start for x in allev if "mul_noise" in x.name]
    t0 = starts[-1] if starts else allev[0].time_range.start
    print("timeline of the last step: start us | dur us | stream | name")
    for x in allev:
        if x.time_range.start < t0:
            continue
        try:
            sid = x.device_resource_id
        except Exception:
            sid = -1
        print(f"{x.time_range.start - t0:9.1f} {x.time_range.end - x.time_range.start:8.1f} {sid:4d}  {x.name.split('(')[0][:90]}")

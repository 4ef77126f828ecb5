"""Driver for ncu: the device pack path (ragged utterances -> CMVN + frame zeroing -> padded batch) on a full-size bucket."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200.dataloader import DevicePacker
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
lens = rng.integers(1601, 1681, 32)
utts = [rng.standard_normal((int(n), 40), dtype=np.float32) for n in lens]
keep = [(rng.random(int(n)) > 0.1).astype(np.uint8) for n in lens]
cmvn = (np.ones(40, np.float32) * 0.5, np.zeros(40, np.float32) + 0.1)
p = DevicePacker(dev, async_copy=False)
for _ in range(3):
    X = p.pack(utts, 1680, keep, cmvn=cmvn, noise_sigma=0.25, seed=3)
torch.cuda.synchronize()
print("packed", tuple(X.shape))

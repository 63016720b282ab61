"""Profiling driver: BASELINE config 4 on one GPU - random 3-regular graph of 10^6 sites, one
ladder of 64 betas, 20 sweeps with a swap step every 10 (general-graph sweep kernel + the fused
post-sweep tempering kernel)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

rng = np.random.default_rng(2026)
n, d = 1_000_000, 3
while True:
    stubs = np.repeat(np.arange(n, dtype=np.int64), d)
    rng.shuffle(stubs)
    a, b = stubs[0::2], stubs[1::2]
    key = np.minimum(a, b) * n + np.maximum(a, b)
    if not (a == b).any() and len(np.unique(key)) == len(key):
        break
ctx = nat.Context.get(0)
g = nat.Graph.from_edges(ctx, n, a.astype(np.uint64), b.astype(np.uint64), np.full(len(a), -1.0))
pt = nat.Tempering(g, np.geomspace(0.1, 1.5, 64), seed=7)
pt.timesteps_sample(20, 10, 21)
print("ok", pt.sim_stats()["kernel_launches"])

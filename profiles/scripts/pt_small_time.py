import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat
ctx = nat.Context.get(0)
for dims, R in (((8, 8, 8), 64), ((16, 16), 32), ((12, 12, 12), 128)):
    g = nat.Graph.torus(ctx, dims, j0=1.0, pmj=True, j_seed=3)
    pt = nat.Tempering(g, np.geomspace(0.1, 1.5, R), seed=5)
    pt.timesteps_sample(100, 10, 100)
    t0 = time.perf_counter(); pt.timesteps_sample(2000, 10, 2000); dt = time.perf_counter() - t0
    print(f"PT {dims} R={R}: {1e6*dt/2000:.2f} us/sweep (swap every 10), swaps={pt.total_swaps()}")

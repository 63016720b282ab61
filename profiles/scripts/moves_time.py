"""Cost of the non-basic moves next to the bit-sliced sweep: device time per timestep on config 3
(64^3 +-J x 1024 replicas) and on config 1 (32x32 x 64) for sweeps only, + one edge pass, + one worm."""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

ctx = nat.Context.get(0)
for dims, pmj, j0, E, T in (((64, 64, 64), True, 1.0, 1024, 20), ((32, 32), False, -1.0, 64, 200)):
    g = nat.Graph.torus(ctx, dims, j0=j0, pmj=pmj, j_seed=2024)
    sim = nat.Sim(g, E, seed=5)
    betas = np.full(T, 0.5)
    for name, mv in (("sweep only", None), ("+ edge pass", dict(edge_passes=1)), ("+ edge pass + 1 worm of 4", dict(edge_passes=1, worms=1)),
                     ("+ 16 worms of 4", dict(worms=16))):
        if mv is None:
            sim.set_moves()
        else:
            sim.set_moves(1, mv.get("edge_passes", 0), mv.get("worms", 0), 4)
        sim.sweeps(betas[:2])
        t0 = time.perf_counter(); sim.sweeps(betas); dt = time.perf_counter() - t0
        print(f"{'x'.join(map(str, dims))} E={E} {name}: {1e6 * dt / T:.1f} us per timestep", flush=True)
    sim.close()

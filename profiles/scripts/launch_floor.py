import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat
ctx = nat.Context.get(0)
g = nat.Graph.torus(ctx, (64, 64, 64), j0=1.0, pmj=True, j_seed=2024)
for E in (32, 64, 128, 256):
    sim = nat.Sim(g, E, seed=1)
    betas = np.linspace(0.1, 1.2, 400)
    sim.sweeps(betas[:50])
    for acc in (False, True):
        sim.reset_stats(); t0 = time.perf_counter(); sim.sweeps(betas, per_sweep_energies=acc); dt = time.perf_counter() - t0
        st = sim.stats()
        print(f"E={E} acc={acc}: device {st['sweep_device_ms']*1e3/400:.2f} us/sweep wall {dt*1e6/400:.2f}")
    sim.close()

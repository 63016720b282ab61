import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat
ctx = nat.Context.get(0)
for dims, E in (((32, 32), 64), ((32, 32), 256), ((32, 32), 512), ((32, 32), 1024), ((16, 16, 16), 64), ((16, 16, 16), 128), ((64, 64), 128)):
    g = nat.Graph.torus(ctx, dims, j0=-1.0)
    sim = nat.Sim(g, E, 1)
    betas = np.full(2000, 0.44)
    sim.sweeps(betas[:50])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); sim.sweeps(betas); best = min(best, time.perf_counter() - t0)
    words = g.nvars // 2 * ((E + 31) // 32)
    print(f"{dims} E={E} words/colour={words}: {1e6*best/2000:.2f} us/sweep  {E*g.nvars*2000/best:.3e} flips/s")
print("with per-sweep energies")
for dims, E in (((32, 32), 64), ((32, 32), 256), ((32, 32), 1024), ((16, 16, 16), 64), ((16, 16, 16), 128)):
    g = nat.Graph.torus(ctx, dims, j0=-1.0)
    sim = nat.Sim(g, E, 1)
    betas = np.full(1000, 0.44)
    sim.sweeps(betas[:50], per_sweep_energies=True)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); sim.sweeps(betas, per_sweep_energies=True); best = min(best, time.perf_counter() - t0)
    print(f"{dims} E={E}: {1e6*best/1000:.2f} us/sweep")

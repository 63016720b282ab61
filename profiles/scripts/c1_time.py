"""Wall time of the public API on BASELINE config 1 (32x32 torus, 1000 sweeps) and a 16x larger
batch; run from the repository root on a B200.  The CPU figure next to it in DESIGN.md comes
from `bench.py --impl reference` / the cpu_baseline leg."""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import pyisingmontecarlo_b200 as pkg

L = 32
edges = [((x * L + y, ((x + 1) % L) * L + y), -1.0) for x in range(L) for y in range(L)] + \
        [((x * L + y, x * L + (y + 1) % L), -1.0) for x in range(L) for y in range(L)]
lat = pkg.Lattice(edges, seed_gen=0)
for E in (64, 1024):
    lat.run_monte_carlo(0.44, 10, E)
    for rep in range(3):
        t0 = time.perf_counter()
        en, st = lat.run_monte_carlo(0.44, 1000, E)
        dt = time.perf_counter() - t0
    print(f"C1 32x32 E={E} 1000 sweeps: {1e3*dt:.2f} ms  ({E*1024*1000/dt:.3e} flips/s)  <e>={en.mean()/1024:.4f}")
    t0 = time.perf_counter()
    lat.run_monte_carlo_annealing_and_get_energies([(0, .1), (1000, .44)], 1000, E)
    dt = time.perf_counter() - t0
    print(f"   annealing+energies: {1e3*dt:.2f} ms")

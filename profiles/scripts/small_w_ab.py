"""Device time per colour phase of BASELINE config 3 (64^3 +-J) at the replica counts of the
1/2/4/8-GPU split (1024 / 512 / 256 / 128 per GPU), with and without per-sweep energies.
ISING_B200_LIB selects the library build under test; run from the repository root on a B200."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

tag = sys.argv[1] if len(sys.argv) > 1 else "stock"
ctx = nat.Context.get(0)
g = nat.Graph.torus(ctx, (64, 64, 64), j0=1.0, pmj=True, j_seed=2024)
n = 64 ** 3
for E in (128, 256, 512, 1024):
    T = 400 if E <= 256 else 200
    betas = np.linspace(0.1, 1.2, T, endpoint=False)
    sim = nat.Sim(g, E, seed=31337)
    out = []
    for acc in (False, True):
        sim.sweeps(betas[:40], per_sweep_energies=acc)
        best = 1e30
        for rep in range(3):
            sim.reset_stats()
            sim.sweeps(betas, per_sweep_energies=acc)
            best = min(best, sim.stats()["sweep_device_ms"] * 1e3 / T)
        out.append(best)
    print(f"{tag} E={E}: sweeps {out[0]:.2f} us/sweep, with energies {out[1]:.2f} us/sweep "
          f"({E * n / out[1] * 1e6:.3e} flips/s)", flush=True)
    sim.close()

"""Device time per sweep of BASELINE config 3 (64^3 +-J x 1024 replicas) at fixed inverse
temperatures and on the annealing ramp, with and without per-sweep energies: the cost of the
tie paths grows with beta.  ISING_B200_LIB selects the build; run from the repository root."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

tag = sys.argv[1] if len(sys.argv) > 1 else "stock"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = nat.Context.get(0)
g = nat.Graph.torus(ctx, (64, 64, 64), j0=1.0, pmj=True, j_seed=2024)
T = 200
sim = nat.Sim(g, E, seed=31337)
sim.sweeps(np.linspace(0.1, 1.2, 300))          # anneal once so that the fixed-beta runs see typical states
for name, betas in (("const0.1", np.full(T, 0.1)), ("ramp", np.linspace(0.1, 1.2, T, endpoint=False)),
                    ("const0.7", np.full(T, 0.7)), ("const1.2", np.full(T, 1.2))):
    out = []
    for acc in (False, True):
        sim.sweeps(betas[:40], per_sweep_energies=acc)
        best = 1e30
        for rep in range(3):
            sim.reset_stats()
            sim.sweeps(betas, per_sweep_energies=acc)
            best = min(best, sim.stats()["sweep_device_ms"] * 1e3 / T)
        out.append(best)
    print(f"{tag} E={E} {name}: sweeps {out[0]:.2f} us/sweep, with energies {out[1]:.2f} us/sweep", flush=True)

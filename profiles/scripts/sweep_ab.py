"""Device time of the stencil sweep on BASELINE config 3 (64^3 +-J) and config 2 (4096^2), for
the replica counts of the 1-GPU and the 8-GPU split, with and without per-sweep energies, Philox
7 / 10 rounds.  Run from the repository root on a B200; ISING_SWEEP_V1=1 selects the round-1
one-row-per-block launch for the A/B."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

ctx = nat.Context.get(0)
tag = "v1" if os.environ.get("ISING_SWEEP_V1") else "rows"
cases = [((64, 64, 64), True, 1.0, 1024, 200), ((64, 64, 64), True, 1.0, 128, 400),
         ((64, 64, 64), True, 1.0, 256, 400), ((4096, 4096), False, -1.0, 1024, 6)]
if len(sys.argv) > 1 and sys.argv[1] == "c3":
    cases = cases[:1]
for dims, pmj, j0, E, T in cases:
    g = nat.Graph.torus(ctx, dims, j0=j0, pmj=pmj, j_seed=2024)
    n = int(np.prod(dims))
    betas = np.linspace(0.1, 1.2, T, endpoint=False) if pmj else np.full(T, 0.43)
    for rounds in (7, 10):
        sim = nat.Sim(g, E, seed=31337, rounds=rounds)
        for acc in (False, True):
            sim.sweeps(betas[: max(2, T // 10)], per_sweep_energies=acc)
            best = 0.0
            for rep in range(3):
                sim.reset_stats()
                sim.sweeps(betas, per_sweep_energies=acc)
                st = sim.stats()
                best = max(best, st["flip_attempts"] / (st["sweep_device_ms"] * 1e-3))
            print(f"{tag} {'x'.join(map(str, dims))} E={E} philox{rounds} {'energies' if acc else 'sweeps  '}: "
                  f"{best:.3e} flips/s  ({1e6 * E * n / best / 2:.2f} us/phase)", flush=True)
        sim.close()

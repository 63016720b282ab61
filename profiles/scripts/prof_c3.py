"""Profiling driver: a few sweeps of BASELINE config 3 (64^3 +-J x 1024 replicas) with per-sweep
energies, so that both instantiations of the sweep kernel (plain colour phase, accumulating colour
phase) launch.  argv: [rounds] [sweeps] [replicas]."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 7
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
E = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ctx = nat.Context.get(0)
g = nat.Graph.torus(ctx, (64, 64, 64), j0=1.0, pmj=True, j_seed=2024)
sim = nat.Sim(g, E, seed=31337, rounds=rounds)
betas = np.linspace(0.6, 0.7, T)
sim.sweeps(betas, per_sweep_energies=True)
print("ok", sim.stats()["kernel_launches"])

"""Profiling driver: three sweeps of BASELINE config 2 (4096^2 ferromagnet x 1024 replicas, 2 GiB
of packed spins: the replica-packed configuration that streams from HBM)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
from pyisingmontecarlo_b200 import _native as nat

ctx = nat.Context.get(0)
g = nat.Graph.torus(ctx, (4096, 4096), j0=-1.0)
sim = nat.Sim(g, 1024, seed=1)
sim.sweeps(np.full(3, 0.43))
print("ok", sim.stats()["kernel_launches"])

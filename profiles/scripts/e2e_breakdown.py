"""Where the end-to-end time of Lattice.run_monte_carlo_annealing_and_get_energies goes beyond the
device-resident sweeps: the same steps through the Sim object, timed one by one (config 3 lattice)."""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
import pyisingmontecarlo_b200 as pkg
from pyisingmontecarlo_b200 import _native as nat

ctx = nat.Context.get(0)
T = 1000
for E in (128, 1024):
    lat = pkg.Lattice.torus((64, 64, 64), j=1.0, pmj=True, j_seed=2024, seed_gen=1)
    init = np.random.default_rng(0).integers(0, 2, 64 ** 3).astype(bool)
    stops = [(0, 0.1), (T, 1.2)]
    for rep in range(3):
        t0 = time.perf_counter()
        lat.set_initial_state(init)
        en, st = lat.run_monte_carlo_annealing_and_get_energies(stops, T, E, only_basic_moves=True)
        t_api = time.perf_counter() - t0
    g = lat.graph()
    betas = nat.schedule_betas(stops, T)
    for rep in range(2):
        t = [time.perf_counter()]
        sim = nat.Sim(g, E, 1); t.append(time.perf_counter())
        sim.set_state(init); t.append(time.perf_counter())
        en2 = sim.sweeps(betas, per_sweep_energies=True); t.append(time.perf_counter())
        dev = sim.stats()["sweep_device_ms"]
        out = nat.PinnedPool.empty((E, 64 ** 3), np.bool_); t.append(time.perf_counter())
        sim.states(out); t.append(time.perf_counter())
        sim.close(); t.append(time.perf_counter())
    names = ["create+randomize", "set_state", "sweeps+energies", "alloc out", "states", "close"]
    print(f"E={E}: API call {1e3 * t_api:.2f} ms; device sweeps {dev:.2f} ms; " +
          ", ".join(f"{n} {1e3 * (b - a):.2f}" for n, a, b in zip(names, t, t[1:])), flush=True)

import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, oracle_lib
import pyisingmontecarlo_b200 as pkg
lat = pkg.Lattice(oracle_lib.square_edges(32), seed_gen=0)
for E in (64, 1024):
    lat.run_monte_carlo(0.44, 10, E)
    for rep in range(3):
        t0 = time.perf_counter(); en, st = lat.run_monte_carlo(0.44, 1000, E); dt = time.perf_counter() - t0
    print(f"C1 32x32 E={E} 1000 sweeps: {1e3*dt:.2f} ms  ({E*1024*1000/dt:.3e} flips/s)  <e>={en.mean()/1024:.4f}")
    t0 = time.perf_counter(); en, st = lat.run_monte_carlo_annealing_and_get_energies([(0, .1), (1000, .44)], 1000, E); dt = time.perf_counter() - t0
    print(f"   annealing+energies: {1e3*dt:.2f} ms")
g = oracle_lib.Graph(oracle_lib.square_edges(32))
oracle_lib.lib().orc_set_num_threads(os.cpu_count())
t0 = time.perf_counter(); g.run_monte_carlo(0.44, 1000, oracle_lib.make_seeds(0, 64)); dt = time.perf_counter() - t0
print(f"CPU oracle ({os.cpu_count()} threads) E=64: {1e3*dt:.1f} ms")

"""Small end-to-end case touching every kernel once, for compute-sanitizer runs:
  compute-sanitizer --tool memcheck python tests/sanitizer_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
import pyisingmontecarlo_b200 as pkg  # noqa: E402
from pyisingmontecarlo_b200 import _native as nat  # noqa: E402

ctx = nat.Context.get(0)
# stencil kernels: 3D +-J (vector width 4, 2, 1), 2D uniform, per-sweep energies (fused), unpack
for dims, E in (((4, 6, 4), 130), ((4, 4, 4), 70), ((6, 4), 33), ((8, 6), 256)):
    g = nat.Graph.torus(ctx, dims, j0=1.0, pmj=len(dims) == 3, j_seed=1)
    sim = nat.Sim(g, E, 5)
    en = sim.sweeps([0.5, 0.9, 1.2], per_sweep_energies=True)
    assert (en[:, -1] == sim.energies()).all()
    sim.magnetization(); sim.states(); sim.packed()
    sim.set_state(np.ones(g.nvars, dtype=bool)); sim.set_states(sim.states())
    sim.run_sampling(0.5, 1, 2, 2)
    sim.close()
# general graph (integer classes), per-experiment betas, tempering
rng = np.random.default_rng(0)
edges = [((i, (i + 1) % 30), -1.0) for i in range(30)] + [((i, (i + 7) % 30), 1.0) for i in range(0, 30, 2)]
lat = pkg.Lattice(edges, seed_gen=1)
lat.run_monte_carlo_annealing_and_get_energies([(0, 0.2), (4, 1.0)], 4, 40)
pt = nat.Tempering(lat.graph(), np.linspace(0.2, 1.0, 9), seed=3)
pt.timesteps_sample(12, 2, 4)
# real couplings + biases
lr = pkg.Lattice([((0, 1), 0.3), ((1, 2), -1.2), ((2, 0), 0.8)], seed_gen=2)
lr.set_individual_bias(1, 0.4)
lr.run_monte_carlo_sampling(0.7, 6, 40, None, 1, 2)
# replay
og = oracle_lib.Graph(oracle_lib.square_edges(4))
sites, u, init, en_o, st_o = og.trace(0.44, oracle_lib.make_seeds(0, 5), 64)
en_r, st_r = pkg.Lattice(oracle_lib.square_edges(4), seed_gen=0).replay(0.44, sites, u, init)
assert (st_r == st_o).all()
# single lattice strips
sl = pkg.SingleLattice2D(128, 8, seed=1)
sl.sweeps([0.4, 0.5]); sl.energy(); sl.magnetization(); sl.local_rows()
# non-basic moves: bit-sliced edge moves (checkerboard and natural layout, runtime / unrolled outer
# degrees), float edge moves with importance weights, worms, the acceptance counter
for dims, gl in (((4, 6, 4), False), ((6, 4), False), ((6, 4), True)):
    g = nat.Graph.torus(ctx, dims, j0=1.0, pmj=True, j_seed=2)
    sim = nat.Sim(g, 70, 6, general_layout=gl)
    sim.set_moves(1, 2, 2, 3)
    sim.sweeps([0.4, 0.9], per_sweep_energies=True)
    sim.step_acceptance(0.5)
    sim.close()
sim = nat.Sim(lat.graph(), 40, 2)
sim.set_moves(1, 1, 1, 4)
sim.sweeps([0.5, 0.5])
sim.close()
lr.non_basic_moves = True
lr.run_monte_carlo(0.7, 3, 40, edge_move_importance_sampling=True)
# the 128-thread shape of the row walk (few site groups per thread) and the 256-thread one
for E in (128, 4096):
    g = nat.Graph.torus(ctx, (16, 16, 16), j0=1.0, pmj=True, j_seed=3)
    sim = nat.Sim(g, E, 7)
    sim.sweeps([0.5, 0.8], per_sweep_energies=True)
    sim.close()
# strips through ising_strip_sweeps; with ISING_STRIP_FUSE=1 ISING_STRIP_FUSE_MIN_ROWS=1 in the
# environment this is the fused two-colour pass (TMA-staged for Lx = 8192, warp loads for 256)
for Lx in (256, 8192):
    st = nat.Strip(ctx, Lx, 24, 0, 24, -1.0, 3, ghost=4)
    st.sweeps([0.4, 0.5, 0.6], None, 2)
    st.rows(); st.close()
print("SANITIZER_CASE_OK")

"""Regression fixtures of the two CPU checkers (the reference cannot be run here, so these pin
the oracle against accidental change rather than against the reference):
  python tests/golden/make_oracle_golden.py  ->  tests/golden/oracle_regression.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as o  # noqa: E402


def cases():
    out = {}
    # reference-algorithm oracle: BASELINE config 1 at reduced length + a biased random graph
    g = o.Graph(o.square_edges(32))
    seeds = o.make_seeds(0, 8)
    en, st = g.run_monte_carlo(0.44, 20, seeds)
    out["c1_energies"], out["c1_states"] = en, np.packbits(st, axis=1)
    sites, u, init, en_t, st_t = g.trace(0.44, seeds[:2], 4096)
    out["c1_trace_sites"], out["c1_trace_u"] = sites[:, :64], u[:, :64]
    # a complete replay case: trace prefix + initial states + the result of consuming it
    en_r, st_r = g.replay(0.44, sites[:, :512], u[:, :512], init)
    out["replay_init"], out["replay_sites"], out["replay_u"] = np.packbits(init, axis=1), sites[:, :512], u[:, :512]
    out["replay_states"], out["replay_energies"] = np.packbits(st_r, axis=1), en_r
    edges = [((0, 1), -1.0), ((1, 2), 0.37), ((2, 3), -2.2), ((3, 0), 1.0), ((0, 2), 0.9)]
    gb = o.Graph(edges, biases=[0.1, 0.0, -0.4, 0.2])
    en, st = gb.run_annealing([(0, 0.2), (30, 1.5)], 30, o.make_seeds(5, 6), per_step_energies=True)
    out["biased_energies"], out["biased_states"] = en, st
    st_pt, en_pt, swaps = o.Graph(o.square_edges(6)).pt_run(np.linspace(0.2, 0.6, 5), 99, 40, 4, 10)
    out["pt_states"], out["pt_energies"], out["pt_swaps"] = np.packbits(st_pt, axis=2), en_pt, np.array([swaps])
    # device-algorithm mirror: 3D +-J stencil colouring, general graph, tempering, single lattice
    L = (4, 4, 6)
    rng = np.random.default_rng(1)
    n = L[0] * L[1] * L[2]
    idx = lambda x, y, z: x + L[0] * (y + L[1] * z)
    a, b, j, col = [], [], [], np.zeros(n, dtype=np.uint32)
    for z in range(L[2]):
        for y in range(L[1]):
            for x in range(L[0]):
                s = idx(x, y, z)
                col[s] = (x + y + z) & 1
                for nb in (idx((x + 1) % L[0], y, z), idx(x, (y + 1) % L[1], z), idx(x, y, (z + 1) % L[2])):
                    a.append(s); b.append(nb); j.append(float(rng.choice([-1.0, 1.0])))
    en, st = o.msc_mirror(a, b, j, n, col, 70, 12345, np.linspace(0.2, 1.3, 5), per_sweep=True, rounds=10)
    out["mirror_edges"] = np.array([a, b]); out["mirror_j"] = np.array(j); out["mirror_colors"] = col
    out["mirror_energies"], out["mirror_states"] = en, np.packbits(st, axis=1)
    st_pt, en_pt, swaps, slots = o.msc_mirror_pt(a, b, j, n, col, np.geomspace(0.2, 1.4, 12), 77, 25, 3, 5, rounds=10)
    out["mirror_pt_states"], out["mirror_pt_energies"] = np.packbits(st_pt, axis=2), en_pt
    out["mirror_pt_swaps"], out["mirror_pt_slots"] = np.array([swaps]), slots
    en, st = o.msc_mirror_single(128, 8, -1.0, 9, [0.4, 0.44, 0.5], rounds=10)
    out["single_energies"], out["single_state"] = en, np.packbits(st, axis=1)
    # the same three with Philox4x32-7, the library default since round 2
    en, st = o.msc_mirror(a, b, j, n, col, 70, 12345, np.linspace(0.2, 1.3, 5), per_sweep=True, rounds=7)
    out["mirror7_energies"], out["mirror7_states"] = en, np.packbits(st, axis=1)
    st_pt, en_pt, swaps, slots = o.msc_mirror_pt(a, b, j, n, col, np.geomspace(0.2, 1.4, 12), 77, 25, 3, 5, rounds=7)
    out["mirror7_pt_states"], out["mirror7_pt_energies"] = np.packbits(st_pt, axis=2), en_pt
    out["mirror7_pt_swaps"], out["mirror7_pt_slots"] = np.array([swaps]), slots
    en, st = o.msc_mirror_single(128, 8, -1.0, 9, [0.4, 0.44, 0.5], rounds=7)
    out["single7_energies"], out["single7_state"] = en, np.packbits(st, axis=1)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **cases())
    print("written")

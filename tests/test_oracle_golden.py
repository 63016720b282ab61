"""The committed fixtures of tests/golden/oracle_regression.npz (made by make_oracle_golden.py)
pin both CPU checkers; test_gpu_golden compares the device with the same fixtures."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = np.load(os.path.join(HERE, "golden", "oracle_regression.npz"))


def test_oracle_outputs_match_committed_fixtures(oracle):
    import make_oracle_golden

    now = make_oracle_golden.cases()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(now[k], GOLD[k], equal_nan=True), k


@pytest.mark.gpu
def test_device_matches_committed_fixtures(native):
    """The same fixtures against the CUDA path through the C ABI (no oracle involved)."""
    import pyisingmontecarlo_b200 as pkg

    a, b = GOLD["mirror_edges"]
    j = GOLD["mirror_j"]
    lat = pkg.Lattice.from_arrays(a, b, j)
    g = lat.graph()
    assert g.kind == native.KIND_STENCIL3D and (g.colors() == GOLD["mirror_colors"]).all()
    for rounds, tag in ((10, ""), (7, "7")):        # 7 = the library default
        sim = native.Sim(g, 70, 12345, rounds=rounds if rounds == 10 else 0)
        en = sim.sweeps(np.linspace(0.2, 1.3, 5), per_sweep_energies=True)
        assert np.array_equal(en, GOLD[f"mirror{tag}_energies"])
        assert np.array_equal(np.packbits(sim.states(), axis=1), GOLD[f"mirror{tag}_states"])
        pt = native.Tempering(g, np.geomspace(0.2, 1.4, 12), seed=77, rounds=rounds if rounds == 10 else 0)
        st, e = pt.timesteps_sample(25, 3, 5)
        assert np.array_equal(np.packbits(st, axis=2), GOLD[f"mirror{tag}_pt_states"])
        assert np.array_equal(e, GOLD[f"mirror{tag}_pt_energies"])
        assert pt.total_swaps() == int(GOLD[f"mirror{tag}_pt_swaps"][0])
        assert (pt.slots() == GOLD[f"mirror{tag}_pt_slots"]).all()
        sl = pkg.SingleLattice2D(128, 8, seed=9, rounds=rounds if rounds == 10 else 0)
        ens = []
        for beta in (0.4, 0.44, 0.5):
            sl.sweeps([beta])
            ens.append(sl.energy())
        assert ens == list(GOLD[f"single{tag}_energies"])
        assert np.array_equal(np.packbits(sl.local_rows(), axis=1), GOLD[f"single{tag}_state"])
    # replay of the reference-algorithm trace fixture (512 attempts of two experiments)
    from oracle_lib import square_edges  # lattice helper only
    lat1 = pkg.Lattice(square_edges(32), seed_gen=0)
    init = np.unpackbits(GOLD["replay_init"], axis=1)[:, :1024].astype(bool)
    en_r, st_r = lat1.replay(0.44, GOLD["replay_sites"], GOLD["replay_u"], init)
    assert np.array_equal(np.packbits(st_r, axis=1), GOLD["replay_states"])
    assert np.array_equal(en_r, GOLD["replay_energies"])

"""Non-basic moves on the GPU (csrc/moves.cu through ising_sim_set_moves and the Lattice /
ClassicIsing faces): two-spin edge moves over a strong edge colouring, worm moves, importance
sampling.  Chains made of those moves reproduce exact enumeration; deterministic limits are
checked bit for bit."""
import warnings

import numpy as np
import pytest

from moves_cases import boltzmann, histogram_z, irregular_graph

pytestmark = pytest.mark.gpu


def _graph(native, edges, nvars, biases=None):
    ctx = native.Context.get(0)
    a = np.array([e[0][0] for e in edges], dtype=np.uint64)
    b = np.array([e[0][1] for e in edges], dtype=np.uint64)
    j = np.array([e[1] for e in edges], dtype=np.float64)
    return native.Graph.from_edges(ctx, nvars, a, b, j, None if biases is None else np.asarray(biases, float))


@pytest.mark.parametrize("moves", [
    dict(spin_sweeps=0, edge_passes=1, worms=2, worm_len=3),
    dict(spin_sweeps=0, edge_passes=1, worms=2, worm_len=3, edge_importance=True),
    dict(spin_sweeps=0, worms=6, worm_len=1),
    dict(spin_sweeps=0, edge_passes=1, worms=4, worm_len=4),
    dict(spin_sweeps=1, edge_passes=2, worms=1, worm_len=2),
])
def test_moves_sample_the_boltzmann_law(native, moves):
    edges, n, biases = irregular_graph()
    beta = 0.8
    g = _graph(native, edges, n, biases)
    E = 32768
    sim = native.Sim(g, E, 2025)
    sim.set_moves(**moves)
    sim.sweeps(np.full(80, beta))
    st = sim.states()
    _, p, en_exact = boltzmann(edges, n, beta, biases)
    if moves["spin_sweeps"] == 0 and moves["worm_len"] % 2 == 0:
        par = st.sum(1) % 2
        parity_of_state = np.array([bin(i).count("1") % 2 for i in range(2 ** n)])
        for sector in (0, 1):
            ps = np.where(parity_of_state == sector, p, 0.0)
            z = histogram_z(st[par == sector], ps / ps.sum())[parity_of_state == sector]
            assert np.abs(z).max() < 4.5, z
    else:
        z = histogram_z(st, p)
        assert np.abs(z).max() < 4.5, z
    # energies reported by the library belong to the returned states
    from moves_cases import state_index
    assert np.allclose(sim.energies(), en_exact[state_index(st)], atol=1e-9)
    stats = sim.stats()
    assert stats["edge_attempts"] == 80 * moves.get("edge_passes", 0) * len(edges) * E
    assert stats["worm_attempts"] == 80 * moves.get("worms", 0) * E


@pytest.mark.parametrize("case", ["torus2d", "torus3d_pmj", "regular3", "irregular_pm", "general_layout"])
def test_bit_sliced_edge_moves_match_the_mirror(native, oracle, case):
    """Graphs with equal |J| and no bias take the bit-sliced edge-move kernel (exact integer
    thresholds): states and per-timestep energies equal the CPU mirror's bit for bit, for the
    checkerboard and the natural spin layout, odd experiment counts, several passes."""
    ctx = native.Context.get(0)
    rng = np.random.default_rng(12)
    general_layout = False
    if case == "torus2d":
        g, E, passes = native.Graph.torus(ctx, (8, 6), j0=-1.0), 70, 1
    elif case == "torus3d_pmj":
        g, E, passes = native.Graph.torus(ctx, (4, 6, 4), j0=1.0, pmj=True, j_seed=5), 33, 2
    elif case == "general_layout":
        g, E, passes, general_layout = native.Graph.torus(ctx, (6, 4), j0=1.0, pmj=True, j_seed=9), 64, 1, True
    elif case == "regular3":
        n = 60
        while True:
            stubs = np.repeat(np.arange(n), 3)
            rng.shuffle(stubs)
            a, b = stubs[0::2], stubs[1::2]
            if not (a == b).any() and len({(min(x, y), max(x, y)) for x, y in zip(a, b)}) == len(a):
                break
        g = native.Graph.from_edges(ctx, n, a.astype(np.uint64), b.astype(np.uint64), np.full(len(a), -1.0))
        E, passes = 96, 1
    else:   # irregular degrees (a leaf, an isolated pair), mixed signs, a multi-edge
        edges = [(0, 1), (1, 2), (2, 0), (2, 3), (3, 4), (4, 5), (5, 6), (6, 3), (7, 8), (1, 4), (1, 4), (9, 5)]
        a = np.array([e[0] for e in edges], dtype=np.uint64)
        b = np.array([e[1] for e in edges], dtype=np.uint64)
        j = np.where(rng.random(len(edges)) < 0.5, 1.0, -1.0)
        g = native.Graph.from_edges(ctx, 10, a, b, j)
        E, passes = 40, 3
    betas = np.array([0.2, 0.5, 0.9, 1.4, 0.7])
    sim = native.Sim(g, E, 31, general_layout=general_layout)
    sim.set_moves(1, passes, 0)
    en = sim.sweeps(betas, per_sweep_energies=True)
    st = sim.states()
    ea, eb, ej = g.edges()
    en_ref, st_ref = oracle.msc_mirror_moves(ea, eb, ej, g.nvars, g.colors(), g.edge_classes(), E, 31, betas,
                                             spin_sweeps=1, edge_passes=passes, per_step=True)
    assert (st == st_ref).all() and (en == en_ref).all()
    # edge moves alone (no sweep), from the state reached
    sim2 = native.Sim(g, E, 77, general_layout=general_layout)
    sim2.set_moves(0, 1, 0)
    sim2.sweeps(betas[:3])
    _, st2 = oracle.msc_mirror_moves(ea, eb, ej, g.nvars, g.colors(), g.edge_classes(), E, 77, betas[:3],
                                     spin_sweeps=0, edge_passes=1)
    assert (sim2.states() == st2).all()


def test_strong_edge_colouring_is_strong(native):
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (6, 4, 4), j0=1.0, pmj=True, j_seed=1)
    ea, eb, _ = g.edges()
    cls = g.edge_classes()
    adj = {}
    for x, y in zip(ea, eb):
        adj.setdefault(int(x), set()).add(int(y))
        adj.setdefault(int(y), set()).add(int(x))
    for c in range(int(cls.max()) + 1):
        members = [(int(ea[i]), int(eb[i])) for i in np.nonzero(cls == c)[0]]
        touched = {}
        for k, (x, y) in enumerate(members):
            for v in (x, y):
                assert v not in touched, "two bonds of a class share a site"
                touched[v] = k
        for k, (x, y) in enumerate(members):
            for v in (x, y):
                for u in adj[v]:
                    assert touched.get(u, k) == k, "a bond joins two bonds of a class"


def test_edge_pass_at_beta_zero_flips_every_site_degree_times(native):
    # at beta = 0 every move is accepted: one pass over the bonds flips a site once per incident
    # bond, whatever the couplings - the strong edge colouring covers every bond exactly once
    rng = np.random.default_rng(5)
    n = 200
    edges = [((i, i + 1), float(rng.normal())) for i in range(n - 1)]
    edges += [((int(u), int(v)), float(rng.normal())) for u, v in rng.integers(0, n, (150, 2)) if u != v]
    g = _graph(native, edges, n)
    deg = np.zeros(n, dtype=int)
    for (u, v), _ in edges:
        deg[u] += 1
        deg[v] += 1
    sim = native.Sim(g, 70, 1)
    before = sim.states()
    sim.set_moves(spin_sweeps=0, edge_passes=1)
    sim.sweeps(np.zeros(1))
    after = sim.states()
    assert ((before ^ after) == (deg % 2 == 1)[None, :]).all()
    sim.set_moves(spin_sweeps=0, edge_passes=3)
    sim.sweeps(np.zeros(1))
    assert ((sim.states() ^ after) == (deg % 2 == 1)[None, :]).all()


def test_moves_on_the_stencil_layout_and_zero_temperature(native, oracle):
    # 2D torus in the checkerboard layout: the moves address sites through the layout map.
    # beta -> infinity: no move may raise the energy, whatever mix of moves runs
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (8, 6), j0=1.0, pmj=True, j_seed=3)
    sim = native.Sim(g, 96, 9)
    sim.set_moves(spin_sweeps=1, edge_passes=1, worms=3, worm_len=4)
    en = sim.sweeps(np.full(25, 1e6), per_sweep_energies=True)
    assert (np.diff(en, axis=1) <= 0).all()
    a, b, j = g.edges()
    og = oracle.Graph(arrays=(a, b, j), nvars=g.nvars)
    st = sim.states()
    assert all(og.energy(st[e]) == en[e, -1] for e in range(0, 96, 7))


def test_torus_4x4_all_moves_vs_enumeration(native):
    edges = [((x + 4 * y, (x + 1) % 4 + 4 * y), -1.0) for x in range(4) for y in range(4)]
    edges += [((x + 4 * y, x + 4 * ((y + 1) % 4)), -1.0) for x in range(4) for y in range(4)]
    beta = 0.35
    _, p, en = boltzmann(edges, 16, beta)
    exact_e = float((p * en).sum())
    var_e = float((p * en * en).sum()) - exact_e ** 2
    g = _graph(native, edges, 16)          # recognised as a torus: stencil layout
    E = 8192
    sim = native.Sim(g, E, 77)
    sim.set_moves(spin_sweeps=0, edge_passes=1, worms=2, worm_len=3)
    sim.sweeps(np.full(150, beta))
    e = sim.energies()
    assert abs(e.mean() - exact_e) < 4 * np.sqrt(var_e / E), (e.mean(), exact_e)


def test_lattice_and_classic_faces(native):
    import pyisingmontecarlo_b200 as pkg

    edges, n, biases = irregular_graph()
    beta = 0.8
    _, p, en = boltzmann(edges, n, beta, biases)
    exact_e = float((p * en).sum())
    sd = np.sqrt(float((p * en * en).sum()) - exact_e ** 2)
    lat = pkg.Lattice(edges, seed_gen=5)
    for i, b in enumerate(biases):
        lat.set_individual_bias(i, b)
    lat.non_basic_moves = True
    with warnings.catch_warnings():
        warnings.simplefilter("error")      # with the flag set nothing is left to warn about
        e1, _ = lat.run_monte_carlo(beta, 60, 8192)
        e2, _ = lat.run_monte_carlo(beta, 60, 8192, edge_move_importance_sampling=True)
        e3, _ = lat.run_monte_carlo_annealing_and_get_energies([(0, beta), (60, beta)], 60, 4096)
    for e in (e1, e2, e3[:, -1]):
        assert abs(e.mean() - exact_e) < 4.5 * sd / np.sqrt(len(e)), (e.mean(), exact_e)
    lat.non_basic_moves = False
    with pytest.raises(NotImplementedError):
        lat.run_monte_carlo(beta, 1, 32, edge_move_importance_sampling=True)
    # only_basic_moves=True: importance sampling has no move to act on (the reference ignores it too)
    lat.run_monte_carlo(beta, 1, 32, only_basic_moves=True, edge_move_importance_sampling=True)

    ci = pkg.ClassicIsing(edges, None, 8192, seed=3)
    ci.worm_len = 3
    ci.run_monte_carlo(beta, 60, nspinupdates=0, nedgeupdates=2 * len(edges), nwormupdates=2)
    e = ci.get_energies()
    exact0 = float((boltzmann(edges, n, beta)[1] * boltzmann(edges, n, beta)[2]).sum())
    sd0 = np.sqrt(float((boltzmann(edges, n, beta)[1] * boltzmann(edges, n, beta)[2] ** 2).sum()) - exact0 ** 2)
    assert abs(e.mean() - exact0) < 4.5 * sd0 / np.sqrt(len(e))
    with pytest.raises(NotImplementedError):
        ci.run_monte_carlo(beta, 1, nedgeupdates=3)


def test_step_acceptance_counts_the_flipped_spins(native):
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (16, 16), j0=-1.0)
    sim = native.Sim(g, 70, 4)
    # beta = 0: every attempt is accepted, every site is attempted once per sweep
    assert (sim.step_acceptance(0.0) == g.nvars).all()
    before = sim.states()
    acc = sim.step_acceptance(0.6)
    assert (acc == (before != sim.states()).sum(axis=1)).all()
    assert 0 < acc.mean() < g.nvars
    # the ordered ferromagnet at a very low temperature does not move
    sim.set_state(np.ones(g.nvars, dtype=bool))
    assert (sim.step_acceptance(50.0) == 0).all()
    assert sim.counter == 3

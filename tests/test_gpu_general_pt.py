"""GPU parity tests for arbitrary graphs (CSR/ELL colour-class kernel) and parallel tempering."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(native):
    import pyisingmontecarlo_b200 as pkg

    native.Context.get(0)
    return pkg


def random_regular(n, d, rng):
    """pairing model with rejection of self loops / multi-edges"""
    while True:
        stubs = np.repeat(np.arange(n), d)
        rng.shuffle(stubs)
        pairs = stubs.reshape(-1, 2)
        if (pairs[:, 0] == pairs[:, 1]).any():
            continue
        key = np.sort(pairs, axis=1)
        if len(np.unique(key, axis=0)) != len(key):
            continue
        return [(int(a), int(b)) for a, b in pairs]


def irregular_graph(n, rng, maxdeg=7):
    edges = set()
    for i in range(n):
        for j in rng.choice(n, rng.integers(1, 4), replace=False):
            if i != j:
                edges.add((min(i, int(j)), max(i, int(j))))
    deg = np.zeros(n, int)
    out = []
    for a, b in sorted(edges):
        if deg[a] < maxdeg and deg[b] < maxdeg:
            out.append((a, b))
            deg[a] += 1
            deg[b] += 1
    return out


def _check_vs_mirror(native, oracle, graph, E, seed, betas, planes=6, rounds=7, general=False):
    a, b, j = graph.edges()
    sim = native.Sim(graph, E, seed, planes=planes, rounds=rounds, general_layout=general)
    en = sim.sweeps(betas, per_sweep_energies=True)
    st = sim.states()
    en_ref, st_ref = oracle.msc_mirror(a, b, j, graph.nvars, graph.colors(), E, seed, betas,
                                       planes=planes, rounds=rounds, per_sweep=True)
    assert (st == st_ref).all()
    assert (en == en_ref).all()
    assert (sim.energies() == en_ref[:, -1]).all()
    return sim


def test_random_regular_graph_matches_mirror(native, oracle, pkg):
    """BASELINE config 4's graph family at test size: random 3-regular, J = -1."""
    rng = np.random.default_rng(1)
    pairs = random_regular(300, 3, rng)
    lat = pkg.Lattice([(p, -1.0) for p in pairs], seed_gen=3)
    g = lat.graph()
    assert g.kind == native.KIND_GENERAL and g.max_degree == 3 and g.ncolors >= 3
    colors = g.colors()
    a, b, _ = g.edges()
    assert (colors[a.astype(int)] != colors[b.astype(int)]).all()   # proper colouring
    _check_vs_mirror(native, oracle, g, 70, 11, np.linspace(0.1, 1.5, 6))
    _check_vs_mirror(native, oracle, g, 33, 12, [0.8, 0.8], planes=5, rounds=10)


def test_irregular_pmj_graph_matches_mirror(native, oracle, pkg):
    rng = np.random.default_rng(2)
    pairs = irregular_graph(150, rng)
    edges = [(p, float(rng.choice([-0.5, 0.5]))) for p in pairs]
    g = pkg.Lattice(edges).graph()
    assert g.kind == native.KIND_GENERAL and g.integer_classes
    _check_vs_mirror(native, oracle, g, 64, 5, [0.3, 1.1, 2.0, 0.6])
    # bipartite irregular graph: a tree gets exactly two colours
    tree = [((i, (i - 1) // 2), 1.0) for i in range(1, 64)]
    gt = pkg.Lattice(tree).graph()
    assert gt.ncolors == 2
    _check_vs_mirror(native, oracle, gt, 40, 6, [0.5, 0.9, 1.4])


def test_isolated_and_leaf_sites(native, oracle, pkg):
    """Index gaps make isolated variables (nvars = max index + 1, lattice.rs:51-55): dE = 0, so the
    Metropolis rule flips them at every attempt, in the mirror and on the device alike."""
    edges = [((0, 1), -1.0), ((1, 2), -1.0), ((5, 6), 1.0), ((9, 2), -1.0)]   # 3, 4, 7, 8 isolated
    lat = pkg.Lattice(edges, seed_gen=3)
    assert lat.nvars == 10
    g = lat.graph()
    _check_vs_mirror(native, oracle, g, 40, 21, [0.5, 0.9, 0.2])
    lat.set_initial_state([True] * 10)
    _, st = lat.run_monte_carlo(0.7, 3, 8)
    assert (~st[:, [3, 4, 7, 8]]).all()        # three forced flips from the all-up start


def test_general_layout_equals_stencil_path(native, oracle, pkg):
    """The same torus through the checkerboard stencil kernels and through the general kernels:
    both implement one algorithm, so the bits must agree."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (6, 4, 8), j0=1.0, pmj=True, j_seed=9)
    betas = np.linspace(0.2, 1.2, 5)
    s1 = native.Sim(g, 96, 77)
    s2 = native.Sim(g, 96, 77, general_layout=True)
    e1 = s1.sweeps(betas, per_sweep_energies=True)
    e2 = s2.sweeps(betas, per_sweep_energies=True)
    assert (s1.states() == s2.states()).all() and (e1 == e2).all()
    assert (s1.magnetization() == s2.magnetization()).all()
    g2 = native.Graph.torus(ctx, (8, 6), j0=-1.0)
    s1, s2 = native.Sim(g2, 40, 5), native.Sim(g2, 40, 5, general_layout=True)
    s1.sweeps([0.44] * 4)
    s2.sweeps([0.44] * 4)
    assert (s1.states() == s2.states()).all()


def test_lattice_api_on_general_graph(pkg, oracle):
    rng = np.random.default_rng(4)
    pairs = random_regular(12, 3, rng)
    edges = [(p, 1.0 if k % 3 else -1.0) for k, p in enumerate(pairs)]
    lat = pkg.Lattice(edges, seed_gen=8)
    en, st = lat.run_monte_carlo(0.7, 400, 4096)
    g = oracle.Graph(edges)
    assert all(en[k] == g.energy(st[k]) for k in range(0, 4096, 512))
    # exact <E> by enumeration
    a = np.array([e[0][0] for e in edges]); b = np.array([e[0][1] for e in edges])
    jj = np.array([e[1] for e in edges])
    s = np.array(list(itertools.product([-1, 1], repeat=12)), dtype=float)
    E = (s[:, a] * s[:, b] * jj).sum(1)
    w = np.exp(-0.7 * (E - E.min())); w /= w.sum()
    exact, var = (w * E).sum(), (w * E * E).sum() - (w * E).sum() ** 2
    assert abs(en.mean() - exact) < 4 * np.sqrt(var / 4096), (en.mean(), exact)
    en2, st2 = lat.run_monte_carlo_annealing_and_get_energies([(0, 0.2), (10, 0.9)], 10, 6)
    assert en2.shape == (6, 10) and all(en2[k, -1] == g.energy(st2[k]) for k in range(6))


def _exact(edges, n, beta, biases=None):
    a = np.array([e[0][0] for e in edges]); b = np.array([e[0][1] for e in edges])
    jj = np.array([e[1] for e in edges], dtype=float)
    bias = np.zeros(n) if biases is None else np.asarray(biases, float)
    s = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=float)
    E = (s[:, a] * s[:, b] * jj).sum(1) - s @ bias
    w = np.exp(-beta * (E - E.min())); w /= w.sum()
    mean = (w * E).sum()
    return mean, np.sqrt((w * E * E).sum() - mean**2), w @ s


@pytest.mark.parametrize("case", ["k4_mixed_bias", "ring_real", "star_degree20", "global_bias_torus"])
def test_real_couplings_and_biases(pkg, oracle, case):
    """Inputs outside the integer-class kernels (real J, biases, degree > 15) run on the
    float-field kernel: energies are those of the returned states, <E> and <s_i> agree with
    exact enumeration."""
    rng = np.random.default_rng(11)
    if case == "k4_mixed_bias":
        edges = [((0, 1), -1.0), ((0, 2), 0.5), ((0, 3), 1.5), ((1, 2), -0.7), ((1, 3), 0.3), ((2, 3), 1.0)]
        n, biases, beta = 4, [0.3, -0.2, 0.0, 0.6], 0.5
    elif case == "ring_real":
        n = 10
        edges = [((i, (i + 1) % n), float(rng.normal())) for i in range(n)]
        biases, beta = None, 0.8
    elif case == "star_degree20":
        n = 15
        edges = [((0, i), -1.0) for i in range(1, n)] + [((1, i), 1.0) for i in range(2, 9)] + \
                [((0, i), -1.0) for i in range(1, 8)]          # multi-edges: hub degree 21
        biases, beta = None, 0.25
    else:
        edges, n, biases, beta = oracle.square_edges(4), 16, [0.15] * 16, 0.35
    lat = pkg.Lattice(edges, seed_gen=5)
    if biases is not None:
        if case == "global_bias_torus":
            lat.set_global_bias(0.15)
        else:
            for v, b in enumerate(biases):
                lat.set_individual_bias(v, b)
    E = 8192
    en, st = lat.run_monte_carlo(beta, 300, E)
    g = oracle.Graph(edges, nvars=n, biases=biases)
    for k in range(0, E, 1024):
        assert abs(en[k] - g.energy(st[k])) < 1e-9 * max(1.0, abs(en[k]))
    mean, sd, mag = _exact(edges, n, beta, biases)
    assert abs(en.mean() - mean) < 4 * sd / np.sqrt(E) + 1e-6, (en.mean(), mean, sd)
    m = (st * 2.0 - 1).mean(0)
    assert np.abs(m - mag).max() < 5 / np.sqrt(E), (m, mag)
    # the other drivers work on this path too
    en2, st2 = lat.run_monte_carlo_annealing_and_get_energies([(0, 0.1), (5, beta)], 5, 64)
    assert en2.shape == (64, 5) and abs(en2[3, -1] - g.energy(st2[3])) < 1e-9 * max(1.0, abs(en2[3, -1]))
    en3, st3 = lat.run_monte_carlo_sampling(beta, 6, 32, None, 1, 2)
    assert en3.shape == (32, 3) and abs(en3[5, 1] - g.energy(st3[5, 1])) < 1e-9 * max(1.0, abs(en3[5, 1]))


def test_per_experiment_betas_match_mirror(native, oracle, pkg):
    """One inverse temperature per replica bit, on the checkerboard kernels and on the general
    kernels: both equal the mirror (and therefore each other)."""
    ctx = native.Context.get(0)
    for dims, E in (((4, 4, 6), 45), ((6, 8), 130), ((4, 6, 4), 256)):
        g = native.Graph.torus(ctx, dims, j0=1.0, pmj=len(dims) == 3, j_seed=21)
        betas = np.linspace(0.1, 1.6, E)
        a, b, j = g.edges()
        en_ref, st_ref = oracle.msc_mirror(a, b, j, g.nvars, g.colors(), E, 99, None,
                                           per_replica_beta=betas, nsweeps=7)
        for general in (False, True):
            sim = native.Sim(g, E, 99, general_layout=general)
            sim.set_betas(betas)
            sim.sweeps(7)
            assert (sim.states() == st_ref).all() and (sim.energies() == en_ref).all()
            if not general:   # fused per-sweep energies with per-replica thresholds
                sim2 = native.Sim(g, E, 99)
                sim2.set_betas(betas)
                en = sim2.sweeps(7, per_sweep_energies=True)
                assert (en[:, -1] == en_ref).all() and (sim2.states() == st_ref).all()
            with pytest.raises(ValueError):
                sim.sweeps([0.5])          # a per-beta sim takes a sweep count
    with pytest.raises(NotImplementedError):
        native.Sim(g, E, 99, planes=7).set_betas(betas)   # lattice tables exist for 6 planes


def test_lattice_tempering_uses_checkerboard_kernels(native, oracle, pkg):
    """Parallel tempering of a 3D +-J lattice (the common spin-glass use) against the mirror."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (4, 4, 4), j0=1.0, pmj=True, j_seed=5)
    betas = np.geomspace(0.2, 1.4, 40)
    pt = native.Tempering(g, betas, seed=2718)
    states, energies = pt.timesteps_sample(30, replica_swap_freq=2, sampling_freq=10)
    a, b, j = g.edges()
    st_ref, en_ref, swaps_ref, slots_ref = oracle.msc_mirror_pt(
        a, b, j, g.nvars, g.colors(), betas, 2718, 30, replica_swap_freq=2, sampling_freq=10)
    assert (states == st_ref).all() and (energies == en_ref).all()
    assert pt.total_swaps() == swaps_ref > 0 and (pt.slots() == slots_ref).all()


def test_parallel_tempering_matches_mirror(native, oracle, pkg):
    rng = np.random.default_rng(6)
    pairs = random_regular(60, 3, rng)
    edges = [(p, -1.0) for p in pairs]
    g = pkg.Lattice(edges).graph()
    betas = np.geomspace(0.1, 1.5, 20)
    pt = native.Tempering(g, betas, seed=314)
    states, energies = pt.timesteps_sample(37, replica_swap_freq=3, sampling_freq=5)
    a, b, j = g.edges()
    st_ref, en_ref, swaps_ref, slots_ref = oracle.msc_mirror_pt(
        a, b, j, g.nvars, g.colors(), betas, 314, 37, replica_swap_freq=3, sampling_freq=5)
    assert states.shape == (20, 7, 60)
    assert (states == st_ref).all()
    assert (energies == en_ref).all()
    assert pt.total_swaps() == swaps_ref and swaps_ref > 0
    assert (pt.slots() == slots_ref).all()


def test_sharded_tempering_equals_single(native, pkg):
    """Two shards holding configurations [0, 12) and [12, 40) stepped in lock-step with a host
    concatenation of their energies (what the all-gather does) == one unsharded ladder."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (4, 4, 4), j0=1.0, pmj=True, j_seed=2)
    betas = np.linspace(0.2, 1.4, 40)
    full = native.Tempering(g, betas, seed=5)
    lo = native.Tempering(g, betas, seed=5, cfg_lo=0, cfg_hi=12)
    hi = native.Tempering(g, betas, seed=5, cfg_lo=12, cfg_hi=40)
    for step in range(6):
        ef = full.sweeps(2)
        es = np.concatenate([lo.sweeps(2), hi.sweeps(2)])
        assert (ef == es).all()
        full.swap_step(ef)
        lo.swap_step(es)
        hi.swap_step(es)
        assert (full.slots() == lo.slots()).all() and (full.slots() == hi.slots()).all()
    assert (full.local_states() == np.concatenate([lo.local_states(), hi.local_states()])).all()
    assert full.total_swaps() == lo.total_swaps() == hi.total_swaps() > 0


def test_lattice_tempering_class_and_statistics(pkg, oracle):
    """tempering.rs surface; time-averaged energies per beta against exact enumeration."""
    edges = oracle.square_edges(4)      # 16 spins: exact by enumeration
    lt = pkg.LatticeTempering(edges, seed=7)
    betas = [0.2, 0.3, 0.4, 0.5, 0.6]
    for b in betas:
        lt.add_graph(0.0, 0.0, b)
    with pytest.raises(NotImplementedError):
        lt.add_graph(1.0, 0.0, 0.5)
    assert lt.get_num_graphs() == 5
    lt.qmc_timesteps(200)
    states, energies = lt.qmc_timesteps_sample(20000, replica_swap_freq=2, sampling_freq=2000)
    assert states.shape == (5, 10, 16) and states.dtype == np.bool_ and energies.shape == (5,)
    assert lt.get_total_swaps() > 100
    a = np.array([e[0][0] for e in edges]); b = np.array([e[0][1] for e in edges])
    s = np.array(list(itertools.product([-1, 1], repeat=16)), dtype=float)
    E = -(s[:, a] * s[:, b]).sum(1)
    for beta, got in zip(betas, energies):
        w = np.exp(-beta * (E - E.min())); w /= w.sum()
        exact = (w * E).sum()
        sd = np.sqrt((w * E * E).sum() - exact**2)
        assert abs(got - exact) < 0.15 * sd + 0.05, (beta, got, exact)
    with pytest.raises(ValueError):
        lt.qmc_timesteps_sample(10, replica_swap_freq=0)


def test_classic_ising_stateful_api(pkg, oracle, native):
    """classicising.rs: state persists between calls; sampling continues from it."""
    edges = oracle.square_edges(8)
    ci = pkg.ClassicIsing(edges, None, 5, 42)
    s0 = ci.get_states()
    assert s0.shape == (5, 64)
    ci.run_monte_carlo(0.4, 10)
    s1 = ci.get_states()
    assert (s1 != s0).any()
    # identical to one uninterrupted sim of the same seed (the state really is resident)
    sim = native.Sim(pkg.Lattice(edges).graph(), 5, 42)
    sim.sweeps([0.4] * 10)
    assert (sim.states() == s1).all()
    en, st = ci.run_monte_carlo_sampling(0.4, 6, None, None, None, None, 2, 3)
    assert en.shape == (5, 2) and st.shape == (5, 2, 64)
    sim.sweeps([0.4] * (2 + 6))
    assert (st[:, 1] == sim.states()).all() and (ci.get_states() == st[:, 1]).all()
    g = oracle.Graph(edges)
    assert en[2, 1] == g.energy(st[2, 1])
    # add_graph keeps the existing experiments and appends one with the given state
    init = np.arange(64) % 2 == 0
    ci.add_graph(init)
    s2 = ci.get_states()
    assert s2.shape == (6, 64) and (s2[:5] == st[:, 1]).all() and (s2[5] == init).all()
    # ... and keeps their random streams running: run -> add_graph -> run equals the
    # uninterrupted run for the old experiments (no replay of the draws of sweeps 0..T-1)
    ci.run_monte_carlo(0.4, 4)
    sim.sweeps([0.4] * 4)
    assert (ci.get_states()[:5] == sim.states()).all()
    with pytest.raises(NotImplementedError):
        ci.run_monte_carlo(0.4, 1, 5)
    cb = pkg.ClassicIsing([((0, 1), 1.0), ((1, 2), 1.0)], 0.7, 2000, 3)   # longitudinal field
    cb.run_monte_carlo(0.5, 200)
    mean, sd, mag = _exact([((0, 1), 1.0), ((1, 2), 1.0)], 3, 0.5, [0.7] * 3)
    assert abs(cb.get_energies().mean() - mean) < 4 * sd / np.sqrt(2000)


def test_checkpoints_resume_bit_for_bit(pkg, oracle, tmp_path):
    """save_to_file / read_from_file (tempering.rs:307-347 shape): the RNG is counter-based, so a
    restored run is indistinguishable from an uninterrupted one."""
    edges = oracle.square_edges(6)
    a = pkg.ClassicIsing(edges, None, 37, 9)
    a.run_monte_carlo(0.45, 7)
    path = str(tmp_path / "classic.npz")
    a.save_to_file(path)
    b = pkg.ClassicIsing.read_from_file(path)
    assert (b.get_states() == a.get_states()).all()
    a.run_monte_carlo(0.5, 6)
    b.run_monte_carlo(0.5, 6)
    assert (b.get_states() == a.get_states()).all() and (b.get_energies() == a.get_energies()).all()
    c = pkg.ClassicIsing.read_from_file(path, reseed=123)     # new streams from the saved state
    c.run_monte_carlo(0.5, 6)
    assert (c.get_states() != a.get_states()).any()

    def ladder():
        lt = pkg.LatticeTempering(edges, seed=4)
        for beta in np.linspace(0.2, 0.7, 9):
            lt.add_graph(0.0, 0.0, beta)
        return lt

    x, y = ladder(), ladder()
    sx1, ex1 = x.qmc_timesteps_sample(20, 3, 5)
    path = str(tmp_path / "pt.npz")
    x.save_to_file(path)
    z = pkg.LatticeTempering.read_from_file(path)
    assert z.get_total_swaps() == x.get_total_swaps() > 0
    sx2, ex2 = x.qmc_timesteps_sample(21, 3, 7)
    sz2, ez2 = z.qmc_timesteps_sample(21, 3, 7)
    assert (sx2 == sz2).all() and (ex2 == ez2).all() and x.get_total_swaps() == z.get_total_swaps()
    with pytest.raises(IOError):
        pkg.ClassicIsing.read_from_file(path)


def test_tempering_on_real_couplings_and_a_field(pkg, oracle, native):
    """A ladder over a graph outside the integer-class kernels (Gaussian couplings, a longitudinal
    field: tempering.rs:70-113 takes a longitudinal field per replica) runs on the float-field kernel
    with one inverse temperature per replica bit: time-averaged energies per beta against exact
    enumeration, samples are valid configurations, and per-experiment betas without swaps agree too."""
    rng = np.random.default_rng(3)
    n = 9
    pairs = [(i, (i + 1) % n) for i in range(n)] + [(0, 4), (2, 7), (3, 8)]
    edges = [(p, float(rng.normal())) for p in pairs]
    field = 0.3
    betas = [0.2, 0.35, 0.5, 0.7, 0.9, 1.2]
    lt = pkg.LatticeTempering(edges, seed=21)
    for b in betas:
        lt.add_graph(0.0, field, b)
    with pytest.raises(NotImplementedError):
        lt.add_graph(0.0, 0.1, 0.5)          # one field per ladder
    lt.qmc_timesteps(300)
    states, energies = lt.qmc_timesteps_sample(40000, replica_swap_freq=2, sampling_freq=4000)
    assert states.shape == (6, 10, n) and energies.shape == (6,)
    assert lt.get_total_swaps() > 1000
    for beta, got in zip(betas, energies):
        mean, sd, _ = _exact(edges, n, beta, [field] * n)
        assert abs(got - mean) < 0.15 * sd + 0.05, (beta, got, mean, sd)
    # a clone continues on the same kernels
    cp = lt.clone()
    s1, e1 = lt.qmc_timesteps_sample(40, 2, 20)
    s2, e2 = cp.qmc_timesteps_sample(40, 2, 20)
    assert s1.shape == s2.shape == (6, 2, n)
    # per-experiment betas without swaps (ising_sim_set_betas on a real-coupling sim)
    lat = pkg.Lattice(edges, seed_gen=2)
    lat.set_global_bias(field)
    E = 4096
    sim = native.Sim(lat.graph(), E, 17)
    per = np.where(np.arange(E) % 2 == 0, 0.35, 0.9)
    sim.set_betas(per)
    sim.sweeps(400)
    en = sim.energies()
    for beta, sel in ((0.35, slice(0, None, 2)), (0.9, slice(1, None, 2))):
        mean, sd, _ = _exact(edges, n, beta, [field] * n)
        assert abs(en[sel].mean() - mean) < 4.5 * sd / np.sqrt(E / 2) + 1e-6, (beta, en[sel].mean(), mean)
    sim.close()


def test_lattice_tempering_clone_and_graph_itime(pkg, oracle):
    """tempering.rs:302-304 (clone) and 119-148 (get_graph_itime, one slice for a classical replica):
    a clone owns its own device state and continues exactly like the original; graph g is the
    configuration currently at beta_g, i.e. row g of a sample taken at the same time."""
    edges = oracle.square_edges(6)
    lt = pkg.LatticeTempering(edges, seed=11)
    for beta in np.linspace(0.2, 0.8, 7):
        lt.add_graph(0.0, 0.0, beta)
    assert lt.clone().get_num_graphs() == 7           # before the first run: nothing on the device yet
    states, _ = lt.qmc_timesteps_sample(12, 2, 12)
    assert lt.get_total_swaps() > 0
    for g in range(7):
        it = lt.get_graph_itime(g)
        assert it.shape == (1, 36) and it.dtype == np.bool_ and (it[0] == states[g, 0]).all()
    with pytest.raises(ValueError, match="Attempted to get graph 7 of 7"):
        lt.get_graph_itime(7)
    cp = lt.clone()
    assert cp.get_total_swaps() == lt.get_total_swaps()
    s1, e1 = lt.qmc_timesteps_sample(15, 3, 5)
    for g in range(7):                                # the clone did not move with the original
        assert (cp.get_graph_itime(g)[0] == states[g, 0]).all()
    s2, e2 = cp.qmc_timesteps_sample(15, 3, 5)
    assert (s1 == s2).all() and (e1 == e2).all() and cp.get_total_swaps() == lt.get_total_swaps()


def test_reference_readme_usage_runs_unchanged(oracle):
    """The usage section of the reference's README.md (lines 44-62), verbatim through the
    `py_monte_carlo` module name: same calls, same return layouts; energies equal the Hamiltonian
    of the returned states (README.md:46, "J*Sza*Szb so positive is antiferromagnetic")."""
    import py_monte_carlo

    edges = [
        ((0, 1), 1.0),
        ((1, 2), -1.0)
    ]
    lat = py_monte_carlo.Lattice(edges)
    beta, timesteps, num_experiments = 1.0, 20, 40
    betas = [(0, 0.1), (timesteps, 2.0)]

    def hamiltonian(states):
        s = states.astype(np.int64) * 2 - 1
        return 1.0 * s[..., 0] * s[..., 1] - 1.0 * s[..., 1] * s[..., 2]

    e, s = lat.run_monte_carlo(beta, timesteps, num_experiments)
    assert e.shape == (40,) and s.shape == (40, 3) and s.dtype == np.bool_ and (e == hamiltonian(s)).all()
    e, s = lat.run_monte_carlo_sampling(beta, timesteps, num_experiments)
    assert e.shape == (40, 20) and s.shape == (40, 20, 3) and (e == hamiltonian(s)).all()
    e, s = lat.run_monte_carlo_annealing(betas, timesteps, num_experiments)
    assert e.shape == (40,) and s.shape == (40, 3) and (e == hamiltonian(s)).all()
    e, s = lat.run_monte_carlo_annealing_and_get_energies(betas, timesteps, num_experiments)
    assert e.shape == (40, 20) and s.shape == (40, 3) and (e[:, -1] == hamiltonian(s)).all()
    # cold end of the anneal: the chain's two ground states have E = -2
    lat.set_seed_gen(1)
    e, _ = lat.run_monte_carlo(6.0, 200, 64)
    assert (e == -2.0).mean() > 0.9
    lat.set_transverse_field(1.0)
    with pytest.raises(ValueError, match="Cannot run classic monte carlo with transverse field"):
        lat.run_monte_carlo(beta, timesteps, num_experiments)
    with pytest.raises(NotImplementedError):
        lat.run_quantum_monte_carlo(beta, timesteps, num_experiments)

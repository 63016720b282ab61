"""GPU parity tests (run with -m gpu on a B200): everything goes through the C ABI of
libising_b200.so and is compared bit for bit with the CPU oracle on the same inputs."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "exact_2d_ising.json")))


@pytest.fixture(scope="module")
def pkg(native):
    import pyisingmontecarlo_b200 as pkg

    native.Context.get(0)  # fails loudly without a B200
    return pkg


# ---------------------------------------------------------------------------------------------
# replay mode: the reference's own (site, uniform) sequence must be reproduced bit-exactly
# ---------------------------------------------------------------------------------------------
def test_replay_config1_bit_exact(pkg, oracle):
    """BASELINE config 1 (32x32 periodic ferromagnet, beta 0.44, 64 experiments), 100 of its
    1000 timesteps (a trace is 12 B per attempt)."""
    L, E, steps = 32, 64, 100
    edges = oracle.square_edges(L)
    g = oracle.Graph(edges)
    lat = pkg.Lattice(edges, seed_gen=0)
    seeds = np.array(lat.make_seeds(E), dtype=np.uint64)
    assert (seeds == oracle.make_seeds(0, E)).all()
    sites, u, init, en_ref, st_ref = g.trace(0.44, seeds, steps * L * L)
    en, st = lat.replay(0.44, sites, u, init)
    assert st.dtype == np.bool_ and en.dtype == np.float64
    assert (st == st_ref).all()
    assert (en == en_ref).all()
    # and the trace is what run_monte_carlo itself does
    en_run, st_run = g.run_monte_carlo(0.44, steps, seeds)
    assert (st_run == st).all() and (en_run == en).all()


def test_replay_real_couplings_and_biases(pkg, oracle):
    rng = np.random.default_rng(5)
    n = 40
    edges = []
    for i in range(n):
        for j in rng.choice(n, 3, replace=False):
            if i != j:
                edges.append(((int(i), int(j)), float(rng.normal())))
    biases = rng.normal(size=n) * 0.3
    g = oracle.Graph(edges, nvars=n, biases=biases)
    lat = pkg.Lattice(edges, seed_gen=1)
    for v, b in enumerate(biases):
        lat.set_individual_bias(v, float(b))
    seeds = oracle.make_seeds(1, 33)
    sites, u, init, en_ref, st_ref = g.trace(0.9, seeds, 5000)
    en, st = lat.replay(0.9, sites, u, init)
    assert (st == st_ref).all()
    assert (en == en_ref).all()


def test_replay_rejects_bad_trace(pkg, oracle):
    lat = pkg.Lattice(oracle.square_edges(4))
    sites = np.full((1, 3), 99, dtype=np.uint32)
    with pytest.raises(ValueError):
        lat.replay(0.4, sites, np.zeros((1, 3)), np.zeros((1, 16), dtype=bool))


# ---------------------------------------------------------------------------------------------
# production mode: CUDA sweep == scalar restatement (oracle/msc_mirror.c), bit for bit
# ---------------------------------------------------------------------------------------------
def _mirror_check(native, oracle, graph, E, seed, betas, planes, rounds, offset=0, init=None):
    a, b, j = graph.edges()
    colors = graph.colors()
    sim = native.Sim(graph, E, seed, replica_offset=offset, planes=planes, rounds=rounds)
    if init is not None:
        sim.set_state(init)
    en_gpu = sim.sweeps(betas, per_sweep_energies=True)
    st_gpu = sim.states()
    fin_gpu = sim.energies()
    en_ref, st_ref = oracle.msc_mirror(a, b, j, graph.nvars, colors, E, seed, betas,
                                       replica_offset=offset, planes=planes, rounds=rounds,
                                       init_state=init, per_sweep=True)
    assert (st_gpu == st_ref).all()
    assert (en_gpu == en_ref).all()
    assert (fin_gpu == en_ref[:, -1]).all()
    sim.close()


@pytest.mark.parametrize("planes,rounds", [(6, 7), (6, 10), (5, 10), (5, 7), (7, 10), (7, 7)])
def test_msc_3d_pmj_matches_mirror(native, oracle, pkg, planes, rounds):
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (4, 6, 4), j0=1.0, pmj=True, j_seed=77)
    assert g.kind == native.KIND_STENCIL3D and g.ncolors == 2
    betas = np.linspace(0.2, 1.3, 6)
    _mirror_check(native, oracle, g, 70, 0xDEADBEEF12345, betas, planes, rounds)


def test_msc_2d_variants_match_mirror(native, oracle, pkg):
    ctx = native.Context.get(0)
    betas = [0.44, 0.44, 0.3, 0.9]
    for j0 in (-1.0, 1.0, -0.37):
        g = native.Graph.torus(ctx, (8, 6), j0=j0)
        assert g.kind == native.KIND_STENCIL2D
        _mirror_check(native, oracle, g, 33, 42, betas, 6, 7)
    g = native.Graph.torus(ctx, (6, 10), j0=2.0, pmj=True, j_seed=3)
    _mirror_check(native, oracle, g, 64, 43, betas, 6, 7)
    init = np.arange(60) % 3 == 0
    _mirror_check(native, oracle, g, 5, 44, betas, 6, 7, init=init)


def test_many_replica_words_match_mirror(native, oracle, pkg):
    """More replica words than one block covers (chunk loop over words, all vector widths):
    E = 4100 -> 129 words (scalar path), 4128 -> 129.. words, 8192 -> 256 words (128-bit path)."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (4, 4, 4), j0=1.0, pmj=True, j_seed=8)
    for E in (4100, 4160, 8192):
        _mirror_check(native, oracle, g, E, 17, [0.6, 1.0], 6, 7)
    g2 = native.Graph.torus(ctx, (4, 6), j0=-1.0)
    _mirror_check(native, oracle, g2, 2080, 18, [0.44, 0.5], 6, 7)


@pytest.mark.parametrize("dims,pmj,E,rounds,one_launch", [
    ((32, 32), False, 64, 7, True),      # BASELINE config 1: 1024 site-words per colour, one per thread
    ((8, 6), False, 33, 10, True),       # 2 words, ragged last word, Philox-10
    ((6, 10), True, 96, 7, True),        # 3 words (scalar path), +-J
    ((4, 6, 4), True, 70, 7, True),      # 3D +-J
    ((8, 8, 8), True, 256, 7, True),     # 2048 words: exactly one word per thread of the full cluster
    ((8, 8, 8), False, 512, 7, True),    # 4096 words: 2 words per thread
    ((16, 8, 8), True, 512, 7, True),    # 8192 words: 4 words per thread, more rows than one pass
    ((16, 16, 8), True, 512, 7, False),  # 16384 words: above the cluster limit (one launch per colour phase)
])
def test_cluster_kernel_matches_mirror(native, oracle, pkg, dims, pmj, E, rounds, one_launch):
    """Sweeps without per-sweep energies on small lattices run inside one thread-block cluster
    (hardware barrier between the colour phases): same bits as the scalar mirror, at every
    thread / word mapping the launcher can choose."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, dims, j0=1.0 if pmj else -1.0, pmj=pmj, j_seed=5)
    betas = np.linspace(0.2, 1.1, 7)
    sim = native.Sim(g, E, 99, rounds=rounds)
    sim.sweeps(betas[:4])
    sim.sweeps(betas[4:])          # second chunk continues the sweep counter
    a, b, j = g.edges()
    en_ref, st_ref = oracle.msc_mirror(a, b, j, g.nvars, g.colors(), E, 99, betas, planes=6, rounds=rounds)
    assert (sim.states() == st_ref).all()
    assert (sim.energies() == en_ref).all()
    if one_launch:
        assert sim.stats()["kernel_launches"] <= 8     # 2 chunk launches + init / read-back kernels


def test_edge_list_torus_is_recognised_and_equal(native, oracle, pkg):
    """The same lattice through Lattice(edges) (config-1 labelling) and through the additive
    torus constructor takes the stencil path and gives identical bits."""
    ctx = native.Context.get(0)
    edges = oracle.square_edges(8)
    lat = pkg.Lattice(edges, seed_gen=9)
    g = lat.graph()
    assert g.kind == native.KIND_STENCIL2D and g.dims[:2] == (8, 8)
    _mirror_check(native, oracle, g, 40, 9, [0.5] * 3, 6, 7)
    # 3D +-J from an explicit edge list
    rng = np.random.default_rng(0)
    sign = rng.integers(0, 2, size=(4 * 4 * 6, 3)) * 2.0 - 1.0
    L = (4, 4, 6)
    idx = lambda x, y, z: x + L[0] * (y + L[1] * z)
    e3 = []
    for z in range(L[2]):
        for y in range(L[1]):
            for x in range(L[0]):
                n = idx(x, y, z)
                e3.append(((n, idx((x + 1) % L[0], y, z)), sign[n, 0]))
                e3.append(((n, idx(x, (y + 1) % L[1], z)), sign[n, 1]))
                e3.append(((n, idx(x, y, (z + 1) % L[2])), sign[n, 2]))
    rng.shuffle(e3)
    e3 = [((int(a), int(b)) if k % 2 else (int(b), int(a)), float(j)) for k, ((a, b), j) in enumerate(e3)]
    g3 = pkg.Lattice(e3).graph()
    assert g3.kind == native.KIND_STENCIL3D and g3.dims == L
    _mirror_check(native, oracle, g3, 64, 5, [0.8, 0.2, 1.1], 6, 7)


def test_sharded_run_reproduces_unsharded(native, pkg):
    """replica_offset only enters the Philox counters: experiments 64..95 of a 96-experiment
    run are the same bits as a 32-experiment shard at offset 64 (multi-GPU partitioning)."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (6, 4, 4), j0=1.0, pmj=True, j_seed=1)
    betas = [0.7] * 5
    full = native.Sim(g, 96, 1234)
    full.sweeps(betas)
    lo = native.Sim(g, 64, 1234, replica_offset=0)
    hi = native.Sim(g, 32, 1234, replica_offset=64)
    lo.sweeps(betas)
    hi.sweeps(betas)
    st = full.states()
    assert (st[:64] == lo.states()).all() and (st[64:] == hi.states()).all()
    assert (full.energies() == np.concatenate([lo.energies(), hi.energies()])).all()


def test_observables_match_numpy(native, pkg):
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (8, 8, 6), j0=1.0, pmj=True, j_seed=11)
    sim = native.Sim(g, 100, 7)
    sim.sweeps([0.5] * 3)
    st = sim.states().astype(np.int64) * 2 - 1
    a, b, j = g.edges()
    e_np = (st[:, a.astype(np.int64)] * st[:, b.astype(np.int64)] * j).sum(1)
    assert (sim.energies() == e_np).all()
    assert (sim.magnetization() == st.sum(1)).all()
    packed = sim.packed()
    bits = ((packed[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.int64)
    assert ((bits.reshape(g.nvars, -1)[:, :100].T * 2 - 1) == st).all()


def test_set_states_roundtrip(native, pkg):
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (6, 8), j0=-1.0)
    rng = np.random.default_rng(2)
    states = rng.integers(0, 2, size=(45, 48)).astype(bool)
    sim = native.Sim(g, 45, 0)
    sim.set_states(states)
    assert (sim.states() == states).all()


# ---------------------------------------------------------------------------------------------
# the Lattice API surface (lattice.rs:171-470): shapes, dtypes, errors, determinism
# ---------------------------------------------------------------------------------------------
def test_lattice_api_shapes_and_errors(pkg, oracle):
    edges = oracle.square_edges(8)
    with pytest.raises(ValueError, match="Must supply some edges"):
        pkg.Lattice([])
    lat = pkg.Lattice(edges, seed_gen=3)
    en, st = lat.run_monte_carlo(0.4, 10, 7)
    assert en.shape == (7,) and en.dtype == np.float64
    assert st.shape == (7, 64) and st.dtype == np.bool_
    en2, st2 = lat.run_monte_carlo(0.4, 10, 7, True, False)
    assert (st == st2).all() and (en == en2).all()          # seed_gen set -> reruns repeat
    g = oracle.Graph(edges)
    assert all(en[k] == g.energy(st[k]) for k in range(7))
    lat.set_seed_gen(None)
    _, st3 = lat.run_monte_carlo(0.4, 10, 7)
    assert (st3 != st).any()                                  # entropy seeds differ
    lat.set_seed_gen(3)

    en, st = lat.run_monte_carlo_sampling(0.4, 10, 5, None, 2, 3)
    assert en.shape == (5, 3) and st.shape == (5, 3, 64) and st.dtype == np.bool_
    assert all(en[1, k] == g.energy(st[1, k]) for k in range(3))
    en, st = lat.run_monte_carlo_annealing([(0, 0.1), (10, 0.9)], 10, 4)
    assert en.shape == (4,) and st.shape == (4, 64)
    en_t, st_t = lat.run_monte_carlo_annealing_and_get_energies([(0, 0.1), (10, 0.9)], 10, 4)
    assert en_t.shape == (4, 10) and (st_t == st).all() and (en_t[:, -1] == en).all()
    # quirk Q1: identical to a constant-beta run at the last stop
    beta_last = (0.9 - 0.1) * 1.0 + 0.1
    en_c, st_c = lat.run_monte_carlo(beta_last, 10, 4)
    assert (st_c == st).all()

    init = [True] * 64
    lat.set_initial_state(init)
    _, st0 = lat.run_monte_carlo(0.4, 0, 3)
    assert st0.all()
    with pytest.raises(ValueError):
        lat.set_initial_state([True] * 5)
    lat.set_initial_state([])
    with pytest.raises(ValueError):
        lat.set_individual_bias(64, 1.0)
    with pytest.raises(ValueError, match="Transverse field must be positive"):
        lat.set_transverse_field(-1.0)
    lat.set_transverse_field(0.5)
    with pytest.raises(ValueError, match="Cannot run classic monte carlo with transverse field"):
        lat.run_monte_carlo(0.4, 1, 1)
    lat.set_transverse_field(0.0)
    with pytest.raises(NotImplementedError):
        lat.run_monte_carlo(0.4, 1, 1, None, True)            # deviation D2
    with pytest.raises(NotImplementedError):
        lat.run_quantum_monte_carlo(0.4, 1, 1)
    e0, s0 = lat.run_monte_carlo(0.4, 3, 0)
    assert e0.shape == (0,) and s0.shape == (0, 64)


def test_sampling_equals_plain_runs(pkg, oracle):
    lat = pkg.Lattice(oracle.square_edges(8), seed_gen=21)
    en, st = lat.run_monte_carlo_sampling(0.45, 12, 40, None, 3, 4)
    for k in range(3):
        e_k, s_k = lat.run_monte_carlo(0.45, 3 + 4 * (k + 1), 40)
        assert (st[:, k] == s_k).all() and (en[:, k] == e_k).all()


def test_sampling_double_buffered_slabs(pkg, oracle, monkeypatch):
    """Many small slabs (two device buffers draining on the copy stream) give the same arrays as
    one slab; the strided device-to-host copies land every sample at [e, k]."""
    lat = pkg.Lattice(oracle.square_edges(8), seed_gen=5)
    en1, st1 = lat.run_monte_carlo_sampling(0.4, 23, 70, None, 2, 1)
    monkeypatch.setenv("ISING_SAMPLING_SLAB_BYTES", str(70 * 64 * 3))  # 3 samples per slab
    en2, st2 = lat.run_monte_carlo_sampling(0.4, 23, 70, None, 2, 1)
    assert en1.shape == (70, 23) and st1.shape == (70, 23, 64)
    assert (en1 == en2).all() and (st1 == st2).all()


def test_packed_sampling_equals_bool_sampling(native, pkg, monkeypatch):
    """Opt-in packed samples uint32[n_s, nvars, ceil(E/32)] hold the same bits as bool[E, n_s, nvars]
    (one slab and two-samples-per-slab double buffering)."""
    ctx = native.Context.get(0)
    g = native.Graph.torus(ctx, (6, 8), j0=-1.0)
    for slab in (None, str(48 * 3 * 4 * 2)):
        if slab:
            monkeypatch.setenv("ISING_SAMPLING_SLAB_BYTES", slab)
        a, b = native.Sim(g, 70, 5), native.Sim(g, 70, 5)
        en_b, st = a.run_sampling(0.4, 2, 3, 5)
        en_p, words = b.run_sampling(0.4, 2, 3, 5, packed=True)
        assert words.shape == (5, 48, 3) and words.dtype == np.uint32
        bits = ((words[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
        bits = bits.reshape(5, 48, 96)[:, :, :70]          # [n_s, nvars, E]
        assert (bits.transpose(2, 0, 1) == st).all() and (en_p == en_b).all()
        assert (b.packed() == words[-1]).all()


@pytest.mark.parametrize("kind", ["2d", "3d_pmj", "general"])
def test_device_observables_match_sampled_states(pkg, oracle, kind):
    """run_monte_carlo_observables follows the same trajectories as run_monte_carlo_sampling;
    E, M and the pair overlaps Q reduced on the device equal the ones computed from the states."""
    if kind == "2d":
        lat = pkg.Lattice(oracle.square_edges(10), seed_gen=3)
    elif kind == "3d_pmj":
        lat = pkg.Lattice.torus((6, 4, 8), pmj=True, j_seed=2, seed_gen=3)
    else:
        rng = np.random.default_rng(0)
        edges = [((i, (i + 1) % 37), -1.0) for i in range(37)]
        edges += [((int(a), int(b)), 1.0) for a, b in rng.integers(0, 37, size=(20, 2)) if a != b
                  and abs(a - b) not in (1, 36)]
        edges = list({(min(a, b), max(a, b)): ((a, b), j) for (a, b), j in edges}.values())
        lat = pkg.Lattice(edges, seed_gen=3)
    E = 77
    en, st = lat.run_monte_carlo_sampling(0.6, 10, E, None, 4, 2)
    e2, m, q = lat.run_monte_carlo_observables(0.6, 10, E, 4, 2)
    s = st.astype(np.int64) * 2 - 1
    assert e2.shape == (E, 5) and m.shape == (E, 5) and q.shape == (E // 2, 5)
    assert (e2 == en).all()
    assert (m == s.sum(-1)).all()
    assert (q == (s[0:2 * (E // 2):2] * s[1:2 * (E // 2):2]).sum(-1)).all()


# ---------------------------------------------------------------------------------------------
# statistics: production mode must sample the same Boltzmann law as the reference algorithm
# ---------------------------------------------------------------------------------------------
def test_config1_kaufman_and_oracle_agree(pkg, oracle):
    """32x32 ferromagnet at beta = 0.44: GPU mean energy within 3 sigma of Kaufman's exact
    value and of the CPU restatement of the reference algorithm."""
    L, E = 32, 1024
    lat = pkg.Lattice(oracle.square_edges(L), seed_gen=12345)
    en, st = lat.run_monte_carlo(0.44, 3000, E)
    exact = GOLD["kaufman"]["L32_b0.44"]
    e = en / (L * L)
    sigma = np.sqrt(exact["c_per_site"] / (0.44**2 * L * L))
    err = sigma / np.sqrt(E)
    assert abs(e.mean() - exact["e_per_site"]) < 3 * err, (e.mean(), exact["e_per_site"], err)
    assert abs(e.std(ddof=1) / sigma - 1) < 0.15
    # Binder cumulant and |m| against the oracle (reference algorithm, 256 experiments)
    m = (st.sum(1) * 2.0 - L * L) / (L * L)
    g = oracle.Graph(oracle.square_edges(L))
    en_o, st_o = g.run_monte_carlo(0.44, 3000, oracle.make_seeds(1, 256))
    m_o = (st_o.sum(1) * 2.0 - L * L) / (L * L)

    def binder(x):
        return 1 - (x**4).mean() / (3 * (x**2).mean() ** 2)

    def jack(x, f):
        n = len(x)
        blocks = np.array_split(np.arange(n), 32)
        vals = np.array([f(np.delete(x, b)) for b in blocks])
        return f(x), np.sqrt((len(blocks) - 1) * vals.var())

    for f in (lambda x: np.abs(x).mean(), binder):
        a, ea = jack(m, f)
        b, eb = jack(m_o, f)
        assert abs(a - b) < 3 * np.hypot(ea, eb), (a, b, ea, eb)
    eo = en_o / (L * L)
    assert abs(e.mean() - eo.mean()) < 3 * np.hypot(err, eo.std(ddof=1) / np.sqrt(256))


def test_3d_spin_glass_matches_reference_algorithm(pkg, oracle):
    """6^3 +-J sample at beta = 0.6: checkerboard/Philox (GPU) vs random-site/xoshiro (oracle)."""
    L = 6
    rng = np.random.default_rng(3)
    signs = rng.integers(0, 2, size=(L**3, 3)) * 2.0 - 1.0
    edges = oracle.cubic_edges(L, lambda n, d: float(signs[n, d]))
    lat = pkg.Lattice(edges, seed_gen=99)
    assert lat.graph().kind == 3
    en, _ = lat.run_monte_carlo(0.6, 2000, 2048)
    g = oracle.Graph(edges)
    en_o, _ = g.run_monte_carlo(0.6, 2000, oracle.make_seeds(4, 512))
    n = L**3
    a, b = en / n, en_o / n
    err = np.hypot(a.std(ddof=1) / np.sqrt(len(a)), b.std(ddof=1) / np.sqrt(len(b)))
    assert abs(a.mean() - b.mean()) < 3 * err, (a.mean(), b.mean(), err)
    assert abs(a.std(ddof=1) / b.std(ddof=1) - 1) < 0.12


# ---------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------------------------
def test_config3_full_size_properties(native, pkg):
    """3D +-J L=64 x 1024 replicas: (i) energies reported per sweep equal the energy of the
    returned states recomputed in numpy; (ii) a 128-replica shard at offset 896 reproduces
    replicas 896..1023 of the full run (the 8-GPU partitioning); (iii) annealing lowers E."""
    ctx = native.Context.get(0)
    L = 64
    g = native.Graph.torus(ctx, (L, L, L), j0=1.0, pmj=True, j_seed=2024)
    betas = np.linspace(0.1, 1.2, 12)
    sim = native.Sim(g, 1024, 31337)
    en = sim.sweeps(betas, per_sweep_energies=True)
    assert en.shape == (1024, 12)
    assert en[:, -1].mean() < en[:, 0].mean() < 0
    st = sim.states()
    assert st.shape == (1024, L**3)
    a, b, j = g.edges()
    sub = st[::97].astype(np.int8) * 2 - 1       # 11 replicas are enough for an exact check
    e_np = (sub[:, a.astype(np.int64)].astype(np.int32) * sub[:, b.astype(np.int64)] * j.astype(np.int32)).sum(1)
    assert (en[::97, -1] == e_np).all()
    shard = native.Sim(g, 128, 31337, replica_offset=896)
    shard.sweeps(betas)
    assert (shard.states() == st[896:]).all()
    stats = sim.stats()
    assert stats["sweeps"] == 12 and stats["flip_attempts"] == 12 * 1024 * L**3
    assert stats["kernel_launches"] >= 24


def test_tma_staged_kernel_matches_direct_loads(native):
    """The opt-in TMA-staged colour phase (ISING_TMA=1: neighbour rows by cp.async.bulk into shared
    memory, mbarrier ring) computes the same bits as the default direct-load row walk.  The switch
    is read once per process, so the staged run happens in a child process."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from pyisingmontecarlo_b200 import _native as nat\n"
        "ctx = nat.Context.get(0)\n"
        "out = []\n"
        "for dims, E in (((16, 16, 16), 1024), ((16, 8, 8), 128), ((32, 16), 256)):\n"
        "    g = nat.Graph.torus(ctx, dims, j0=1.0, pmj=True, j_seed=4)\n"
        "    sim = nat.Sim(g, E, 5)\n"
        "    en = sim.sweeps(np.linspace(0.2, 1.2, 9), per_sweep_energies=True)\n"
        "    sim.sweeps(np.linspace(0.5, 0.6, 4))\n"
        "    out.append(en); out.append(sim.packed()); out.append(sim.energies())\n"
        "np.savez(sys.argv[1], *out)\n" % root)
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        res = {}
        for tag, env in (("direct", {}), ("tma", {"ISING_TMA": "1"})):
            path = os.path.join(tmp, tag + ".npz")
            e = dict(os.environ)
            e.pop("ISING_TMA", None)
            e.update(env)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=e, timeout=600)
            with np.load(path) as d:
                res[tag] = [d[k] for k in d.files]
        assert len(res["direct"]) == len(res["tma"]) == 9
        for x, y in zip(res["direct"], res["tma"]):
            assert x.shape == y.shape and (x == y).all()


def test_default_moves_warn_once(pkg, oracle):
    """only_basic_moves=None/False (the reference's default) runs single-spin sweeps and says so
    once per process (deviation D1); only_basic_moves=True is silent."""
    import warnings

    from pyisingmontecarlo_b200 import lattice as lattice_mod

    lat = pkg.Lattice(oracle.square_edges(4), seed_gen=1)
    lattice_mod._warned_basic_moves = False
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        lat.run_monte_carlo(0.4, 2, 4, True)            # explicit: no warning
    with pytest.warns(UserWarning, match="single-spin Metropolis"):
        lat.run_monte_carlo(0.4, 2, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        lat.run_monte_carlo(0.4, 2, 4)                  # once per process


def test_philox7_statistical_regression(pkg, native, oracle):
    """Philox4x32-7 is the default since round 2 (Random123's Crush-resistant minimum).  With 4096
    experiments - twice to four times the statistics of the parity tests above - the default
    streams must still hit (i) Kaufman's exact <E> of the 32 x 32 torus at beta = 0.44 within 3
    sigma and its exact specific heat within 8 %, (ii) the reference algorithm's <E> and spread on
    one 6^3 +-J sample at beta = 0.6 within 3 sigma, (iii) and agree with the 10-round streams."""
    L, E = 32, 4096
    lat = pkg.Lattice(oracle.square_edges(L), seed_gen=20261018)
    en, st = lat.run_monte_carlo(0.44, 3000, E, True)
    exact = GOLD["kaufman"]["L32_b0.44"]
    e = en / (L * L)
    sigma = np.sqrt(exact["c_per_site"] / (0.44**2 * L * L))
    assert abs(e.mean() - exact["e_per_site"]) < 3 * sigma / np.sqrt(E), (e.mean(), exact["e_per_site"])
    assert abs(e.var(ddof=1) / sigma**2 - 1) < 0.08, (e.var(ddof=1), sigma**2)
    m = (st.sum(1) * 2.0 - L * L) / (L * L)
    assert abs(m.mean()) < 3 * np.sqrt((m**2).mean() / E)          # Z2 symmetry of the sample
    # the same lattice with 10 rounds: the two estimates of <E> agree within their errors
    ctx = native.Context.get(0)
    g = lat.graph()
    s10 = native.Sim(g, E, 7, rounds=10)
    s10.sweeps(np.full(3000, 0.44))
    e10 = s10.energies() / (L * L)
    assert abs(e.mean() - e10.mean()) < 3 * np.hypot(e.std(ddof=1), e10.std(ddof=1)) / np.sqrt(E)
    s10.close()
    # 6^3 +-J glass against the reference algorithm (random-site, xoshiro256++), 4096 vs 1024
    Lg = 6
    rng = np.random.default_rng(3)
    signs = rng.integers(0, 2, size=(Lg**3, 3)) * 2.0 - 1.0
    edges = oracle.cubic_edges(Lg, lambda n, d: float(signs[n, d]))
    glass = pkg.Lattice(edges, seed_gen=5)
    en_g, _ = glass.run_monte_carlo(0.6, 3000, E, True)
    en_o, _ = oracle.Graph(edges).run_monte_carlo(0.6, 3000, oracle.make_seeds(11, 1024))
    a, b = en_g / Lg**3, en_o / Lg**3
    err = np.hypot(a.std(ddof=1) / np.sqrt(len(a)), b.std(ddof=1) / np.sqrt(len(b)))
    assert abs(a.mean() - b.mean()) < 3 * err, (a.mean(), b.mean(), err)
    assert abs(a.std(ddof=1) / b.std(ddof=1) - 1) < 0.08

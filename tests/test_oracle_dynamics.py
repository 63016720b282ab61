"""Pins the oracle's GraphState / driver restatement with independent exact results:
brute-force enumeration (N <= 16) and Kaufman's finite-lattice 2D Ising energy."""
import itertools
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "exact_2d_ising.json")))


def exact_moments(edges, nvars, beta, biases=None):
    """<E>, <E^2>, <|m|> by summing all 2^N states; E = sum J s s - sum b s, s = +-1."""
    a = np.array([e[0][0] for e in edges]); b = np.array([e[0][1] for e in edges])
    j = np.array([e[1] for e in edges], dtype=float)
    bias = np.zeros(nvars) if biases is None else np.asarray(biases, float)
    states = np.array(list(itertools.product([-1, 1], repeat=nvars)), dtype=float)
    E = (states[:, a] * states[:, b] * j).sum(1) - states @ bias
    w = np.exp(-beta * (E - E.min()))
    w /= w.sum()
    m = np.abs(states.sum(1)) / nvars
    return float((w * E).sum()), float((w * E * E).sum()), float((w * m).sum())


def test_energy_convention(oracle):
    # README.md:45-46: J*Sza*Szb, positive is antiferromagnetic
    g = oracle.Graph([((0, 1), 1.0), ((1, 2), -1.0)])
    assert g.energy([1, 1, 1]) == 0.0
    assert g.energy([1, 0, 0]) == -2.0
    assert g.energy([1, 1, 0]) == 2.0
    gb = oracle.Graph([((0, 1), 1.0)], biases=[0.5, -0.25])
    assert gb.energy([1, 1]) == 1.0 - 0.5 + 0.25
    assert gb.energy([0, 1]) == -1.0 + 0.5 + 0.25


def test_kaufman_matches_brute_force_L4(oracle):
    edges = oracle.square_edges(4)
    e, _, _ = exact_moments(edges, 16, 0.44)
    assert abs(e / 16 - GOLD["kaufman"]["L4_b0.44"]["e_per_site"]) < 1e-10


@pytest.mark.parametrize("case", ["frustrated_triangle", "k4_mixed_bias", "ring6_afm"])
def test_oracle_dynamics_vs_enumeration(oracle, case):
    if case == "frustrated_triangle":
        edges, n, bias, beta = [((0, 1), 1.0), ((1, 2), 1.0), ((0, 2), 1.0)], 3, None, 0.7
    elif case == "k4_mixed_bias":
        edges = [((0, 1), -1.0), ((0, 2), 0.5), ((0, 3), 1.5), ((1, 2), -0.7), ((1, 3), 0.3), ((2, 3), 1.0)]
        n, bias, beta = 4, [0.3, -0.2, 0.0, 0.6], 0.5
    else:
        edges, n, bias, beta = [((i, (i + 1) % 6), 1.0) for i in range(6)], 6, None, 0.6
    g = oracle.Graph(edges, biases=bias)
    seeds = oracle.make_seeds(7, 256)
    # every experiment contributes samples after a short burn-in; errors from the spread of means
    en, _ = g.run_sampling(beta, 400, seeds, thermalization=50, sampling_freq=2, attempts_per_step=n)
    per_exp = en.mean(axis=1)
    mean, err = per_exp.mean(), per_exp.std(ddof=1) / np.sqrt(len(per_exp))
    exact, _, _ = exact_moments(edges, n, beta, bias)
    assert abs(mean - exact) < 4 * err + 1e-12, (mean, exact, err)


def test_oracle_c1_kaufman(oracle):
    # BASELINE config 1: 32x32 periodic ferromagnet, beta = 0.44, 1000 sweeps, 64 experiments
    L = 32
    g = oracle.Graph(oracle.square_edges(L))
    seeds = oracle.make_seeds(0, 64)
    en, st = g.run_monte_carlo(0.44, 1000, seeds)
    assert st.shape == (64, L * L) and st.dtype == bool
    e = en / (L * L)
    exact = GOLD["kaufman"]["L32_b0.44"]
    sigma = np.sqrt(exact["c_per_site"] / (0.44**2 * L * L))  # sqrt(Var(E/N)) = sqrt(c/(beta^2 N))
    err = sigma / np.sqrt(64)
    assert abs(e.mean() - exact["e_per_site"]) < 3.5 * err, (e.mean(), exact["e_per_site"], err)
    assert 0.6 * sigma < e.std(ddof=1) < 1.5 * sigma
    # energies are those of the returned states
    assert en[3] == g.energy(st[3])


def test_seed_determinism_and_initial_state(oracle):
    g = oracle.Graph(oracle.square_edges(8))
    seeds = oracle.make_seeds(3, 4)
    e1, s1 = g.run_monte_carlo(0.3, 5, seeds)
    e2, s2 = g.run_monte_carlo(0.3, 5, seeds)
    assert (s1 == s2).all() and (e1 == e2).all()
    # zero timesteps returns the initial state: random from the rng, or the one supplied
    init = np.arange(64) % 3 == 0
    _, s0 = g.run_monte_carlo(0.3, 0, seeds, initial_state=init)
    assert (s0 == init).all()
    _, sr = g.run_monte_carlo(0.3, 0, seeds)
    rng = oracle.Rng(seed=int(seeds[1]))
    assert [rng.gen_bool() for _ in range(64)] == list(sr[1])


def test_schedule_quirk_q1(oracle):
    # lattice.rs:331/359-365: every timestep runs at the beta of the last user stop
    b = oracle.schedule_betas([(0, 0.1), (1000, 1.2)], 1000, q1_compat=True)
    assert (b == (1.2 - 0.1) * 1.0 + 0.1).all()
    b = oracle.schedule_betas([(500, 2.0), (10, 0.5)], 1000, q1_compat=True)
    assert (b == (2.0 - 0.5) * 1.0 + 0.5).all()
    b = oracle.schedule_betas([], 10, q1_compat=True)
    assert (b == 1.0).all()
    b = oracle.schedule_betas([(0, 0.7)], 10, q1_compat=True)   # i = 0 -> va
    assert (b == 0.7).all()
    b = oracle.schedule_betas([(0, 0.2), (0, 0.9)], 10, q1_compat=True)  # 0/0
    assert np.isnan(b).all()
    # documented behaviour: linear interpolation in t
    b = oracle.schedule_betas([(0, 0.0), (10, 1.0)], 10, q1_compat=False)
    assert np.allclose(b, np.arange(10) / 10.0)
    b = oracle.schedule_betas([(4, 1.0), (8, 3.0)], 12, q1_compat=False)
    assert np.allclose(b, [1, 1, 1, 1, 1, 1.5, 2, 2.5, 3, 3, 3, 3])


def test_annealing_drivers(oracle):
    g = oracle.Graph(oracle.square_edges(8))
    seeds = oracle.make_seeds(11, 6)
    stops = [(0, 0.1), (20, 0.6)]
    e_fin, s_fin = g.run_annealing(stops, 20, seeds)
    e_all, s_all = g.run_annealing(stops, 20, seeds, per_step_energies=True)
    assert e_all.shape == (6, 20) and (s_fin == s_all).all() and (e_all[:, -1] == e_fin).all()
    # Q1: identical to a constant-beta run at the last stop's beta
    beta = oracle.schedule_betas(stops, 20)[0]
    e_c, s_c = g.run_monte_carlo(beta, 20, seeds)
    assert (s_c == s_fin).all() and (e_c == e_fin).all()


def test_sampling_driver(oracle):
    g = oracle.Graph(oracle.square_edges(8))
    seeds = oracle.make_seeds(5, 3)
    en, st = g.run_sampling(0.4, 10, seeds, thermalization=2, sampling_freq=3)
    assert en.shape == (3, 3) and st.shape == (3, 3, 64)
    # sample k equals a plain run of thermalization + 3(k+1) timesteps
    e_ref, s_ref = g.run_monte_carlo(0.4, 2 + 6, seeds)
    assert (st[:, 1] == s_ref).all() and (en[:, 1] == e_ref).all()


def test_trace_and_replay_roundtrip(oracle):
    edges = [((0, 1), -1.0), ((1, 2), 0.37), ((2, 3), -2.2), ((3, 0), 1.0), ((0, 2), 0.9)]
    g = oracle.Graph(edges, biases=[0.1, 0.0, -0.4, 0.2])
    seeds = oracle.make_seeds(1, 5)
    sites, u, init, en, st = g.trace(0.8, seeds, 300)
    e_run, s_run = g.run_monte_carlo(0.8, 75, seeds)   # 75 timesteps x 4 attempts
    assert (s_run == st).all() and (e_run == en).all()
    e_rep, s_rep = g.replay(0.8, sites, u, init)
    assert (s_rep == st).all() and (e_rep == en).all()
    assert ((u == 2.0) | ((u >= 0) & (u < 1))).all() and (u == 2.0).any() and (u < 1).any()


def test_pt_cadence_and_swaps(oracle):
    g = oracle.Graph(oracle.square_edges(6))
    betas = np.linspace(0.2, 0.6, 5)
    states, en, swaps = g.pt_run(betas, 99, timesteps=40, replica_swap_freq=4, sampling_freq=10)
    assert states.shape == (5, 4, 36) and en.shape == (5,) and swaps > 0
    # colder slots have lower time-averaged energy
    assert en[0] > en[-1]
    # with the swap period longer than the run nothing is ever swapped
    _, _, no_swaps = g.pt_run(betas, 99, timesteps=40, replica_swap_freq=1000, sampling_freq=10)
    assert no_swaps == 0


@pytest.mark.parametrize("opts", [
    dict(uniform_always=1), dict(zero_draws=1), dict(init_draws=0), dict(uniform_always=1, zero_draws=1),
])
def test_named_alternatives_of_the_recalled_semantics(oracle, opts):
    """Every choice the restatement had to make about the absent crates (when a uniform is drawn,
    whether GraphState::new draws under a given state) is a named option of the oracle.  An
    alternative consumes different random numbers - a run diverges from the default one - and
    samples the same Boltzmann law."""
    edges = [((0, 1), -1.0), ((0, 2), 0.5), ((0, 3), 1.5), ((1, 2), -0.7), ((1, 3), 0.3), ((2, 3), 1.0),
             ((3, 4), 1.0), ((4, 0), -1.0)]                  # the last two make dE == 0 possible
    n, bias, beta = 5, None, 0.5
    g = oracle.Graph(edges, biases=bias)
    seeds = oracle.make_seeds(7, 256)
    init = [True, False, True, True, False]
    _, st_default = g.run_monte_carlo(beta, 30, seeds, initial_state=init)
    with oracle.options(**opts):
        assert all(oracle.get_option(k) == v for k, v in opts.items())
        _, st_alt = g.run_monte_carlo(beta, 30, seeds, initial_state=init)
        en, _ = g.run_sampling(beta, 400, seeds, thermalization=50, sampling_freq=2, attempts_per_step=n)
    assert oracle.get_option("uniform_always") == 0 and oracle.get_option("init_draws") == 1   # defaults are back
    assert (st_alt != st_default).any()
    per_exp = en.mean(axis=1)
    mean, err = per_exp.mean(), per_exp.std(ddof=1) / np.sqrt(len(per_exp))
    exact, _, _ = exact_moments(edges, n, beta, bias)
    assert abs(mean - exact) < 4 * err + 1e-12, (opts, mean, exact, err)


def test_bias_sign_and_tempering_pair_options(oracle):
    edges = [((0, 1), -1.0), ((1, 2), 0.5), ((2, 0), 1.5)]
    bias = [0.3, -0.2, 0.6]
    st = np.array([1, 0, 1], dtype=np.uint8)
    plus = oracle.Graph(edges, biases=bias).energy(st)
    with oracle.options(bias_sign=-1):
        minus = oracle.Graph(edges, biases=bias).energy(st)
        flipped = oracle.Graph(edges, biases=[-b for b in bias]).energy(st)
    assert abs(flipped - plus) < 1e-12                   # sign -1 of the negated biases = the default
    assert abs(minus - oracle.Graph(edges, biases=[-b for b in bias]).energy(st)) < 1e-12
    assert abs(plus - minus) > 0.1
    with pytest.raises(ValueError):
        with oracle.options(bias_sign=0):
            pass
    # tempering: one parity per step attempts about half as many swaps as both parities
    g = oracle.Graph(oracle.square_edges(6))
    betas = np.linspace(0.2, 0.6, 6)
    _, en0, both = g.pt_run(betas, 99, timesteps=400, replica_swap_freq=2, sampling_freq=100)
    results = {}
    for rule in (1, 2):
        with oracle.options(pt_pairs=rule):
            _, en, swaps = g.pt_run(betas, 99, timesteps=400, replica_swap_freq=2, sampling_freq=100)
        results[rule] = swaps
        assert 0.3 * both < swaps < 0.75 * both, (rule, swaps, both)
        assert en[0] > en[-1]
    assert results[1] != results[2]

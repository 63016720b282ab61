/*
 * oracle/ising_oracle.c -- CPU restatement of the reference's classical Monte-Carlo path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyisingmontecarlo_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED.  The reference (/root/reference, crate py_monte_carlo 2.20.0) is a pyo3
 * binding whose arithmetic lives in the out-of-tree crates `qmc ^2.20`
 * (qmc::classical::graph::GraphState) and `rand ^0.8` (SmallRng); neither source is in the
 * tree, there is no Cargo.lock, no Rust toolchain here, and the reference has no tests or
 * golden vectors.  The functions below therefore restate the *published* algorithms of those
 * crates at the reference's call sites and are pinned by independent goldens only
 * (xoshiro256++ / SplitMix64 vectors, exact enumeration, Kaufman's finite-lattice 2D Ising
 * energy) -- see tests/test_oracle_*.py.
 *
 * What follows what (reference file:line):
 *   orc_make_seeds ............ src/lattice.rs:83-91      (master SmallRng -> one u64 per run)
 *   orc_run_monte_carlo ....... src/lattice.rs:171-221
 *   orc_run_sampling .......... src/lattice.rs:231-299
 *   orc_run_annealing ......... src/lattice.rs:309-385 and 395-470 (incl. the captured-`i`
 *                               schedule quirk at :331/:359-365 and :417/:445-451)
 *   orc_graph_new ............. qmc GraphState::new_with_state_and_rng as called at
 *                               src/lattice.rs:199, src/classicising.rs:70-74
 *   orc_attempt ............... qmc GraphState::do_spin_flip / should_flip, called through
 *                               do_time_step at src/lattice.rs:205,272,278,366,452
 *   orc_energy ................ qmc GraphState::get_energy, src/lattice.rs:208,284,370,454
 *   orc_run_moves ............. do_time_step with non-basic moves (edge / worm), restated as the
 *                               library defines them (the crate's rules are not readable here)
 *   orc_pt_* .................. src/tempering.rs:156-222 (run / swap / sample cadence) with the
 *                               classical swap rule min(1, exp((b_a-b_b)(E_a-E_b)))
 *   rng ....................... rand 0.8 SmallRng = xoshiro256++ (64-bit targets), seed_from_u64
 *                               = SplitMix64, Standard<bool/f64/u64>, UniformInt::sample_single
 *
 * One "timestep" = `attempts_per_step` single-spin Metropolis attempts at uniformly random
 * sites (the reference's do_time_step(beta, None, None, None, Some(true)); the crate's default
 * attempt count is not verifiable here, so it is an explicit parameter, default nvars).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Named choices of the RECALLED crate semantics (A4-A6, A11).  Nothing in
 * /root/reference can settle them; the defaults are what this restatement believes `qmc` and
 * `rand` do, and every alternative is one switch away for a maintainer who can read the crates.
 * None of them changes the stationary distribution (tests/test_oracle_dynamics.py checks each
 * against exact enumeration); they change which random numbers a run consumes, i.e. whether a
 * trace of the reference could be reproduced draw by draw.
 *   ORC_OPT_UNIFORM_ALWAYS  0*: gen::<f64>() is drawn only when dE > 0        1: once per attempt
 *   ORC_OPT_ZERO_DRAWS      0*: dE == 0 flips without a draw (dE <= 0)        1: dE == 0 draws (and accepts: u < 1)
 *   ORC_OPT_INIT_DRAWS      1*: GraphState::new draws nvars bools even when set_state follows
 *                               (lattice.rs:199-203)                          0: no draw when a state is given
 *   ORC_OPT_BIAS_SIGN       +1*: E = sum J s s' - sum b s                     -1: E = sum J s s' + sum b s
 *   ORC_OPT_PT_PAIRS        0*: a tempering step tries the even pairs, then the odd pairs
 *                           1: one parity per step, alternating from even
 *                           2: one parity per step, chosen by gen::<bool>() of the container rng
 * (* = default).  Process-wide, not thread-safe: test infrastructure.
 * ---------------------------------------------------------------------------------------- */
enum { ORC_OPT_UNIFORM_ALWAYS = 0, ORC_OPT_ZERO_DRAWS = 1, ORC_OPT_INIT_DRAWS = 2, ORC_OPT_BIAS_SIGN = 3,
       ORC_OPT_PT_PAIRS = 4, ORC_OPT_COUNT = 5 };
static int g_opt[ORC_OPT_COUNT] = {0, 0, 1, 1, 0};

ORC_EXPORT int orc_set_option(int which, int value) {
    if (which < 0 || which >= ORC_OPT_COUNT) return -1;
    if (which == ORC_OPT_BIAS_SIGN && value != 1 && value != -1) return -2;
    if (which == ORC_OPT_PT_PAIRS && (value < 0 || value > 2)) return -2;
    g_opt[which] = value;
    return 0;
}

ORC_EXPORT int orc_get_option(int which) { return (which < 0 || which >= ORC_OPT_COUNT) ? -1 : g_opt[which]; }

ORC_EXPORT void orc_reset_options(void) {
    g_opt[ORC_OPT_UNIFORM_ALWAYS] = 0; g_opt[ORC_OPT_ZERO_DRAWS] = 0; g_opt[ORC_OPT_INIT_DRAWS] = 1;
    g_opt[ORC_OPT_BIAS_SIGN] = 1; g_opt[ORC_OPT_PT_PAIRS] = 0;
}

/* ------------------------------------------------------------------------------------------
 * rand 0.8 SmallRng (xoshiro256++), as used at src/lattice.rs:85-90,198
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* SeedableRng::seed_from_u64 for Xoshiro256PlusPlus: four SplitMix64 outputs. */
ORC_EXPORT void orc_seed_from_u64(uint64_t seed, uint64_t st[4]) {
    for (int i = 0; i < 4; ++i) {
        seed += 0x9e3779b97f4a7c15ULL;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        st[i] = z ^ (z >> 31);
    }
}

ORC_EXPORT uint64_t orc_next_u64(uint64_t s[4]) {
    uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* Standard.sample::<bool>: sign bit of next_u32(); xoshiro256++ next_u32 = next_u64 >> 32. */
ORC_EXPORT int orc_gen_bool(uint64_t s[4]) { return (int)(orc_next_u64(s) >> 63); }

/* Standard.sample::<f64>: 53 high bits scaled by 2^-53. */
ORC_EXPORT double orc_gen_f64(uint64_t s[4]) {
    return (double)(orc_next_u64(s) >> 11) * (1.0 / 9007199254740992.0);
}

/* rng.gen_range(0..n) for usize on a 64-bit target: UniformInt::sample_single, widening
 * multiply with the conservative zone (n << lzcnt(n)) - 1. */
ORC_EXPORT uint64_t orc_gen_range(uint64_t s[4], uint64_t n) {
    if (n == 0) return orc_next_u64(s);
    uint64_t zone = (n << __builtin_clzll(n)) - 1;
    for (;;) {
        uint64_t v = orc_next_u64(s);
        unsigned __int128 m = (unsigned __int128)v * (unsigned __int128)n;
        uint64_t lo = (uint64_t)m;
        if (lo <= zone) return (uint64_t)(m >> 64);
    }
}

/* src/lattice.rs:83-91 with seed_gen = Some(seed). */
ORC_EXPORT void orc_make_seeds(uint64_t seed_gen, uint64_t n, uint64_t *out) {
    uint64_t st[4];
    orc_seed_from_u64(seed_gen, st);
    for (uint64_t i = 0; i < n; ++i) out[i] = orc_next_u64(st);
}

/* ------------------------------------------------------------------------------------------
 * GraphState: adjacency lists (both directions, stably sorted by neighbour index) + biases
 * ---------------------------------------------------------------------------------------- */
typedef struct orc_graph {
    uint64_t nvars;
    uint64_t nedges;
    uint64_t *row;  /* nvars+1 offsets into nbr/jv                                  */
    uint64_t *nbr;  /* neighbour index, ascending within a row, ties in edge order  */
    double *jv;     /* coupling of that bond                                        */
    double *bias;   /* nvars                                                        */
} orc_graph;

typedef struct {
    uint64_t nbr;
    double j;
    uint64_t ord;
} orc_adj;

static int adj_cmp(const void *pa, const void *pb) {
    const orc_adj *a = (const orc_adj *)pa, *b = (const orc_adj *)pb;
    if (a->nbr != b->nbr) return a->nbr < b->nbr ? -1 : 1;
    return a->ord < b->ord ? -1 : (a->ord > b->ord);
}

ORC_EXPORT orc_graph *orc_graph_new(uint64_t nvars, uint64_t nedges, const uint64_t *ea,
                                    const uint64_t *eb, const double *ej, const double *biases) {
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->nvars = nvars;
    g->nedges = nedges;
    g->row = (uint64_t *)calloc(nvars + 1, sizeof(uint64_t));
    g->nbr = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (nedges ? nedges : 1));
    g->jv = (double *)malloc(sizeof(double) * 2 * (nedges ? nedges : 1));
    g->bias = (double *)calloc(nvars ? nvars : 1, sizeof(double));
    if (biases)
        for (uint64_t i = 0; i < nvars; ++i) g->bias[i] = g_opt[ORC_OPT_BIAS_SIGN] * biases[i];
    for (uint64_t e = 0; e < nedges; ++e) {
        g->row[ea[e] + 1]++;
        g->row[eb[e] + 1]++;
    }
    for (uint64_t i = 0; i < nvars; ++i) g->row[i + 1] += g->row[i];
    uint64_t *fill = (uint64_t *)calloc(nvars ? nvars : 1, sizeof(uint64_t));
    orc_adj *tmp = (orc_adj *)malloc(sizeof(orc_adj) * 2 * (nedges ? nedges : 1));
    for (uint64_t e = 0; e < nedges; ++e) {
        uint64_t a = ea[e], b = eb[e];
        uint64_t pa = g->row[a] + fill[a]++;
        tmp[pa].nbr = b; tmp[pa].j = ej[e]; tmp[pa].ord = 2 * e;
        uint64_t pb = g->row[b] + fill[b]++;
        tmp[pb].nbr = a; tmp[pb].j = ej[e]; tmp[pb].ord = 2 * e + 1;
    }
    for (uint64_t i = 0; i < nvars; ++i) {
        uint64_t lo = g->row[i], hi = g->row[i + 1];
        qsort(tmp + lo, hi - lo, sizeof(orc_adj), adj_cmp); /* ord makes it a stable sort */
        for (uint64_t k = lo; k < hi; ++k) { g->nbr[k] = tmp[k].nbr; g->jv[k] = tmp[k].j; }
    }
    free(tmp);
    free(fill);
    return g;
}

ORC_EXPORT void orc_graph_free(orc_graph *g) {
    if (!g) return;
    free(g->row); free(g->nbr); free(g->jv); free(g->bias); free(g);
}

/* GraphState::get_energy: sum_i ( sum_adj J*coupling/2 - b_i s_i ), s = +1 for true. */
ORC_EXPORT double orc_energy(const orc_graph *g, const uint8_t *state) {
    double acc = 0.0;
    for (uint64_t i = 0; i < g->nvars; ++i) {
        double total = 0.0;
        for (uint64_t k = g->row[i]; k < g->row[i + 1]; ++k) {
            double coupling = (state[i] == state[g->nbr[k]]) ? 1.0 : -1.0;
            total += g->jv[k] * coupling / 2.0;
        }
        double bias_e = state[i] ? -g->bias[i] : g->bias[i];
        acc = acc + total + bias_e;
    }
    return acc;
}

/* dE of flipping `site`: sum_adj(-2 J coupling) + 2 b s, summed in adjacency order. */
static inline double orc_delta_e(const orc_graph *g, const uint8_t *state, uint64_t site) {
    double de = 0.0;
    uint8_t cur = state[site];
    for (uint64_t k = g->row[site]; k < g->row[site + 1]; ++k) {
        double coupling = (cur == state[g->nbr[k]]) ? 1.0 : -1.0;
        de += -2.0 * g->jv[k] * coupling;
    }
    return de + 2.0 * g->bias[site] * (cur ? 1.0 : -1.0);
}

/* One single-spin Metropolis attempt (do_spin_flip + should_flip).  A uniform is drawn only
 * when dE > 0.  Optionally records (site, u) for the replay trace; u = 2.0 when not drawn. */
static inline void orc_attempt(const orc_graph *g, uint8_t *state, uint64_t rng[4], double beta,
                               uint32_t *site_out, double *u_out) {
    uint64_t site = orc_gen_range(rng, g->nvars);
    double de = orc_delta_e(g, state, site);
    double u = 2.0;
    int flip = 1;
    if (de > 0.0 || g_opt[ORC_OPT_UNIFORM_ALWAYS] || (de == 0.0 && g_opt[ORC_OPT_ZERO_DRAWS])) {
        u = orc_gen_f64(rng);
        if (de > 0.0) flip = u < exp(-beta * de);
    }
    if (flip) state[site] = !state[site];
    if (site_out) *site_out = (uint32_t)site;
    if (u_out) *u_out = u;
}

/* GraphState::new draws a random state (nvars x gen::<bool>) even when set_state replaces it
 * afterwards (src/lattice.rs:199-203), so the rng is advanced either way. */
static void orc_init_state(const orc_graph *g, uint64_t rng[4], const uint8_t *initial_state,
                           uint8_t *state) {
    if (!initial_state || g_opt[ORC_OPT_INIT_DRAWS])
        for (uint64_t i = 0; i < g->nvars; ++i) state[i] = (uint8_t)orc_gen_bool(rng);
    if (initial_state) memcpy(state, initial_state, g->nvars);
}

static inline uint64_t attempts_or_default(const orc_graph *g, uint64_t a) {
    return a ? a : g->nvars;
}

/* src/lattice.rs:171-221 */
ORC_EXPORT int orc_run_monte_carlo(const orc_graph *g, double beta, uint64_t timesteps,
                                   uint64_t nexp, const uint64_t *seeds,
                                   const uint8_t *initial_state, uint64_t attempts_per_step,
                                   double *energies, uint8_t *states) {
    uint64_t aps = attempts_or_default(g, attempts_per_step);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint64_t rng[4];
        uint8_t *st = states + (uint64_t)e * g->nvars;
        orc_seed_from_u64(seeds[e], rng);
        orc_init_state(g, rng, initial_state, st);
        for (uint64_t t = 0; t < timesteps; ++t)
            for (uint64_t a = 0; a < aps; ++a) orc_attempt(g, st, rng, beta, NULL, NULL);
        energies[e] = orc_energy(g, st);
    }
    return 0;
}

/* src/lattice.rs:231-299; energies[nexp, n_samples], states[nexp, n_samples, nvars]. */
ORC_EXPORT int orc_run_sampling(const orc_graph *g, double beta, uint64_t timesteps, uint64_t nexp,
                                const uint64_t *seeds, const uint8_t *initial_state,
                                uint64_t attempts_per_step, uint64_t thermalization,
                                uint64_t sampling_freq, double *energies, uint8_t *states) {
    if (sampling_freq == 0) return -1; /* reference: integer division by zero panics */
    uint64_t aps = attempts_or_default(g, attempts_per_step);
    uint64_t ns = timesteps / sampling_freq;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint64_t rng[4];
        uint8_t *st = (uint8_t *)malloc(g->nvars ? g->nvars : 1);
        orc_seed_from_u64(seeds[e], rng);
        orc_init_state(g, rng, initial_state, st);
        for (uint64_t t = 0; t < thermalization; ++t)
            for (uint64_t a = 0; a < aps; ++a) orc_attempt(g, st, rng, beta, NULL, NULL);
        for (uint64_t k = 0; k < ns; ++k) {
            for (uint64_t t = 0; t < sampling_freq; ++t)
                for (uint64_t a = 0; a < aps; ++a) orc_attempt(g, st, rng, beta, NULL, NULL);
            memcpy(states + ((uint64_t)e * ns + k) * g->nvars, st, g->nvars);
            energies[(uint64_t)e * ns + k] = orc_energy(g, st);
        }
        free(st);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Non-basic moves of GraphState::do_time_step(beta, nspinupdates, nedgeupdates, nwormupdates,
 * only_basic_moves = false), called at src/lattice.rs:205,272,278,366,452 and
 * src/classicising.rs:100-106.  [RECALLED + RESTATED]  The crate's rules are not readable here;
 * what follows is the sequential form of the moves the CUDA library implements (csrc/moves.cu),
 * pinned by exact enumeration (tests/test_oracle_moves.py):
 *   edge move: a bond (a, b) chosen uniformly - or, with importance sampling
 *     (enable_edge_importance_sampling, src/lattice.rs:200), with probability |J| / sum |J| -
 *     flips both of its spins with min(1, exp(-beta dE)); bonds between a and b keep their energy;
 *   worm move: a chain of `len` distinct sites grown from a uniform site along uniform adjacency
 *     entries (given up when it bites itself) flips as a whole with
 *     min(1, deg(first)/deg(last) exp(-beta dE)).
 * ---------------------------------------------------------------------------------------- */
static inline double orc_delta_e_excluding(const orc_graph *g, const uint8_t *state, uint64_t site,
                                           const uint64_t *skip, uint64_t nskip) {
    double de = 0.0;
    uint8_t cur = state[site];
    for (uint64_t k = g->row[site]; k < g->row[site + 1]; ++k) {
        int inside = 0;
        for (uint64_t q = 0; q < nskip; ++q) inside |= skip[q] == g->nbr[k];
        if (inside) continue;
        double coupling = (cur == state[g->nbr[k]]) ? 1.0 : -1.0;
        de += -2.0 * g->jv[k] * coupling;
    }
    return de + 2.0 * g->bias[site] * (cur ? 1.0 : -1.0);
}

static inline void orc_edge_attempt(const orc_graph *g, const uint64_t *ea, const uint64_t *eb,
                                    const double *cumw, uint8_t *state, uint64_t rng[4], double beta) {
    uint64_t e;
    if (cumw) {   /* cumulative absolute weights, binary search of u * total */
        double x = orc_gen_f64(rng) * cumw[g->nedges - 1];
        uint64_t lo = 0, hi = g->nedges - 1;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (cumw[mid] > x) hi = mid; else lo = mid + 1;
        }
        e = lo;
    } else {
        e = orc_gen_range(rng, g->nedges);
    }
    uint64_t pair[2] = {ea[e], eb[e]};
    double de = orc_delta_e_excluding(g, state, pair[0], pair, 2) +
                orc_delta_e_excluding(g, state, pair[1], pair, 2);
    int flip = 1;
    if (de > 0.0) flip = orc_gen_f64(rng) < exp(-beta * de);
    if (flip) {
        state[pair[0]] = !state[pair[0]];
        state[pair[1]] = !state[pair[1]];
    }
}

#define ORC_WORM_MAX 8
static inline void orc_worm_attempt(const orc_graph *g, uint8_t *state, uint64_t rng[4], double beta,
                                    uint64_t len) {
    uint64_t path[ORC_WORM_MAX];
    path[0] = orc_gen_range(rng, g->nvars);
    for (uint64_t t = 1; t < len; ++t) {
        uint64_t head = path[t - 1];
        uint64_t deg = g->row[head + 1] - g->row[head];
        if (deg == 0) return;
        uint64_t nxt = g->nbr[g->row[head] + orc_gen_range(rng, deg)];
        for (uint64_t q = 0; q < t; ++q)
            if (path[q] == nxt) return;
        path[t] = nxt;
    }
    double de = 0.0;
    for (uint64_t t = 0; t < len; ++t) de += orc_delta_e_excluding(g, state, path[t], path, len);
    double ratio = 1.0;
    if (len > 1)
        ratio = (double)(g->row[path[0] + 1] - g->row[path[0]]) /
                (double)(g->row[path[len - 1] + 1] - g->row[path[len - 1]]);
    double p = ratio * exp(-beta * de);
    int flip = p >= 1.0 || orc_gen_f64(rng) < p;
    if (flip)
        for (uint64_t t = 0; t < len; ++t) state[path[t]] = !state[path[t]];
}

/* The loop of src/lattice.rs:204-211 with explicit move counts per timestep. */
ORC_EXPORT int orc_run_moves(const orc_graph *g, const uint64_t *ea, const uint64_t *eb, const double *ej,
                             double beta, uint64_t timesteps, uint64_t nexp, const uint64_t *seeds,
                             const uint8_t *initial_state, uint64_t nspin, uint64_t nedge, uint64_t nworm,
                             uint64_t worm_len, int importance, double *energies, uint8_t *states) {
    if (worm_len < 1 || worm_len > ORC_WORM_MAX) return -1;
    double *cumw = NULL;
    if (importance) {
        cumw = (double *)malloc(sizeof(double) * (g->nedges ? g->nedges : 1));
        double acc = 0.0;
        for (uint64_t e = 0; e < g->nedges; ++e) { acc += fabs(ej[e]); cumw[e] = acc; }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint64_t rng[4];
        uint8_t *st = states + (uint64_t)e * g->nvars;
        orc_seed_from_u64(seeds[e], rng);
        orc_init_state(g, rng, initial_state, st);
        for (uint64_t t = 0; t < timesteps; ++t) {
            for (uint64_t a = 0; a < nspin; ++a) orc_attempt(g, st, rng, beta, NULL, NULL);
            for (uint64_t a = 0; a < nedge; ++a) orc_edge_attempt(g, ea, eb, cumw, st, rng, beta);
            for (uint64_t a = 0; a < nworm; ++a) orc_worm_attempt(g, st, rng, beta, worm_len);
        }
        energies[e] = orc_energy(g, st);
    }
    free(cumw);
    return 0;
}

/* Schedule normalisation of src/lattice.rs:320-334 (= 406-420) followed by the per-timestep
 * beta of :357-365 (= 445-451).  `sched_t/sched_b` hold n stops; out_beta gets `timesteps`
 * values.  q1_compat != 0 reproduces the reference exactly: the closure ignores its timestep
 * argument and uses the captured time `i` of the last user stop, so every timestep runs at
 * (vb-va)*((i-ia)/(ib-ia))+va of the segment that ends at that stop (NaN if 0/0).
 * q1_compat == 0 implements the documented behaviour (linear interpolation at time t). */
ORC_EXPORT int orc_schedule_betas(const uint64_t *sched_t, const double *sched_b, uint64_t n,
                                  uint64_t timesteps, int q1_compat, double *out_beta) {
    uint64_t cap = n + 3;
    uint64_t *T = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    double *B = (double *)malloc(sizeof(double) * cap);
    uint64_t m = 0;
    /* stable sort by time (sort_by_key is stable) */
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t p = m++;
        while (p > 0 && T[p - 1] > sched_t[k]) { T[p] = T[p - 1]; B[p] = B[p - 1]; --p; }
        T[p] = sched_t[k]; B[p] = sched_b[k];
    }
    if (m == 0) { T[0] = 0; B[0] = 1.0; T[1] = timesteps; B[1] = 1.0; m = 2; }
    if (T[0] > 0) {
        memmove(T + 1, T, sizeof(uint64_t) * m); memmove(B + 1, B, sizeof(double) * m);
        T[0] = 0; B[0] = B[1]; ++m;
    }
    uint64_t i_cap = T[m - 1]; /* the `i` captured at :331 / :417 */
    if (i_cap < timesteps) { T[m] = timesteps; B[m] = B[m - 1]; ++m; }
    if (m < 2) { free(T); free(B); return -1; } /* reference would index out of bounds */
    uint64_t idx = 0;
    for (uint64_t t = 0; t < timesteps; ++t) {
        uint64_t i = q1_compat ? i_cap : t;
        while (idx + 2 < m + 0 && i > T[idx + 1]) ++idx;
        uint64_t ia = T[idx], ib = T[idx + 1];
        double va = B[idx], vb = B[idx + 1];
        out_beta[t] = (vb - va) * ((double)(i - ia) / (double)(ib - ia)) + va;
    }
    free(T); free(B);
    return 0;
}

/* src/lattice.rs:309-385 (per_step_energies == 0: energies[nexp]) and 395-470
 * (per_step_energies != 0: energies[nexp, timesteps]).  betas[timesteps] from
 * orc_schedule_betas. */
ORC_EXPORT int orc_run_annealing(const orc_graph *g, const double *betas, uint64_t timesteps,
                                 uint64_t nexp, const uint64_t *seeds,
                                 const uint8_t *initial_state, uint64_t attempts_per_step,
                                 int per_step_energies, double *energies, uint8_t *states) {
    uint64_t aps = attempts_or_default(g, attempts_per_step);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint64_t rng[4];
        uint8_t *st = states + (uint64_t)e * g->nvars;
        orc_seed_from_u64(seeds[e], rng);
        orc_init_state(g, rng, initial_state, st);
        for (uint64_t t = 0; t < timesteps; ++t) {
            for (uint64_t a = 0; a < aps; ++a) orc_attempt(g, st, rng, betas[t], NULL, NULL);
            if (per_step_energies) energies[(uint64_t)e * timesteps + t] = orc_energy(g, st);
        }
        if (!per_step_energies) energies[e] = orc_energy(g, st);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Replay contract: trace = init[nvars] + nattempts x (site:u32, u:f64)
 * ---------------------------------------------------------------------------------------- */
/* Runs the reference algorithm at constant beta and records what it drew.  init_out[nexp,nvars]
 * is the state the attempts start from; sites/u are [nexp, nattempts]. */
ORC_EXPORT int orc_trace(const orc_graph *g, double beta, uint64_t nexp, const uint64_t *seeds,
                         const uint8_t *initial_state, uint64_t nattempts, uint32_t *sites,
                         double *u, uint8_t *init_out, double *energies, uint8_t *states) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint64_t rng[4];
        uint8_t *st = states + (uint64_t)e * g->nvars;
        orc_seed_from_u64(seeds[e], rng);
        orc_init_state(g, rng, initial_state, st);
        memcpy(init_out + (uint64_t)e * g->nvars, st, g->nvars);
        for (uint64_t a = 0; a < nattempts; ++a)
            orc_attempt(g, st, rng, beta, sites + (uint64_t)e * nattempts + a,
                        u + (uint64_t)e * nattempts + a);
        energies[e] = orc_energy(g, st);
    }
    return 0;
}

/* Consumes a trace: accept iff !(dE > 0) || u < exp(-beta dE). */
ORC_EXPORT int orc_replay(const orc_graph *g, double beta, uint64_t nexp, uint64_t nattempts,
                          const uint32_t *sites, const double *u, const uint8_t *init,
                          double *energies, uint8_t *states) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t e = 0; e < (int64_t)nexp; ++e) {
        uint8_t *st = states + (uint64_t)e * g->nvars;
        memcpy(st, init + (uint64_t)e * g->nvars, g->nvars);
        for (uint64_t a = 0; a < nattempts; ++a) {
            uint64_t site = sites[(uint64_t)e * nattempts + a];
            double de = orc_delta_e(g, st, site);
            int flip = 1;
            if (de > 0.0) flip = u[(uint64_t)e * nattempts + a] < exp(-beta * de);
            if (flip) st[site] = !st[site];
        }
        energies[e] = orc_energy(g, st);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Classical parallel tempering with the cadence of src/tempering.rs:156-222
 * ---------------------------------------------------------------------------------------- */
/* R replicas, slot r keeps betas[r]; a swap exchanges configurations (tempering.rs:182,192).
 * Swap step: even pairs (0,1),(2,3).. then odd pairs (1,2),(3,4)..; pair (a,b) swaps with
 * probability min(1, exp((beta_a-beta_b)(E_a-E_b))), the uniform drawn from the container rng
 * only when that probability is < 1.  Replica seeds are container_rng.gen::<u64>() in slot
 * order (tempering.rs:85); every replica starts from GraphState::new's random state.
 * Outputs follow tempering.rs:165-170,213-221: states[R, timesteps/sampling_freq, nvars] and
 * energies[R] = sum over chunks of (E after the chunk) * chunk_len / timesteps. */
ORC_EXPORT int orc_pt_run(const orc_graph *g, uint64_t R, const double *betas,
                          uint64_t container_seed, uint64_t timesteps, uint64_t replica_swap_freq,
                          uint64_t sampling_freq, uint64_t attempts_per_step, uint8_t *states,
                          double *energies, uint64_t *total_swaps) {
    if (replica_swap_freq == 0 || sampling_freq == 0) return -1; /* reference never terminates */
    uint64_t aps = attempts_or_default(g, attempts_per_step);
    uint64_t N = g->nvars, ns = timesteps / sampling_freq;
    uint64_t crng[4];
    orc_seed_from_u64(container_seed, crng);
    uint64_t(*rngs)[4] = (uint64_t(*)[4])malloc(sizeof(uint64_t[4]) * (R ? R : 1));
    uint8_t **cfg = (uint8_t **)malloc(sizeof(uint8_t *) * (R ? R : 1));
    double *ecur = (double *)calloc(R ? R : 1, sizeof(double));
    for (uint64_t r = 0; r < R; ++r) {
        orc_seed_from_u64(orc_next_u64(crng), rngs[r]);
        cfg[r] = (uint8_t *)malloc(N ? N : 1);
        orc_init_state(g, rngs[r], NULL, cfg[r]);
        energies[r] = 0.0;
    }
    uint64_t remaining = timesteps, to_swap = replica_swap_freq, to_sample = sampling_freq;
    uint64_t sample_idx = 0, swaps = 0, pt_step = 0;
    while (remaining > 0) {
        uint64_t t = to_sample < to_swap ? to_sample : to_swap;
        if (remaining < t) t = remaining;
#pragma omp parallel for schedule(dynamic, 1)
        for (int64_t r = 0; r < (int64_t)R; ++r) {
            for (uint64_t k = 0; k < t * aps; ++k)
                orc_attempt(g, cfg[r], rngs[r], betas[r], NULL, NULL);
            ecur[r] = orc_energy(g, cfg[r]);
            energies[r] += ecur[r] * (double)t;
        }
        to_sample -= t; to_swap -= t; remaining -= t;
        if (to_swap == 0) {
            /* ORC_OPT_PT_PAIRS: both parities (default) / one, alternating / one, drawn */
            int p_lo = 0, p_hi = 2;
            if (g_opt[ORC_OPT_PT_PAIRS] == 1) { p_lo = (int)(pt_step & 1u); p_hi = p_lo + 1; }
            if (g_opt[ORC_OPT_PT_PAIRS] == 2) { p_lo = orc_gen_bool(crng); p_hi = p_lo + 1; }
            ++pt_step;
            for (int parity = p_lo; parity < p_hi; ++parity)
                for (uint64_t a = parity; a + 1 < R; a += 2) {
                    double d = (betas[a] - betas[a + 1]) * (ecur[a] - ecur[a + 1]);
                    int acc = 1;
                    if (d < 0.0) acc = orc_gen_f64(crng) < exp(d);
                    if (acc) {
                        uint8_t *tc = cfg[a]; cfg[a] = cfg[a + 1]; cfg[a + 1] = tc;
                        double te = ecur[a]; ecur[a] = ecur[a + 1]; ecur[a + 1] = te;
                        ++swaps;
                    }
                }
            to_swap = replica_swap_freq;
        }
        if (to_sample == 0) {
            if (sample_idx < ns)
                for (uint64_t r = 0; r < R; ++r)
                    memcpy(states + (r * ns + sample_idx) * N, cfg[r], N);
            ++sample_idx;
            to_sample = sampling_freq;
        }
    }
    for (uint64_t r = 0; r < R; ++r) { energies[r] /= (double)timesteps; free(cfg[r]); }
    if (total_swaps) *total_swaps = swaps;
    free(cfg); free(rngs); free(ecur);
    return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline leg wants every host core */
ORC_EXPORT void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

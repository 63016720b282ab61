/*
 * oracle/msc_mirror.c -- scalar CPU restatement of the PRODUCTION-RNG sweep of libising_b200
 * (DESIGN.md "Production sweep"), one spin at a time, for bit-exact checks of the CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY (see ising_oracle.c).  This file does not restate the reference: the
 * reference updates uniformly random sites sequentially with xoshiro256++ (restated in
 * ising_oracle.c), whereas the device path sweeps colour classes with Philox randomness.  The
 * two agree in distribution (tests/test_statistics*.py), not draw by draw; this mirror pins the
 * device path's exact bits so that a kernel optimisation cannot silently change results.
 *
 * Algorithm (graphs with all |J| equal and no bias):
 *   for sweep t, colour c = 0..C-1, every site n of colour c, every replica word w:
 *     n_sat = satisfied bonds (J s s' < 0) of n, d = degree;  dE = 2|J|(2 n_sat - d)
 *     dE <= 0 -> flip.  dE > 0 -> flip iff U < T(dE), T = floor(exp(-beta dE) 2^(K+32)):
 *       words R_m = output m%4 of Philox4x32-R(key = seed, ctr = (n, w, t, m/4));
 *       bit b of R_0..R_{K-1} are the K most significant bits of U for replica 32w+b, compared
 *       MSB first; bits of the word still tied after K planes are taken in ascending b and the
 *       j-th of them accepts iff R_{K+j} < T mod 2^32.
 *   initial state: bit b of Philox4x32-10(key, ctr = (n, w, 0, 1<<24)).x
 *   energy: |J| (n_edges - 2 n_sat_total)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

static void philox4x32(int rounds, uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += W0;
        k1 += W1;
    }
}

/* known-answer hook for the tests */
ORC_EXPORT void msc_philox4x32(int rounds, const uint32_t ctr[4], const uint32_t key[2],
                               uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32(rounds, c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}

static uint32_t stream_word(int rounds, uint32_t site, uint32_t gw, uint32_t sweep, uint32_t m,
                            uint32_t k0, uint32_t k1) {
    uint32_t c[4] = {site, gw, sweep, m >> 2};
    philox4x32(rounds, c, k0, k1);
    return c[m & 3];
}

static uint64_t threshold(double beta, double de, int K) {
    const double scaled = ldexp(exp(-beta * de), K + 32);
    const uint64_t tmax = (1ull << (K + 32)) - 1;
    if (!(scaled >= 0.0)) return 0;
    if (scaled >= (double)tmax) return tmax;
    return (uint64_t)floor(scaled);
}

/* states: bool[E, nvars], in/out (filled from Philox when randomize != 0, or broadcast from
 * init_state when given).  energies_per_sweep: double[E, nsweeps] or NULL.  final_energies:
 * double[E] or NULL. */
ORC_EXPORT int msc_mirror_run(uint64_t nvars, uint64_t nedges, const uint64_t *ea,
                              const uint64_t *eb, const double *ej, const uint32_t *colors,
                              uint32_t ncolors, uint64_t E, uint64_t seed,
                              uint64_t replica_offset, int K, int rounds, int randomize,
                              const uint8_t *init_state, const double *betas, uint64_t nsweeps,
                              uint64_t sweep0, uint8_t *states, double *energies_per_sweep,
                              double *final_energies) {
    if (nedges == 0 || K < 1 || K > 8) return -1;
    const double jabs = fabs(ej[0]);
    for (uint64_t e = 0; e < nedges; ++e)
        if (fabs(ej[e]) != jabs) return -2;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t gw0 = (uint32_t)(replica_offset / 32);
    /* adjacency with antiferro flag */
    uint64_t *row = (uint64_t *)calloc(nvars + 1, sizeof(uint64_t));
    for (uint64_t e = 0; e < nedges; ++e) { row[ea[e] + 1]++; row[eb[e] + 1]++; }
    for (uint64_t i = 0; i < nvars; ++i) row[i + 1] += row[i];
    uint64_t *nbr = (uint64_t *)malloc(sizeof(uint64_t) * 2 * nedges);
    uint8_t *anti = (uint8_t *)malloc(2 * nedges);
    uint64_t *fill = (uint64_t *)calloc(nvars, sizeof(uint64_t));
    for (uint64_t e = 0; e < nedges; ++e) {
        uint64_t a = ea[e], b = eb[e];
        nbr[row[a] + fill[a]] = b; anti[row[a] + fill[a]++] = ej[e] > 0;
        nbr[row[b] + fill[b]] = a; anti[row[b] + fill[b]++] = ej[e] > 0;
    }
    free(fill);
    const uint64_t W = (E + 31) / 32;
    if (randomize)
        for (uint64_t n = 0; n < nvars; ++n)
            for (uint64_t w = 0; w < W; ++w) {
                uint32_t c[4] = {(uint32_t)n, gw0 + (uint32_t)w, 0u, 1u << 24};
                philox4x32(10, c, k0, k1);
                for (uint64_t b = 0; b < 32 && w * 32 + b < E; ++b)
                    states[(w * 32 + b) * nvars + n] = (c[0] >> b) & 1u;
            }
    else if (init_state)
        for (uint64_t e = 0; e < E; ++e) memcpy(states + e * nvars, init_state, nvars);

    for (uint64_t t = 0; t < nsweeps; ++t) {
        const double beta = betas[t];
        const uint32_t sweep = (uint32_t)(sweep0 + t);
        for (uint32_t col = 0; col < ncolors; ++col)
            for (uint64_t n = 0; n < nvars; ++n) {
                if (colors[n] != col) continue;
                const int d = (int)(row[n + 1] - row[n]);
                for (uint64_t w = 0; w < W; ++w) {
                    int j = 0; /* resolver rank within the word */
                    for (uint64_t b = 0; b < 32 && w * 32 + b < E; ++b) {
                        uint8_t *st = states + (w * 32 + b) * nvars;
                        int nsat = 0;
                        for (uint64_t k = row[n]; k < row[n + 1]; ++k) {
                            const int equal = st[n] == st[nbr[k]];
                            nsat += anti[k] ? !equal : equal;
                        }
                        const int cls = 2 * nsat - d;
                        if (cls <= 0) { st[n] ^= 1; continue; }
                        const uint64_t T = threshold(beta, 2.0 * jabs * (double)cls, K);
                        int decided = 0, accept = 0;
                        for (int p = 0; p < K && !decided; ++p) {
                            const uint32_t rb =
                                (stream_word(rounds, (uint32_t)n, gw0 + (uint32_t)w, sweep,
                                             (uint32_t)p, k0, k1) >> b) & 1u;
                            const uint32_t tb = (uint32_t)((T >> (K + 31 - p)) & 1ull);
                            if (rb != tb) { decided = 1; accept = rb < tb; }
                        }
                        if (!decided) {
                            const uint32_t v = stream_word(rounds, (uint32_t)n, gw0 + (uint32_t)w,
                                                           sweep, (uint32_t)(K + j), k0, k1);
                            accept = v < (uint32_t)(T & 0xFFFFFFFFull);
                            ++j;
                        }
                        if (accept) st[n] ^= 1;
                    }
                }
            }
        if (energies_per_sweep || (final_energies && t + 1 == nsweeps))
            for (uint64_t e = 0; e < E; ++e) {
                const uint8_t *st = states + e * nvars;
                long long nsat = 0;
                for (uint64_t k = 0; k < nedges; ++k) {
                    const int equal = st[ea[k]] == st[eb[k]];
                    nsat += (ej[k] > 0) ? !equal : equal;
                }
                const double en = jabs * (double)((long long)nedges - 2 * nsat);
                if (energies_per_sweep) energies_per_sweep[e * nsweeps + t] = en;
                if (final_energies && t + 1 == nsweeps) final_energies[e] = en;
            }
    }
    if (final_energies && nsweeps == 0)
        for (uint64_t e = 0; e < E; ++e) {
            const uint8_t *st = states + e * nvars;
            long long nsat = 0;
            for (uint64_t k = 0; k < nedges; ++k) {
                const int equal = st[ea[k]] == st[eb[k]];
                nsat += (ej[k] > 0) ? !equal : equal;
            }
            final_energies[e] = jabs * (double)((long long)nedges - 2 * nsat);
        }
    free(row); free(nbr); free(anti);
    return 0;
}

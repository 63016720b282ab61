// K2/K3, large lattices: one colour phase of a checkerboard sweep as a persistent row walk.
//
// A block owns a balanced, contiguous range of work units (unit = NRS consecutive rows of one z
// plane x one tile of half-row positions x one tile of replica-word groups) and a thread keeps its
// (xh, word group) as long as the tile does not change.  The row geometry of a chunk of units
// (row offsets of the five neighbour rows, Philox site base, parity) is computed once per chunk
// into shared memory, so the per-site work is two broadcast LDS, the seven vector loads, the
// Philox calls, the bit-sliced neighbour count and the Metropolis mask - no per-site index
// arithmetic on the ALU pipe, which is the pipe this kernel saturates
// (profiles/r01_sweep_metrics.md: ALU 60-67 %, FMA heavy 42 %).  Two more things move work from
// the ALU pipe to the FMA pipe:
//   * the class select of the threshold planes: with one-hot class masks m1, m2 (disjoint) the
//     plane word is t = c0 + m1 * d1 + m2 * d2 in 32-bit arithmetic (d = difference of the 0/1
//     threshold bits, c0 = all-ones or 0) - one IMAD per class instead of one LOP3;
//   * the tie resolver's comparisons r < low_c: the high word of r + (2^64 - low_c), one
//     IMAD.WIDE, is the all-ones / zero accept mask of class c.
// Philox: the V words x 2 calls of a site share the counter words (site, sweep), so rounds 1-3
// need 1 + (V + 2) + (V + 2) multiplications instead of 3 * 4V (written out in philox_site).
// Results are bit-identical to sweep_colour_phase / oracle/msc_mirror.c.
#pragma once
#include "msc_device.cuh"

// tuning knobs (profiles/microbench/rows_variants.sh builds and times the alternatives)
#ifndef ISING_ROWS_THREADS
#define ISING_ROWS_THREADS 256   // threads per block of the row walk
#endif
#ifndef ISING_ROWS_MINB
#define ISING_ROWS_MINB 3        // resident blocks per SM the plain colour phase is compiled for
#endif
#ifndef ISING_ROWS_TMA_MINB
#define ISING_ROWS_TMA_MINB 4    // ... the TMA-staged plain colour phase (no registers for loads in flight)
#endif
#ifndef ISING_ROWS_ACC_MINB
#define ISING_ROWS_ACC_MINB 2    // ... the accumulating colour phase
#endif
#ifndef ISING_ROWS_DEFER_RARE
// third-and-later ties of a word (their words come from continuation rounds of the word's second
// Philox block, msc_device.cuh): 0 = divergent loop inside the word, where that block is still in
// registers; 1 = all V words after the word loop (the blocks are kept: 4 V registers); 2 = form 1 in
// the plain colour phase, form 0 in the accumulating phase and in the 128-thread shape, which
// have no registers to spare.
// Config 3 on the annealing ramp, us per sweep without / with energies (profiles/
// r02_beta_dependence.log): form 0 63.6 / 72.0, form 1 62.3 / 72.2; with the third Philox call
// these words used to cost (r02_rare_path_ab.log): 68.9 / 79.0.
#define ISING_ROWS_DEFER_RARE 2
#endif
#ifndef ISING_ROWS_SPLIT_ACC_DEFAULT
#define ISING_ROWS_SPLIT_ACC_DEFAULT 0  // 1: per-sweep energies by a count-only pass instead of the fused phase
#endif
#ifndef ISING_ROWS_LT_SEL
#define ISING_ROWS_LT_SEL 0      // 1: tie compare with ISETP + SEL instead of the multiply-add carry
#endif
#ifndef ISING_ROWS_ZSKIP
// 1: the leading planes whose threshold bit is zero in every class (floor(5.77 beta |J|) of them on a
// cubic lattice: all six from beta = 1.04) skip the class select: a uniform is below such a threshold
// prefix only if its own bits are zero there, so Z planes cost ceil((Z - 1) / 2) + 2 LOP3 instead of
// 2 Z LOP3 + 2 Z IMAD.  The plane count is a launch constant (MscMux::z): a uniform switch per word.
#define ISING_ROWS_ZSKIP 0
#endif
#ifndef ISING_ROWS_MUL_SPLIT
#define ISING_ROWS_MUL_SPLIT 1   // 1: Philox products as mul.hi + mul.lo instead of one 32x32->64 multiply
#endif

namespace ising {

struct MscMux {
    uint32_t c0[8];   // plane p of class 0 (all-ones / 0)
    uint32_t d1[8];   // bit(class 1) - bit(class 0)  in {0, 1, 0xFFFFFFFF}
    uint32_t d2[8];   // bit(class 2) - bit(class 0)
    uint32_t low[3];
    uint32_t one;     // 1 (keeps r * one + c an IMAD.WIDE)
    uint32_t z;       // planes 0 .. z-1 (the most significant ones) are zero in every class
};

inline MscMux make_mux(const MscThresholds& th) {
    MscMux m;
    for (int p = 0; p < 8; ++p) {
        m.c0[p] = th.plane[0][p];
        m.d1[p] = (th.plane[1][p] & 1u) - (th.plane[0][p] & 1u);
        m.d2[p] = (th.plane[2][p] & 1u) - (th.plane[0][p] & 1u);
    }
    for (int c = 0; c < 3; ++c) m.low[c] = th.low[c];
    m.one = 1u;
    m.z = 0;
    while (m.z < 8 && (th.plane[0][m.z] | th.plane[1][m.z] | th.plane[2][m.z]) == 0u) ++m.z;
    return m;
}

struct RowsArgs {
    uint32_t* own;          // colour being updated   [rows][Lxh][W]
    const uint32_t* oth;    // the other colour
    const uint4* jm8;       // +-J: bond masks of this colour [rows][Lxh][2] uint4 (k = 0..5, 2 pad)
    uint32_t Lx, Ly, Lz, Lxh, W;
    uint32_t c, sweep, gw0, antiferro;
    uint32_t bxh_log;       // threadIdx.y = rsub << bxh_log | xh_local
    uint32_t nrs_log;       // rows per unit = 1 << nrs_log (divides Ly)
    uint32_t ygroups, xtiles;
    uint32_t units;         // tiles * Lz * ygroups, tile = wt * xtiles + xt slowest
    uint32_t uq, urem;      // units = uq * gridDim.x + urem: block b gets uq (+1 if b < urem) units
    unsigned long long* nsat;
    uint32_t nsat_copies, nsat_stride;   // blocks add into copy (block % copies), see SweepArgs
    PhiloxKeys pk;
    MscMux mx;
};

// geometry of one row of one unit (shared memory, built once per chunk of units)
struct RowDesc {
    uint32_t e_row, e_ym, e_yp, e_zm, e_zp;  // row base offsets in vector elements (V words)
    uint32_t site0;                          // row * Lx + parity
    uint32_t par_tile;                       // parity | tile << 1;  0xFFFFFFFF = no row (y >= Ly)
    uint32_t jrow;                           // row * Lxh * 2 (uint4 index of the row's bond masks)
};
constexpr int ROWS_DESC_CHUNK = 64;          // RowDesc entries per chunk (2 KiB)

__device__ __forceinline__ void mulhilo(uint32_t m, uint32_t x, uint32_t& hi, uint32_t& lo) {
#if ISING_ROWS_MUL_SPLIT
    hi = __umulhi(m, x);
    lo = m * x;
#else
    const uint64_t p = (uint64_t)m * x;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

// Two Philox4x32 calls (q = 0, 1) for each of V replica words of one site: counter
// (site, gw0 + v, sweep, q | TAG_ACCEPT << 24).  r[v][4 q + i] = output word i of call q.
template <int ROUNDS, int V>
struct PhiloxSite {
    // state after round 3, shared parts
    uint32_t c1v[V], c3q[2], hq0[2], hq1v[V], c1q[2], c3v[V];
    __device__ __forceinline__ void prepare(uint32_t site, uint32_t gw0v, uint32_t sweep,
                                            const PhiloxKeys& pk) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        // round 1
        const uint64_t p0 = (uint64_t)M0 * site;
        const uint64_t p1 = (uint64_t)M1 * sweep;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
        const uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
        // round 2: c0 = h1 ^ gw ^ k0 (per word), c1 = l1, c2 = h0 ^ cq ^ k1 (per call), c3 = l0
        uint32_t c0q[2], c2v[V];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t c2 = h0 ^ ((uint32_t)q | (TAG_ACCEPT << 24)) ^ pk.k[1];
            const uint64_t P1 = (uint64_t)M1 * c2;
            c0q[q] = (uint32_t)(P1 >> 32) ^ l1 ^ pk.k[2];
            c1q[q] = (uint32_t)P1;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t c0 = h1 ^ (gw0v + v) ^ pk.k[0];
            const uint64_t P0 = (uint64_t)M0 * c0;
            c2v[v] = (uint32_t)(P0 >> 32) ^ l0 ^ pk.k[3];
            c3v[v] = (uint32_t)P0;
        }
        // round 3 products: M0 * c0 (per call), M1 * c2 (per word)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint64_t Q0 = (uint64_t)M0 * c0q[q];
            hq0[q] = (uint32_t)(Q0 >> 32);
            c3q[q] = (uint32_t)Q0;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint64_t Q1 = (uint64_t)M1 * c2v[v];
            hq1v[v] = (uint32_t)(Q1 >> 32);
            c1v[v] = (uint32_t)Q1;
        }
    }
    // rounds 4 .. ROUNDS of call q of word v
    __device__ __forceinline__ void finish(int v, int q, const PhiloxKeys& pk, uint32_t* out) const {
        rounds_from_4(hq1v[v] ^ c1q[q] ^ pk.k[4], c1v[v], hq0[q] ^ c3v[v] ^ pk.k[5], c3q[q], pk, out);
    }
    __device__ __forceinline__ static void rounds_from_4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         const PhiloxKeys& pk, uint32_t* out) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
        for (int r = 3; r < ROUNDS; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(M0, c0, h0, l0);
            mulhilo(M1, c2, h1, l1);
            const uint32_t n0 = h1 ^ c1 ^ pk.k[2 * r];
            const uint32_t n2 = h0 ^ c3 ^ pk.k[2 * r + 1];
            c1 = l1;
            c3 = l0;
            c0 = n0;
            c2 = n2;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// all-ones iff r < lo, on the FMA pipe: high word of r * 1 + (2^64 - lo)
__device__ __forceinline__ uint32_t lt_mask(uint32_t r, uint32_t lo, uint32_t one) {
#if ISING_ROWS_LT_SEL
    return r < lo ? 0xFFFFFFFFu : 0u;
#else
    const uint64_t neg = 0ull - (uint64_t)lo;
    return (uint32_t)(((uint64_t)r * one + neg) >> 32);
#endif
}

// The K compare steps of one word with the Z most significant planes known to be zero in every class
// (borrow of U_top - T_top and the mask of the bits that tie): planes K-1 .. Z in full, then the
// leading ones at once - U_top < T_top and U_top == T_top both need the uniform's bits to be zero there.
template <int NCLS, int K, int Z>
__device__ __forceinline__ void msc_compare_planes(uint32_t m1, uint32_t m2, const MscMux& mx,
                                                   const uint32_t (&r)[8], uint32_t& eq, uint32_t& borrow) {
#pragma unroll
    for (int p = K - 1; p >= Z; --p) {
        uint32_t t = m1 * mx.d1[p] + mx.c0[p];
        if (NCLS == 3) t += m2 * mx.d2[p];
        borrow = maj3(~r[p], t, borrow);
        eq &= ~(r[p] ^ t);
    }
    if (Z > 0) {
        uint32_t nr = r[0];
#pragma unroll
        for (int p = 1; p < Z; ++p) nr |= r[p];
        borrow &= ~nr;
        eq &= ~nr;
    }
}

// Metropolis mask of one word from eight random words r (two Philox calls).
//   up: dE > 0;  m1, m2: one-hot masks of uphill classes 1 and 2 (class 0 = up & ~m1 & ~m2)
template <int NCLS, int K, int ROUNDS>
__device__ __forceinline__ uint32_t msc_flip_mask_mux(uint32_t up, uint32_t m1, uint32_t m2,
                                                      const MscMux& mx, const uint32_t (&r)[8],
                                                      uint32_t site, uint32_t gw, uint32_t sweep,
                                                      const PhiloxKeys& pk, uint32_t* eq_left = nullptr) {
    static_assert(K >= 4 && K <= 7, "two Philox calls: K planes + (8 - K) resolver words");
    uint32_t eq = up, borrow = 0;
#if ISING_ROWS_ZSKIP
    switch (mx.z < (uint32_t)K ? mx.z : (uint32_t)K) {   // launch constant: a uniform branch
        case 0: msc_compare_planes<NCLS, K, 0>(m1, m2, mx, r, eq, borrow); break;
        case 1: msc_compare_planes<NCLS, K, 1>(m1, m2, mx, r, eq, borrow); break;
        case 2: msc_compare_planes<NCLS, K, 2>(m1, m2, mx, r, eq, borrow); break;
        case 3: msc_compare_planes<NCLS, K, 3>(m1, m2, mx, r, eq, borrow); break;
        case 4: msc_compare_planes<NCLS, K, 4>(m1, m2, mx, r, eq, borrow); break;
        case 5: msc_compare_planes<NCLS, K, (K >= 5 ? 5 : K)>(m1, m2, mx, r, eq, borrow); break;
        case 6: msc_compare_planes<NCLS, K, (K >= 6 ? 6 : K)>(m1, m2, mx, r, eq, borrow); break;
        default: msc_compare_planes<NCLS, K, K>(m1, m2, mx, r, eq, borrow); break;
    }
#else
    msc_compare_planes<NCLS, K, 0>(m1, m2, mx, r, eq, borrow);
#endif
    uint32_t flip = ~up | (borrow & ~eq);
    // Tied bits (2^-K each): the first SPARE of a word, in ascending bit position, compare the
    // words left over from the two calls against the low threshold bits of their class.
    constexpr int SPARE = (8 - K) < 2 ? (8 - K) : 2;
#pragma unroll
    for (int j = 0; j < SPARE; ++j) {
        const uint32_t bit = eq & (0u - eq);
        uint32_t acc = lt_mask(r[K + j], mx.low[0], mx.one);
        acc = (m1 & lt_mask(r[K + j], mx.low[1], mx.one)) | (~m1 & acc);
        if (NCLS == 3) acc = (m2 & lt_mask(r[K + j], mx.low[2], mx.one)) | (~m2 & acc);
        flip |= bit & acc;
        eq -= bit;
    }
    if (eq_left) {  // caller resolves the remaining ties after its word loop (msc_resolve_rest)
        *eq_left = eq;
        return flip;
    }
    if (eq) {  // third tie of a word (rare): continuation rounds of the second block, as msc_flip_mask
        int j = K + SPARE;
        u32x4 cur = {r[4], r[5], r[6], r[7]};
        do {
            const int b = __ffs((int)eq) - 1;
            if ((j & 3) == 0 && j >= 8)
                cur = philox4x32_more(cur, (uint32_t)(ROUNDS + (j >> 2) - 2), pk.k[0], pk.k[1]);
            const int m = j & 3;
            const uint32_t v = m == 0 ? cur.x : (m == 1 ? cur.y : (m == 2 ? cur.z : cur.w));
            uint32_t lo = ((m1 >> b) & 1u) ? mx.low[1] : mx.low[0];
            if (NCLS == 3 && ((m2 >> b) & 1u)) lo = mx.low[2];
            if (v < lo) flip |= 1u << b;
            eq &= eq - 1;
            ++j;
        } while (eq);
    }
    return flip;
}

// third and later ties of a word (deferred form): resolver words 8, 9, ... = continuation rounds
// of the word's second Philox block `cur`
template <int NCLS, int K, int ROUNDS>
__device__ __noinline__ uint32_t msc_resolve_rest(uint32_t eq, uint32_t m1, uint32_t m2, uint32_t low0,
                                                  uint32_t low1, uint32_t low2, u32x4 cur,
                                                  const PhiloxKeys& pk) {
    constexpr int SPARE = (8 - K) < 2 ? (8 - K) : 2;
    uint32_t flip = 0;
    int j = K + SPARE;   // below 8: words of the second block that the straight-line ties left over
    do {
        const int b = __ffs((int)eq) - 1;
        if ((j & 3) == 0 && j >= 8)
            cur = philox4x32_more(cur, (uint32_t)(ROUNDS + (j >> 2) - 2), pk.k[0], pk.k[1]);
        const int m = j & 3;
        const uint32_t v = m == 0 ? cur.x : (m == 1 ? cur.y : (m == 2 ? cur.z : cur.w));
        uint32_t lo = ((m1 >> b) & 1u) ? low1 : low0;
        if (NCLS == 3 && ((m2 >> b) & 1u)) lo = low2;
        if (v < lo) flip |= 1u << b;
        eq &= eq - 1;
        ++j;
    } while (eq);
    return flip;
}

// The update of one site for V replica words: two Philox calls per word, bit-sliced count of the
// satisfied bonds, Metropolis mask, (ACC) accumulation of the post-flip count.
//   s: the site's words (updated in place);  n[k]: neighbour words;  m[k]: bond masks
//   NPC: planes of the per-thread counters (ACC)
template <int DIM, int K, int ROUNDS, int V, bool ACC, int NPC = SW_NP>
__device__ __forceinline__ void update_site(uint32_t (&s)[V], const uint32_t (&n)[2 * DIM][V],
                                            const uint32_t (&m)[2 * DIM], uint32_t site, uint32_t gw0w,
                                            uint32_t sweep, const PhiloxKeys& pk, const MscMux& mx,
                                            VCount<ACC ? NPC : 1> (&vc)[V]) {
    uint32_t s0[V];
#pragma unroll
    for (int v = 0; v < V; ++v) s0[v] = s[v];
    PhiloxSite<ROUNDS, V> ph;
    ph.prepare(site, gw0w, sweep, pk);
    static_assert(ISING_ROWS_DEFER_RARE >= 0 && ISING_ROWS_DEFER_RARE <= 2, "see the knob's description");
    // (NPC != SW_NP marks the 128-thread shape: compiled for 73 registers, it would spill the kept blocks)
    constexpr bool kDefer = ISING_ROWS_DEFER_RARE == 1 || (ISING_ROWS_DEFER_RARE == 2 && !ACC && NPC == SW_NP);
    uint32_t left[V], lm1[V], lm2[V];
    u32x4 blk1[kDefer ? V : 1];   // deferred form: the second Philox block of every word
    uint32_t any_left = 0;
    uint32_t nb0[V], nb1[V], nb2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        uint32_t r[8];
        ph.finish(v, 0, pk, r);
        ph.finish(v, 1, pk, r + 4);
        uint32_t av[2 * DIM];
#pragma unroll
        for (int k2 = 0; k2 < 2 * DIM; ++k2) av[k2] = ~(s[v] ^ n[k2][v] ^ m[k2]);
        uint32_t b0, b1, b2;
        count_sat<DIM>(av, b0, b1, b2);
        // 3D: n_sat 4, 5, 6 -> dE = 4, 8, 12 |J|;  2D: n_sat 3, 4 -> dE = 4, 8 |J|
        const uint32_t up = DIM == 3 ? b2 : (b2 | (b1 & b0));
        const uint32_t m1 = DIM == 3 ? (b2 & b0) : b2;
        const uint32_t m2 = DIM == 3 ? (b2 & b1) : 0u;
        left[v] = 0;
        uint32_t flip = msc_flip_mask_mux<DIM == 3 ? 3 : 2, K, ROUNDS>(
            up, m1, m2, mx, r, site, gw0w + v, sweep, pk, kDefer ? &left[v] : nullptr);
        s[v] ^= flip;
        if constexpr (kDefer) {
            blk1[v] = u32x4{r[4], r[5], r[6], r[7]};
            lm1[v] = m1;
            lm2[v] = m2;
            any_left |= left[v];
            nb0[v] = b0; nb1[v] = b1; nb2[v] = b2;
        }
        if constexpr (ACC && !kDefer) {
            // a flipped spin turns its n_sat satisfied bonds into 2*DIM - n_sat
            uint32_t c1, c2;
            if (DIM == 3) {
                c1 = (flip & ~(b1 ^ b0)) | (~flip & b1);
                c2 = (flip & ~b2 & ~(b1 & b0)) | (~flip & b2);
            } else {
                c1 = (flip & (b1 ^ b0)) | (~flip & b1);
                c2 = (flip & ~(b2 | b1 | b0)) | (~flip & b2);
            }
            vc[v].add3(b0, c1, c2);
        }
    }
    if constexpr (kDefer) {
        uint32_t extra[V];
#pragma unroll
        for (int v = 0; v < V; ++v) extra[v] = 0;
        if (any_left) {  // a word with three or more ties (rare)
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (left[v])
                    extra[v] = msc_resolve_rest<DIM == 3 ? 3 : 2, K, ROUNDS>(
                        left[v], lm1[v], lm2[v], mx.low[0], mx.low[1], mx.low[2], blk1[v], pk);
#pragma unroll
            for (int v = 0; v < V; ++v) s[v] ^= extra[v];
        }
        if constexpr (ACC) {
            // s0 = spins before the update: flip = s ^ s0 is not kept, recompute from
            // the stored word: flipped bits = bits where the final s differs
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const uint32_t b0 = nb0[v], b1 = nb1[v], b2 = nb2[v];
                const uint32_t flip = s[v] ^ s0[v];
                uint32_t c1, c2;
                if (DIM == 3) {
                    c1 = (flip & ~(b1 ^ b0)) | (~flip & b1);
                    c2 = (flip & ~b2 & ~(b1 & b0)) | (~flip & b2);
                } else {
                    c1 = (flip & (b1 ^ b0)) | (~flip & b1);
                    c2 = (flip & ~(b2 | b1 | b0)) | (~flip & b2);
                }
                vc[v].add3(b0, c1, c2);
            }
        }
    }
}

// COUNT_ONLY (with ACC): no update, nsat[e] += satisfied bonds seen from the sites of this colour
// (every bond once) - get_energy() of the current configuration with the same row walk.
// NPC / NRC: planes of the per-thread and of the block-level counters of the accumulating phase.  The
// 128-thread shape runs with 5 / 12 (at most 5 site groups between two reductions, 128 * 31 < 2^12)
// instead of 7 / 20: at one or two site groups per thread the block reduction is a third of the
// phase's instructions (profiles/r02_sweep_small_metrics.md), and it shrinks with the plane counts.
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool MULTIROW, bool COUNT_ONLY = false,
          int NPC = SW_NP, int NRC = NS_NR>
__device__ __forceinline__ void sweep_rows_phase(const RowsArgs& a, uint32_t* sm) {
    constexpr int MAX_ITEMS = ((1 << NPC) - 1) / 6;
    typedef typename WordVec<V>::type VecT;
    __shared__ RowDesc s_desc[ROWS_DESC_CHUNK];
    const uint32_t Lxh = a.Lxh, W = a.W, Ly = a.Ly, Lz = a.Lz;
    const uint32_t rowlenV = Lxh * (W / V);  // vector elements per colour row (W % V == 0)
    const uint32_t wx = blockDim.x;
    const uint32_t nthreads = blockDim.x * blockDim.y;
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint32_t nrs_log = MULTIROW ? a.nrs_log : 0u;
    const uint32_t bxh = MULTIROW ? (1u << a.bxh_log) : blockDim.y;
    const uint32_t xh_l = MULTIROW ? (threadIdx.y & (bxh - 1u)) : threadIdx.y;
    const uint32_t rsub = MULTIROW ? (threadIdx.y >> a.bxh_log) : 0u;
    // balanced contiguous unit range of this block
    const uint32_t b = blockIdx.x;
    const uint32_t u0 = b * a.uq + (b < a.urem ? b : a.urem);
    const uint32_t u1 = u0 + a.uq + (b < a.urem ? 1u : 0u);
    const VecT* __restrict__ othv = reinterpret_cast<const VecT*>(a.oth);
    VecT* __restrict__ ownv = reinterpret_cast<VecT*>(a.own);
    // Programmatic dependent launch: the next colour phase may be scheduled while this one
    // drains (its blocks start as SMs free up and run their preamble), and this phase must not
    // touch the spins before the previous one has completed.  No-ops on an ordinary launch.
    // The wait sits behind the first chunk's row geometry (which touches no spins): that part of the
    // preamble overlaps the tail of the previous phase as well.
    asm volatile("griddepcontrol.launch_dependents;");
    bool waited = false;

    VCount<ACC ? NPC : 1> vc[V];
    if constexpr (ACC) {
#pragma unroll
        for (int v = 0; v < V; ++v) vc[v].clear();
    }
    unsigned long long* const nsat = ACC ? a.nsat + (size_t)(blockIdx.x % a.nsat_copies) * a.nsat_stride : nullptr;
    int pending = 0;
    uint32_t cur_tile = 0xFFFFFFFFu;
    uint32_t w = 0, xh2 = 0, toff = 0, offP = 0, offM = 0, xsite = 0;
    bool col_ok = false;
    const uint32_t chunk_units = ROWS_DESC_CHUNK >> nrs_log;

    for (uint32_t uc = u0; uc < u1; uc += chunk_units) {
        const uint32_t nu = u1 - uc < chunk_units ? u1 - uc : chunk_units;
        __syncthreads();
        for (uint32_t i = tid; i < (nu << nrs_log); i += nthreads) {
            const uint32_t u = uc + (i >> nrs_log), rs = i & ((1u << nrs_log) - 1u);
            const uint32_t yg = u % a.ygroups;
            uint32_t t = u / a.ygroups;
            const uint32_t z = t % Lz, tile = t / Lz;
            const uint32_t y = (yg << nrs_log) + rs;
            RowDesc d;
            if (y < Ly) {
                const uint32_t ym = y == 0 ? Ly - 1 : y - 1, yp = y + 1 == Ly ? 0 : y + 1;
                const uint32_t zm = z == 0 ? Lz - 1 : z - 1, zp = z + 1 == Lz ? 0 : z + 1;
                const uint32_t row = z * Ly + y, p = (y + z + a.c) & 1u;
                d.e_row = row * rowlenV;
                d.e_ym = (z * Ly + ym) * rowlenV;
                d.e_yp = (z * Ly + yp) * rowlenV;
                d.e_zm = (zm * Ly + y) * rowlenV;
                d.e_zp = (zp * Ly + y) * rowlenV;
                d.site0 = row * a.Lx + p;
                d.par_tile = p | (tile << 1);
                d.jrow = row * Lxh * 2u;
            } else {
                d.e_row = d.e_ym = d.e_yp = d.e_zm = d.e_zp = d.site0 = d.jrow = 0u;
                d.par_tile = 0xFFFFFFFFu;
            }
            s_desc[i] = d;
        }
        if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            waited = true;
        }
        __syncthreads();
        for (uint32_t k = 0; k < nu; ++k) {
            const uint4 da = reinterpret_cast<const uint4*>(&s_desc[(k << nrs_log) + rsub])[0];
            const uint4 db = reinterpret_cast<const uint4*>(&s_desc[(k << nrs_log) + rsub])[1];
            // da = (e_row, e_ym, e_yp, e_zm), db = (e_zp, site0, par_tile, jrow)
            // the tile is the same for every row of a unit, so this branch is block-uniform
            const uint32_t tile = reinterpret_cast<const uint32_t*>(&s_desc[k << nrs_log])[6] >> 1;
            if (tile != cur_tile) {
                if constexpr (ACC) {
                    if (pending) {
                        block_reduce_vcount<NPC, V, NRC>(vc, sm, nsat, (cur_tile / a.xtiles) * wx * V, W);
#pragma unroll
                        for (int v = 0; v < V; ++v) vc[v].clear();
                        pending = 0;
                    }
                }
                cur_tile = tile;
                const uint32_t xt = tile % a.xtiles, wt = tile / a.xtiles;
                w = (wt * wx + threadIdx.x) * V;
                const uint32_t xh = xt * bxh + xh_l;
                col_ok = w < W && xh < Lxh;
                const uint32_t wv = w / V;
                toff = xh * (W / V) + wv;
                offP = (xh + 1 == Lxh ? 0u : xh + 1) * (W / V) + wv;
                offM = (xh == 0 ? Lxh - 1 : xh - 1) * (W / V) + wv;
                xh2 = 2u * xh;
                xsite = 2u * xh;
            }
            if (col_ok && db.z != 0xFFFFFFFFu) {
                const uint32_t p = db.z & 1u;
                const uint32_t xsoff = p ? offP : offM;
                const uint32_t e_own = da.x + toff;
                uint32_t s[V], n[2 * DIM][V];
                {
                    const VecT t0 = ownv[e_own];
                    const VecT t1 = othv[e_own];
                    const VecT t2 = othv[da.x + xsoff];
                    const VecT t3 = othv[da.y + toff];
                    const VecT t4 = othv[da.z + toff];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        s[v] = reinterpret_cast<const uint32_t*>(&t0)[v];
                        n[0][v] = reinterpret_cast<const uint32_t*>(&t1)[v];
                        n[1][v] = reinterpret_cast<const uint32_t*>(&t2)[v];
                        n[2][v] = reinterpret_cast<const uint32_t*>(&t3)[v];
                        n[3][v] = reinterpret_cast<const uint32_t*>(&t4)[v];
                    }
                    if (DIM == 3) {
                        const VecT t5 = othv[da.w + toff];
                        const VecT t6 = othv[db.x + toff];
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            n[4][v] = reinterpret_cast<const uint32_t*>(&t5)[v];
                            n[5][v] = reinterpret_cast<const uint32_t*>(&t6)[v];
                        }
                    }
                }
                uint32_t m[2 * DIM];
                if (PMJ) {
                    const uint4* jp = a.jm8 + (db.w + xh2);
                    const uint4 j0 = __ldg(jp);
                    m[0] = j0.x; m[1] = j0.y; m[2] = j0.z; m[3] = j0.w;
                    if (DIM == 3) {
                        const uint2 j1 = __ldg(reinterpret_cast<const uint2*>(jp + 1));
                        m[4] = j1.x; m[5] = j1.y;
                    }
                } else {
#pragma unroll
                    for (int k2 = 0; k2 < 2 * DIM; ++k2) m[k2] = a.antiferro;
                }
                if constexpr (COUNT_ONLY) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        uint32_t av[2 * DIM];
#pragma unroll
                        for (int k2 = 0; k2 < 2 * DIM; ++k2) av[k2] = ~(s[v] ^ n[k2][v] ^ m[k2]);
                        uint32_t b0, b1, b2;
                        count_sat<DIM>(av, b0, b1, b2);
                        vc[v].add3(b0, b1, b2);
                    }
                } else {
                    const uint32_t site = db.y + xsite;
                    update_site<DIM, K, ROUNDS, V, ACC, NPC>(s, n, m, site, a.gw0 + w, a.sweep, a.pk, a.mx, vc);
                    VecT o;
#pragma unroll
                    for (int v = 0; v < V; ++v) reinterpret_cast<uint32_t*>(&o)[v] = s[v];
                    ownv[e_own] = o;
                }
            }
            if constexpr (ACC) {
                if (++pending == MAX_ITEMS) {  // counters full: reduce and start over
                    block_reduce_vcount<NPC, V, NRC>(vc, sm, nsat, (cur_tile / a.xtiles) * wx * V, W);
#pragma unroll
                    for (int v = 0; v < V; ++v) vc[v].clear();
                    pending = 0;
                }
            }
        }
    }
    if constexpr (ACC) {
        if (pending) block_reduce_vcount<NPC, V, NRC>(vc, sm, nsat, (cur_tile / a.xtiles) * wx * V, W);
    }
}

// SMALL: the same kernel compiled for blocks of 128 threads, 7 (plain) / 4 (accumulating) of them per
// SM.  When a colour phase has only one or two site groups per resident thread (the 8-GPU split of
// config 3: 128 replicas per GPU), 1024 units of 128 threads fit the 148 x 7 slots in ONE wave where
// 512 units of 256 threads need a second, mostly empty one (profiles/r02_small_w_ab.log: 20.9 ->
// 18.6 us per sweep with energies at 128 replicas; at 512 and more the 256-thread shape is faster).
constexpr int ROWS_SMALL_THREADS = 128;
constexpr int ROWS_SMALL_NP = 5, ROWS_SMALL_NR = 12;   // counter planes of the small shape (sweep_rows_phase)
constexpr int ROWS_NR = 15;                             // 256 threads x (2^7 - 1) < 2^15
static_assert(ISING_ROWS_THREADS * ((1 << SW_NP) - 1) < (1 << ROWS_NR), "block counters too narrow");
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool MULTIROW, bool SMALL = false>
__global__ void __launch_bounds__(SMALL ? ROWS_SMALL_THREADS : ISING_ROWS_THREADS,
                                  SMALL ? (ACC ? 4 : 7) : (ACC ? ISING_ROWS_ACC_MINB : ISING_ROWS_MINB))
k_sweep_rows(const __grid_constant__ RowsArgs a) {
    extern __shared__ uint32_t sm[];
    sweep_rows_phase<DIM, PMJ, K, ROUNDS, V, ACC, MULTIROW, false, SMALL ? ROWS_SMALL_NP : SW_NP,
                     SMALL ? ROWS_SMALL_NR : ROWS_NR>(a, sm);
}

template <int DIM, bool PMJ, int V, bool MULTIROW>
__global__ void __launch_bounds__(ISING_ROWS_THREADS, ISING_ROWS_MINB) k_nsat_rows(const __grid_constant__ RowsArgs a) {
    extern __shared__ uint32_t sm[];
    sweep_rows_phase<DIM, PMJ, 6, 7, V, true, MULTIROW, true>(a, sm);
}

// ------------------------------------------------------------------------------------------
// The same colour phase with the neighbour rows staged in shared memory by the TMA unit
// (cp.async.bulk global -> shared, completion on an mbarrier): whole-row units only (one tile
// covers the row: Lxh * W / V <= threads per block / rows per unit), W % 4 == 0.
//
// A unit = R consecutive rows y0 .. y0+R-1 of plane z.  One stage of the ring holds
//   OWN  R rows of the colour being updated
//   OTH  R + 2 rows of the other colour: y0-1 (periodic), y0 .. y0+R-1, y0+R (periodic)
//   ZM, ZP  R rows of planes z-1, z+1 (3D)
//   JM   R rows of bond masks (+-J)
// filled by seven bulk copies of >= 512 B issued by thread 0, two units ahead of their use, so a
// warp never waits for global memory and needs no registers for in-flight loads nor 64-bit
// addresses: a thread's shared-memory offsets are constants of the launch.  Consumers release a
// stage by one mbarrier arrive per warp; only thread 0 waits for that before it refills the stage.
// ------------------------------------------------------------------------------------------
struct RowsTmaArgs {
    RowsArgs r;
    uint32_t row_bytes;     // Lxh * W * 4
    uint32_t jrow_bytes;    // Lxh * 32 (0 without bond masks)
    uint32_t stage_bytes;   // one stage of the ring (multiple of 128)
    uint32_t off_oth, off_zm, off_zp, off_jm;   // byte offsets inside a stage (OWN at 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1D bulk copy global -> shared::cta of `bytes` (multiple of 16), completion counted on `bar`
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct RowsStageMeta {   // written by the producer thread before it arms the stage's barrier
    uint32_t row0;       // z * Ly + y0
    uint32_t par0;       // (y0 + z + colour) & 1
    uint32_t pad[2];
};

template <int DIM, bool PMJ, int K, int ROUNDS, bool ACC>
__device__ __forceinline__ void sweep_rows_tma_phase(const RowsTmaArgs& ta, unsigned char* smem) {
    constexpr int V = 4;
    constexpr int NSTAGE = 2;
    const RowsArgs& a = ta.r;
    __shared__ __align__(8) unsigned long long s_full[NSTAGE], s_empty[NSTAGE];
    __shared__ RowsStageMeta s_meta[NSTAGE];
    const uint32_t Lxh = a.Lxh, W = a.W, Ly = a.Ly, Lz = a.Lz;
    const uint32_t wx = blockDim.x;
    const uint32_t nthreads = blockDim.x * blockDim.y;
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint32_t bxh = 1u << a.bxh_log;
    const uint32_t xh = threadIdx.y & (bxh - 1u);
    const uint32_t rsub = threadIdx.y >> a.bxh_log;
    const uint32_t R = 1u << a.nrs_log;
    const uint32_t b = blockIdx.x;
    const uint32_t u0 = b * a.uq + (b < a.urem ? b : a.urem);
    const uint32_t nsteps = a.uq + (b < a.urem ? 1u : 0u);
    const uint32_t w = threadIdx.x * V;
    const bool col_ok = w < W && xh < Lxh;
    // shared-memory byte offsets of this thread's words inside a stage
    const uint32_t o_site = ((rsub * Lxh + xh) * W + w) * 4u;
    const uint32_t o_xp = (((rsub + 1u) * Lxh + (xh + 1 == Lxh ? 0u : xh + 1)) * W + w) * 4u;
    const uint32_t o_xm = (((rsub + 1u) * Lxh + (xh == 0 ? Lxh - 1 : xh - 1)) * W + w) * 4u;
    const uint32_t o_jm = (rsub * Lxh + xh) * 32u;
    const uint32_t rowlenV = Lxh * (W / V);
    const uint32_t toffV = xh * (W / V) + threadIdx.x;
    uint4* __restrict__ ownv = reinterpret_cast<uint4*>(a.own);
    const uint32_t smem_base = smem_u32(smem);

    // producer (thread 0): fill stage `st` with the rows of unit u0 + k
    auto produce = [&](uint32_t k, uint32_t st) {
        const uint32_t u = u0 + k;
        const uint32_t yg = u % a.ygroups, z = u / a.ygroups;
        const uint32_t y0 = yg << a.nrs_log;
        const uint32_t ym = y0 == 0 ? Ly - 1 : y0 - 1, yp = y0 + R == Ly ? 0u : y0 + R;
        const uint32_t zm = z == 0 ? Lz - 1 : z - 1, zp = z + 1 == Lz ? 0u : z + 1;
        const uint32_t row0 = z * Ly + y0;
        s_meta[st].row0 = row0;
        s_meta[st].par0 = (y0 + z + a.c) & 1u;
        const uint32_t bar = smem_u32(&s_full[st]);
        const uint32_t dst = smem_base + st * ta.stage_bytes;
        const uint32_t rb = ta.row_bytes;
        const unsigned char* own = reinterpret_cast<const unsigned char*>(a.own);
        const unsigned char* oth = reinterpret_cast<const unsigned char*>(a.oth);
        uint32_t total = (2u * R + 2u) * rb;
        if (DIM == 3) total += 2u * R * rb;
        if (PMJ) total += R * ta.jrow_bytes;
        mbar_expect_tx(bar, total);
        tma_load_1d(dst, own + (size_t)row0 * rb, R * rb, bar);
        tma_load_1d(dst + ta.off_oth, oth + (size_t)(z * Ly + ym) * rb, rb, bar);
        tma_load_1d(dst + ta.off_oth + rb, oth + (size_t)row0 * rb, R * rb, bar);
        tma_load_1d(dst + ta.off_oth + (R + 1u) * rb, oth + (size_t)(z * Ly + yp) * rb, rb, bar);
        if (DIM == 3) {
            tma_load_1d(dst + ta.off_zm, oth + (size_t)(zm * Ly + y0) * rb, R * rb, bar);
            tma_load_1d(dst + ta.off_zp, oth + (size_t)(zp * Ly + y0) * rb, R * rb, bar);
        }
        if (PMJ)
            tma_load_1d(dst + ta.off_jm, reinterpret_cast<const unsigned char*>(a.jm8) + (size_t)row0 * ta.jrow_bytes,
                        R * ta.jrow_bytes, bar);
    };

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < NSTAGE; ++st) {
            mbar_init(smem_u32(&s_full[st]), 1u);
            mbar_init(smem_u32(&s_empty[st]), nthreads / 32u);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        for (uint32_t k = 0; k < NSTAGE && k < nsteps; ++k) produce(k, k);
    }

    VCount<ACC ? SW_NP : 1> vc[V];
    if constexpr (ACC) {
#pragma unroll
        for (int v = 0; v < V; ++v) vc[v].clear();
    }
    int pending = 0;
    unsigned long long* const nsat = ACC ? a.nsat + (size_t)(blockIdx.x % a.nsat_copies) * a.nsat_stride : nullptr;
    uint32_t* red = reinterpret_cast<uint32_t*>(smem + NSTAGE * ta.stage_bytes);  // ACC: block reduction area

    for (uint32_t k = 0; k < nsteps; ++k) {
        const uint32_t st = k & (NSTAGE - 1), ph = (k / NSTAGE) & 1u;
        mbar_wait(smem_u32(&s_full[st]), ph);
        const unsigned char* sb = smem + st * ta.stage_bytes;
        if (col_ok) {
            const uint32_t row = s_meta[st].row0 + rsub;
            const uint32_t p = (s_meta[st].par0 ^ rsub) & 1u;
            uint32_t s[V], n[2 * DIM][V];
            auto ld = [&](uint32_t off, uint32_t (&out)[V]) {
                const uint4 t = *reinterpret_cast<const uint4*>(sb + off);
                out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
            };
            ld(o_site, s);
            ld(ta.off_oth + ta.row_bytes + o_site, n[0]);
            ld(ta.off_oth + (p ? o_xp : o_xm), n[1]);
            ld(ta.off_oth + o_site, n[2]);
            ld(ta.off_oth + 2u * ta.row_bytes + o_site, n[3]);
            if (DIM == 3) {
                ld(ta.off_zm + o_site, n[4]);
                ld(ta.off_zp + o_site, n[5]);
            }
            uint32_t m[2 * DIM];
            if (PMJ) {
                const uint4 j0 = *reinterpret_cast<const uint4*>(sb + ta.off_jm + o_jm);
                m[0] = j0.x; m[1] = j0.y; m[2] = j0.z; m[3] = j0.w;
                if (DIM == 3) {
                    const uint2 j1 = *reinterpret_cast<const uint2*>(sb + ta.off_jm + o_jm + 16u);
                    m[4] = j1.x; m[5] = j1.y;
                }
            } else {
#pragma unroll
                for (int k2 = 0; k2 < 2 * DIM; ++k2) m[k2] = a.antiferro;
            }
            const uint32_t site = row * a.Lx + 2u * xh + p;
            update_site<DIM, K, ROUNDS, V, ACC>(s, n, m, site, a.gw0 + w, a.sweep, a.pk, a.mx, vc);
            ownv[row * rowlenV + toffV] = make_uint4(s[0], s[1], s[2], s[3]);
        }
        // this warp is done with the stage
        __syncwarp();
        if ((tid & 31u) == 0) mbar_arrive(smem_u32(&s_empty[st]));
        if (tid == 0 && k + NSTAGE < nsteps) {
            mbar_wait(smem_u32(&s_empty[st]), ph);   // every warp has read step k's rows
            produce(k + NSTAGE, st);
        }
        if constexpr (ACC) {
            if (++pending == SW_MAX_ITEMS) {
                block_reduce_vcount<SW_NP, V>(vc, red, nsat, 0u, W);
#pragma unroll
                for (int v = 0; v < V; ++v) vc[v].clear();
                pending = 0;
            }
        }
    }
    if constexpr (ACC) {
        if (pending) block_reduce_vcount<SW_NP, V>(vc, red, nsat, 0u, W);
    }
}

template <int DIM, bool PMJ, int K, int ROUNDS, bool ACC>
__global__ void __launch_bounds__(ISING_ROWS_THREADS, ACC ? ISING_ROWS_ACC_MINB : ISING_ROWS_TMA_MINB)
k_sweep_rows_tma(const __grid_constant__ RowsTmaArgs ta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    sweep_rows_tma_phase<DIM, PMJ, K, ROUNDS, ACC>(ta, smem_raw);
}

}  // namespace ising

// sm_100a kernels of the classical Ising engine.  See DESIGN.md for the algorithm; every
// kernel here is integer / bitwise work on replica-bit-packed words (no tensor cores: the path
// is not a contraction).
#include "kernels.h"

#include <cooperative_groups.h>

#include <stdlib.h>

#include "../../include/ising_b200.h"
#include "philox.h"

namespace cg = cooperative_groups;

namespace ising {

// ------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t site_word_base(const Layout& L, uint64_t n) {
    if (L.kind == ISING_KIND_GENERAL) return (size_t)n * L.W;
    const uint32_t x = (uint32_t)(n % L.Lx);
    const uint32_t r = (uint32_t)(n / L.Lx);  // row = z * Ly + y
    const uint32_t y = r % L.Ly, z = r / L.Ly;
    const uint32_t c = (x + y + z) & 1u;
    return (((size_t)c * L.rows + r) * L.Lxh + (x >> 1)) * L.W;
}

// ------------------------------------------------------------------------------------------
// bit-sliced satisfied-bond count of one word: planes (b0, b1, b2) of n_sat in 0..2*DIM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    return (a & b) | (c & (a ^ b));
}

template <int DIM>
__device__ __forceinline__ void count_sat(const uint32_t (&a)[2 * DIM], uint32_t& b0,
                                          uint32_t& b1, uint32_t& b2) {
    if (DIM == 3) {
        const uint32_t s0 = a[0] ^ a[1] ^ a[2], c0 = maj3(a[0], a[1], a[2]);
        const uint32_t s1 = a[3] ^ a[4] ^ a[5], c1 = maj3(a[3], a[4], a[5]);
        const uint32_t k = s0 & s1;
        b0 = s0 ^ s1;
        b1 = c0 ^ c1 ^ k;
        b2 = maj3(c0, c1, k);
    } else {
        const uint32_t s0 = a[0] ^ a[1] ^ a[2], c0 = maj3(a[0], a[1], a[2]);
        const uint32_t k = s0 & a[3];
        b0 = s0 ^ a[3];
        b1 = c0 ^ k;
        b2 = c0 & k;
    }
}

// ------------------------------------------------------------------------------------------
// vertical (bit-sliced) counters: plane l holds bit l of 32 independent per-replica counters
// ------------------------------------------------------------------------------------------
template <int NP>
struct VCount {
    uint32_t v[NP];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int l = 0; l < NP; ++l) v[l] = 0;
    }
    // counters += b0 + 2 b1 + 4 b2
    __device__ __forceinline__ void add3(uint32_t b0, uint32_t b1, uint32_t b2) {
        uint32_t c = v[0] & b0;
        v[0] ^= b0;
        uint32_t t = v[1] ^ b1, c2 = (v[1] & b1) | (c & t);
        v[1] = t ^ c;
        c = c2;
        t = v[2] ^ b2;
        c2 = (v[2] & b2) | (c & t);
        v[2] = t ^ c;
        c = c2;
#pragma unroll
        for (int l = 3; l < NP; ++l) {
            t = v[l] & c;
            v[l] ^= c;
            c = t;
        }
    }
    __device__ __forceinline__ void add1(uint32_t b0) {
        uint32_t c = b0;
#pragma unroll
        for (int l = 0; l < NP; ++l) {
            const uint32_t t = v[l] & c;
            v[l] ^= c;
            c = t;
        }
    }
    // sm is int[32][nthreads]; adds this thread's 32 counters to its column
    __device__ __forceinline__ void flush(int* sm, int tid, int nthreads) {
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
            int cnt = 0;
#pragma unroll
            for (int l = 0; l < NP; ++l) cnt |= (int)((v[l] >> b) & 1u) << l;
            sm[b * nthreads + tid] += cnt;
        }
        clear();
    }
};

// --- bit-sliced helpers for the cross-thread reduction of vertical counters --------------------
// acc (NR planes) += x (NX planes), both little-endian bit-sliced integers
template <int NR, int NX>
__device__ __forceinline__ void vadd(uint32_t (&acc)[NR], const uint32_t (&x)[NX]) {
    uint32_t c = 0;
#pragma unroll
    for (int l = 0; l < NR; ++l) {
        const uint32_t xi = l < NX ? x[l] : 0u;
        const uint32_t t = acc[l] ^ xi;
        const uint32_t c2 = (acc[l] & xi) | (c & t);
        acc[l] = t ^ c;
        c = c2;
    }
}

constexpr int NS_NR = 20;  // block-level counter planes: 256 threads x 1023 fits 18 bits

// Block-wide reduction of per-thread vertical counters (NP planes, V replica words per thread,
// block = (wx, by)) into per-experiment integers: bit-sliced tree through shared memory, then
// one SWAR bit-transpose per word column and 32 integer atomics per column.
//   sm: max(NP * V, NS_NR) * nthreads words;  out[(w0 + column) * 32 + bit] += count
template <int NP, int V>
__device__ __forceinline__ void block_reduce_vcount(const VCount<NP> (&vc)[V], uint32_t* sm,
                                                    unsigned long long* __restrict__ out,
                                                    uint32_t w0, uint32_t W) {
    const uint32_t wx = blockDim.x, by = blockDim.y, nthreads = wx * by;
    const uint32_t tid = threadIdx.y * wx + threadIdx.x;
    const uint32_t C = wx * V;  // word columns of this chunk (C divides nthreads)
    __syncthreads();
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int l = 0; l < NP; ++l) sm[(l * V + v) * nthreads + tid] = vc[v].v[l];
    __syncthreads();
    // stage A: thread (column c, part q) adds the counters of every Q-th row-thread
    const uint32_t Q = nthreads / C;
    const uint32_t c = tid % C, q = tid / C;
    const uint32_t cx = c / V, cv = c % V;
    uint32_t acc[NS_NR];
#pragma unroll
    for (int l = 0; l < NS_NR; ++l) acc[l] = 0;
    for (uint32_t ty = q; ty < by; ty += Q) {
        uint32_t x[NP];
#pragma unroll
        for (int l = 0; l < NP; ++l) x[l] = sm[(l * V + cv) * nthreads + ty * wx + cx];
        vadd<NS_NR, NP>(acc, x);
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NS_NR; ++l) sm[l * nthreads + tid] = acc[l];  // [plane][q][c]
    __syncthreads();
    // stage B1: tree over the Q parts of every column (all threads of the surviving parts work)
    for (uint32_t half = Q >> 1; half >= 1; half >>= 1) {
        if (q < half) {
            uint32_t x[NS_NR];
#pragma unroll
            for (int l = 0; l < NS_NR; ++l) x[l] = sm[l * nthreads + (q + half) * C + c];
            vadd<NS_NR, NS_NR>(acc, x);
#pragma unroll
            for (int l = 0; l < NS_NR; ++l) sm[l * nthreads + tid] = acc[l];
        }
        __syncthreads();
    }
    // stage B2: SWAR bit-transpose of the column totals (part 0), byte-lane group g per thread:
    // bits g, g+8, g+16, g+24 of the planes land in four byte lanes; 4 integer atomics each
    if (w0 + c < W) {
        if (q != 0) {
#pragma unroll
            for (int l = 0; l < NS_NR; ++l) acc[l] = sm[l * nthreads + c];
        }
        unsigned long long* o = out + (size_t)(w0 + c) * 32;
        for (uint32_t g = q; g < 8; g += Q) {
            uint32_t lo = 0, hi = 0, top = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                lo += ((acc[l] >> g) & 0x01010101u) << l;
                hi += ((acc[l + 8] >> g) & 0x01010101u) << l;
                if (l + 16 < NS_NR) top += ((acc[l + 16] >> g) & 0x01010101u) << l;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t cnt = ((lo >> (8 * k)) & 0xFFu) | (((hi >> (8 * k)) & 0xFFu) << 8) |
                                     (((top >> (8 * k)) & 0xFFu) << 16);
                if (cnt) atomicAdd(o + g + 8 * k, (unsigned long long)cnt);
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Metropolis acceptance of the uphill bits of one word.
//   up   : bits with dE > 0;  (sel1, sel0) select the class of each such bit
//   K bit-planes R_0..R_{K-1} are the K most significant bits of a uniform U per replica and
//   are compared against the class threshold T: U_top < T_top is the borrow of the bit-sliced
//   subtraction U_top - T_top (one majority LOP3 per plane, LSB first); bits whose top K bits
//   tie (probability 2^-K) are resolved by a fresh 32-bit word against the low 32 threshold
//   bits, tied bits taken in ascending position.  Word R_m is output m%4 of Philox call m/4
//   on counter (site, replica word, sweep, call).
// returns the flip mask (downhill bits always flip)
// ------------------------------------------------------------------------------------------
// PERBETA: every replica bit has its own inverse temperature (parallel tempering): the plane
// masks are words tp[cls * 8 + p] whose bit b is the threshold bit of replica b of this word,
// the resolver thresholds tl[b * 3 + cls].
template <int NCLS, int K, int ROUNDS, bool PERBETA = false>
__device__ __forceinline__ uint32_t msc_flip_mask(uint32_t up, uint32_t sel0, uint32_t sel1,
                                                  const MscThresholds& th, uint32_t site,
                                                  uint32_t gw, uint32_t sweep, const PhiloxKeys& pk,
                                                  const uint32_t* __restrict__ tp = nullptr,
                                                  const uint32_t* __restrict__ tl = nullptr) {
    constexpr int NCALL = K / 4 + 1;
    uint32_t r[NCALL * 4];
#pragma unroll
    for (int q = 0; q < NCALL; ++q) {
        const u32x4 o = philox4x32_keys<ROUNDS>(site, gw, sweep, (uint32_t)q | (TAG_ACCEPT << 24), pk);
        r[4 * q + 0] = o.x;
        r[4 * q + 1] = o.y;
        r[4 * q + 2] = o.z;
        r[4 * q + 3] = o.w;
    }
    uint32_t eq = up, borrow = 0;
#pragma unroll
    for (int p = K - 1; p >= 0; --p) {
        const uint32_t P0 = PERBETA ? __ldg(tp + 0 * 8 + p) : th.plane[0][p];
        const uint32_t P1 = PERBETA ? __ldg(tp + 1 * 8 + p) : th.plane[1][p];
        uint32_t t = (sel0 & P1) | (~sel0 & P0);
        if (NCLS == 3) {
            const uint32_t P2 = PERBETA ? __ldg(tp + 2 * 8 + p) : th.plane[2][p];
            t = (sel1 & P2) | (~sel1 & t);
        }
        borrow = maj3(~r[p], t, borrow);
        eq &= ~(r[p] ^ t);
    }
    uint32_t flip = ~up | (borrow & ~eq);
    // Tied bits (2^-K each).  The first SPARE of them use the words left over from the calls
    // above in straight-line predicated code (no divergent loop for the common case); anything
    // beyond that, rare, draws further Philox calls in a loop.
    constexpr int SPARE = (4 * NCALL - K) < 2 ? (4 * NCALL - K) : 2;
#pragma unroll
    for (int j = 0; j < SPARE; ++j) {
        const uint32_t bit = eq & (0u - eq);  // lowest tied bit, 0 when nothing is tied
        uint32_t lo;
        if (PERBETA) {
            const int b = (__ffs((int)eq) - 1) & 31;
            const uint32_t cls = (NCLS == 3 && (sel1 & bit)) ? 2u : ((sel0 & bit) ? 1u : 0u);
            lo = __ldg(tl + b * 3 + cls);
        } else {
            lo = (sel0 & bit) ? th.low[1] : th.low[0];
            if (NCLS == 3 && (sel1 & bit)) lo = th.low[2];
        }
        if (r[K + j] < lo) flip |= bit;
        eq ^= bit;
    }
    if (eq) {
        int j = K + SPARE;
        u32x4 cur = {r[4 * (NCALL - 1)], r[4 * (NCALL - 1) + 1], r[4 * (NCALL - 1) + 2],
                     r[4 * (NCALL - 1) + 3]};
        do {
            const int b = __ffs((int)eq) - 1;
            if ((j & 3) == 0 && j >= 4 * NCALL)
                cur = philox4x32_keys<ROUNDS>(site, gw, sweep, (uint32_t)(j >> 2) | (TAG_ACCEPT << 24), pk);
            const int m = j & 3;
            const uint32_t v = m == 0 ? cur.x : (m == 1 ? cur.y : (m == 2 ? cur.z : cur.w));
            uint32_t lo;
            if (PERBETA) {
                const uint32_t cls = (NCLS == 3 && ((sel1 >> b) & 1u)) ? 2u : (((sel0 >> b) & 1u) ? 1u : 0u);
                lo = __ldg(tl + b * 3 + cls);
            } else {
                lo = ((sel0 >> b) & 1u) ? th.low[1] : th.low[0];
                if (NCLS == 3 && ((sel1 >> b) & 1u)) lo = th.low[2];
            }
            if (v < lo) flip |= 1u << b;
            eq &= eq - 1;
            ++j;
        } while (eq);
    }
    return flip;
}

// V consecutive replica words as one vector load / store (128-bit when V == 4)
template <int V> struct WordVec;
template <> struct WordVec<1> { typedef uint32_t type; };
template <> struct WordVec<2> { typedef uint2 type; };
template <> struct WordVec<4> { typedef uint4 type; };

template <int V>
__device__ __forceinline__ void load_words(const uint32_t* p, uint32_t (&out)[V]) {
    typedef typename WordVec<V>::type T;
    const T v = *reinterpret_cast<const T*>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = w[k];
}

template <int V>
__device__ __forceinline__ void store_words(uint32_t* p, const uint32_t (&in)[V]) {
    typedef typename WordVec<V>::type T;
    T v;
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < V; ++k) w[k] = in[k];
    *reinterpret_cast<T*>(p) = v;
}

// ------------------------------------------------------------------------------------------
// K2: one colour phase of a checkerboard sweep on a square / cubic torus
// block = (WX lanes over groups of V replica words, BY over half-row positions); one row of
// the colour-compacted lattice per iteration; all seven spin loads are coalesced vector loads
// ------------------------------------------------------------------------------------------
#ifndef ISING_SWEEP_MIN_BLOCKS
#define ISING_SWEEP_MIN_BLOCKS 3
#endif
#ifndef ISING_SWEEP_UNROLL_V
#define ISING_SWEEP_UNROLL_V 4
#endif
#ifndef ISING_SWEEP_MAXV
#define ISING_SWEEP_MAXV 4
#endif
#ifndef ISING_SW_NP
#define ISING_SW_NP 7
#endif
#ifndef ISING_ACC_MIN_BLOCKS
#define ISING_ACC_MIN_BLOCKS 2
#endif
constexpr int SW_NP = ISING_SW_NP;                       // fused n_sat counter planes per thread
constexpr int SW_MAX_ITEMS = ((1 << SW_NP) - 1) / 6;    // sites a thread may accumulate (n_sat <= 6)

// ACC: this phase also accumulates the post-flip satisfied-bond count of every replica into
// nsat[] (used for the second colour: its sites see every bond once, so after the phase
// nsat[e] is the total of experiment e and E = |J| (n_bonds - 2 nsat), lattice.rs:454).
// GRID2D: one row per block, (y, z) = 2D block index; otherwise blocks walk the rows with stride
// row_step (persistent launch).
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool GRID2D, bool PERBETA = false>
__device__ __forceinline__ void sweep_colour_phase(
    uint32_t* __restrict__ own, const uint32_t* __restrict__ oth, const uint32_t* __restrict__ jm,
    const Layout& L, uint32_t c, uint32_t sweep, const PhiloxKeys& pk, uint32_t gw0,
    uint32_t antiferro, const MscThresholds& th, unsigned long long* __restrict__ nsat,
    uint32_t row_step, uint32_t step_y, uint32_t step_z, uint32_t* sm,
    const uint32_t* __restrict__ tplane = nullptr, const uint32_t* __restrict__ tlow = nullptr) {
    constexpr int kUnrollV = ISING_SWEEP_UNROLL_V;
    const uint32_t Lxh = L.Lxh, W = L.W, Ly = L.Ly, Lz = L.Lz;
    const uint32_t rowlen = Lxh * W;  // words per colour row (< 2^32: checked on the host)
    for (uint32_t w0 = 0; w0 < W; w0 += V * blockDim.x) {
        const uint32_t w = w0 + V * threadIdx.x;
        VCount<ACC ? SW_NP : 1> vc[V];
        if constexpr (ACC) {
#pragma unroll
            for (int v = 0; v < V; ++v) vc[v].clear();
        }
        int pending = 0;  // block-uniform count of accumulated sites per thread
        // Row walk without per-row integer division: the one-row-per-block launch reads (y, z)
        // from its 2D block index; the persistent (ACC) launch divides once and then steps by
        // the grid size with a carry.
        uint32_t y, z, row;
        if constexpr (GRID2D) {
            y = blockIdx.x;
            z = blockIdx.y;
            row = z * Ly + y;
        } else {
            row = blockIdx.x;
            z = row / Ly;
            y = row - z * Ly;
        }
        for (; row < L.rows; row += row_step, y += step_y, z += step_z) {
            if (y >= Ly) {
                y -= Ly;
                ++z;
            }
            const uint32_t p = (y + z + c) & 1u;
            const uint32_t ym = y == 0 ? Ly - 1 : y - 1, yp = y + 1 == Ly ? 0 : y + 1;
            uint32_t* __restrict__ o_c = own + (size_t)row * rowlen;
            const uint32_t* __restrict__ n_x = oth + (size_t)row * rowlen;
            const uint32_t* __restrict__ n_ym = oth + (size_t)(z * Ly + ym) * rowlen;
            const uint32_t* __restrict__ n_yp = oth + (size_t)(z * Ly + yp) * rowlen;
            const uint32_t* __restrict__ n_zm = nullptr;
            const uint32_t* __restrict__ n_zp = nullptr;
            if (DIM == 3) {
                const uint32_t zm = z == 0 ? Lz - 1 : z - 1, zp = z + 1 == Lz ? 0 : z + 1;
                n_zm = oth + (size_t)(zm * Ly + y) * rowlen;
                n_zp = oth + (size_t)(zp * Ly + y) * rowlen;
            }
            for (uint32_t xh0 = 0; xh0 < Lxh; xh0 += blockDim.y) {
                const uint32_t xh = xh0 + threadIdx.y;
                if (xh < Lxh && w < W) {
                const uint32_t xs = p ? (xh + 1 == Lxh ? 0 : xh + 1) : (xh == 0 ? Lxh - 1 : xh - 1);
                uint32_t m[2 * DIM];
#pragma unroll
                for (int k = 0; k < 2 * DIM; ++k)
                    m[k] = PMJ ? __ldg(jm + (size_t)k * L.halfN + (size_t)row * Lxh + xh) : antiferro;
                const uint32_t site = row * L.Lx + 2 * xh + p;
                const uint32_t i = xh * W + w;
                uint32_t s[V], n[2 * DIM][V];
                load_words<V>(o_c + i, s);
                load_words<V>(n_x + i, n[0]);
                load_words<V>(n_x + xs * W + w, n[1]);
                load_words<V>(n_ym + i, n[2]);
                load_words<V>(n_yp + i, n[3]);
                if (DIM == 3) {
                    load_words<V>(n_zm + i, n[4]);
                    load_words<V>(n_zp + i, n[5]);
                }
#pragma unroll(kUnrollV)
                for (int v = 0; v < V; ++v) {
                    uint32_t a[2 * DIM];
#pragma unroll
                    for (int k = 0; k < 2 * DIM; ++k) a[k] = ~(s[v] ^ n[k][v] ^ m[k]);
                    uint32_t b0, b1, b2;
                    count_sat<DIM>(a, b0, b1, b2);
                    uint32_t flip;
                    if (DIM == 3)  // n_sat 4,5,6 -> dE = 4,8,12 |J|
                        flip = msc_flip_mask<3, K, ROUNDS, PERBETA>(
                            b2, b0, b1, th, site, gw0 + w + v, sweep, pk,
                            PERBETA ? tplane + (size_t)(w + v) * 24 : nullptr,
                            PERBETA ? tlow + (size_t)(w + v) * 96 : nullptr);
                    else  // n_sat 3,4 -> dE = 4,8 |J|
                        flip = msc_flip_mask<2, K, ROUNDS, PERBETA>(
                            b2 | (b1 & b0), b2, 0u, th, site, gw0 + w + v, sweep, pk,
                            PERBETA ? tplane + (size_t)(w + v) * 24 : nullptr,
                            PERBETA ? tlow + (size_t)(w + v) * 96 : nullptr);
                    s[v] ^= flip;
                    if constexpr (ACC) {
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
                store_words<V>(o_c + i, s);
                }
                if constexpr (ACC) {
                    if (++pending == SW_MAX_ITEMS) {  // counters full: reduce and start over
                        block_reduce_vcount<SW_NP, V>(vc, sm, nsat, w0, W);
#pragma unroll
                        for (int v = 0; v < V; ++v) vc[v].clear();
                        pending = 0;
                    }
                }
            }
        }
        if constexpr (ACC) block_reduce_vcount<SW_NP, V>(vc, sm, nsat, w0, W);
    }
}

template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC>
__global__ void __launch_bounds__(256, ACC ? ISING_ACC_MIN_BLOCKS : ISING_SWEEP_MIN_BLOCKS)
k_sweep_stencil(uint32_t* __restrict__ own, const uint32_t* __restrict__ oth,
                const uint32_t* __restrict__ jm, Layout L, uint32_t c, uint32_t sweep,
                PhiloxKeys pk, uint32_t gw0, uint32_t antiferro, MscThresholds th,
                unsigned long long* __restrict__ nsat, uint32_t row_step, uint32_t step_y,
                uint32_t step_z) {
    extern __shared__ uint32_t sm[];
    sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, ACC, !ACC>(own, oth, jm, L, c, sweep, pk, gw0, antiferro,
                                                         th, nsat, row_step, step_y, step_z, sm);
}

// per-replica inverse temperatures (parallel tempering on lattices): thresholds from tables
//   tplane[(w * 3 + cls) * 8 + p], tlow[(w * 32 + b) * 3 + cls]
template <int DIM, bool PMJ, int ROUNDS, int V, bool ACC>
__global__ void __launch_bounds__(256, 2)
k_sweep_stencil_perbeta(uint32_t* __restrict__ own, const uint32_t* __restrict__ oth,
                        const uint32_t* __restrict__ jm, Layout L, uint32_t c, uint32_t sweep,
                        PhiloxKeys pk, uint32_t gw0, uint32_t antiferro,
                        const uint32_t* __restrict__ tplane, const uint32_t* __restrict__ tlow,
                        unsigned long long* __restrict__ nsat, uint32_t row_step, uint32_t step_y,
                        uint32_t step_z) {
    extern __shared__ uint32_t sm[];
    MscThresholds unused{};
    sweep_colour_phase<DIM, PMJ, 6, ROUNDS, V, ACC, !ACC, true>(own, oth, jm, L, c, sweep, pk, gw0, antiferro,
                                                               unused, nsat, row_step, step_y, step_z, sm,
                                                               tplane, tlow);
}

// Small lattices are launch-bound (a colour phase of config 1 is ~2 us of work): one cooperative
// launch runs a whole chunk of sweeps, both colours, with a grid barrier between phases.  The
// per-sweep thresholds come from a table in global memory, staged in shared memory.
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC>
__global__ void __launch_bounds__(256, 2)
k_sweep_stencil_coop(uint32_t* __restrict__ spins, const uint32_t* __restrict__ jmask, Layout L,
                     uint32_t sweep0, uint32_t nsweeps, PhiloxKeys pk, uint32_t gw0,
                     uint32_t antiferro, const MscThresholds* __restrict__ th_table,
                     unsigned long long* __restrict__ nsat_hist, uint32_t cw) {
    extern __shared__ uint32_t sm[];
    __shared__ MscThresholds th;
    cg::grid_group grid = cg::this_grid();
    const size_t csz = (size_t)L.halfN * L.W;
    const size_t jsz = (size_t)2 * DIM * L.halfN;
    const uint32_t g = gridDim.x, tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (uint32_t t = 0; t < nsweeps; ++t) {
        __syncthreads();
        if (tid < sizeof(MscThresholds) / 4)
            reinterpret_cast<uint32_t*>(&th)[tid] = reinterpret_cast<const uint32_t*>(th_table + t)[tid];
        __syncthreads();
        sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, false, false>(
            spins, spins + csz, PMJ ? jmask : nullptr, L, 0u, sweep0 + t, pk, gw0, antiferro, th,
            nullptr, g, g % L.Ly, g / L.Ly, sm);
        grid.sync();
        sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, ACC, false>(
            spins + csz, spins, PMJ ? jmask + jsz : nullptr, L, 1u, sweep0 + t, pk, gw0, antiferro, th,
            ACC ? nsat_hist + (size_t)t * cw : nullptr, g, g % L.Ly, g / L.Ly, sm);
        grid.sync();
    }
}

template <int DIM, bool PMJ, int K, int ROUNDS, int V>
static void sweep_launch_phase(const SweepArgs& a, cudaStream_t st, dim3 grid, dim3 block,
                               uint32_t c, bool acc) {
    const Layout& L = a.lay;
    const size_t csz = (size_t)L.halfN * L.W;
    const size_t jsz = (size_t)2 * DIM * L.halfN;
    uint32_t* own = a.spins + c * csz;
    const uint32_t* oth = a.spins + (1 - c) * csz;
    const uint32_t* jm = a.jmask ? a.jmask + c * jsz : nullptr;
    const PhiloxKeys pk = philox_round_keys(a.key0, a.key1);
    if (a.tplane) {  // per-replica betas (K == 6 checked by the caller)
        if constexpr (K == 6) {
            if (!acc) {
                const dim3 grid2(L.Ly, L.Lz > 65535u ? 65535u : L.Lz, 1);
                k_sweep_stencil_perbeta<DIM, PMJ, ROUNDS, V, false><<<grid2, block, 0, st>>>(
                    own, oth, jm, L, c, a.sweep, pk, a.gw0, a.antiferro, a.tplane, a.tlow, nullptr, L.rows,
                    0u, 0u);
            } else {
                if (block.y < (unsigned)V) block.y = V;
                uint32_t g = 148u * ISING_ACC_MIN_BLOCKS;
                if (g > L.rows) g = L.rows;
                const int nthreads = block.x * block.y;
                const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
                const size_t smem = (size_t)planes * nthreads * sizeof(uint32_t);
                k_sweep_stencil_perbeta<DIM, PMJ, ROUNDS, V, true><<<dim3(g, 1, 1), block, smem, st>>>(
                    own, oth, jm, L, c, a.sweep, pk, a.gw0, a.antiferro, a.tplane, a.tlow, a.nsat_out, g,
                    g % L.Ly, g / L.Ly);
            }
        }
        return;
    }
    if (!acc) {
        const dim3 grid2(L.Ly, L.Lz > 65535u ? 65535u : L.Lz, 1);
        k_sweep_stencil<DIM, PMJ, K, ROUNDS, V, false><<<grid2, block, 0, st>>>(
            own, oth, jm, L, c, a.sweep, pk, a.gw0, a.antiferro, a.th, nullptr, L.rows, 0u, 0u);
        return;
    }
    // fused accumulation: persistent blocks so that the per-block reduction is amortised, but
    // never more sites per thread than the SW_NP-plane counters can hold
    if (block.y < (unsigned)V) block.y = V;
    uint32_t g = 148u * ISING_ACC_MIN_BLOCKS;
    if (g > L.rows) g = L.rows;
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
    const size_t smem = (size_t)planes * nthreads * sizeof(uint32_t);
    k_sweep_stencil<DIM, PMJ, K, ROUNDS, V, true><<<dim3(g, 1, 1), block, smem, st>>>(
        own, oth, jm, L, c, a.sweep, pk, a.gw0, a.antiferro, a.th, a.nsat_out, g, g % L.Ly, g / L.Ly);
}

template <int DIM, bool PMJ, int K, int V>
static int sweep_dispatch_rounds(const SweepArgs& a, cudaStream_t st, dim3 grid, dim3 block) {
    for (uint32_t c = 0; c < 2; ++c) {
        const bool acc = a.nsat_out != nullptr && c == 1;
        if (a.rounds == 7) sweep_launch_phase<DIM, PMJ, K, 7, V>(a, st, grid, block, c, acc);
        else sweep_launch_phase<DIM, PMJ, K, 10, V>(a, st, grid, block, c, acc);
    }
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

template <int DIM, bool PMJ, int V>
static int sweep_dispatch_planes(const SweepArgs& a, cudaStream_t st, dim3 grid, dim3 block) {
    switch (a.planes) {
        case 5: return sweep_dispatch_rounds<DIM, PMJ, 5, V>(a, st, grid, block);
        case 6: return sweep_dispatch_rounds<DIM, PMJ, 6, V>(a, st, grid, block);
        case 7: return sweep_dispatch_rounds<DIM, PMJ, 7, V>(a, st, grid, block);
        default: return -1;
    }
}

static uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void stencil_block_shape(const Layout& L, uint32_t V, dim3* grid, dim3* block,
                                bool persistent) {
    const uint32_t groups = (L.W + V - 1) / V;  // vector groups of replica words per site
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    uint32_t threads = 256;
    if (const char* env = getenv("ISING_BLOCK_THREADS")) threads = (uint32_t)atoi(env);  // tuning knob
    if (threads < 32 || threads > 256 || (threads & (threads - 1))) threads = 256;
    uint32_t by = threads / wx;
    if (by < 1) by = 1;
    const uint32_t need = pow2_ceil(L.Lxh);
    if (by > need) by = need;
    if (wx * by < 32) by = 32 / wx;
    *block = dim3(wx, by, 1);
    uint32_t g = L.rows;
    if (persistent && g > 148u * 8u) g = 148u * 8u;
    *grid = dim3(g, 1, 1);
}

template <int V>
static int sweep_dispatch_kind(const SweepArgs& a, cudaStream_t st) {
    dim3 grid, block;
    stencil_block_shape(a.lay, V, &grid, &block, false);
    const bool pmj = a.jmask != nullptr;
    if (a.lay.kind == ISING_KIND_STENCIL3D)
        return pmj ? sweep_dispatch_planes<3, true, V>(a, st, grid, block)
                   : sweep_dispatch_planes<3, false, V>(a, st, grid, block);
    if (a.lay.kind == ISING_KIND_STENCIL2D)
        return pmj ? sweep_dispatch_planes<2, true, V>(a, st, grid, block)
                   : sweep_dispatch_planes<2, false, V>(a, st, grid, block);
    return -1;
}

int launch_sweep_stencil(const SweepArgs& a, cudaStream_t st) {
    if (a.tplane && a.planes != 6) return -1;  // per-replica tables are built for K = 6
    // widest vector the replica-word count allows (rows then stay 16-byte aligned)
    if (ISING_SWEEP_MAXV >= 4 && a.lay.W % 4 == 0) return sweep_dispatch_kind<4>(a, st);
    if (ISING_SWEEP_MAXV >= 2 && a.lay.W % 2 == 0) return sweep_dispatch_kind<2>(a, st);
    return sweep_dispatch_kind<1>(a, st);
}

// ---- cooperative multi-sweep launch (small lattices) -------------------------------------------
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC>
static int coop_launch(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                       unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    dim3 grid, block;
    stencil_block_shape(a.lay, V, &grid, &block, false);
    if (ACC && block.y < (unsigned)V) block.y = V;
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
    const size_t smem = ACC ? (size_t)planes * nthreads * sizeof(uint32_t) : 0;
    auto kern = k_sweep_stencil_coop<DIM, PMJ, K, ROUNDS, V, ACC>;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem) != cudaSuccess ||
        per_sm < 1)
        return -1;
    uint32_t g = a.lay.rows;
    const uint32_t resident = (uint32_t)per_sm * (uint32_t)sms;
    if (g > resident) g = resident;
    Layout L = a.lay;
    uint32_t* spins = a.spins;
    const uint32_t* jmask = a.jmask;
    uint32_t sweep0 = a.sweep, gw0 = a.gw0, antiferro = a.antiferro;
    PhiloxKeys pk = philox_round_keys(a.key0, a.key1);
    void* params[] = {&spins, &jmask, &L, &sweep0, &nsweeps, &pk, &gw0, &antiferro, &th_dev, &hist, &cw};
    if (cudaLaunchCooperativeKernel((void*)kern, dim3(g, 1, 1), block, params, smem, st) != cudaSuccess)
        return -1;
    return 1;
}

template <int DIM, bool PMJ, int V>
static int coop_dispatch(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                         unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    // the cooperative path is an optimisation for launch-bound sizes: default planes / rounds only
    if (a.planes != 6 || a.rounds != 10) return 0;
    return hist ? coop_launch<DIM, PMJ, 6, 10, V, true>(a, th_dev, nsweeps, hist, cw, st)
                : coop_launch<DIM, PMJ, 6, 10, V, false>(a, th_dev, nsweeps, nullptr, cw, st);
}

// returns 1 when the chunk was launched cooperatively, 0 when this configuration has no
// cooperative variant (caller falls back to one launch per colour phase), -1 on error
int launch_sweeps_stencil_coop(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                               unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    const bool pmj = a.jmask != nullptr;
    const bool d3 = a.lay.kind == ISING_KIND_STENCIL3D;
    if (!d3 && a.lay.kind != ISING_KIND_STENCIL2D) return 0;
#define COOP_V(VV)                                                                              \
    (d3 ? (pmj ? coop_dispatch<3, true, VV>(a, th_dev, nsweeps, hist, cw, st)                     \
               : coop_dispatch<3, false, VV>(a, th_dev, nsweeps, hist, cw, st))                   \
        : (pmj ? coop_dispatch<2, true, VV>(a, th_dev, nsweeps, hist, cw, st)                     \
               : coop_dispatch<2, false, VV>(a, th_dev, nsweeps, hist, cw, st)))
    if (a.lay.W % 4 == 0) return COOP_V(4);
    if (a.lay.W % 2 == 0) return COOP_V(2);
    return COOP_V(1);
#undef COOP_V
}

// ------------------------------------------------------------------------------------------
// K3: positional popcount (per-experiment integer observables from packed words)
// vertical counters: plane l of VCount holds bit l of 32 independent counters
// ------------------------------------------------------------------------------------------
constexpr int VC_PLANES = 12;           // counters up to 4095
constexpr int VC_FLUSH_ADD1 = 4095;

// reduce sm[32][nthreads] over threadIdx.y and add to out[(w0 + tx) * 32 + b]
__device__ __forceinline__ void block_reduce_counts(int* sm, unsigned long long* out, uint32_t w0,
                                                    uint32_t W) {
    const int wx = blockDim.x, by = blockDim.y, nthreads = wx * by;
    const int tid = threadIdx.y * wx + threadIdx.x;
    __syncthreads();
    for (int idx = tid; idx < 32 * wx; idx += nthreads) {
        const int b = idx / wx, tx = idx - b * wx;
        if (w0 + tx >= W) continue;
        long long sum = 0;
        for (int ty = 0; ty < by; ++ty) sum += sm[b * nthreads + ty * wx + tx];
        if (sum) atomicAdd(out + (size_t)(w0 + tx) * 32 + b, (unsigned long long)sum);
    }
    __syncthreads();
}

constexpr int NS_NP = 10;   // per-thread counter planes: up to 1023 = 146 sites x 7
constexpr int NS_MAX_ITEMS = 1023 / 7;

// n_sat[e] += satisfied bonds of experiment e.  Colour-0 sites see every bond exactly once.
// Per thread: V replica words, carry-save vertical counters over all its sites (no per-site
// integer work); per block: bit-sliced tree reduction through shared memory, one SWAR
// bit-transpose per word column, 32 integer atomics per column.
template <int DIM, bool PMJ, int V>
__global__ void __launch_bounds__(256)
k_nsat_stencil(const uint32_t* __restrict__ spins, const uint32_t* __restrict__ jm, Layout L,
               uint32_t antiferro, unsigned long long* __restrict__ nsat) {
    extern __shared__ uint32_t sm[];  // [max(NS_NP * V, NS_NR)][256]
    const uint32_t Lxh = L.Lxh, W = L.W, Ly = L.Ly, Lz = L.Lz;
    const uint32_t rowlen = Lxh * W;
    const size_t csz = (size_t)L.halfN * W;
    const uint32_t* __restrict__ own = spins;
    const uint32_t* __restrict__ oth = spins + csz;
    const uint32_t wx = blockDim.x, by = blockDim.y, nthreads = wx * by;
    const uint32_t tid = threadIdx.y * wx + threadIdx.x;
    const uint32_t C = wx * V;  // word columns handled per chunk
    for (uint32_t w0 = 0; w0 < W; w0 += C) {
        const uint32_t w = w0 + V * threadIdx.x;
        VCount<NS_NP> vc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) vc[v].clear();
        int pending = 0;
        {
            for (uint32_t row = blockIdx.x; row < L.rows; row += gridDim.x) {
                const uint32_t z = row / Ly, y = row - z * Ly;
                const uint32_t p = (y + z) & 1u;
                const uint32_t ym = y == 0 ? Ly - 1 : y - 1, yp = y + 1 == Ly ? 0 : y + 1;
                const uint32_t* o_c = own + (size_t)row * rowlen;
                const uint32_t* n_x = oth + (size_t)row * rowlen;
                const uint32_t* n_ym = oth + (size_t)(z * Ly + ym) * rowlen;
                const uint32_t* n_yp = oth + (size_t)(z * Ly + yp) * rowlen;
                const uint32_t* n_zm = nullptr;
                const uint32_t* n_zp = nullptr;
                if (DIM == 3) {
                    const uint32_t zm = z == 0 ? Lz - 1 : z - 1, zp = z + 1 == Lz ? 0 : z + 1;
                    n_zm = oth + (size_t)(zm * Ly + y) * rowlen;
                    n_zp = oth + (size_t)(zp * Ly + y) * rowlen;
                }
                for (uint32_t xh0 = 0; xh0 < Lxh; xh0 += by) {
                    const uint32_t xh = xh0 + threadIdx.y;
                    if (xh < Lxh && w < W) {
                    const uint32_t xs =
                        p ? (xh + 1 == Lxh ? 0 : xh + 1) : (xh == 0 ? Lxh - 1 : xh - 1);
                    const uint32_t i = xh * W + w;
                    uint32_t m[2 * DIM];
#pragma unroll
                    for (int k = 0; k < 2 * DIM; ++k)
                        m[k] = PMJ ? __ldg(jm + (size_t)k * L.halfN + (size_t)row * Lxh + xh)
                                   : antiferro;
                    uint32_t s[V], n[2 * DIM][V];
                    load_words<V>(o_c + i, s);
                    load_words<V>(n_x + i, n[0]);
                    load_words<V>(n_x + xs * W + w, n[1]);
                    load_words<V>(n_ym + i, n[2]);
                    load_words<V>(n_yp + i, n[3]);
                    if (DIM == 3) {
                        load_words<V>(n_zm + i, n[4]);
                        load_words<V>(n_zp + i, n[5]);
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        uint32_t a[2 * DIM];
#pragma unroll
                        for (int k = 0; k < 2 * DIM; ++k) a[k] = ~(s[v] ^ n[k][v] ^ m[k]);
                        uint32_t b0, b1, b2;
                        count_sat<DIM>(a, b0, b1, b2);
                        vc[v].add3(b0, b1, b2);
                    }
                    }
                    if (++pending == NS_MAX_ITEMS) {
                        block_reduce_vcount<NS_NP, V>(vc, sm, nsat, w0, W);
#pragma unroll
                        for (int v = 0; v < V; ++v) vc[v].clear();
                        pending = 0;
                    }
                }
            }
        }
        block_reduce_vcount<NS_NP, V>(vc, sm, nsat, w0, W);
    }
}

template <int V>
static int nsat_dispatch(const uint32_t* spins, const uint32_t* jmask, const Layout& lay,
                         uint32_t antiferro, unsigned long long* nsat, cudaStream_t st) {
    dim3 grid, block;
    stencil_block_shape(lay, V, &grid, &block, false);
    if (block.y < (unsigned)V) block.y = V;  // the reduction needs >= one thread per word column
    uint64_t g = 148ull * 2;  // persistent; counters are reduced every NS_MAX_ITEMS sites
    if (g > lay.rows) g = lay.rows;
    grid = dim3((unsigned)g, 1, 1);
    const int nthreads = block.x * block.y;
    const int planes = NS_NP * V > NS_NR ? NS_NP * V : NS_NR;
    const size_t smem = (size_t)planes * nthreads * sizeof(uint32_t);
    const bool pmj = jmask != nullptr;
#define NSAT_LAUNCH(D, P)                                                                     \
    do {                                                                                      \
        if (smem > 48 * 1024)                                                                 \
            cudaFuncSetAttribute(k_nsat_stencil<D, P, V>,                                     \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        k_nsat_stencil<D, P, V><<<grid, block, smem, st>>>(spins, jmask, lay, antiferro, nsat); \
    } while (0)
    if (lay.kind == ISING_KIND_STENCIL3D) {
        if (pmj) NSAT_LAUNCH(3, true); else NSAT_LAUNCH(3, false);
    } else if (lay.kind == ISING_KIND_STENCIL2D) {
        if (pmj) NSAT_LAUNCH(2, true); else NSAT_LAUNCH(2, false);
    } else {
        return -1;
    }
#undef NSAT_LAUNCH
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_nsat_stencil(const uint32_t* spins, const uint32_t* jmask, const Layout& lay,
                        uint32_t antiferro, unsigned long long* nsat, cudaStream_t st) {
    if (lay.W % 4 == 0) return nsat_dispatch<4>(spins, jmask, lay, antiferro, nsat, st);
    if (lay.W % 2 == 0) return nsat_dispatch<2>(spins, jmask, lay, antiferro, nsat, st);
    return nsat_dispatch<1>(spins, jmask, lay, antiferro, nsat, st);
}

// up-spin count over all sites (any layout: the sum runs over every stored site word)
__global__ void __launch_bounds__(256)
k_count_up(const uint32_t* __restrict__ spins, uint64_t nsites, uint32_t W,
           unsigned long long* __restrict__ up, uint32_t pair) {
    __shared__ int sm[32 * 256];
    const int nthreads = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (uint32_t w0 = 0; w0 < W; w0 += blockDim.x) {
        for (int b = 0; b < 32; ++b) sm[b * nthreads + tid] = 0;
        const uint32_t w = w0 + threadIdx.x;
        VCount<VC_PLANES> vc;
        vc.clear();
        int pending = 0;
        if (w < W) {
            for (uint64_t n = (uint64_t)blockIdx.x * blockDim.y + threadIdx.y; n < nsites;
                 n += (uint64_t)gridDim.x * blockDim.y) {
                uint32_t x = spins[(size_t)n * W + w];
                // pair mode: bit 2p = experiments 2p and 2p+1 disagree on this site
                if (pair) x = (x ^ (x >> 1)) & 0x55555555u;
                vc.add1(x);
                if (++pending == VC_FLUSH_ADD1) {
                    vc.flush(sm, tid, nthreads);
                    pending = 0;
                }
            }
            vc.flush(sm, tid, nthreads);
        }
        block_reduce_counts(sm, up, w0, W);
    }
}

int launch_count_up(const uint32_t* spins, const Layout& lay, unsigned long long* up,
                    cudaStream_t st, bool pair) {
    const uint32_t wx = lay.W >= 32 ? 32 : pow2_ceil(lay.W);
    dim3 block(wx, 256 / wx, 1);
    uint64_t g = (lay.nvars + block.y - 1) / block.y;
    if (g > 148u * 8u) g = 148u * 8u;
    if (g == 0) g = 1;
    k_count_up<<<dim3((unsigned)g), block, 0, st>>>(spins, lay.nvars, lay.W, up, pair ? 1u : 0u);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// overlap of the experiment pair (2p, 2p+1) from the pair-mode counts: q = N - 2 * disagreements
__global__ void k_overlap_from_counts(const unsigned long long* __restrict__ dis, uint64_t P,
                                      uint64_t nsites, double* __restrict__ out, uint64_t stride,
                                      uint64_t off) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    out[p * stride + off] = (double)((long long)nsites - 2ll * (long long)dis[2 * p]);
}

int launch_overlap_from_counts(const unsigned long long* dis, uint64_t P, uint64_t nsites,
                               double* out_dev, uint64_t stride, uint64_t off, cudaStream_t st) {
    const unsigned g = (unsigned)((P + 255) / 256);
    k_overlap_from_counts<<<g ? g : 1, 256, 0, st>>>(dis, P, nsites, out_dev, stride, off);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_energy_from_nsat(const unsigned long long* __restrict__ nsat, uint64_t E,
                                   double scale, uint64_t nbonds, int mult,
                                   double* __restrict__ out, uint64_t estride, uint64_t eoff) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long v = (long long)nbonds - (long long)mult * (long long)nsat[e];
    out[e * estride + eoff] = scale * (double)v;
}

// energies[e * nt + t] = scale * (nbonds - 2 * hist[t * cw + e]) for a chunk of nt sweeps
__global__ void k_energy_from_hist(const unsigned long long* __restrict__ hist, uint64_t E,
                                   uint64_t cw, uint64_t nt, double scale, uint64_t nbonds,
                                   int mult, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * nt) return;
    const uint64_t e = i / nt, t = i - e * nt;
    const long long v = (long long)nbonds - (long long)mult * (long long)hist[t * cw + e];
    out[i] = scale * (double)v;
}

int launch_energy_from_hist(const unsigned long long* hist, uint64_t E, uint64_t cw, uint64_t nt,
                            double scale, uint64_t nbonds, int mult, double* out_dev,
                            cudaStream_t st) {
    const uint64_t n = E * nt;
    const unsigned g = (unsigned)((n + 255) / 256);
    k_energy_from_hist<<<g ? g : 1, 256, 0, st>>>(hist, E, cw, nt, scale, nbonds, mult, out_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// out[e * estride + eoff] = in[e]   /   out[e * nt + t] = hist[t * cw + e]  (f64 energies)
__global__ void k_copy_strided_f64(const double* __restrict__ in, uint64_t E,
                                   double* __restrict__ out, uint64_t estride, uint64_t eoff) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) out[e * estride + eoff] = in[e];
}

__global__ void k_transpose_hist_f64(const double* __restrict__ hist, uint64_t E, uint64_t cw,
                                     uint64_t nt, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * nt) return;
    const uint64_t e = i / nt, t = i - e * nt;
    out[i] = hist[t * cw + e];
}

int launch_copy_strided_f64(const double* in, uint64_t E, double* out, uint64_t estride,
                            uint64_t eoff, cudaStream_t st) {
    const unsigned g = (unsigned)((E + 255) / 256);
    k_copy_strided_f64<<<g ? g : 1, 256, 0, st>>>(in, E, out, estride, eoff);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_transpose_hist_f64(const double* hist, uint64_t E, uint64_t cw, uint64_t nt, double* out,
                              cudaStream_t st) {
    const uint64_t n = E * nt;
    const unsigned g = (unsigned)((n + 255) / 256);
    k_transpose_hist_f64<<<g ? g : 1, 256, 0, st>>>(hist, E, cw, nt, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_energy_from_nsat(const unsigned long long* nsat, uint64_t E, double scale,
                            uint64_t nbonds, int mult, double* out_dev, uint64_t estride,
                            uint64_t eoff, cudaStream_t st) {
    const unsigned g = (unsigned)((E + 255) / 256);
    k_energy_from_nsat<<<g ? g : 1, 256, 0, st>>>(nsat, E, scale, nbonds, mult, out_dev, estride, eoff);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// K4: colour-class sweep on an arbitrary graph (CSR/ELL), all |J| equal, no bias.
// Same decision rule and Philox stream as the stencil kernel (DESIGN.md "Production sweep"),
// generic in the degree: n_sat is a 4-plane vertical counter, the uphill classes are
// n_sat = deg/2+1 .. deg.  PERBETA: thresholds differ per replica (parallel tempering).
// ------------------------------------------------------------------------------------------
// DEG > 0: compile-time degree (neighbour loads unrolled and in flight together); DEG = 0: runtime.
// V consecutive replica words of a site per thread: one index load / address computation and one
// 4V-byte gather per neighbour for V words (needs W % V == 0).
template <int K, int ROUNDS, bool PERBETA, int DEG, int V>
__global__ void __launch_bounds__(256)
k_sweep_general(uint32_t* __restrict__ spins, GenGroup g, uint32_t W, uint32_t sweep, PhiloxKeys pk,
                uint32_t gw0, GenThresholds th, GenTables tab) {
    constexpr int NCALL = K / 4 + 1;
    // planes of the n_sat counter: enough for DEG when it is known at compile time
    constexpr int NPL = DEG == 0 ? 4 : (DEG < 2 ? 1 : (DEG < 4 ? 2 : (DEG < 8 ? 3 : 4)));
    const uint32_t deg = DEG > 0 ? (uint32_t)DEG : g.deg;
    const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
    // block = (wx lanes over replica word groups, by over sites): no division to split an item index
    for (uint32_t i = blockIdx.x * blockDim.y + threadIdx.y; i < g.count; i += gridDim.x * blockDim.y)
    for (uint32_t w0 = threadIdx.x * V; w0 < W; w0 += blockDim.x * V) {
        const uint32_t n = g.sites[i];
        const uint32_t ab = g.anti[i];
        uint32_t sv[V];
        load_words<V>(spins + (size_t)n * W + w0, sv);
        uint32_t cntv[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) cntv[v][l] = 0;
        if constexpr (DEG > 0) {
            uint32_t x[DEG][V];
#pragma unroll
            for (int k = 0; k < DEG; ++k)
                load_words<V>(spins + (size_t)g.nbr[(size_t)k * g.count + i] * W + w0, x[k]);
#pragma unroll
            for (int k = 0; k < DEG; ++k) {
                const uint32_t m = 0u - ((ab >> k) & 1u);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    uint32_t c = ~(sv[v] ^ x[k][v] ^ m);  // satisfied bond
#pragma unroll
                    for (int l = 0; l < NPL; ++l) {
                        const uint32_t t = cntv[v][l] & c;
                        cntv[v][l] ^= c;
                        c = t;
                    }
                }
            }
        } else {
            for (uint32_t k = 0; k < deg; ++k) {
                const uint32_t nb = g.nbr[(size_t)k * g.count + i];
                uint32_t x[V];
                load_words<V>(spins + (size_t)nb * W + w0, x);
                const uint32_t m = 0u - ((ab >> k) & 1u);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    uint32_t c = ~(sv[v] ^ x[v] ^ m);  // satisfied bond
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const uint32_t t = cntv[v][l] & c;
                        cntv[v][l] ^= c;
                        c = t;
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
        const uint32_t w = w0 + v;
        const uint32_t (&cnt)[4] = cntv[v];
        // one-hot masks of the uphill classes
        uint32_t oh[GEN_MAX_CLS];
        uint32_t up = 0;
#pragma unroll
        for (int j = 0; j < GEN_MAX_CLS; ++j) {
            oh[j] = 0;
            if ((uint32_t)j < ncls) {
                const uint32_t val = cmin + j;
                uint32_t o = 0xFFFFFFFFu;
#pragma unroll
                for (int l = 0; l < NPL; ++l) o &= ((val >> l) & 1u) ? cnt[l] : ~cnt[l];
                oh[j] = o;
                up |= o;
            }
        }
        uint32_t r[NCALL * 4];
#pragma unroll
        for (int q = 0; q < NCALL; ++q) {
            const u32x4 o = philox4x32_keys<ROUNDS>(n, gw0 + w, sweep, (uint32_t)q | (TAG_ACCEPT << 24), pk);
            r[4 * q + 0] = o.x; r[4 * q + 1] = o.y; r[4 * q + 2] = o.z; r[4 * q + 3] = o.w;
        }
        const uint32_t* tp = PERBETA ? tab.plane + ((size_t)deg * W + w) * GEN_MAX_CLS * 8 : nullptr;
        uint32_t eq = up, borrow = 0;
#pragma unroll
        for (int p = K - 1; p >= 0; --p) {
            uint32_t t = 0;
#pragma unroll
            for (int j = 0; j < GEN_MAX_CLS; ++j)
                if ((uint32_t)j < ncls) t |= oh[j] & (PERBETA ? __ldg(tp + j * 8 + p) : th.plane[j][p]);
            borrow = maj3(~r[p], t, borrow);
            eq &= ~(r[p] ^ t);
        }
        uint32_t flip = ~up | (borrow & ~eq);
        // tied bits: the first SPARE in straight-line code on the words left over from the calls
        // above (as in msc_flip_mask), the rare rest in a loop
        constexpr int SPARE = (4 * NCALL - K) < 2 ? (4 * NCALL - K) : 2;
#pragma unroll
        for (int j2 = 0; j2 < SPARE; ++j2) {
            const uint32_t bit = eq & (0u - eq);
            const int b = (__ffs((int)eq) - 1) & 31;
            uint32_t cls = 0;
#pragma unroll
            for (int j = 1; j < GEN_MAX_CLS; ++j)
                if ((uint32_t)j < ncls && (oh[j] & bit)) cls = j;
            uint32_t lo;
            if (PERBETA) {
                lo = __ldg(tab.low + ((size_t)deg * 32 * W + (size_t)w * 32 + b) * GEN_MAX_CLS + cls);
            } else {
                lo = th.low[0];
#pragma unroll
                for (int j = 1; j < GEN_MAX_CLS; ++j)
                    if (cls == (uint32_t)j) lo = th.low[j];
            }
            if (r[K + j2] < lo) flip |= bit;
            eq ^= bit;
        }
        if (eq) {
            int jj = K + SPARE;
            u32x4 cur = {r[4 * (NCALL - 1)], r[4 * (NCALL - 1) + 1], r[4 * (NCALL - 1) + 2],
                         r[4 * (NCALL - 1) + 3]};
            do {
                const int b = __ffs((int)eq) - 1;
                if ((jj & 3) == 0 && jj >= 4 * NCALL)
                    cur = philox4x32_keys<ROUNDS>(n, gw0 + w, sweep, (uint32_t)(jj >> 2) | (TAG_ACCEPT << 24), pk);
                const int m = jj & 3;
                const uint32_t val = m == 0 ? cur.x : (m == 1 ? cur.y : (m == 2 ? cur.z : cur.w));
                uint32_t cls = 0;
#pragma unroll
                for (int j = 1; j < GEN_MAX_CLS; ++j)
                    if ((oh[j] >> b) & 1u) cls = j;
                const uint32_t lo = PERBETA
                    ? __ldg(tab.low + ((size_t)deg * 32 * W + (size_t)w * 32 + b) * GEN_MAX_CLS + cls)
                    : th.low[cls];
                if (val < lo) flip |= 1u << b;
                eq &= eq - 1;
                ++jj;
            } while (eq);
        }
        sv[v] ^= flip;
        }
        store_words<V>(spins + (size_t)n * W + w0, sv);
    }
}

template <int K, int ROUNDS, int DEG, int V>
static void gen_launch(const GenSweepArgs& a, const GenGroup& g, cudaStream_t st) {
    const uint32_t groups = a.W / V;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)g.count + block.y - 1) / block.y;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    const dim3 grid((unsigned)blocks);
    const PhiloxKeys pk = philox_round_keys(a.key0, a.key1);
    if (a.tables.plane != nullptr)
        k_sweep_general<K, ROUNDS, true, DEG, V><<<grid, block, 0, st>>>(a.spins, g, a.W, a.sweep, pk, a.gw0,
                                                                         a.th, a.tables);
    else
        k_sweep_general<K, ROUNDS, false, DEG, V><<<grid, block, 0, st>>>(a.spins, g, a.W, a.sweep, pk, a.gw0,
                                                                          a.th, a.tables);
}

// degree-specialised kernels only for the default (K, rounds)
template <int K, int ROUNDS, int V>
static void gen_launch_degree(const GenSweepArgs& a, const GenGroup& g, cudaStream_t st) {
    if constexpr (K == 6 && ROUNDS == 10) {
        if (g.deg == 3) return gen_launch<K, ROUNDS, 3, V>(a, g, st);
        if (g.deg == 4) return gen_launch<K, ROUNDS, 4, V>(a, g, st);
        if (g.deg == 6) return gen_launch<K, ROUNDS, 6, V>(a, g, st);
    }
    gen_launch<K, ROUNDS, 0, V>(a, g, st);
}

template <int K, int ROUNDS>
static void gen_launch_vec(const GenSweepArgs& a, const GenGroup& g, cudaStream_t st) {
    if (a.W % 2 == 0) gen_launch_degree<K, ROUNDS, 2>(a, g, st);
    else gen_launch_degree<K, ROUNDS, 1>(a, g, st);
}

int launch_sweep_general(const GenSweepArgs& a, const GenGroup& g, cudaStream_t st) {
    if (g.count == 0) return 0;
    if (g.deg > (uint32_t)GEN_MAX_DEG) return -1;
#define GEN_ROUNDS(KK)                                                                            \
    do { if (a.rounds == 7) gen_launch_vec<KK, 7>(a, g, st); else gen_launch_vec<KK, 10>(a, g, st); } while (0)
    switch (a.planes) {
        case 5: GEN_ROUNDS(5); break;
        case 6: GEN_ROUNDS(6); break;
        case 7: GEN_ROUNDS(7); break;
        default: return -1;
    }
#undef GEN_ROUNDS
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// per-replica threshold tables from host-computed 64-bit thresholds (integer work only, so the
// bits are exactly the host's): T64[(slot * (GEN_MAX_DEG+1) + deg) * GEN_MAX_CLS + cls]
// one warp per (degree, word, class); lane b = replica bit b, plane masks by ballot
__global__ void k_build_tables(const unsigned long long* __restrict__ t64,
                               const uint32_t* __restrict__ slot_of_replica, uint32_t W, int K,
                               uint32_t* __restrict__ plane_out, uint32_t* __restrict__ low_out) {
    const uint32_t total = (GEN_MAX_DEG + 1) * W * GEN_MAX_CLS;
    const uint32_t b = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < total; idx += warps) {
        const uint32_t cls = idx % GEN_MAX_CLS;
        const uint32_t w = (idx / GEN_MAX_CLS) % W;
        const uint32_t deg = idx / (GEN_MAX_CLS * W);
        const uint32_t e = w * 32 + b;
        const uint32_t slot = slot_of_replica[e];
        const unsigned long long T = t64[((size_t)slot * (GEN_MAX_DEG + 1) + deg) * GEN_MAX_CLS + cls];
        low_out[((size_t)deg * 32 * W + e) * GEN_MAX_CLS + cls] = (uint32_t)(T & 0xFFFFFFFFull);
        uint32_t mine = 0;  // lane p keeps plane p
        for (int p = 0; p < 8; ++p) {
            const uint32_t m = p < K ? __ballot_sync(0xFFFFFFFFu, (T >> (K + 31 - p)) & 1ull) : 0u;
            if (b == (uint32_t)p) mine = m;
        }
        if (b < 8) plane_out[(((size_t)deg * W + w) * GEN_MAX_CLS + cls) * 8 + b] = mine;
    }
}

// stencil variant: T64[e * 3 + cls] per replica -> tplane[(w * 3 + cls) * 8 + p], tlow[(e) * 3 + cls]
__global__ void k_build_tables_stencil(const unsigned long long* __restrict__ t64, uint32_t W, int K,
                                       uint32_t* __restrict__ plane_out, uint32_t* __restrict__ low_out) {
    const uint32_t b = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < W * 3; idx += warps) {
        const uint32_t cls = idx % 3, w = idx / 3;
        const uint32_t e = w * 32 + b;
        const unsigned long long T = t64[(size_t)e * 3 + cls];
        low_out[(size_t)e * 3 + cls] = (uint32_t)(T & 0xFFFFFFFFull);
        uint32_t mine = 0;
        for (int p = 0; p < 8; ++p) {
            const uint32_t m = p < K ? __ballot_sync(0xFFFFFFFFu, (T >> (K + 31 - p)) & 1ull) : 0u;
            if (b == (uint32_t)p) mine = m;
        }
        if (b < 8) plane_out[((size_t)w * 3 + cls) * 8 + b] = mine;
    }
}

int launch_build_tables_stencil(const unsigned long long* t64, uint32_t W, int K, uint32_t* plane_out,
                                uint32_t* low_out, cudaStream_t st) {
    const uint32_t blocks = (W * 3 + 3) / 4;   // 4 warps per block
    k_build_tables_stencil<<<blocks < 1184u ? blocks : 1184u, 128, 0, st>>>(t64, W, K, plane_out, low_out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_build_tables(const unsigned long long* t64, const uint32_t* slot_of_replica, uint32_t W,
                        int K, uint32_t* plane_out, uint32_t* low_out, cudaStream_t st) {
    const uint32_t total = (GEN_MAX_DEG + 1) * W * GEN_MAX_CLS;
    const uint32_t blocks = (total + 3) / 4;
    k_build_tables<<<blocks < 1184u ? blocks : 1184u, 128, 0, st>>>(t64, slot_of_replica, W, K, plane_out, low_out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// satisfied bonds per experiment on a general graph (each bond seen from both ends)
__global__ void __launch_bounds__(256)
k_nsat_general(const uint32_t* __restrict__ spins, uint64_t nvars, uint32_t W,
               const uint32_t* __restrict__ row, const uint32_t* __restrict__ nbr,
               const uint8_t* __restrict__ anti, unsigned long long* __restrict__ nsat2) {
    __shared__ int sm[32 * 256];
    const int nthreads = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (uint32_t w0 = 0; w0 < W; w0 += blockDim.x) {
        for (int b = 0; b < 32; ++b) sm[b * nthreads + tid] = 0;
        const uint32_t w = w0 + threadIdx.x;
        VCount<VC_PLANES> vc;
        vc.clear();
        int pending = 0;
        if (w < W) {
            for (uint64_t n = (uint64_t)blockIdx.x * blockDim.y + threadIdx.y; n < nvars;
                 n += (uint64_t)gridDim.x * blockDim.y) {
                const uint32_t s = spins[(size_t)n * W + w];
                const uint32_t lo = row[n], hi = row[n + 1];
                // satisfied bonds of this site in a 4-plane counter (degree <= 15 on this path),
                // neighbours four at a time so that the gathers are in flight together
                uint32_t cnt[4] = {0, 0, 0, 0};
                for (uint32_t k = lo; k < hi; k += 4) {
                    uint32_t c4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool ok = k + u < hi;
                        const uint32_t x = spins[(size_t)(ok ? nbr[k + u] : n) * W + w];
                        const uint32_t m = (ok && anti[k + u]) ? 0xFFFFFFFFu : 0u;
                        c4[u] = ok ? ~(s ^ x ^ m) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t c = c4[u];
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            const uint32_t t = cnt[l] & c;
                            cnt[l] ^= c;
                            c = t;
                        }
                    }
                }
                vadd<VC_PLANES, 4>(vc.v, cnt);
                pending += 15;
                if (pending > VC_FLUSH_ADD1 - 15) {
                    vc.flush(sm, tid, nthreads);
                    pending = 0;
                }
            }
            vc.flush(sm, tid, nthreads);
        }
        block_reduce_counts(sm, nsat2, w0, W);
    }
}

int launch_nsat_general(const uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row,
                        const uint32_t* nbr, const uint8_t* anti, unsigned long long* nsat2,
                        cudaStream_t st) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    dim3 block(wx, 256 / wx, 1);
    uint64_t g = (nvars + block.y - 1) / block.y;
    if (g > 148u * 8u) g = 148u * 8u;
    if (g == 0) g = 1;
    k_nsat_general<<<dim3((unsigned)g), block, 0, st>>>(spins, nvars, W, row, nbr, anti, nsat2);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// Colour-class sweep for arbitrary real couplings and biases (lattice.rs:31, 104-131): the local
// field is not an integer class, so every replica bit gets its own float field, its own
// exp(-beta dE) and its own 32-bit uniform (word b%4 of Philox call b/4 on the usual counter).
// dE = -2 s_i sum_k J_ik s_k + 2 b_i s_i  (qmc GraphState::do_spin_flip); accept iff dE <= 0 or
// R < floor(exp(-beta dE) 2^32).  Validated statistically (f32 field / __expf), not bit-exactly.
// ------------------------------------------------------------------------------------------
template <int ROUNDS>
__global__ void __launch_bounds__(256)
k_sweep_real(RealSweepArgs a) {
    for (uint32_t i = blockIdx.x * blockDim.y + threadIdx.y; i < a.count; i += gridDim.x * blockDim.y)
    for (uint32_t w = threadIdx.x; w < a.W; w += blockDim.x) {
        const uint32_t n = a.sites[i];
        const uint32_t s = a.spins[(size_t)n * a.W + w];
        const uint32_t lo = a.row[n], hi = a.row[n + 1];
        const float bias = a.biasf[n];
        uint32_t flip = 0;
        for (uint32_t b0 = 0; b0 < 32; b0 += 4) {
            const u32x4 r = philox4x32<ROUNDS>(n, a.gw0 + w, a.sweep, (b0 >> 2) | (TAG_ACCEPT << 24),
                                               a.key0, a.key1);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            float h[4] = {0.f, 0.f, 0.f, 0.f};
            for (uint32_t k = lo; k < hi; ++k) {
                const uint32_t x = a.spins[(size_t)a.nbr[k] * a.W + w] >> b0;
                const uint32_t jb = __float_as_uint(a.jf[k]);
#pragma unroll
                for (int q = 0; q < 4; ++q)  // J * s_k: flip the sign bit where the spin is down
                    h[q] += __uint_as_float(jb ^ ((~(x >> q) & 1u) << 31));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float si = ((s >> (b0 + q)) & 1u) ? 1.f : -1.f;
                const float de = 2.f * si * (bias - h[q]);
                bool acc = true;
                if (de > 0.f) {
                    const float pth = __expf(-a.beta * de) * 4294967296.f;
                    acc = rr[q] < __float2uint_rz(pth);  // saturating conversion
                }
                if (acc) flip |= 1u << (b0 + q);
            }
        }
        a.spins[(size_t)n * a.W + w] = s ^ flip;
    }
}

int launch_sweep_real(const RealSweepArgs& a, cudaStream_t st) {
    if (a.count == 0) return 0;
    const uint32_t wx = a.W >= 32 ? 32 : pow2_ceil(a.W);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)a.count + block.y - 1) / block.y;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    if (a.rounds == 7) k_sweep_real<7><<<(unsigned)blocks, block, 0, st>>>(a);
    else k_sweep_real<10><<<(unsigned)blocks, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void __launch_bounds__(256)
k_energy_real(const uint32_t* __restrict__ spins, uint64_t nvars, uint32_t W,
              const uint32_t* __restrict__ row, const uint32_t* __restrict__ nbr,
              const double* __restrict__ jv, const double* __restrict__ bias,
              double* __restrict__ energies) {
    // block = (wx word columns, by site lanes); each thread keeps 32 f64 partial energies
    const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.0;
    for (uint64_t n = (uint64_t)blockIdx.x * blockDim.y + threadIdx.y; n < nvars;
         n += (uint64_t)gridDim.x * blockDim.y) {
        const uint32_t s = spins[(size_t)n * W + w];
        const double bi = bias[n];
        for (uint32_t k = row[n]; k < row[n + 1]; ++k) {
            const uint32_t eqm = ~(s ^ spins[(size_t)nbr[k] * W + w]);  // 1 where s_i == s_k
            const double hj = 0.5 * jv[k];
#pragma unroll
            for (int b = 0; b < 32; ++b) acc[b] += ((eqm >> b) & 1u) ? hj : -hj;
        }
#pragma unroll
        for (int b = 0; b < 32; ++b) acc[b] += ((s >> b) & 1u) ? -bi : bi;
    }
#pragma unroll
    for (int b = 0; b < 32; ++b) atomicAdd(energies + (size_t)w * 32 + b, acc[b]);
}

int launch_energy_real(const uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row,
                       const uint32_t* nbr, const double* jv, const double* bias, double* energies,
                       cudaStream_t st) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    dim3 block(wx, 128 / wx, 1);
    uint64_t g = (nvars + block.y - 1) / block.y;
    if (g > 148u * 4u) g = 148u * 4u;
    if (g == 0) g = 1;
    dim3 grid((unsigned)g, (W + wx - 1) / wx, 1);
    k_energy_real<<<grid, block, 0, st>>>(spins, nvars, W, row, nbr, jv, bias, energies);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// K7: one large 2D lattice bit-packed along x (config 5).  Same decision rule as the replica-
// packed kernels; here the 32 bits of a word are 32 same-colour sites of one row, so the two
// x neighbours are the other-colour word at the same index and that word funnel-shifted by one
// bit (carry from the adjacent word).  Philox counter = (global row, colour << 30 | word, sweep,
// call): a draw does not depend on how rows are split into strips.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t strip_off(const StripGeom& g, uint32_t c, uint32_t r, uint32_t j) {
    return ((size_t)c * (g.rows + 2 * g.ghost) + r) * g.Wr + j;
}

// V consecutive words of a row per thread (128-bit loads when V == 4; needs Wr % V == 0)
template <int K, int ROUNDS, int V>
__global__ void __launch_bounds__(256)
k_strip_phase(uint32_t* __restrict__ spins, StripGeom g, uint32_t c, uint32_t sweep, PhiloxKeys pk,
              uint32_t antiferro, MscThresholds th, uint32_t r_begin, uint32_t r_count) {
    const uint32_t groups = g.Wr / V;
    // block = (x over the word groups of a row, y over rows): storage rows [r_begin, r_begin + r_count)
    for (uint32_t rr = blockIdx.y * blockDim.y + threadIdx.y; rr < r_count; rr += gridDim.y * blockDim.y)
    for (uint32_t jg = blockIdx.x * blockDim.x + threadIdx.x; jg < groups; jg += gridDim.x * blockDim.x) {
        const uint32_t j = jg * V;
        const uint32_t r = r_begin + rr;
        const uint32_t y = strip_global_row(g, r);
        const uint32_t p = (y + c) & 1u;
        const uint32_t o = 1u - c;
        uint32_t s[V], nx[V], nu[V], nd[V];
        load_words<V>(spins + strip_off(g, c, r, j), s);
        load_words<V>(spins + strip_off(g, o, r, j), nx);
        load_words<V>(spins + strip_off(g, o, r - 1, j), nu);
        load_words<V>(spins + strip_off(g, o, r + 1, j), nd);
        // the x neighbour one bit over: funnel shift with carry from the adjacent word
        const uint32_t edge = p ? spins[strip_off(g, o, r, j + V == g.Wr ? 0 : j + V)]
                                : spins[strip_off(g, o, r, j == 0 ? g.Wr - 1 : j - 1)];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            uint32_t nsh;
            if (p) nsh = __funnelshift_r(nx[v], v + 1 < V ? nx[v + 1 < V ? v + 1 : v] : edge, 1);
            else nsh = __funnelshift_l(v > 0 ? nx[v > 0 ? v - 1 : 0] : edge, nx[v], 1);
            uint32_t a[4] = {~(s[v] ^ nx[v] ^ antiferro), ~(s[v] ^ nsh ^ antiferro),
                             ~(s[v] ^ nu[v] ^ antiferro), ~(s[v] ^ nd[v] ^ antiferro)};
            uint32_t b0, b1, b2;
            count_sat<2>(a, b0, b1, b2);
            s[v] ^= msc_flip_mask<2, K, ROUNDS>(b2 | (b1 & b0), b2, 0u, th, y, (c << 30) | (j + v), sweep, pk);
        }
        store_words<V>(spins + strip_off(g, c, r, j), s);
    }
}

template <int V>
static int strip_phase_dispatch(const StripSweepArgs& a, cudaStream_t st) {
    const uint32_t groups = a.g.Wr / V;
    const uint32_t bx = groups >= 256 ? 256 : pow2_ceil(groups);
    const dim3 block(bx, 256 / bx, 1);
    uint32_t gx = (groups + bx - 1) / bx;
    if (gx > 64) gx = 64;
    uint32_t gy = (a.r_count + block.y - 1) / block.y;
    const uint32_t gy_cap = (148u * 32u + gx - 1) / gx;
    if (gy > gy_cap) gy = gy_cap;
    const dim3 grid(gx, gy, 1);
#define STRIP_LAUNCH(KK, RR)                                                                     \
    k_strip_phase<KK, RR, V><<<grid, block, 0, st>>>(a.spins, a.g, a.colour, a.sweep,              \
                                                     philox_round_keys(a.key0, a.key1), a.antiferro, \
                                                     a.th, a.r_begin, a.r_count)
#define STRIP_ROUNDS(KK)                                                                         \
    do { if (a.rounds == 7) STRIP_LAUNCH(KK, 7); else STRIP_LAUNCH(KK, 10); } while (0)
    switch (a.planes) {
        case 5: STRIP_ROUNDS(5); break;
        case 6: STRIP_ROUNDS(6); break;
        case 7: STRIP_ROUNDS(7); break;
        default: return -1;
    }
#undef STRIP_ROUNDS
#undef STRIP_LAUNCH
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_strip_phase(const StripSweepArgs& a, cudaStream_t st) {
    if (a.r_count == 0 || a.g.Wr == 0) return 0;
    // rows are 16-byte aligned when Wr % 4 == 0 (every row starts at a multiple of Wr words)
    if (a.g.Wr % 4 == 0) return strip_phase_dispatch<4>(a, st);
    if (a.g.Wr % 2 == 0) return strip_phase_dispatch<2>(a, st);
    return strip_phase_dispatch<1>(a, st);
}

__global__ void k_strip_init_random(uint32_t* __restrict__ spins, StripGeom g, uint32_t k0,
                                    uint32_t k1) {
    const uint64_t total = 2ull * g.rows * g.Wr;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i / ((uint64_t)g.rows * g.Wr));
        const uint64_t rem = i - (uint64_t)c * g.rows * g.Wr;
        const uint32_t r = (uint32_t)(rem / g.Wr) + g.ghost, j = (uint32_t)(rem % g.Wr);
        const u32x4 v = philox4x32<10>(g.row0 + r - g.ghost, (c << 30) | j, 0u, TAG_INIT << 24, k0, k1);
        spins[strip_off(g, c, r, j)] = v.x;
    }
}

int launch_strip_init_random(uint32_t* spins, const StripGeom& g, uint32_t key0, uint32_t key1,
                             cudaStream_t st) {
    k_strip_init_random<<<148 * 8, 256, 0, st>>>(spins, g, key0, key1);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// per-lattice observables are plain popcounts in this layout
__global__ void __launch_bounds__(256)
k_strip_observables(const uint32_t* __restrict__ spins, StripGeom g, uint32_t antiferro,
                    unsigned long long* __restrict__ acc) {
    unsigned long long nsat = 0, up = 0;
    const uint64_t total = (uint64_t)g.rows * g.Wr;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(item / g.Wr) + g.ghost, j = (uint32_t)(item % g.Wr);
        const uint32_t y = g.row0 + r - g.ghost;
        const uint32_t p = y & 1u;  // colour 0
        const uint32_t s = spins[strip_off(g, 0, r, j)];
        const uint32_t nx = spins[strip_off(g, 1, r, j)];
        uint32_t nsh;
        if (p) nsh = __funnelshift_r(nx, spins[strip_off(g, 1, r, j + 1 == g.Wr ? 0 : j + 1)], 1);
        else nsh = __funnelshift_l(spins[strip_off(g, 1, r, j == 0 ? g.Wr - 1 : j - 1)], nx, 1);
        const uint32_t nu = spins[strip_off(g, 1, r - 1, j)];
        const uint32_t nd = spins[strip_off(g, 1, r + 1, j)];
        nsat += __popc(~(s ^ nx ^ antiferro)) + __popc(~(s ^ nsh ^ antiferro)) +
                __popc(~(s ^ nu ^ antiferro)) + __popc(~(s ^ nd ^ antiferro));
        up += __popc(s) + __popc(nx);
    }
    for (int off = 16; off; off >>= 1) {
        nsat += __shfl_xor_sync(0xFFFFFFFFu, nsat, off);
        up += __shfl_xor_sync(0xFFFFFFFFu, up, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nsat) atomicAdd(acc, nsat);
        if (up) atomicAdd(acc + 1, up);
    }
}

int launch_strip_observables(const uint32_t* spins, const StripGeom& g, uint32_t antiferro,
                             unsigned long long* acc, cudaStream_t st) {
    k_strip_observables<<<148 * 4, 256, 0, st>>>(spins, g, antiferro, acc);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_strip_unpack(const uint32_t* __restrict__ spins, StripGeom g,
                               uint8_t* __restrict__ out) {
    const uint64_t Lx = 64ull * g.Wr;
    const uint64_t total = (uint64_t)g.rows * Lx;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / Lx) + g.ghost;
        const uint32_t x = (uint32_t)(i % Lx);
        const uint32_t y = g.row0 + r - g.ghost;
        const uint32_t c = (x + y) & 1u, xh = x >> 1;
        out[i] = (uint8_t)((spins[strip_off(g, c, r, xh >> 5)] >> (xh & 31u)) & 1u);
    }
}

int launch_strip_unpack(const uint32_t* spins, const StripGeom& g, uint8_t* out_dev, cudaStream_t st) {
    k_strip_unpack<<<148 * 8, 256, 0, st>>>(spins, g, out_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// state initialisation / import / export (not hot)
// ------------------------------------------------------------------------------------------
__global__ void k_init_random(uint32_t* __restrict__ spins, Layout L, uint32_t k0, uint32_t k1,
                              uint32_t gw0) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        const u32x4 r = philox4x32<10>((uint32_t)n, gw0 + w, 0u, TAG_INIT << 24, k0, k1);
        spins[site_word_base(L, n) + w] = r.x;
    }
}

int launch_init_random(uint32_t* spins, const Layout& lay, uint32_t key0, uint32_t key1,
                       uint32_t gw0, cudaStream_t st) {
    k_init_random<<<148 * 8, 256, 0, st>>>(spins, lay, key0, key1, gw0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_init_broadcast(uint32_t* __restrict__ spins, Layout L,
                                 const uint8_t* __restrict__ state) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        spins[site_word_base(L, n) + w] = state[n] ? 0xFFFFFFFFu : 0u;
    }
}

int launch_init_broadcast(uint32_t* spins, const Layout& lay, const uint8_t* state_dev,
                          cudaStream_t st) {
    k_init_broadcast<<<148 * 8, 256, 0, st>>>(spins, lay, state_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// bool[E, N] -> packed; lanes run over sites so the byte reads coalesce
__global__ void k_pack_states(uint32_t* __restrict__ spins, Layout L,
                              const uint8_t* __restrict__ states, uint64_t E) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(i / L.nvars);
        const uint64_t n = i - (uint64_t)w * L.nvars;
        uint32_t word = 0;
        for (int b = 0; b < 32; ++b) {
            const uint64_t e = (uint64_t)w * 32 + b;
            if (e < E && states[e * L.nvars + n]) word |= 1u << b;
        }
        spins[site_word_base(L, n) + w] = word;
    }
}

int launch_pack_states(uint32_t* spins, const Layout& lay, const uint8_t* states_dev, uint64_t E,
                       cudaStream_t st) {
    k_pack_states<<<148 * 8, 256, 0, st>>>(spins, lay, states_dev, E);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// K8: packed -> bool[E, N].  A warp takes 128 consecutive sites of one replica word: lane l
// holds sites 4l..4l+3 and writes one 32-bit store (4 bools) per experiment, so each warp
// store covers 128 contiguous bytes of one output row.
__global__ void __launch_bounds__(256)
k_unpack_states(const uint32_t* __restrict__ spins, Layout L, uint8_t* __restrict__ out,
                uint64_t E, uint64_t out_stride) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t chunks = (L.nvars + 127) / 128;
    const bool vec_ok = (L.nvars % 4 == 0) && (out_stride % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
    for (uint64_t item = warp; item < chunks * L.W; item += nwarps) {
        const uint64_t chunk = item / L.W;
        const uint32_t w = (uint32_t)(item - chunk * L.W);
        const uint64_t n0 = chunk * 128 + 4 * lane;
        uint32_t word[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            word[k] = (n0 + k < L.nvars) ? spins[site_word_base(L, n0 + k) + w] : 0u;
        const uint64_t e0 = (uint64_t)w * 32;
        const int nb = (int)(E - e0 < 32 ? E - e0 : 32);
        if (vec_ok) {
            if (n0 < L.nvars)
                for (int b = 0; b < nb; ++b) {
                    const uint32_t v = ((word[0] >> b) & 1u) | (((word[1] >> b) & 1u) << 8) |
                                       (((word[2] >> b) & 1u) << 16) |
                                       (((word[3] >> b) & 1u) << 24);
                    *reinterpret_cast<uint32_t*>(out + (e0 + b) * out_stride + n0) = v;
                }
        } else {
            for (int b = 0; b < nb; ++b)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (n0 + k < L.nvars)
                        out[(e0 + b) * out_stride + n0 + k] = (uint8_t)((word[k] >> b) & 1u);
        }
    }
}

int launch_unpack_states(const uint32_t* spins, const Layout& lay, uint8_t* out_dev, uint64_t E,
                         uint64_t out_stride, cudaStream_t st) {
    k_unpack_states<<<148 * 8, 256, 0, st>>>(spins, lay, out_dev, E, out_stride);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_export_natural(const uint32_t* __restrict__ spins, Layout L,
                                 uint32_t* __restrict__ out) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        out[i] = spins[site_word_base(L, n) + w];
    }
}

__global__ void k_import_natural(uint32_t* __restrict__ spins, Layout L,
                                 const uint32_t* __restrict__ in) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        spins[site_word_base(L, n) + w] = in[i];
    }
}

int launch_import_natural(uint32_t* spins, const Layout& lay, const uint32_t* in_dev, cudaStream_t st) {
    k_import_natural<<<148 * 8, 256, 0, st>>>(spins, lay, in_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_export_natural(const uint32_t* spins, const Layout& lay, uint32_t* out_dev,
                          cudaStream_t st) {
    k_export_natural<<<148 * 8, 256, 0, st>>>(spins, lay, out_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// K1: replay of the reference's (site, uniform) sequence; one thread per experiment, f64,
// no fused multiply-add so every rounding matches the CPU restatement
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_replay(ReplayArgs a) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    uint8_t* st = a.states + e * a.N;
    const uint32_t* sites = a.sites + e * a.A;
    const double* u = a.u + e * a.A;
    unsigned int amb = 0;
    for (uint64_t t = 0; t < a.A; ++t) {
        const uint32_t site = sites[t];
        const uint8_t cur = st[site];
        double de = 0.0;
        for (uint64_t k = a.row[site]; k < a.row[site + 1]; ++k) {
            const double coupling = (cur == st[a.nbr[k]]) ? 1.0 : -1.0;
            de = __dadd_rn(de, __dmul_rn(__dmul_rn(-2.0, a.jv[k]), coupling));
        }
        de = __dadd_rn(de, __dmul_rn(__dmul_rn(2.0, a.bias[site]), cur ? 1.0 : -1.0));
        bool flip = true;
        if (de > 0.0) {
            const double chance = exp(__dmul_rn(-a.beta, de));
            const double uu = u[t];
            flip = uu < chance;
            // device exp and the host libm may differ in the last place: refuse to certify a
            // decision that close to the threshold instead of guessing
            if (fabs(uu - chance) <= chance * 4.0e-15) ++amb;
        }
        if (flip) st[site] = cur ^ 1;
    }
    // GraphState::get_energy order: per site sum_adj(J*coupling/2), then + bias term
    double acc = 0.0;
    for (uint64_t i = 0; i < a.N; ++i) {
        double total = 0.0;
        const uint8_t si = st[i];
        for (uint64_t k = a.row[i]; k < a.row[i + 1]; ++k) {
            const double coupling = (si == st[a.nbr[k]]) ? 1.0 : -1.0;
            total = __dadd_rn(total, __dmul_rn(a.jv[k], coupling) / 2.0);
        }
        const double bias_e = si ? -a.bias[i] : a.bias[i];
        acc = __dadd_rn(__dadd_rn(acc, total), bias_e);
    }
    a.energies[e] = acc;
    if (amb) atomicAdd(a.ambiguous, amb);
}

int launch_replay(const ReplayArgs& a, cudaStream_t st) {
    const unsigned g = (unsigned)((a.E + 127) / 128);
    k_replay<<<g ? g : 1, 128, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ising

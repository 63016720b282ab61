// sm_100a kernels: one large 2D lattice bit-packed along x in row strips (K7).
#include "sweep_rows.cuh"

namespace ising {

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

// V consecutive words of a row per thread (128-bit loads when V == 4; needs Wr % V == 0).  The
// site update is the row walk's (sweep_rows.cuh: update_site): Philox rounds 1-3 shared over the
// V words and both calls (counter = (row, colour << 30 | word, sweep, call), i.e. the words of a
// thread differ in counter word 1 only), threshold planes selected with IMAD, rare ties deferred.
template <int K, int ROUNDS, int V>
__global__ void __launch_bounds__(256, 3)
k_strip_phase(uint32_t* __restrict__ spins, const __grid_constant__ StripGeom g, uint32_t c, uint32_t sweep,
              const __grid_constant__ PhiloxKeys pk, uint32_t antiferro, const __grid_constant__ MscMux mx,
              uint32_t r_begin, uint32_t r_count) {
    const uint32_t groups = g.Wr / V;
    const uint32_t o = 1u - c;
    VCount<1> unused[V];
    // block = (x over the word groups of a row, y over rows): storage rows [r_begin, r_begin + r_count)
    for (uint32_t rr = blockIdx.y * blockDim.y + threadIdx.y; rr < r_count; rr += gridDim.y * blockDim.y)
    for (uint32_t jg = blockIdx.x * blockDim.x + threadIdx.x; jg < groups; jg += gridDim.x * blockDim.x) {
        const uint32_t j = jg * V;
        const uint32_t r = r_begin + rr;
        const uint32_t y = strip_global_row(g, r);
        const uint32_t p = (y + c) & 1u;
        uint32_t s[V], n[4][V];
        uint32_t* own = spins + strip_off(g, c, r, j);
        const uint32_t* orow = spins + strip_off(g, o, r, 0);
        load_words<V>(own, s);
        load_words<V>(orow + j, n[0]);
        load_words<V>(orow + j - g.Wr, n[2]);
        load_words<V>(orow + j + g.Wr, n[3]);
        // the x neighbour one bit over: funnel shift with carry from the adjacent word
        const uint32_t edge = p ? orow[j + V == g.Wr ? 0 : j + V] : orow[j == 0 ? g.Wr - 1 : j - 1];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (p) n[1][v] = __funnelshift_r(n[0][v], v + 1 < V ? n[0][v + 1 < V ? v + 1 : v] : edge, 1);
            else n[1][v] = __funnelshift_l(v > 0 ? n[0][v > 0 ? v - 1 : 0] : edge, n[0][v], 1);
        }
        const uint32_t m[4] = {antiferro, antiferro, antiferro, antiferro};
        update_site<2, K, ROUNDS, V, false>(s, n, m, y, (c << 30) | j, sweep, pk, mx, unused);
        store_words<V>(own, s);
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
    const uint32_t gy_cap = (device_sms() * 32u + gx - 1) / gx;
    if (gy > gy_cap) gy = gy_cap;
    const dim3 grid(gx, gy, 1);
#define STRIP_LAUNCH(KK, RR)                                                                     \
    k_strip_phase<KK, RR, V><<<grid, block, 0, st>>>(a.spins, a.g, a.colour, a.sweep,              \
                                                     philox_round_keys(a.key0, a.key1), a.antiferro, \
                                                     make_mux(a.th), a.r_begin, a.r_count)
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
    k_strip_init_random<<<device_sms() * 8, 256, 0, st>>>(spins, g, key0, key1);
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
    k_strip_observables<<<device_sms() * 4, 256, 0, st>>>(spins, g, antiferro, acc);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_strip_unpack(const uint32_t* __restrict__ spins, StripGeom g,
                               uint8_t* __restrict__ out, uint32_t l0, uint32_t nrows) {
    const uint64_t Lx = 64ull * g.Wr;
    const uint64_t total = (uint64_t)nrows * Lx;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / Lx) + l0 + g.ghost;
        const uint32_t x = (uint32_t)(i % Lx);
        const uint32_t y = g.row0 + r - g.ghost;
        const uint32_t c = (x + y) & 1u, xh = x >> 1;
        out[i] = (uint8_t)((spins[strip_off(g, c, r, xh >> 5)] >> (xh & 31u)) & 1u);
    }
}

int launch_strip_unpack(const uint32_t* spins, const StripGeom& g, uint8_t* out_dev, cudaStream_t st,
                        uint32_t l0, uint32_t nrows) {
    if (nrows == 0xFFFFFFFFu) nrows = g.rows - l0;
    k_strip_unpack<<<device_sms() * 8, 256, 0, st>>>(spins, g, out_dev, l0, nrows);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ising

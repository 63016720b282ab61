// TEST INFRASTRUCTURE (CPU suite only).  Runs the library's strip kernel (k_strip_phase of
// csrc/strip.cu: BASELINE config 5, one large 2D lattice bit-packed along x in row strips) from
// its own source on the host (cuda_on_host.h) for tests/test_device_source_on_host.py, which
// compares the result with oracle/msc_mirror.c (msc_mirror_single).  Block and grid shape are
// strip_phase_dispatch()'s; the thresholds restate fill_thresholds() of api_sim.cu for 2D.
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include "prepared/strip_phase_kernel.cuh"
#include "prepared/strip_aux_kernels.cuh"
#include "prepared/strip_fused_kernel.cuh"

using namespace ising;

namespace {

void thresholds2d(double jabs, double beta, int K, MscThresholds* th) {
    memset(th, 0, sizeof *th);
    for (int c = 0; c < 2; ++c) {
        const double scaled = ldexp(exp(-beta * 4.0 * (c + 1) * jabs), K + 32);
        const uint64_t tmax = (1ull << (K + 32)) - 1;
        uint64_t T;
        if (!(scaled >= 0.0)) T = 0;
        else if (scaled >= (double)tmax) T = tmax;
        else T = (uint64_t)floor(scaled);
        for (int pl = 0; pl < K; ++pl) th->plane[c][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[c] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

template <int K, int ROUNDS, int V>
void run(uint32_t* spins, const StripGeom& g, uint32_t c, uint32_t sweep, uint64_t seed, uint32_t antiferro,
         const MscThresholds& th, uint32_t r_begin, uint32_t r_count, uint32_t gy_cap) {
    const uint32_t groups = g.Wr / V;
    const uint32_t bx = groups >= 256 ? 256 : pow2_ceil(groups);
    const dim3 block(bx, 256 / bx, 1);
    uint32_t gx = (groups + bx - 1) / bx;
    if (gx > 64) gx = 64;
    uint32_t gy = (r_count + block.y - 1) / block.y;
    if (gy > gy_cap) gy = gy_cap;                 // the kernel strides over the rows by the grid
    emu::launch_v(k_strip_phase<K, ROUNDS, V>, dim3(gx, gy, 1), block, 0, spins, g, c, sweep,
                  philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32)), antiferro, make_mux(th), r_begin, r_count);
}

}  // namespace

// One colour phase on the storage rows [r_begin, r_begin + r_count) of spins[2][rows + 2 ghost][Wr].
extern "C" int emu_strip_phase(uint32_t* spins, uint32_t Wr, uint32_t rows, uint32_t row0, uint32_t Ly,
                               uint32_t ghost, uint32_t colour, uint32_t sweep, uint64_t seed, uint32_t antiferro,
                               double beta, double jabs, int K, int rounds, uint32_t r_begin, uint32_t r_count,
                               uint32_t gy_cap) {
    if (Wr == 0 || r_count == 0) return 0;
    StripGeom g{Wr, rows, row0, Ly, ghost};
    MscThresholds th;
    thresholds2d(jabs, beta, K, &th);
#define GO(KK, RR)                                                                                      \
    do {                                                                                                \
        if (Wr % 4 == 0) run<KK, RR, 4>(spins, g, colour, sweep, seed, antiferro, th, r_begin, r_count, gy_cap);      \
        else if (Wr % 2 == 0) run<KK, RR, 2>(spins, g, colour, sweep, seed, antiferro, th, r_begin, r_count, gy_cap); \
        else run<KK, RR, 1>(spins, g, colour, sweep, seed, antiferro, th, r_begin, r_count, gy_cap);                  \
        return 0;                                                                                       \
    } while (0)
    if (K == 6 && rounds == 7) GO(6, 7);
    if (K == 5 && rounds == 7) GO(5, 7);
    if (K == 7 && rounds == 10) GO(7, 10);
#undef GO
    return -2;
}

// the strip's Philox initial state, its observables (satisfied bonds, up spins: acc[2]) and its
// conversion to bool rows: k_strip_init_random / k_strip_observables / k_strip_unpack
extern "C" void emu_strip_init_random(uint32_t* spins, uint32_t Wr, uint32_t rows, uint32_t row0, uint32_t Ly,
                                      uint32_t ghost, uint64_t seed, unsigned blocks) {
    emu::launch_v(k_strip_init_random, dim3(blocks), dim3(256), 0, spins, StripGeom{Wr, rows, row0, Ly, ghost},
                  (uint32_t)seed, (uint32_t)(seed >> 32));
}

extern "C" void emu_strip_observables(const uint32_t* spins, uint32_t Wr, uint32_t rows, uint32_t row0, uint32_t Ly,
                                      uint32_t ghost, uint32_t antiferro, unsigned long long* acc, unsigned blocks) {
    emu::launch_v(k_strip_observables, dim3(blocks), dim3(256), 0, spins, StripGeom{Wr, rows, row0, Ly, ghost}, antiferro, acc);
}

extern "C" void emu_strip_unpack(const uint32_t* spins, uint32_t Wr, uint32_t rows, uint32_t row0, uint32_t Ly,
                                 uint32_t ghost, uint8_t* out, uint32_t l0, uint32_t nrows, unsigned blocks) {
    emu::launch_v(k_strip_unpack, dim3(blocks), dim3(256), 0, spins, StripGeom{Wr, rows, row0, Ly, ghost}, out, l0, nrows);
}

// Both colour phases of a sweep in one out-of-place pass (k_strip_sweep_fused, the opt-in variant whose
// rows are loaded by the warps): colour 0 on storage rows [r0, r0 + n0), colour 1 on [r0 + 1, r0 + n0 - 1),
// src -> dst; block shape as launch_strip_sweep_fused() chooses it.  Returns 0, or 1 where that launcher declines.
extern "C" int emu_strip_fused(const uint32_t* src, uint32_t* dst, uint32_t Wr, uint32_t rows, uint32_t row0, uint32_t Ly,
                               uint32_t ghost, uint32_t sweep, uint64_t seed, uint32_t antiferro, double beta, double jabs,
                               uint32_t r0, uint32_t n0, uint32_t nbands) {
    constexpr int V = 4;
    if (Wr % V || Wr / V > 256u || n0 < 4) return 1;
    const uint32_t groups = Wr / V;
    const uint32_t bx = pow2_ceil(groups) < 32u ? 32u : pow2_ceil(groups);
    const uint32_t by = 256u / bx;
    const size_t smem = (size_t)by * 4u * Wr * sizeof(uint32_t);
    if (nbands > n0 / 2) nbands = n0 / 2;
    if (nbands < 1) return 1;
    MscThresholds th;
    thresholds2d(jabs, beta, 6, &th);
    emu::launch_v(k_strip_sweep_fused<6, kDefaultRounds, V>, dim3((nbands + by - 1) / by), dim3(bx, by, 1), smem, src, dst,
                  StripGeom{Wr, rows, row0, Ly, ghost}, sweep, philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32)),
                  antiferro, make_mux(th), r0, n0, nbands);
    return 0;
}

// One colour phase of a checkerboard sweep (K2 / K3), shared by the per-phase, cooperative and
// cluster kernels.
#pragma once
#include "msc_device.cuh"

namespace ising {

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
#ifndef ISING_ACC_MIN_BLOCKS
#define ISING_ACC_MIN_BLOCKS 2
#endif

// ACC: this phase also accumulates the post-flip satisfied-bond count of every replica into
// nsat[] (used for the second colour: its sites see every bond once, so after the phase
// nsat[e] is the total of experiment e and E = |J| (n_bonds - 2 nsat), lattice.rs:454).
// GRID2D: one row per block, (y, z) = 2D block index; otherwise blocks walk the rows with stride
// row_step (persistent launch).
// FOLD (small lattices, one thread-block cluster): threadIdx.y = (row within the block, xh), by_row
// threads per row, so that one block works on blockDim.y / by_row rows at a time.
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool GRID2D, bool PERBETA = false,
          bool FOLD = false>
__device__ __forceinline__ void sweep_colour_phase(
    uint32_t* __restrict__ own, const uint32_t* __restrict__ oth, const uint32_t* __restrict__ jm,
    const Layout& L, uint32_t c, uint32_t sweep, const PhiloxKeys& pk, uint32_t gw0,
    uint32_t antiferro, const MscThresholds& th, unsigned long long* __restrict__ nsat,
    uint32_t row_step, uint32_t step_y, uint32_t step_z, uint32_t* sm,
    const uint32_t* __restrict__ tplane = nullptr, const uint32_t* __restrict__ tlow = nullptr,
    uint32_t by_row = 0) {
    // FOLD + ACC: the launcher guarantees fewer than SW_MAX_ITEMS sites per thread, so the
    // (block-wide) mid-loop flush below is never taken although trip counts differ per thread
    static_assert(!(FOLD && GRID2D), "folded rows: persistent row walk only");
    constexpr int kUnrollV = ISING_SWEEP_UNROLL_V;
    uint32_t ty_row = 0, xh_t = threadIdx.y, xh_step = blockDim.y, rpb = 1;
    if constexpr (FOLD) {
        ty_row = threadIdx.y / by_row;
        xh_t = threadIdx.y - ty_row * by_row;
        xh_step = by_row;
        rpb = blockDim.y / by_row;
    }
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
            row = blockIdx.x * rpb + ty_row;
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
            for (uint32_t xh0 = 0; xh0 < Lxh; xh0 += xh_step) {
                const uint32_t xh = xh0 + xh_t;
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

}  // namespace ising

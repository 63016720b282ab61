// sm_100a kernels: checkerboard sweeps on square / cubic tori (K2), fused energy accumulation
// (K3), cooperative multi-sweep launch, n_sat of a configuration.
#include "sweep_phase.cuh"

namespace ising {

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
                uint32_t g = device_sms() * ISING_ACC_MIN_BLOCKS;
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
    uint32_t g = device_sms() * ISING_ACC_MIN_BLOCKS;
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
    if (persistent && g > device_sms() * 8u) g = device_sms() * 8u;
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
    static const bool v1 = getenv("ISING_SWEEP_V1") != nullptr;  // A/B knob: round-1 launch shape
    if (!v1 && !a.tplane && a.planes == 6) {
        int rc = 0;
        if (a.lay.kind == ISING_KIND_STENCIL3D) rc = launch_sweep_rows_3d(a, st);
        else if (a.lay.kind == ISING_KIND_STENCIL2D) rc = launch_sweep_rows_2d(a, st);
        if (rc != 0) return rc;
    }
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
    if (a.planes != 6 || a.rounds != kDefaultRounds) return 0;
    return hist ? coop_launch<DIM, PMJ, 6, kDefaultRounds, V, true>(a, th_dev, nsweeps, hist, cw, st)
                : coop_launch<DIM, PMJ, 6, kDefaultRounds, V, false>(a, th_dev, nsweeps, nullptr, cw, st);
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
    uint64_t g = (uint64_t)device_sms() * 2;  // persistent; counters are reduced every NS_MAX_ITEMS sites
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

}  // namespace ising

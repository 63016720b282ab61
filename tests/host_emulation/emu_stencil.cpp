// TEST INFRASTRUCTURE (CPU suite only).  Runs the remaining checkerboard sweep kernels from their
// own source on the host (cuda_on_host.h) for tests/test_device_source_on_host.py:
//   k_sweep_stencil / k_sweep_stencil_perbeta (csrc/sweep_stencil.cu)  one launch per colour phase:
//       the path of non-default planes / rounds and of per-replica betas (tempering on lattices)
//   k_sweep_stencil_coop                                              whole chunks of sweeps behind a grid barrier
//   k_sweep_stencil_cluster (csrc/sweep_cluster.cu)                    whole chunks inside one thread-block
//       cluster: what small lattices (BASELINE config 1 in production mode, tempering on small
//       lattices) run
//   k_nsat_stencil                                                    get_energy of a configuration
// Block shapes: the library's own stencil_block_shape() (taken from the source text) and the
// arithmetic of cluster_launch_n() restated; thresholds as fill_thresholds() of api_sim.cu.
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include <vector>

namespace ising { unsigned device_sms(); }
#include "prepared/stencil_kernels.cuh"
#include "prepared/cluster_kernel.cuh"

namespace ising {
static unsigned g_sms = 2;
unsigned device_sms() { return g_sms; }
}  // namespace ising

using namespace ising;

namespace {

void thresholds(int dim, double jabs, double beta, int K, MscThresholds* th) {
    memset(th, 0, sizeof *th);
    for (int c = 0; c < dim; ++c) {
        const double scaled = ldexp(exp(-beta * 4.0 * (c + 1) * jabs), K + 32);
        const uint64_t tmax = (1ull << (K + 32)) - 1;
        const uint64_t T = !(scaled >= 0.0) ? 0 : (scaled >= (double)tmax ? tmax : (uint64_t)floor(scaled));
        for (int pl = 0; pl < K; ++pl) th->plane[c][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[c] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

Layout make_layout(int dim, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint32_t W) {
    Layout L;
    memset(&L, 0, sizeof L);
    L.kind = dim == 3 ? ISING_KIND_STENCIL3D : ISING_KIND_STENCIL2D;
    L.Lx = Lx; L.Ly = Ly; L.Lz = dim == 3 ? Lz : 1; L.Lxh = Lx / 2; L.rows = L.Ly * L.Lz; L.W = W;
    L.nvars = (uint64_t)Lx * L.Ly * L.Lz;
    L.halfN = L.nvars / 2;
    return L;
}

struct Run {
    Layout L;
    uint32_t* spins;
    const uint32_t* jmask;
    uint32_t antiferro, sweep, gw0;
    uint64_t seed;
    PhiloxKeys pk() const { return philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32)); }
};

// sweep_launch_phase() of sweep_stencil.cu
template <int DIM, bool PMJ, int K, int ROUNDS, int V>
void phase(const Run& r, uint32_t c, bool acc, const MscThresholds& th, const uint32_t* tplane, const uint32_t* tlow,
           unsigned long long* nsat) {
    const Layout& L = r.L;
    dim3 grid, block;
    stencil_block_shape(L, V, &grid, &block, false);
    const size_t csz = (size_t)L.halfN * L.W, jsz = (size_t)2 * DIM * L.halfN;
    uint32_t* own = r.spins + c * csz;
    const uint32_t* oth = r.spins + (1 - c) * csz;
    const uint32_t* jm = r.jmask ? r.jmask + c * jsz : nullptr;
    if (!acc) {
        const dim3 grid2(L.Ly, L.Lz, 1);
        if (tplane) {
            if constexpr (K == 6)
                emu::launch_v(k_sweep_stencil_perbeta<DIM, PMJ, ROUNDS, V, false>, grid2, block, 0, own, oth, jm, L, c,
                              r.sweep, r.pk(), r.gw0, r.antiferro, tplane, tlow, (unsigned long long*)nullptr, L.rows, 0u, 0u);
        } else {
            emu::launch_v(k_sweep_stencil<DIM, PMJ, K, ROUNDS, V, false>, grid2, block, 0, own, oth, jm, L, c, r.sweep,
                          r.pk(), r.gw0, r.antiferro, th, (unsigned long long*)nullptr, L.rows, 0u, 0u);
        }
        return;
    }
    if (block.y < (unsigned)V) block.y = V;
    uint32_t g = device_sms() * ISING_ACC_MIN_BLOCKS;
    if (g > L.rows) g = L.rows;
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
    const size_t smem = (size_t)planes * nthreads * sizeof(uint32_t);
    if (tplane) {
        if constexpr (K == 6)
            emu::launch_v(k_sweep_stencil_perbeta<DIM, PMJ, ROUNDS, V, true>, dim3(g), block, smem, own, oth, jm, L, c,
                          r.sweep, r.pk(), r.gw0, r.antiferro, tplane, tlow, nsat, g, g % L.Ly, g / L.Ly);
    } else {
        emu::launch_v(k_sweep_stencil<DIM, PMJ, K, ROUNDS, V, true>, dim3(g), block, smem, own, oth, jm, L, c, r.sweep,
                      r.pk(), r.gw0, r.antiferro, th, nsat, g, g % L.Ly, g / L.Ly);
    }
}

// coop_launch() of sweep_stencil.cu
template <int DIM, bool PMJ, int V, bool ACC>
void coop(const Run& r, const MscThresholds* th_table, uint32_t nsweeps, unsigned long long* hist, uint32_t cw,
          uint32_t resident) {
    dim3 grid, block;
    stencil_block_shape(r.L, V, &grid, &block, false);
    if (ACC && block.y < (unsigned)V) block.y = V;
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
    const size_t smem = ACC ? (size_t)planes * nthreads * sizeof(uint32_t) : 0;
    uint32_t g = r.L.rows;
    if (g > resident) g = resident;
    emu::launch_v<true>(k_sweep_stencil_coop<DIM, PMJ, 6, kDefaultRounds, V, ACC>, dim3(g), block, smem, r.spins, r.jmask,
                        r.L, r.sweep, nsweeps, r.pk(), r.gw0, r.antiferro, th_table, hist, cw);
}

// cluster_launch_n() of sweep_cluster.cu; returns 0 when that launcher would decline the shape
template <int DIM, bool PMJ, int V, bool ACC, bool PERBETA>
int cluster(const Run& r, const MscThresholds* th_table, uint32_t nsweeps, unsigned long long* hist, uint32_t cw,
            const uint32_t* tplane, const uint32_t* tlow, uint32_t ncta) {
    const Layout& L = r.L;
    const uint32_t groups = L.W / V;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    uint32_t by_row = pow2_ceil(L.Lxh);
    if (wx * by_row > 256) by_row = 256 / wx;
    const uint32_t per_row = wx * by_row;
    const uint64_t want = ((uint64_t)L.rows * per_row + ncta - 1) / ncta;
    uint32_t threads = pow2_ceil((uint32_t)(want > 256 ? 256 : want));
    if (threads < per_row) threads = per_row;
    if (threads < 32) threads = 32;
    const uint32_t rpb = threads / per_row;
    if (rpb == 0) return 0;
    uint32_t g = (L.rows + rpb - 1) / rpb;
    if (g > ncta) g = ncta;
    const uint32_t row_step = g * rpb;
    const dim3 block(wx, by_row * rpb, 1);
    size_t smem = 0;
    if (ACC) {
        const uint32_t items = ((L.rows + row_step - 1) / row_step) * ((L.Lxh + by_row - 1) / by_row);
        if (items >= (uint32_t)SW_MAX_ITEMS || block.y < (unsigned)V) return 0;
        const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
        smem = (size_t)planes * block.x * block.y * sizeof(uint32_t);
    }
    emu::launch_v<true>(k_sweep_stencil_cluster<DIM, PMJ, 6, kDefaultRounds, V, ACC, PERBETA>, dim3(g), block, smem,
                        r.spins, r.jmask, L, r.sweep, nsweeps, r.pk(), r.gw0, r.antiferro, th_table, by_row, row_step,
                        row_step % L.Ly, row_step / L.Ly, hist, cw, tplane, tlow);
    return 1;
}

template <int DIM, bool PMJ, int V>
void nsat(const Run& r, unsigned long long* out) {
    dim3 grid, block;
    stencil_block_shape(r.L, V, &grid, &block, false);
    if (block.y < (unsigned)V) block.y = V;
    uint64_t g = (uint64_t)device_sms() * 2;
    if (g > r.L.rows) g = r.L.rows;
    const int nthreads = block.x * block.y;
    const int planes = NS_NP * V > NS_NR ? NS_NP * V : NS_NR;
    emu::launch_v(k_nsat_stencil<DIM, PMJ, V>, dim3((unsigned)g), block, (size_t)planes * nthreads * 4, (const uint32_t*)r.spins,
                  r.jmask, r.L, r.antiferro, out);
}

// run `body` with the compile-time (DIM, PMJ, V) that match the run-time values
#define FOR_SHAPE(dim, pmj, V, CALL)                                         \
    do {                                                                     \
        if (dim == 3 && pmj && V == 4) { CALL(3, true, 4); }                 \
        else if (dim == 3 && pmj && V == 1) { CALL(3, true, 1); }            \
        else if (dim == 3 && !pmj && V == 2) { CALL(3, false, 2); }          \
        else if (dim == 2 && !pmj && V == 4) { CALL(2, false, 4); }          \
        else if (dim == 2 && !pmj && V == 1) { CALL(2, false, 1); }          \
        else if (dim == 2 && pmj && V == 2) { CALL(2, true, 2); }            \
        else return -9;                                                      \
    } while (0)

}  // namespace

// mode: 0 one colour phase (acc = add the post-flip satisfied-bond counts to hist[W * 32]);
//       1 cooperative chunk of nsweeps (betas[nsweeps]; acc: hist[nsweeps][cw]);
//       2 cluster chunk (tplane == NULL: betas[nsweeps], acc: hist[nsweeps][cw];
//                        tplane != NULL: per-replica tables, acc: hist[W * 32] of the LAST sweep)
//       3 count only
// jmask: [2][2 dim][halfN] bond masks or NULL.  Returns 0 (ran), 1 (the launcher declines this shape), < 0 error.
extern "C" int emu_stencil(int mode, int dim, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint32_t W, int V,
                           const uint32_t* jmask, uint32_t antiferro, uint32_t* spins, uint32_t colour, uint64_t seed,
                           uint32_t sweep, uint32_t nsweeps, uint32_t gw0, int K, int rounds, const double* betas,
                           double jabs, const uint32_t* tplane, const uint32_t* tlow, int acc,
                           unsigned long long* hist, uint32_t cw, uint32_t sms_or_ncta) {
    if (W % V) return -1;
    Run r{make_layout(dim, Lx, Ly, Lz, W), spins, jmask, antiferro, sweep, gw0, seed};
    const bool pmj = jmask != nullptr;
    g_sms = sms_or_ncta;
    std::vector<MscThresholds> tab(nsweeps ? nsweeps : 1);
    for (uint32_t t = 0; t < nsweeps; ++t) thresholds(dim, jabs, betas[t], K, &tab[t]);
    if (mode == 0) {
#define CALL(D, P, VV)                                                                                    \
        if (K == 6 && rounds == 7) phase<D, P, 6, 7, VV>(r, colour, acc, tab[0], tplane, tlow, hist);      \
        else if (K == 5 && rounds == 7 && !tplane) phase<D, P, 5, 7, VV>(r, colour, acc, tab[0], tplane, tlow, hist);  \
        else if (K == 7 && rounds == 10 && !tplane) phase<D, P, 7, 10, VV>(r, colour, acc, tab[0], tplane, tlow, hist); \
        else return -2
        FOR_SHAPE(dim, pmj, V, CALL);
#undef CALL
        return 0;
    }
    if ((mode == 1 || mode == 2) && (K != 6 || rounds != 7)) return -2;
    if (mode == 1) {
#define CALL(D, P, VV) if (acc) coop<D, P, VV, true>(r, tab.data(), nsweeps, hist, cw, sms_or_ncta); \
                       else coop<D, P, VV, false>(r, tab.data(), nsweeps, nullptr, cw, sms_or_ncta)
        FOR_SHAPE(dim, pmj, V, CALL);
#undef CALL
        return 0;
    }
    if (mode == 2) {
        int rc = 0;
#define CALL(D, P, VV)                                                                                               \
        if (tplane) rc = acc ? cluster<D, P, VV, true, true>(r, tab.data(), nsweeps, hist, cw, tplane, tlow, sms_or_ncta)  \
                             : cluster<D, P, VV, false, true>(r, tab.data(), nsweeps, nullptr, cw, tplane, tlow, sms_or_ncta); \
        else rc = acc ? cluster<D, P, VV, true, false>(r, tab.data(), nsweeps, hist, cw, nullptr, nullptr, sms_or_ncta)    \
                      : cluster<D, P, VV, false, false>(r, tab.data(), nsweeps, nullptr, cw, nullptr, nullptr, sms_or_ncta)
        FOR_SHAPE(dim, pmj, V, CALL);
#undef CALL
        return rc == 1 ? 0 : 1;
    }
    if (mode == 3) {
#define CALL(D, P, VV) nsat<D, P, VV>(r, hist)
        FOR_SHAPE(dim, pmj, V, CALL);
#undef CALL
        return 0;
    }
    return -3;
}

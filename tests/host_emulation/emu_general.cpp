// TEST INFRASTRUCTURE (CPU suite only).  Runs the library's general-graph sweep kernel
// (k_sweep_general of csrc/sweep_general.cu: BASELINE config 4 and every graph Lattice takes that
// is not a torus) from its own source on the host (cuda_on_host.h) for
// tests/test_device_source_on_host.py, which compares the result with oracle/msc_mirror.c.
// The block shape is gen_launch()'s; the thresholds restate fill_gen_thresholds() of api_sim.cu.
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include "prepared/sweep_general_kernel.cuh"

using namespace ising;

namespace {

uint64_t threshold64(double beta, double de, int K) {
    const double scaled = ldexp(exp(-beta * de), K + 32);
    const uint64_t tmax = (1ull << (K + 32)) - 1;
    if (!(scaled >= 0.0)) return 0;
    if (scaled >= (double)tmax) return tmax;
    return (uint64_t)floor(scaled);
}

void fill_gen_thresholds(double jabs, double beta, int K, uint32_t deg, GenThresholds* th) {
    memset(th, 0, sizeof *th);
    const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
    for (uint32_t j = 0; j < ncls && j < (uint32_t)GEN_MAX_CLS; ++j) {
        const int cls = 2 * (int)(cmin + j) - (int)deg;
        const uint64_t T = threshold64(beta, 2.0 * jabs * (double)cls, K);
        for (int pl = 0; pl < K; ++pl) th->plane[j][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[j] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

template <int K, int ROUNDS, int DEG, int V>
void run(uint32_t* spins, const GenGroup& g, uint32_t W, uint32_t sweep, uint64_t seed, uint32_t gw0,
         const GenThresholds& th, const GenTables& tab, unsigned max_blocks) {
    const uint32_t groups = W / V;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)g.count + block.y - 1) / block.y;
    if (blocks > max_blocks) blocks = max_blocks;      // the kernel strides over the sites by the grid
    const dim3 grid((unsigned)blocks);
    const PhiloxKeys pk = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    if (tab.plane) emu::launch_v(k_sweep_general<K, ROUNDS, true, DEG, V>, grid, block, 0, spins, g, W, sweep, pk, gw0, th, tab);
    else emu::launch_v(k_sweep_general<K, ROUNDS, false, DEG, V>, grid, block, 0, spins, g, W, sweep, pk, gw0, th, tab);
}

}  // namespace

// One (colour, degree) group of a colour-class sweep on spins[nvars][W] (natural site order).
//   sites[count], nbr[deg][count], anti[count]: the group in the library's ELL form
//   tplane / tlow: per-replica threshold tables (GenTables layout, kernels.h) or NULL: uniform beta
//   specialise: use the kernel compiled for this degree (3, 4, 6) as the launcher does for K = 6, 7 rounds
extern "C" int emu_general_group(uint32_t* spins, uint32_t W, int V, const uint32_t* sites, const uint32_t* nbr,
                                 const uint32_t* anti, uint32_t count, uint32_t deg, uint32_t sweep,
                                 uint64_t seed, uint32_t gw0, int K, int rounds, double beta, double jabs,
                                 const uint32_t* tplane, const uint32_t* tlow, int specialise,
                                 unsigned max_blocks) {
    if (count == 0) return 0;
    if (deg > (uint32_t)GEN_MAX_DEG || W % V || (V != 1 && V != 2)) return -1;
    GenGroup g{sites, nbr, anti, count, deg};
    GenThresholds th;
    fill_gen_thresholds(jabs, beta, K, deg, &th);
    if (tplane) memset(&th, 0, sizeof th);
    GenTables tab{tplane, tlow};
#define GO(KK, RR, DD)                                                                     \
    do {                                                                                   \
        if (V == 2) run<KK, RR, DD, 2>(spins, g, W, sweep, seed, gw0, th, tab, max_blocks); \
        else run<KK, RR, DD, 1>(spins, g, W, sweep, seed, gw0, th, tab, max_blocks);        \
        return 0;                                                                          \
    } while (0)
    if (K == 6 && rounds == 7) {
        if (specialise && deg == 3) GO(6, 7, 3);
        if (specialise && deg == 4) GO(6, 7, 4);
        if (specialise && deg == 6) GO(6, 7, 6);
        GO(6, 7, 0);
    }
    if (K == 5 && rounds == 7) GO(5, 7, 0);
    if (K == 7 && rounds == 10) GO(7, 10, 0);
    if (K == 6 && rounds == 10) GO(6, 10, 0);
#undef GO
    return -2;
}

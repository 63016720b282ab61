// TEST INFRASTRUCTURE (CPU suite only).  Runs the bit-sliced two-spin edge-move kernel
// (k_edge_general of csrc/moves.cu) from its own source on the host (cuda_on_host.h) for
// tests/test_device_source_on_host.py, which compares the result with oracle/msc_mirror.c
// (msc_mirror_moves).  The block shape is edge_general_launch()'s; the thresholds are those of a
// site whose degree is the pair's number of outer bonds (fill_gen_thresholds of api_sim.cu).
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include "prepared/edge_general_kernel.cuh"

using namespace ising;

namespace {

void fill_gen_thresholds(double jabs, double beta, int K, uint32_t deg, GenThresholds* th) {
    memset(th, 0, sizeof *th);
    const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
    for (uint32_t j = 0; j < ncls && j < (uint32_t)GEN_MAX_CLS; ++j) {
        const int cls = 2 * (int)(cmin + j) - (int)deg;
        const double scaled = ldexp(exp(-beta * 2.0 * jabs * (double)cls), K + 32);
        const uint64_t tmax = (1ull << (K + 32)) - 1;
        const uint64_t T = !(scaled >= 0.0) ? 0 : (scaled >= (double)tmax ? tmax : (uint64_t)floor(scaled));
        for (int pl = 0; pl < K; ++pl) th->plane[j][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[j] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

template <int K, int ROUNDS, int DEG>
void run(uint32_t* spins, const EdgeGroup& g, uint32_t W, uint32_t sweep, uint32_t pass, uint64_t seed, uint32_t gw0,
         const GenThresholds& th, unsigned max_blocks) {
    const bool v2 = W % 2 == 0;
    const uint32_t groups = v2 ? W / 2 : W;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)g.count + block.y - 1) / block.y;
    if (blocks > max_blocks) blocks = max_blocks;
    const PhiloxKeys pk = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    if (v2) emu::launch_v(k_edge_general<K, ROUNDS, 2, DEG>, dim3((unsigned)blocks), block, 0, spins, g, W, sweep, pass, pk, gw0, th);
    else emu::launch_v(k_edge_general<K, ROUNDS, 1, DEG>, dim3((unsigned)blocks), block, 0, spins, g, W, sweep, pass, pk, gw0, th);
}

}  // namespace

// One (class, outer degree) group of a pass of edge moves on spins[slots][W].
extern "C" int emu_edge_group(uint32_t* spins, uint32_t W, const uint32_t* sa, const uint32_t* sb, const uint32_t* eid,
                              const uint32_t* anti, const uint32_t* endp, const uint32_t* nbr, uint32_t count,
                              uint32_t deg, uint32_t sweep, uint32_t pass, uint64_t seed, uint32_t gw0, int K,
                              int rounds, double beta, double jabs, int specialise, unsigned max_blocks) {
    if (count == 0) return 0;
    if (deg > (uint32_t)GEN_MAX_DEG) return -1;
    EdgeGroup g{sa, sb, eid, anti, endp, nbr, count, deg};
    GenThresholds th;
    fill_gen_thresholds(jabs, beta, K, deg, &th);
#define GO(KK, RR, DD) do { run<KK, RR, DD>(spins, g, W, sweep, pass, seed, gw0, th, max_blocks); return 0; } while (0)
    if (K == 6 && rounds == 7) {
        if (specialise && deg == 4) GO(6, 7, 4);
        if (specialise && deg == 6) GO(6, 7, 6);
        if (specialise && deg == 10) GO(6, 7, 10);
        GO(6, 7, 0);
    }
    if (K == 5 && rounds == 10) GO(5, 10, 0);
    if (K == 7 && rounds == 7) GO(7, 7, 0);
#undef GO
    return -2;
}

// sm_100a kernel: whole chunks of sweeps of a SMALL lattice inside one thread-block cluster.
// A colour phase of BASELINE config 1 (32x32, 64 experiments) is 1024 site-words of work: any
// launch or grid-wide barrier costs more than the update itself.  Here up to 8 CTAs (one
// portable cluster) hold the lattice, every thread owns at most a few site-words, and the two
// colour phases of every sweep are separated by the hardware cluster barrier (~0.2 us) instead
// of a cooperative grid barrier (~2.5 us) or a kernel boundary.  Same decision rule and Philox
// stream as every other sweep kernel, so the results are bit-identical.
#include "sweep_phase.cuh"

namespace ising {

// ACC: the second colour phase of sweep t adds its post-flip satisfied-bond counts to
// nsat_hist[t * cw + e] (per-sweep energies, lattice.rs:454), reduced per CTA.
// PERBETA: every replica bit at its own inverse temperature (parallel tempering between two
// swap steps): thresholds come from the bit-sliced tables instead of th_table.
// ACC && PERBETA: tempering reads energies at the end of a chunk only - the LAST sweep's second
// phase adds its counts to nsat_hist[e] (the swap cycle's satisfied-bond counters), which saves
// the separate count pass between the sweeps and the swap step.
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool PERBETA>
__global__ void __launch_bounds__(256)
k_sweep_stencil_cluster(uint32_t* __restrict__ spins, const uint32_t* __restrict__ jmask, Layout L,
                        uint32_t sweep0, uint32_t nsweeps, PhiloxKeys pk, uint32_t gw0,
                        uint32_t antiferro, const MscThresholds* __restrict__ th_table,
                        uint32_t by_row, uint32_t row_step, uint32_t step_y, uint32_t step_z,
                        unsigned long long* __restrict__ nsat_hist, uint32_t cw,
                        const uint32_t* __restrict__ tplane, const uint32_t* __restrict__ tlow) {
    constexpr bool ACC_EVERY = ACC && !PERBETA, ACC_LAST = ACC && PERBETA;
    extern __shared__ uint32_t sm[];  // ACC: reduction scratch
    __shared__ MscThresholds th[2];  // this sweep's thresholds / the next sweep's, prefetched
    cg::cluster_group cluster = cg::this_cluster();
    const size_t csz = (size_t)L.halfN * L.W;
    const size_t jsz = (size_t)2 * DIM * L.halfN;
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    constexpr uint32_t TW = sizeof(MscThresholds) / 4;
    // programmatic dependent launch (no-ops on an ordinary launch): between two chunks of a tempering
    // run sits one small kernel (k_pt_cycle) - each is scheduled while its predecessor drains
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!PERBETA && tid < TW)
        reinterpret_cast<uint32_t*>(&th[0])[tid] = reinterpret_cast<const uint32_t*>(th_table)[tid];
    __syncthreads();
    for (uint32_t t = 0; t < nsweeps; ++t) {
        const MscThresholds& cur = th[t & 1u];
        sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, false, false, PERBETA, true>(
            spins, spins + csz, PMJ ? jmask : nullptr, L, 0u, sweep0 + t, pk, gw0, antiferro, cur,
            nullptr, row_step, step_y, step_z, nullptr, tplane, tlow, by_row);
        // the other buffer was last read before the previous barrier: refill it now, so that the
        // load overlaps the barrier (made visible to the block by the barrier itself)
        if (!PERBETA && t + 1 < nsweeps && tid < TW)
            reinterpret_cast<uint32_t*>(&th[(t + 1) & 1u])[tid] =
                reinterpret_cast<const uint32_t*>(th_table + t + 1)[tid];
        cluster.sync();  // release / acquire at cluster scope: the other colour is complete
        if (ACC_LAST && t + 1 == nsweeps)
            sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, ACC_LAST, false, PERBETA, true>(
                spins + csz, spins, PMJ ? jmask + jsz : nullptr, L, 1u, sweep0 + t, pk, gw0, antiferro, cur,
                nsat_hist, row_step, step_y, step_z, sm, tplane, tlow, by_row);
        else
            sweep_colour_phase<DIM, PMJ, K, ROUNDS, V, ACC_EVERY, false, PERBETA, true>(
                spins + csz, spins, PMJ ? jmask + jsz : nullptr, L, 1u, sweep0 + t, pk, gw0, antiferro, cur,
                ACC_EVERY ? nsat_hist + (size_t)t * cw : nullptr, row_step, step_y, step_z, sm, tplane, tlow,
                by_row);
        cluster.sync();
    }
}

static_assert(sizeof(MscThresholds) / 4 <= 32, "threshold block is staged by the first warp");

// ncta = CTAs of the cluster: 8 (portable) or 16 (opt-in size, when the GPC has room for it)
template <int DIM, bool PMJ, int ROUNDS, int V, bool ACC, bool PERBETA>
static int cluster_launch_n(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                            unsigned long long* hist, uint32_t cw, cudaStream_t st, uint32_t ncta) {
    const Layout& L = a.lay;
    const uint32_t groups = L.W / V;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    uint32_t by_row = pow2_ceil(L.Lxh);
    if (wx * by_row > 256) by_row = 256 / wx;
    const uint32_t per_row = wx * by_row;                              // threads working on one row
    const uint64_t want = ((uint64_t)L.rows * per_row + ncta - 1) / ncta;  // threads per CTA
    uint32_t threads = pow2_ceil((uint32_t)(want > 256 ? 256 : want));
    if (threads < per_row) threads = per_row;
    if (threads < 32) threads = 32;
    const uint32_t rpb = threads / per_row;                             // rows a block works on at a time
    if (rpb == 0) return 0;
    uint32_t g = (L.rows + rpb - 1) / rpb;
    if (g > ncta) g = ncta;
    const uint32_t row_step = g * rpb;
    const dim3 block(wx, by_row * rpb, 1);
    size_t smem = 0;
    if (ACC) {
        // the fused counters are flushed once per phase: every thread must stay below their capacity
        const uint32_t items = ((L.rows + row_step - 1) / row_step) * ((L.Lxh + by_row - 1) / by_row);
        if (items >= (uint32_t)SW_MAX_ITEMS || block.y < (unsigned)V) return 0;
        const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
        smem = (size_t)planes * block.x * block.y * sizeof(uint32_t);
    }
    auto kernel = k_sweep_stencil_cluster<DIM, PMJ, 6, ROUNDS, V, ACC, PERBETA>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g, 1, 1);
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    static const bool no_pdl = getenv("ISING_NO_PDL") != nullptr;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (PERBETA && !no_pdl) ? 2 : 1;   // tempering chunks alternate with k_pt_cycle
    if (g > 8) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
            return -1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1)
            return -1;
    }
    const cudaError_t e = cudaLaunchKernelEx(
        &cfg, kernel, a.spins, a.jmask, L, a.sweep, nsweeps, philox_round_keys(a.key0, a.key1), a.gw0,
        a.antiferro, th_dev, by_row, row_step, row_step % L.Ly, row_step / L.Ly, hist, cw, a.tplane,
        a.tlow);
    return e == cudaSuccess ? 1 : -1;
}

template <int DIM, bool PMJ, int ROUNDS, int V, bool ACC, bool PERBETA>
static int cluster_launch(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                          unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    static const int max_cta = getenv("ISING_CLUSTER_MAX") ? atoi(getenv("ISING_CLUSTER_MAX")) : 16;  // A/B knob
    // more than one word per thread of a portable cluster: try the 16-CTA cluster first
    if ((uint64_t)a.lay.halfN * a.lay.W > 2048 && max_cta >= 16) {
        const int rc = cluster_launch_n<DIM, PMJ, ROUNDS, V, ACC, PERBETA>(a, th_dev, nsweeps, hist, cw, st, 16);
        if (rc > 0) return rc;
        cudaGetLastError();
    }
    return cluster_launch_n<DIM, PMJ, ROUNDS, V, ACC, PERBETA>(a, th_dev, nsweeps, hist, cw, st,
                                                               max_cta >= 8 ? 8u : (uint32_t)(max_cta < 1 ? 1 : max_cta));
}

template <int DIM, bool PMJ, int V>
static int cluster_rounds(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                          unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    if (a.tplane && hist)   // tempering chunk that ends with the satisfied-bond counts of its last sweep
        return a.rounds == 7 ? cluster_launch<DIM, PMJ, 7, V, true, true>(a, th_dev, nsweeps, hist, cw, st)
                             : cluster_launch<DIM, PMJ, 10, V, true, true>(a, th_dev, nsweeps, hist, cw, st);
    if (a.tplane)
        return a.rounds == 7 ? cluster_launch<DIM, PMJ, 7, V, false, true>(a, th_dev, nsweeps, nullptr, cw, st)
                             : cluster_launch<DIM, PMJ, 10, V, false, true>(a, th_dev, nsweeps, nullptr, cw, st);
    if (hist)
        return a.rounds == 7 ? cluster_launch<DIM, PMJ, 7, V, true, false>(a, th_dev, nsweeps, hist, cw, st)
                             : cluster_launch<DIM, PMJ, 10, V, true, false>(a, th_dev, nsweeps, hist, cw, st);
    return a.rounds == 7 ? cluster_launch<DIM, PMJ, 7, V, false, false>(a, th_dev, nsweeps, nullptr, cw, st)
                         : cluster_launch<DIM, PMJ, 10, V, false, false>(a, th_dev, nsweeps, nullptr, cw, st);
}

// Largest lattice taken: 16384 site-words per colour (4 per thread of a 16-CTA cluster).
// Returns 1 if launched, 0 if this configuration is not handled here, -1 on a launch error.
int launch_sweeps_stencil_cluster(const SweepArgs& a, const MscThresholds* th_dev, uint32_t nsweeps,
                                  unsigned long long* hist, uint32_t cw, cudaStream_t st) {
    const Layout& L = a.lay;
    const bool d3 = L.kind == ISING_KIND_STENCIL3D;
    if (!d3 && L.kind != ISING_KIND_STENCIL2D) return 0;
    if (a.planes != 6 || a.nsat_out) return 0;
    const uint64_t words = (uint64_t)L.halfN * L.W;
    static const uint64_t max_words = getenv("ISING_CLUSTER_WORDS") ? strtoull(getenv("ISING_CLUSTER_WORDS"), nullptr, 10) : 16384;
    if (words > max_words) return 0;
    // with fused energies the per-CTA reduction dominates beyond 8192 words (measured: 12.9 vs
    // 12.5 us/sweep for 32^2 x 1024 experiments against one launch per colour phase)
    if (hist && words > 8192) return 0;
    const bool pmj = a.jmask != nullptr;
    // fewest words per thread that still gives every site-word its own thread (4096 threads)
    int V = 1;
    if (words > 4096 && L.W % 2 == 0) V = 2;
    if (words > 8192 && L.W % 4 == 0) V = 4;
#define CLUSTER_V(VV)                                                                            \
    (d3 ? (pmj ? cluster_rounds<3, true, VV>(a, th_dev, nsweeps, hist, cw, st)                    \
               : cluster_rounds<3, false, VV>(a, th_dev, nsweeps, hist, cw, st))                  \
        : (pmj ? cluster_rounds<2, true, VV>(a, th_dev, nsweeps, hist, cw, st)                    \
               : cluster_rounds<2, false, VV>(a, th_dev, nsweeps, hist, cw, st)))
    if (V == 4) return CLUSTER_V(4);
    if (V == 2) return CLUSTER_V(2);
    return CLUSTER_V(1);
#undef CLUSTER_V
}

}  // namespace ising

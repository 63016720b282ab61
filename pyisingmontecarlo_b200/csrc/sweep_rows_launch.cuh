// Host-side launcher of the persistent row-walk sweep (sweep_rows.cuh), included by one
// translation unit per lattice dimension (sweep_rows2d.cu / sweep_rows3d.cu) to keep the
// compile parallel.
#pragma once
#include "sweep_rows.cuh"

namespace ising {

// ---- persistent row walk (sweep_rows.cuh): the default for lattices that fill the GPU ----------
template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool MULTIROW>
static int rows_launch(RowsArgs& ra, dim3 block, int sms, cudaStream_t st) {
    auto kern = k_sweep_rows<DIM, PMJ, K, ROUNDS, V, ACC, MULTIROW>;
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * V > NS_NR ? SW_NP * V : NS_NR;
    const size_t smem = ACC ? (size_t)planes * nthreads * sizeof(uint32_t) : 0;
    static int per_sm = 0, per_sm_threads = 0;  // per instantiation
    if (per_sm == 0 || per_sm_threads != nthreads) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, nthreads, smem) != cudaSuccess || n < 1)
            return -1;
        per_sm = n;
        per_sm_threads = nthreads;
    }
    // one resident wave; every block gets a balanced, contiguous range of units
    uint32_t g = (uint32_t)(per_sm * sms);
    if (g > ra.units) g = ra.units;
    ra.uq = ra.units / g;
    ra.urem = ra.units % g;
    kern<<<dim3(g, 1, 1), block, smem, st>>>(ra);
    return 1;
}

static uint32_t log2_exact(uint32_t v) {
    uint32_t lg = 0;
    while ((1u << lg) < v) ++lg;
    return lg;
}

template <int DIM, bool PMJ, int K, int ROUNDS, int V>
static int rows_phase(const SweepArgs& a, cudaStream_t st, uint32_t c, bool acc) {
    const Layout& L = a.lay;
    const size_t csz = (size_t)L.halfN * L.W;
    if (csz / V > 0xFFFFFFFFull || L.nvars > 0xFFFFFFFFull) return 0;  // 32-bit element offsets
    RowsArgs ra;
    ra.own = a.spins + c * csz;
    ra.oth = a.spins + (1 - c) * csz;
    ra.jm8 = PMJ ? reinterpret_cast<const uint4*>(a.jmask8 + (size_t)c * L.halfN * 8) : nullptr;
    ra.Lx = L.Lx; ra.Ly = L.Ly; ra.Lz = L.Lz; ra.Lxh = L.Lxh; ra.W = L.W;
    ra.c = c; ra.sweep = a.sweep; ra.gw0 = a.gw0; ra.antiferro = a.antiferro;
    // thread decomposition: x = word group, y = (row within the unit, half-row position)
    const uint32_t groups = L.W / V;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    uint32_t by = 256 / wx;
    uint32_t bxh = pow2_ceil(L.Lxh);
    if (bxh > by) bxh = by;
    uint32_t nrs = by / bxh;
    while (nrs > 1 && L.Ly % nrs) nrs >>= 1;
    if ((uint32_t)ROWS_DESC_CHUNK < nrs) return 0;
    if (acc && bxh * nrs < (uint32_t)V) return 0;  // the block reduction wants >= V thread rows
    ra.bxh_log = log2_exact(bxh);
    ra.nrs_log = log2_exact(nrs);
    ra.ygroups = L.Ly / nrs;
    ra.xtiles = (L.Lxh + bxh - 1) / bxh;
    const uint32_t wtiles = (groups + wx - 1) / wx;
    const uint64_t units = (uint64_t)L.Lz * ra.ygroups * ra.xtiles * wtiles;
    if (units > 0x7FFFFFFFull) return 0;
    ra.units = (uint32_t)units;
    ra.nsat = acc ? a.nsat_out : nullptr;
    ra.pk = philox_round_keys(a.key0, a.key1);
    ra.mx = make_mux(a.th);
    const dim3 block(wx, bxh * nrs, 1);
    const int sms = a.sm_count > 0 ? a.sm_count : (int)device_sms();
    if (nrs > 1)
        return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, true>(ra, block, sms, st)
                   : rows_launch<DIM, PMJ, K, ROUNDS, V, false, true>(ra, block, sms, st);
    return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, false>(ra, block, sms, st)
               : rows_launch<DIM, PMJ, K, ROUNDS, V, false, false>(ra, block, sms, st);
}

template <int DIM, bool PMJ, int V>
static int rows_sweep(const SweepArgs& a, cudaStream_t st) {
    int n = 0;
    for (uint32_t c = 0; c < 2; ++c) {
        const bool acc = a.nsat_out != nullptr && c == 1;
        int rc;
        if (a.rounds == 7) rc = rows_phase<DIM, PMJ, 6, 7, V>(a, st, c, acc);
        else rc = rows_phase<DIM, PMJ, 6, 10, V>(a, st, c, acc);
        if (rc <= 0) return c == 0 ? rc : -1;
        n += rc;
    }
    return cudaGetLastError() == cudaSuccess ? n : -1;
}

// returns launches made, 0 when this configuration is not covered (caller uses the per-row
// launch), -1 on error
template <int DIM, int V>
static int rows_dispatch(const SweepArgs& a, cudaStream_t st) {
    const bool pmj = a.jmask != nullptr;
    if (pmj && !a.jmask8) return 0;
    return pmj ? rows_sweep<DIM, true, V>(a, st) : rows_sweep<DIM, false, V>(a, st);
}

template <int DIM>
static int launch_sweep_rows_dim(const SweepArgs& a, cudaStream_t st) {
    if (a.lay.W % 4 == 0) return rows_dispatch<DIM, 4>(a, st);
    if (a.lay.W % 2 == 0) return rows_dispatch<DIM, 2>(a, st);
    return rows_dispatch<DIM, 1>(a, st);
}


}  // namespace ising

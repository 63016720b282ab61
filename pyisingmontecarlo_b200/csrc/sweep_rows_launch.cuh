// Host-side launcher of the persistent row-walk sweep (sweep_rows.cuh), included by one
// translation unit per lattice dimension (sweep_rows2d.cu / sweep_rows3d.cu) to keep the
// compile parallel.
#pragma once
#include "sweep_rows.cuh"

namespace ising {

static uint32_t log2_exact(uint32_t v) {
    uint32_t lg = 0;
    while ((1u << lg) < v) ++lg;
    return lg;
}

// Thread decomposition of a colour phase: block = (wx word groups) x (bxh half-row positions x nrs
// rows); unit = nrs consecutive rows of one z plane x one (xh, word) tile.
struct RowsShape {
    uint32_t wx, bxh, nrs, xtiles, wtiles;
    uint64_t units;
};

static bool rows_shape(const Layout& L, uint32_t V, RowsShape* out, uint32_t threads = ISING_ROWS_THREADS) {
    const uint32_t groups = L.W / V;
    RowsShape s;
    s.wx = groups >= 32 ? 32 : pow2_ceil(groups);
    const uint32_t by = threads / s.wx;
    s.bxh = pow2_ceil(L.Lxh);
    if (s.bxh > by) s.bxh = by;
    s.nrs = by / s.bxh;
    while (s.nrs > 1 && L.Ly % s.nrs) s.nrs >>= 1;
    if ((uint32_t)ROWS_DESC_CHUNK < s.nrs) return false;
    s.xtiles = (L.Lxh + s.bxh - 1) / s.bxh;
    s.wtiles = (groups + s.wx - 1) / s.wx;
    s.units = (uint64_t)L.Lz * (L.Ly / s.nrs) * s.xtiles * s.wtiles;
    if (s.units > 0x7FFFFFFFull) return false;
    *out = s;
    return true;
}

// one resident wave of blocks; every block gets a balanced, contiguous range of units
static void rows_partition(RowsArgs& ra, int per_sm, int sms, uint32_t* grid) {
    uint32_t g = (uint32_t)(per_sm * sms);
    // A/B knob: a grid that fills only part of the resident slots lets the blocks of the next colour
    // phase (programmatic dependent launch) become resident and run their preamble while this
    // phase still computes
    static const int grid_pct = getenv("ISING_ROWS_GRID_PCT") ? atoi(getenv("ISING_ROWS_GRID_PCT")) : 100;
    if (grid_pct > 0 && grid_pct < 100) g = (uint32_t)((uint64_t)g * grid_pct / 100);
    if (g < 1) g = 1;
    if (g > ra.units) g = ra.units;
    ra.uq = ra.units / g;
    ra.urem = ra.units % g;
    *grid = g;
}

template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool MULTIROW, bool COUNT, bool SMALL = false>
static int rows_launch(RowsArgs& ra, const RowsShape& sh, int sms, cudaStream_t st) {
    void (*kern)(const RowsArgs);
    if constexpr (COUNT) kern = k_nsat_rows<DIM, PMJ, V, MULTIROW>;
    else kern = k_sweep_rows<DIM, PMJ, K, ROUNDS, V, ACC, MULTIROW, SMALL>;
    const dim3 block(sh.wx, sh.bxh * sh.nrs, 1);
    const int nthreads = block.x * block.y;
    constexpr int np = SMALL ? ROWS_SMALL_NP : SW_NP, nr = SMALL ? ROWS_SMALL_NR : (COUNT ? NS_NR : ROWS_NR);
    const int planes = np * V > nr ? np * V : nr;
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
    uint32_t g = 0;
    rows_partition(ra, per_sm, sms, &g);
    if (launch_pdl(kern, dim3(g, 1, 1), block, smem, st, ra) != cudaSuccess) return -1;
    return 1;
}

// ---- TMA-staged variant (opt-in, ISING_TMA=1): whole-row units, W % 4 == 0 ---------------------
template <int DIM, bool PMJ, int K, int ROUNDS, bool ACC>
static int rows_tma_launch(RowsTmaArgs& ta, dim3 block, size_t smem, int sms, cudaStream_t st) {
    auto kern = k_sweep_rows_tma<DIM, PMJ, K, ROUNDS, ACC>;
    const int nthreads = block.x * block.y;
    static int per_sm = 0, per_sm_threads = 0;
    static size_t per_sm_smem = 0;
    if (per_sm == 0 || per_sm_threads != nthreads || per_sm_smem != smem) {
        int n = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, nthreads, smem) != cudaSuccess || n < 1) {
            cudaGetLastError();
            return 0;
        }
        per_sm = n;
        per_sm_threads = nthreads;
        per_sm_smem = smem;
    }
    uint32_t g = 0;
    rows_partition(ta.r, per_sm, sms, &g);
    if (launch_pdl(kern, dim3(g, 1, 1), block, smem, st, ta) != cudaSuccess) return -1;
    return 1;
}

// returns 1 when launched, 0 when the shape does not fit the staged kernel (caller uses the
// direct-load row walk)
template <int DIM, bool PMJ, int K, int ROUNDS>
static int rows_tma_phase(const Layout& L, const RowsArgs& ra, const RowsShape& sh, bool acc, int sms,
                          cudaStream_t st) {
    if (L.W % 4 || sh.xtiles != 1 || sh.wtiles != 1) return 0;
    RowsTmaArgs ta;
    ta.r = ra;
    ta.row_bytes = L.Lxh * L.W * 4u;
    ta.jrow_bytes = PMJ ? L.Lxh * 32u : 0u;
    auto up128 = [](uint32_t v) { return (v + 127u) & ~127u; };
    uint32_t off = up128(sh.nrs * ta.row_bytes);
    ta.off_oth = off;
    off += up128((sh.nrs + 2u) * ta.row_bytes);
    ta.off_zm = off;
    if (DIM == 3) off += up128(sh.nrs * ta.row_bytes);
    ta.off_zp = off;
    if (DIM == 3) off += up128(sh.nrs * ta.row_bytes);
    ta.off_jm = off;
    if (PMJ) off += up128(sh.nrs * ta.jrow_bytes);
    ta.stage_bytes = off;
    const dim3 block(sh.wx, sh.bxh * sh.nrs, 1);
    const int nthreads = block.x * block.y;
    const int planes = SW_NP * 4 > NS_NR ? SW_NP * 4 : NS_NR;
    const size_t smem = 2 * (size_t)ta.stage_bytes + (acc ? (size_t)planes * nthreads * sizeof(uint32_t) : 0);
    if (smem > 200 * 1024 || nthreads % 32) return 0;
    return acc ? rows_tma_launch<DIM, PMJ, K, ROUNDS, true>(ta, block, smem, sms, st)
               : rows_tma_launch<DIM, PMJ, K, ROUNDS, false>(ta, block, smem, sms, st);
}

// mode: 0 = plain colour phase, 1 = colour phase + accumulation of the post-flip satisfied-bond
// counts into a.nsat_out, 2 = count only (no update).  Returns launches made, 0 when this
// configuration is not covered, -1 on error.
template <int DIM, bool PMJ, int K, int ROUNDS, int V>
static int rows_phase(const SweepArgs& a, cudaStream_t st, uint32_t c, int mode) {
    const Layout& L = a.lay;
    const size_t csz = (size_t)L.halfN * L.W;
    if (csz / V > 0xFFFFFFFFull || L.nvars > 0xFFFFFFFFull) return 0;  // 32-bit element offsets
    RowsShape sh;
    // few site groups per resident thread: the 128-thread shape (k_sweep_rows<..., SMALL>)
    static const bool no_small = getenv("ISING_ROWS_NO_SMALL") != nullptr;   // A/B knob
    const int sms0 = a.sm_count > 0 ? a.sm_count : (int)device_sms();
    static const bool tma_env = getenv("ISING_TMA") != nullptr;
    bool small = V == 4 && mode != 2 && !no_small && !tma_env && ISING_ROWS_THREADS == 256 &&
                 (uint64_t)L.halfN * (L.W / V) * 2u < (uint64_t)sms0 * 768u * (mode == 1 ? 5u : 3u);   // < 2.5 / 1.5 per thread
    if (small && (!rows_shape(L, V, &sh, ROWS_SMALL_THREADS) || sh.bxh * sh.nrs < (uint32_t)V)) small = false;
    if (!small && !rows_shape(L, V, &sh)) return 0;
    const bool acc = mode != 0;
    if (acc && sh.bxh * sh.nrs < (uint32_t)V) return 0;  // the block reduction wants >= V thread rows
    RowsArgs ra;
    ra.own = a.spins + c * csz;
    ra.oth = a.spins + (1 - c) * csz;
    ra.jm8 = PMJ ? reinterpret_cast<const uint4*>(a.jmask8 + (size_t)c * L.halfN * 8) : nullptr;
    ra.Lx = L.Lx; ra.Ly = L.Ly; ra.Lz = L.Lz; ra.Lxh = L.Lxh; ra.W = L.W;
    ra.c = c; ra.sweep = a.sweep; ra.gw0 = a.gw0; ra.antiferro = a.antiferro;
    ra.nsat = acc ? a.nsat_out : nullptr;
    ra.nsat_copies = a.nsat_copies ? a.nsat_copies : 1;
    ra.nsat_stride = a.nsat_stride;
    ra.pk = philox_round_keys(a.key0, a.key1);
    ra.mx = make_mux(a.th);
    ra.bxh_log = log2_exact(sh.bxh);
    ra.nrs_log = log2_exact(sh.nrs);
    ra.ygroups = L.Ly / sh.nrs;
    ra.xtiles = sh.xtiles;
    ra.units = (uint32_t)sh.units;
    ra.uq = ra.urem = 0;
    const int sms = a.sm_count > 0 ? a.sm_count : (int)device_sms();
    if (mode == 2)
        return sh.nrs > 1 ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, true, true>(ra, sh, sms, st)
                          : rows_launch<DIM, PMJ, K, ROUNDS, V, true, false, true>(ra, sh, sms, st);
#ifndef ISING_ROWS_NO_TMA
    static const bool use_tma = getenv("ISING_TMA") != nullptr;   // opt-in: measured slower
    if (V == 4 && use_tma) {
        const int rc = rows_tma_phase<DIM, PMJ, K, ROUNDS>(L, ra, sh, acc, sms, st);
        if (rc != 0) return rc;
    }
#endif
    if constexpr (V == 4) {
        if (small) {
            if (sh.nrs > 1)
                return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, true, false, true>(ra, sh, sms, st)
                           : rows_launch<DIM, PMJ, K, ROUNDS, V, false, true, false, true>(ra, sh, sms, st);
            return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, false, false, true>(ra, sh, sms, st)
                       : rows_launch<DIM, PMJ, K, ROUNDS, V, false, false, false, true>(ra, sh, sms, st);
        }
    }
    if (sh.nrs > 1)
        return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, true, false>(ra, sh, sms, st)
                   : rows_launch<DIM, PMJ, K, ROUNDS, V, false, true, false>(ra, sh, sms, st);
    return acc ? rows_launch<DIM, PMJ, K, ROUNDS, V, true, false, false>(ra, sh, sms, st)
               : rows_launch<DIM, PMJ, K, ROUNDS, V, false, false, false>(ra, sh, sms, st);
}

// Both colour phases of one sweep.  With a.nsat_out the per-experiment satisfied-bond count after
// the sweep is added to it, fused into the second phase.  ISING_ROWS_SPLIT_ACC=1 (A/B knob) runs
// two plain phases and a count-only pass instead: measured slower on BASELINE config 3 (92.9 vs
// 76.3 us per sweep) - re-reading the lattice through L2 costs more than the fused accumulation.
template <int DIM, bool PMJ, int V>
static int rows_sweep(const SweepArgs& a, cudaStream_t st) {
    if (a.rounds != kDefaultRounds) return 0;   // other round counts: the one-row-per-block launch
    static const int split_env = getenv("ISING_ROWS_SPLIT_ACC") ? atoi(getenv("ISING_ROWS_SPLIT_ACC")) : -1;
    const bool split = a.nsat_out != nullptr && (split_env >= 0 ? split_env != 0 : ISING_ROWS_SPLIT_ACC_DEFAULT != 0);
    int n = 0;
    for (uint32_t c = 0; c < 2; ++c) {
        const int mode = (a.nsat_out != nullptr && c == 1 && !split) ? 1 : 0;
        const int rc = rows_phase<DIM, PMJ, 6, kDefaultRounds, V>(a, st, c, mode);
        if (rc <= 0) return c == 0 ? rc : -1;
        n += rc;
    }
    if (split) {
        const int rc = rows_phase<DIM, PMJ, 6, kDefaultRounds, V>(a, st, 0u, 2);
        if (rc <= 0) return -1;
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

// TEST INFRASTRUCTURE (CPU suite only).  Runs the library's row-walk sweep kernel - the kernel
// that dominates BASELINE configs 2 and 3 - from its own source on the host (cuda_on_host.h) and
// exports one C function for tests/test_device_source_on_host.py, which compares the result with
// oracle/msc_mirror.c.  The argument set-up below restates rows_phase() of
// csrc/sweep_rows_launch.cuh (which cannot be included whole: it asks the CUDA runtime for
// occupancies); the block shape and the unit partition come from that header's own rows_shape() /
// rows_partition().
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include "prepared/sweep_rows_launch_shape.cuh"   // includes prepared/sweep_rows.cuh

using namespace ising;

namespace {

// fill_thresholds() of csrc/api_sim.cu: dE of the uphill classes = 4 (c + 1) |J|
void thresholds(int dim, double jabs, double beta, int K, MscThresholds* th) {
    memset(th, 0, sizeof *th);
    for (int c = 0; c < dim; ++c) {
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

template <int DIM, bool PMJ, int V, bool ACC, bool MULTIROW, bool COUNT, bool SMALL>
void run(RowsArgs& ra, const RowsShape& sh, int grid_blocks) {
    void (*kern)(const RowsArgs);
    if constexpr (COUNT) kern = k_nsat_rows<DIM, PMJ, V, MULTIROW>;
    else kern = k_sweep_rows<DIM, PMJ, 6, kDefaultRounds, V, ACC, MULTIROW, SMALL>;
    const dim3 block(sh.wx, sh.bxh * sh.nrs, 1);
    const int nthreads = block.x * block.y;
    constexpr int np = SMALL ? ROWS_SMALL_NP : SW_NP, nr = SMALL ? ROWS_SMALL_NR : (COUNT ? NS_NR : ROWS_NR);
    const int planes = np * V > nr ? np * V : nr;
    const size_t smem = ACC ? (size_t)planes * nthreads * sizeof(uint32_t) : 0;
    uint32_t g = 0;
    rows_partition(ra, grid_blocks, 1, &g);
    emu::launch(kern, dim3(g, 1, 1), block, smem, ra);
}

template <int DIM, bool PMJ, int V>
int phase(RowsArgs& ra, const RowsShape& sh, int mode, bool small, int grid_blocks) {
    const bool multirow = sh.nrs > 1;
    if (mode == 2) {
        if (multirow) run<DIM, PMJ, V, true, true, true, false>(ra, sh, grid_blocks);
        else run<DIM, PMJ, V, true, false, true, false>(ra, sh, grid_blocks);
        return 0;
    }
    const bool acc = mode == 1;
    if constexpr (V == 4) {
        if (small) {
            if (multirow) acc ? run<DIM, PMJ, V, true, true, false, true>(ra, sh, grid_blocks)
                              : run<DIM, PMJ, V, false, true, false, true>(ra, sh, grid_blocks);
            else acc ? run<DIM, PMJ, V, true, false, false, true>(ra, sh, grid_blocks)
                     : run<DIM, PMJ, V, false, false, false, true>(ra, sh, grid_blocks);
            return 0;
        }
    }
    if (small) return -3;
    if (multirow) acc ? run<DIM, PMJ, V, true, true, false, false>(ra, sh, grid_blocks)
                      : run<DIM, PMJ, V, false, true, false, false>(ra, sh, grid_blocks);
    else acc ? run<DIM, PMJ, V, true, false, false, false>(ra, sh, grid_blocks)
             : run<DIM, PMJ, V, false, false, false, false>(ra, sh, grid_blocks);
    return 0;
}

}  // namespace

// One colour phase of the row walk on spins[2][rows][Lxh][W] (the library's stencil layout).
//   mode 0 plain, 1 update + post-flip satisfied-bond counts into nsat[W * 32], 2 count only
//   jm8: [2][halfN][8] bond masks of a +-J lattice or NULL (then antiferro = all-ones iff J > 0)
//   small: the 128-thread shape (V = 4 only);  grid_blocks: blocks the units are dealt to
// returns 0, or a negative code when the library's launcher would not take this shape either
extern "C" int emu_rows_phase(int dim, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint32_t W, int V,
                              const uint32_t* jm8, uint32_t antiferro, uint32_t* spins, uint32_t colour,
                              uint64_t seed, uint32_t sweep, uint32_t gw0, double beta, double jabs,
                              int mode, unsigned long long* nsat, int small, int grid_blocks) {
    if ((dim != 2 && dim != 3) || W % V || (V != 1 && V != 2 && V != 4) || Lx % 2) return -1;
    if (dim == 2) Lz = 1;
    Layout L;
    memset(&L, 0, sizeof L);
    L.kind = dim == 3 ? ISING_KIND_STENCIL3D : ISING_KIND_STENCIL2D;
    L.Lx = Lx; L.Ly = Ly; L.Lz = Lz; L.Lxh = Lx / 2; L.rows = Ly * Lz; L.W = W;
    L.nvars = (uint64_t)Lx * Ly * Lz;
    L.halfN = L.nvars / 2;
    RowsShape sh;
    if (!rows_shape(L, (uint32_t)V, &sh, small ? (uint32_t)ROWS_SMALL_THREADS : (uint32_t)ISING_ROWS_THREADS)) return -2;
    if (mode != 0 && sh.bxh * sh.nrs < (uint32_t)V) return -2;
    const size_t csz = (size_t)L.halfN * W;
    MscThresholds th;
    thresholds(dim, jabs, beta, 6, &th);
    RowsArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.own = spins + colour * csz;
    ra.oth = spins + (1 - colour) * csz;
    ra.jm8 = jm8 ? reinterpret_cast<const uint4*>(jm8 + (size_t)colour * L.halfN * 8) : nullptr;
    ra.Lx = Lx; ra.Ly = Ly; ra.Lz = Lz; ra.Lxh = L.Lxh; ra.W = W;
    ra.c = colour; ra.sweep = sweep; ra.gw0 = gw0; ra.antiferro = antiferro;
    ra.nsat = mode != 0 ? nsat : nullptr;
    ra.nsat_copies = 1;
    ra.nsat_stride = W * 32;
    ra.pk = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    ra.mx = make_mux(th);
    ra.bxh_log = log2_exact(sh.bxh);
    ra.nrs_log = log2_exact(sh.nrs);
    ra.ygroups = Ly / sh.nrs;
    ra.xtiles = sh.xtiles;
    ra.units = (uint32_t)sh.units;
    const bool pmj = jm8 != nullptr;
#define DISPATCH(D, P, VV) return phase<D, P, VV>(ra, sh, mode, small != 0, grid_blocks)
    if (dim == 3) {
        if (pmj) { if (V == 4) DISPATCH(3, true, 4); if (V == 2) DISPATCH(3, true, 2); DISPATCH(3, true, 1); }
        if (V == 4) DISPATCH(3, false, 4); if (V == 2) DISPATCH(3, false, 2); DISPATCH(3, false, 1);
    }
    if (pmj) { if (V == 4) DISPATCH(2, true, 4); if (V == 2) DISPATCH(2, true, 2); DISPATCH(2, true, 1); }
    if (V == 4) DISPATCH(2, false, 4); if (V == 2) DISPATCH(2, false, 2); DISPATCH(2, false, 1);
#undef DISPATCH
}

// TEST INFRASTRUCTURE (CPU suite only).  The smaller kernels around the sweeps, from their own
// source on the host (cuda_on_host.h), for tests/test_device_source_on_host.py: the warp-voted
// threshold-table builders and the satisfied-bond count of a general graph (csrc/sweep_general.cu),
// the positional popcounts behind magnetisations and overlaps and the conversions of bond counts to
// energies (csrc/observables.cu).  Block / grid shapes as the launch_* wrappers choose them.
#include "cuda_on_host.h"

#include <string.h>

#include "prepared/general_aux_kernels.cuh"
#include "prepared/observables_kernels.cuh"

using namespace ising;

extern "C" void emu_build_tables(const unsigned long long* t64, const uint32_t* slot_of_replica, uint32_t W, int K,
                                 uint32_t* plane_out, uint32_t* low_out, int stencil, unsigned cap) {
    const uint32_t total = stencil ? W * 3 : (GEN_MAX_DEG + 1) * W * GEN_MAX_CLS;
    const uint32_t blocks = (total + 3) / 4;
    const dim3 grid(blocks < cap ? blocks : cap);
    if (stencil) emu::launch_v(k_build_tables_stencil, grid, dim3(128), 0, t64, slot_of_replica, W, K, plane_out, low_out);
    else emu::launch_v(k_build_tables, grid, dim3(128), 0, t64, slot_of_replica, W, K, plane_out, low_out);
}

extern "C" void emu_nsat_general(const uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row,
                                 const uint32_t* nbr, const uint8_t* anti, unsigned long long* nsat2, unsigned cap) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t g = (nvars + block.y - 1) / block.y;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    emu::launch_v(k_nsat_general, dim3((unsigned)g), block, 0, spins, nvars, W, row, nbr, anti, nsat2);
}

extern "C" void emu_count_up(const uint32_t* spins, uint64_t nsites, uint32_t W, unsigned long long* up, int pair,
                             unsigned cap) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t g = (nsites + block.y - 1) / block.y;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    emu::launch_v(k_count_up, dim3((unsigned)g), block, 0, spins, nsites, W, up, pair ? 1u : 0u);
}

extern "C" void emu_overlap_from_counts(const unsigned long long* dis, uint64_t P, uint64_t nsites, double* out,
                                        uint64_t stride, uint64_t off) {
    const unsigned g = (unsigned)((P + 255) / 256);
    emu::launch_v(k_overlap_from_counts, dim3(g ? g : 1), dim3(256), 0, dis, P, nsites, out, stride, off);
}

extern "C" void emu_energy_from_hist(const unsigned long long* hist, uint64_t E, uint64_t cw, uint64_t nt, double scale,
                                     uint64_t nbonds, int mult, double* out, uint32_t copies) {
    const unsigned g = (unsigned)((E * nt + 255) / 256);
    emu::launch_v(k_energy_from_hist, dim3(g ? g : 1), dim3(256), 0, hist, E, cw, nt, scale, nbonds, mult, out, copies);
}

extern "C" void emu_energy_from_nsat(const unsigned long long* nsat, uint64_t E, double scale, uint64_t nbonds, int mult,
                                     double* out, uint64_t estride, uint64_t eoff) {
    const unsigned g = (unsigned)((E + 255) / 256);
    emu::launch_v(k_energy_from_nsat, dim3(g ? g : 1), dim3(256), 0, nsat, E, scale, nbonds, mult, out, estride, eoff);
}

// TEST INFRASTRUCTURE (CPU suite only).  The float-field kernels - colour-class sweeps on real
// couplings and biases (k_sweep_real, k_energy_real of csrc/sweep_general.cu), the float two-spin
// edge moves and the worm moves (k_edge_moves, k_worm_moves of csrc/moves.cu) - from their own
// source on the host (cuda_on_host.h), for the statistical check against exact enumeration in
// tests/test_device_source_on_host.py (their gate on the GPU is statistical too: f32 fields and
// the fast exponential).
#include "cuda_on_host.h"

#include "prepared/real_kernels.cuh"
#include "prepared/float_moves_kernels.cuh"

using namespace ising;

namespace {
Layout natural_layout(uint64_t nvars, uint32_t W) {
    Layout L;
    memset(&L, 0, sizeof L);
    L.kind = ISING_KIND_GENERAL;
    L.nvars = nvars;
    L.W = W;
    return L;
}
dim3 site_block(uint32_t W) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    return dim3(wx, 256 / wx, 1);
}
}  // namespace

extern "C" void emu_sweep_real(uint32_t* spins, const uint32_t* sites, uint32_t count, const uint32_t* row,
                               const uint32_t* nbr, const float* jf, const float* biasf, uint32_t W, float beta,
                               uint32_t sweep, uint64_t seed, uint32_t gw0) {
    if (count == 0) return;
    RealSweepArgs a;
    memset(&a, 0, sizeof a);
    a.spins = spins; a.sites = sites; a.count = count; a.row = row; a.nbr = nbr; a.jf = jf; a.biasf = biasf;
    a.W = W; a.beta = beta; a.sweep = sweep; a.key0 = (uint32_t)seed; a.key1 = (uint32_t)(seed >> 32); a.gw0 = gw0;
    a.rounds = 7;
    const dim3 block = site_block(W);
    emu::launch_v(k_sweep_real<7>, dim3((count + block.y - 1) / block.y), block, 0, a);
}

extern "C" void emu_energy_real(const uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row,
                                const uint32_t* nbr, const double* jv, const double* bias, double* energies) {
    const uint32_t wx = W >= 32 ? 32 : pow2_ceil(W);
    const dim3 block(wx, 128 / wx, 1);
    uint64_t g = (nvars + block.y - 1) / block.y;
    if (g > 3) g = 3;
    emu::launch_v(k_energy_real, dim3((unsigned)g, (W + wx - 1) / wx, 1), block, 0, spins, nvars, W, row, nbr, jv, bias,
                  energies);
}

extern "C" void emu_edge_moves(uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row, const uint32_t* nbr,
                               const float* jf, const float* biasf, const uint32_t* ea, const uint32_t* eb,
                               const uint32_t* eid, const float* wrel, uint32_t count, float beta, uint32_t sweep,
                               uint64_t seed, uint32_t gw0, uint32_t pass) {
    if (count == 0) return;
    EdgeMoveArgs a;
    memset(&a, 0, sizeof a);
    a.spins = spins; a.lay = natural_layout(nvars, W); a.g = MoveGraph{row, nbr, jf, biasf};
    a.ea = ea; a.eb = eb; a.eid = eid; a.wrel = wrel; a.count = count; a.beta = beta;
    a.sweep = sweep; a.key0 = (uint32_t)seed; a.key1 = (uint32_t)(seed >> 32); a.gw0 = gw0; a.pass = pass; a.rounds = 7;
    const dim3 block = site_block(W);
    emu::launch_v(k_edge_moves<7>, dim3((count + block.y - 1) / block.y), block, 0, a);
}

extern "C" void emu_worm_moves(uint32_t* spins, uint64_t nvars, uint32_t W, const uint32_t* row, const uint32_t* nbr,
                               const float* jf, const float* biasf, uint64_t E, uint32_t nworms, uint32_t len, float beta,
                               uint32_t sweep, uint64_t seed) {
    WormArgs a;
    memset(&a, 0, sizeof a);
    a.spins = spins; a.lay = natural_layout(nvars, W); a.g = MoveGraph{row, nbr, jf, biasf};
    a.E = E; a.replica_offset = 0; a.nworms = nworms; a.worm0 = 0; a.len = len; a.beta = beta;
    a.sweep = sweep; a.key0 = (uint32_t)seed; a.key1 = (uint32_t)(seed >> 32); a.rounds = 7;
    emu::launch_v(k_worm_moves<7>, dim3((unsigned)((E + 127) / 128)), dim3(128), 0, a);
}

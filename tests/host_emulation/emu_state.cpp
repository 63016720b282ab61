// TEST INFRASTRUCTURE (CPU suite only).  Runs the kernels of csrc/state_io.cu from their own
// source on the host (cuda_on_host.h): the Philox initial state, bool <-> packed conversion in both
// spin layouts, and K1, the replay of the reference algorithm's (site, uniform) trace
// (north-star correctness check 1), for tests/test_device_source_on_host.py.  On the host
// k_replay's exp is libm's, the oracle's own, so no decision can be ambiguous here; what the
// test pins is the kernel's control flow and summation order.
#include "cuda_on_host.h"

#include <math.h>
#include <string.h>

#include "prepared/state_io_kernels.cuh"

using namespace ising;

namespace {
Layout make_layout(int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint64_t nvars, uint32_t W) {
    Layout L;
    memset(&L, 0, sizeof L);
    L.kind = kind;
    L.W = W;
    L.nvars = nvars;
    if (kind != ISING_KIND_GENERAL) {
        L.Lx = Lx; L.Ly = Ly; L.Lz = kind == ISING_KIND_STENCIL3D ? Lz : 1;
        L.Lxh = Lx / 2; L.rows = L.Ly * L.Lz;
        L.halfN = nvars / 2;
    }
    return L;
}
}  // namespace

extern "C" int emu_kind(int dim) {
    return dim == 0 ? ISING_KIND_GENERAL : (dim == 3 ? ISING_KIND_STENCIL3D : ISING_KIND_STENCIL2D);
}

extern "C" void emu_init_random(uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint64_t nvars,
                                uint32_t W, uint64_t seed, uint32_t gw0, unsigned blocks) {
    emu::launch_v(k_init_random, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W),
                  (uint32_t)seed, (uint32_t)(seed >> 32), gw0);
}

extern "C" void emu_pack_states(uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint64_t nvars,
                                uint32_t W, const uint8_t* states, uint64_t E, unsigned blocks) {
    emu::launch_v(k_pack_states, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W), states, E);
}

extern "C" void emu_unpack_states(const uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz,
                                  uint64_t nvars, uint32_t W, uint8_t* out, uint64_t E, uint64_t out_stride,
                                  unsigned blocks) {
    emu::launch_v(k_unpack_states, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W), out, E,
                  out_stride);
}

extern "C" void emu_init_broadcast(uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint64_t nvars,
                                   uint32_t W, const uint8_t* state, unsigned blocks) {
    emu::launch_v(k_init_broadcast, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W), state);
}

extern "C" unsigned emu_replay(uint64_t E, uint64_t N, uint64_t A, const uint64_t* row, const uint32_t* nbr,
                               const double* jv, const double* bias, const uint32_t* sites, const double* u,
                               uint8_t* states, double* energies, double beta) {
    unsigned int ambiguous = 0;
    ReplayArgs a{E, N, A, row, nbr, jv, bias, sites, u, states, energies, beta, &ambiguous};
    const unsigned g = (unsigned)((E + 127) / 128);
    emu::launch_v(k_replay, dim3(g ? g : 1), dim3(128), 0, a);
    return ambiguous;
}

// packed words between the sim's layout and natural site order (checkpoints, ising_sim_get_packed)
extern "C" void emu_export_natural(const uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz,
                                   uint64_t nvars, uint32_t W, uint32_t* out, unsigned blocks) {
    emu::launch_v(k_export_natural, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W), out);
}

extern "C" void emu_import_natural(uint32_t* spins, int kind, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint64_t nvars,
                                   uint32_t W, const uint32_t* in, unsigned blocks) {
    emu::launch_v(k_import_natural, dim3(blocks), dim3(256), 0, spins, make_layout(kind, Lx, Ly, Lz, nvars, W), in);
}

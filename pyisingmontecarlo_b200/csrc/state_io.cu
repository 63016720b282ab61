// sm_100a kernels: state initialisation, bool <-> packed conversion (K8), replay (K1).
#include "msc_device.cuh"

namespace ising {

unsigned device_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int v = cached[dev];
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 1;
        cached[dev] = v;   // benign race: every writer stores the same value
    }
    return (unsigned)v;
}


// ------------------------------------------------------------------------------------------
// state initialisation / import / export (not hot)
// ------------------------------------------------------------------------------------------
__global__ void k_init_random(uint32_t* __restrict__ spins, Layout L, uint32_t k0, uint32_t k1,
                              uint32_t gw0) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        const u32x4 r = philox4x32<10>((uint32_t)n, gw0 + w, 0u, TAG_INIT << 24, k0, k1);
        spins[site_word_base(L, n) + w] = r.x;
    }
}

int launch_init_random(uint32_t* spins, const Layout& lay, uint32_t key0, uint32_t key1,
                       uint32_t gw0, cudaStream_t st) {
    k_init_random<<<device_sms() * 8, 256, 0, st>>>(spins, lay, key0, key1, gw0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_init_broadcast(uint32_t* __restrict__ spins, Layout L,
                                 const uint8_t* __restrict__ state) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        spins[site_word_base(L, n) + w] = state[n] ? 0xFFFFFFFFu : 0u;
    }
}

int launch_init_broadcast(uint32_t* spins, const Layout& lay, const uint8_t* state_dev,
                          cudaStream_t st) {
    k_init_broadcast<<<device_sms() * 8, 256, 0, st>>>(spins, lay, state_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// bool[E, N] -> packed; lanes run over sites so the byte reads coalesce
__global__ void k_pack_states(uint32_t* __restrict__ spins, Layout L,
                              const uint8_t* __restrict__ states, uint64_t E) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(i / L.nvars);
        const uint64_t n = i - (uint64_t)w * L.nvars;
        uint32_t word = 0;
        for (int b = 0; b < 32; ++b) {
            const uint64_t e = (uint64_t)w * 32 + b;
            if (e < E && states[e * L.nvars + n]) word |= 1u << b;
        }
        spins[site_word_base(L, n) + w] = word;
    }
}

int launch_pack_states(uint32_t* spins, const Layout& lay, const uint8_t* states_dev, uint64_t E,
                       cudaStream_t st) {
    k_pack_states<<<device_sms() * 8, 256, 0, st>>>(spins, lay, states_dev, E);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// K8: packed -> bool[E, N].  A warp takes 128 consecutive sites of one replica word: lane l
// holds sites 4l..4l+3 and writes one 32-bit store (4 bools) per experiment, so each warp
// store covers 128 contiguous bytes of one output row.
__global__ void __launch_bounds__(256)
k_unpack_states(const uint32_t* __restrict__ spins, Layout L, uint8_t* __restrict__ out,
                uint64_t E, uint64_t out_stride) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t chunks = (L.nvars + 127) / 128;
    const bool vec_ok = (L.nvars % 4 == 0) && (out_stride % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
    for (uint64_t item = warp; item < chunks * L.W; item += nwarps) {
        const uint64_t chunk = item / L.W;
        const uint32_t w = (uint32_t)(item - chunk * L.W);
        const uint64_t n0 = chunk * 128 + 4 * lane;
        uint32_t word[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            word[k] = (n0 + k < L.nvars) ? spins[site_word_base(L, n0 + k) + w] : 0u;
        const uint64_t e0 = (uint64_t)w * 32;
        const int nb = (int)(E - e0 < 32 ? E - e0 : 32);
        if (vec_ok) {
            if (n0 < L.nvars)
                for (int b = 0; b < nb; ++b) {
                    const uint32_t v = ((word[0] >> b) & 1u) | (((word[1] >> b) & 1u) << 8) |
                                       (((word[2] >> b) & 1u) << 16) |
                                       (((word[3] >> b) & 1u) << 24);
                    *reinterpret_cast<uint32_t*>(out + (e0 + b) * out_stride + n0) = v;
                }
        } else {
            for (int b = 0; b < nb; ++b)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (n0 + k < L.nvars)
                        out[(e0 + b) * out_stride + n0 + k] = (uint8_t)((word[k] >> b) & 1u);
        }
    }
}

int launch_unpack_states(const uint32_t* spins, const Layout& lay, uint8_t* out_dev, uint64_t E,
                         uint64_t out_stride, cudaStream_t st) {
    k_unpack_states<<<device_sms() * 8, 256, 0, st>>>(spins, lay, out_dev, E, out_stride);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_export_natural(const uint32_t* __restrict__ spins, Layout L,
                                 uint32_t* __restrict__ out) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        out[i] = spins[site_word_base(L, n) + w];
    }
}

__global__ void k_import_natural(uint32_t* __restrict__ spins, Layout L,
                                 const uint32_t* __restrict__ in) {
    const uint64_t total = L.nvars * L.W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = i / L.W;
        const uint32_t w = (uint32_t)(i - n * L.W);
        spins[site_word_base(L, n) + w] = in[i];
    }
}

int launch_import_natural(uint32_t* spins, const Layout& lay, const uint32_t* in_dev, cudaStream_t st) {
    k_import_natural<<<device_sms() * 8, 256, 0, st>>>(spins, lay, in_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_export_natural(const uint32_t* spins, const Layout& lay, uint32_t* out_dev,
                          cudaStream_t st) {
    k_export_natural<<<device_sms() * 8, 256, 0, st>>>(spins, lay, out_dev);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// K1: replay of the reference's (site, uniform) sequence; one thread per experiment, f64,
// no fused multiply-add so every rounding matches the CPU restatement
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_replay(ReplayArgs a) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    uint8_t* st = a.states + e * a.N;
    const uint32_t* sites = a.sites + e * a.A;
    const double* u = a.u + e * a.A;
    unsigned int amb = 0;
    for (uint64_t t = 0; t < a.A; ++t) {
        const uint32_t site = sites[t];
        const uint8_t cur = st[site];
        double de = 0.0;
        for (uint64_t k = a.row[site]; k < a.row[site + 1]; ++k) {
            const double coupling = (cur == st[a.nbr[k]]) ? 1.0 : -1.0;
            de = __dadd_rn(de, __dmul_rn(__dmul_rn(-2.0, a.jv[k]), coupling));
        }
        de = __dadd_rn(de, __dmul_rn(__dmul_rn(2.0, a.bias[site]), cur ? 1.0 : -1.0));
        bool flip = true;
        if (de > 0.0) {
            const double chance = exp(__dmul_rn(-a.beta, de));
            const double uu = u[t];
            flip = uu < chance;
            // device exp and the host libm may differ in the last place: refuse to certify a
            // decision that close to the threshold instead of guessing
            if (fabs(uu - chance) <= chance * 4.0e-15) ++amb;
        }
        if (flip) st[site] = cur ^ 1;
    }
    // GraphState::get_energy order: per site sum_adj(J*coupling/2), then + bias term
    double acc = 0.0;
    for (uint64_t i = 0; i < a.N; ++i) {
        double total = 0.0;
        const uint8_t si = st[i];
        for (uint64_t k = a.row[i]; k < a.row[i + 1]; ++k) {
            const double coupling = (si == st[a.nbr[k]]) ? 1.0 : -1.0;
            total = __dadd_rn(total, __dmul_rn(a.jv[k], coupling) / 2.0);
        }
        const double bias_e = si ? -a.bias[i] : a.bias[i];
        acc = __dadd_rn(__dadd_rn(acc, total), bias_e);
    }
    a.energies[e] = acc;
    if (amb) atomicAdd(a.ambiguous, amb);
}

int launch_replay(const ReplayArgs& a, cudaStream_t st) {
    const unsigned g = (unsigned)((a.E + 127) / 128);
    k_replay<<<g ? g : 1, 128, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ising

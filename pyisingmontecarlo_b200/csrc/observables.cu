// sm_100a kernels: per-experiment integer observables (positional popcount) and their conversion to
// energies / magnetisations / overlaps.
#include "msc_device.cuh"

namespace ising {

// up-spin count over all sites (any layout: the sum runs over every stored site word)
__global__ void __launch_bounds__(256)
k_count_up(const uint32_t* __restrict__ spins, uint64_t nsites, uint32_t W,
           unsigned long long* __restrict__ up, uint32_t pair) {
    __shared__ int sm[32 * 256];
    const int nthreads = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (uint32_t w0 = 0; w0 < W; w0 += blockDim.x) {
        for (int b = 0; b < 32; ++b) sm[b * nthreads + tid] = 0;
        const uint32_t w = w0 + threadIdx.x;
        VCount<VC_PLANES> vc;
        vc.clear();
        int pending = 0;
        if (w < W) {
            for (uint64_t n = (uint64_t)blockIdx.x * blockDim.y + threadIdx.y; n < nsites;
                 n += (uint64_t)gridDim.x * blockDim.y) {
                uint32_t x = spins[(size_t)n * W + w];
                // pair mode: bit 2p = experiments 2p and 2p+1 disagree on this site
                if (pair) x = (x ^ (x >> 1)) & 0x55555555u;
                vc.add1(x);
                if (++pending == VC_FLUSH_ADD1) {
                    vc.flush(sm, tid, nthreads);
                    pending = 0;
                }
            }
            vc.flush(sm, tid, nthreads);
        }
        block_reduce_counts(sm, up, w0, W);
    }
}

int launch_count_up(const uint32_t* spins, const Layout& lay, unsigned long long* up,
                    cudaStream_t st, bool pair) {
    const uint32_t wx = lay.W >= 32 ? 32 : pow2_ceil(lay.W);
    dim3 block(wx, 256 / wx, 1);
    uint64_t g = (lay.nvars + block.y - 1) / block.y;
    if (g > device_sms() * 8u) g = device_sms() * 8u;
    if (g == 0) g = 1;
    k_count_up<<<dim3((unsigned)g), block, 0, st>>>(spins, lay.nvars, lay.W, up, pair ? 1u : 0u);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// a ^= b over n words (flipped spins of a sweep: before ^ after)
__global__ void k_xor_words(uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        a[i] ^= b[i];
}

int launch_xor_words(uint32_t* a, const uint32_t* b, uint64_t n, cudaStream_t st) {
    uint64_t g = (n + 255) / 256;
    if (g > device_sms() * 8u) g = device_sms() * 8u;
    if (g == 0) g = 1;
    k_xor_words<<<(unsigned)g, 256, 0, st>>>(a, b, n);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// overlap of the experiment pair (2p, 2p+1) from the pair-mode counts: q = N - 2 * disagreements
__global__ void k_overlap_from_counts(const unsigned long long* __restrict__ dis, uint64_t P,
                                      uint64_t nsites, double* __restrict__ out, uint64_t stride,
                                      uint64_t off) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    out[p * stride + off] = (double)((long long)nsites - 2ll * (long long)dis[2 * p]);
}

int launch_overlap_from_counts(const unsigned long long* dis, uint64_t P, uint64_t nsites,
                               double* out_dev, uint64_t stride, uint64_t off, cudaStream_t st) {
    const unsigned g = (unsigned)((P + 255) / 256);
    k_overlap_from_counts<<<g ? g : 1, 256, 0, st>>>(dis, P, nsites, out_dev, stride, off);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void k_energy_from_nsat(const unsigned long long* __restrict__ nsat, uint64_t E,
                                   double scale, uint64_t nbonds, int mult,
                                   double* __restrict__ out, uint64_t estride, uint64_t eoff) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long v = (long long)nbonds - (long long)mult * (long long)nsat[e];
    out[e * estride + eoff] = scale * (double)v;
}

// energies[e * nt + t] = scale * (nbonds - 2 * hist[t * cw + e]) for a chunk of nt sweeps
__global__ void k_energy_from_hist(const unsigned long long* __restrict__ hist, uint64_t E,
                                   uint64_t cw, uint64_t nt, double scale, uint64_t nbonds,
                                   int mult, double* __restrict__ out, uint32_t copies) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * nt) return;
    const uint64_t e = i / nt, t = i - e * nt;
    unsigned long long n = 0;
    for (uint32_t c = 0; c < copies; ++c) n += hist[(t * copies + c) * cw + e];
    const long long v = (long long)nbonds - (long long)mult * (long long)n;
    out[i] = scale * (double)v;
}

int launch_energy_from_hist(const unsigned long long* hist, uint64_t E, uint64_t cw, uint64_t nt,
                            double scale, uint64_t nbonds, int mult, double* out_dev,
                            cudaStream_t st, uint32_t copies) {
    const uint64_t n = E * nt;
    const unsigned g = (unsigned)((n + 255) / 256);
    k_energy_from_hist<<<g ? g : 1, 256, 0, st>>>(hist, E, cw, nt, scale, nbonds, mult, out_dev, copies);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// out[e * estride + eoff] = in[e]   /   out[e * nt + t] = hist[t * cw + e]  (f64 energies)
__global__ void k_copy_strided_f64(const double* __restrict__ in, uint64_t E,
                                   double* __restrict__ out, uint64_t estride, uint64_t eoff) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) out[e * estride + eoff] = in[e];
}

__global__ void k_transpose_hist_f64(const double* __restrict__ hist, uint64_t E, uint64_t cw,
                                     uint64_t nt, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * nt) return;
    const uint64_t e = i / nt, t = i - e * nt;
    out[i] = hist[t * cw + e];
}

int launch_copy_strided_f64(const double* in, uint64_t E, double* out, uint64_t estride,
                            uint64_t eoff, cudaStream_t st) {
    const unsigned g = (unsigned)((E + 255) / 256);
    k_copy_strided_f64<<<g ? g : 1, 256, 0, st>>>(in, E, out, estride, eoff);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_transpose_hist_f64(const double* hist, uint64_t E, uint64_t cw, uint64_t nt, double* out,
                              cudaStream_t st) {
    const uint64_t n = E * nt;
    const unsigned g = (unsigned)((n + 255) / 256);
    k_transpose_hist_f64<<<g ? g : 1, 256, 0, st>>>(hist, E, cw, nt, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_energy_from_nsat(const unsigned long long* nsat, uint64_t E, double scale,
                            uint64_t nbonds, int mult, double* out_dev, uint64_t estride,
                            uint64_t eoff, cudaStream_t st) {
    const unsigned g = (unsigned)((E + 255) / 256);
    k_energy_from_nsat<<<g ? g : 1, 256, 0, st>>>(nsat, E, scale, nbonds, mult, out_dev, estride, eoff);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ising

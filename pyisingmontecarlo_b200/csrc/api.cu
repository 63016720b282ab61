// C ABI of libising_b200.so (include/ising_b200.h).  Host orchestration only: argument
// checks with the reference's error behaviour, device buffers, kernel launches, CUDA-event
// timing.  There is deliberately no CPU implementation behind these entry points: without a
// CUDA device every compute call fails with ISING_E_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ising_b200.h"
#include "graph.h"
#include "kernels.h"
#include "philox.h"

using namespace ising;

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
static thread_local std::string g_global_error;

struct ising_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int sm_count = 0;
    // grow-only device scratch (staging of outputs), so that repeated calls do not pay
    // cudaMalloc/cudaFree of hundreds of MB every time
    void* scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
    // second stream + events for the double-buffered device-to-host copies of the sampling path
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_filled[2] = {nullptr, nullptr}, ev_drained[2] = {nullptr, nullptr};
    // free list of device buffers released by destroyed sims: a stateless Lattice run creates
    // and destroys a sim per call, and cudaMalloc/cudaFree (device-wide synchronising, tens
    // of ms with large pinned regions mapped) must not be on that path
    std::vector<std::pair<void*, size_t>> free_bufs;
    size_t free_bytes = 0;
};

static cudaError_t ctx_buf_get(ising_ctx* ctx, size_t bytes, void** out) {
    bytes = std::max<size_t>(bytes, 256);
    for (size_t i = 0; i < ctx->free_bufs.size(); ++i)
        if (ctx->free_bufs[i].second >= bytes && ctx->free_bufs[i].second <= bytes + bytes / 4 + 4096) {
            *out = ctx->free_bufs[i].first;
            ctx->free_bytes -= ctx->free_bufs[i].second;
            ctx->free_bufs.erase(ctx->free_bufs.begin() + i);
            return cudaSuccess;
        }
    return cudaMalloc(out, bytes);
}


static void ctx_buf_put(ising_ctx* ctx, void* p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>(bytes, 256);
    const size_t cap = (size_t)8 << 30;
    if (ctx->free_bytes + bytes > cap || ctx->free_bufs.size() >= 64) {
        cudaFree(p);
        return;
    }
    ctx->free_bufs.emplace_back(p, bytes);
    ctx->free_bytes += bytes;
}

static cudaError_t ctx_scratch(ising_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        const size_t want = std::max<size_t>(bytes, 1 << 20);
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) return e;
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return cudaSuccess;
}

struct ising_graph {
    ising_ctx* ctx = nullptr;
    HostGraph h;
    // device copies
    uint32_t* d_jmask = nullptr;   // stencil +-J bond masks [2][2*dim][halfN]
    uint64_t* d_row = nullptr;     // CSR for replay / general kernels (uploaded on demand)
    uint32_t* d_nbr = nullptr;
    double* d_jv = nullptr;
    double* d_bias = nullptr;
    // general-graph sweep data (built on demand): colour x degree groups in ELL form
    bool gen_built = false;
    std::vector<GenGroup> gen_groups;
    std::vector<int> gen_group_color;
    int gen_ncolors = 0;
    uint32_t* d_gsites = nullptr;
    uint32_t* d_gnbr = nullptr;
    uint32_t* d_ganti = nullptr;
    uint32_t* d_row32 = nullptr;   // CSR for the energy kernel
    uint32_t* d_nbr32 = nullptr;
    uint8_t* d_anti8 = nullptr;
    // real couplings / biases: sites ordered by colour + float CSR values
    bool real_built = false;
    uint32_t* d_csites = nullptr;
    std::vector<uint32_t> color_off;   // ncolors + 1 offsets into d_csites
    float* d_jf = nullptr;
    float* d_biasf = nullptr;
};

struct ising_sim {
    ising_ctx* ctx = nullptr;
    const ising_graph* g = nullptr;
    uint64_t E = 0;
    uint64_t seed = 0;
    uint64_t replica_offset = 0;
    Layout lay{};
    uint32_t* d_spins = nullptr;
    unsigned long long* d_counts = nullptr;  // per-experiment integer accumulator [W*32]
    size_t spins_bytes = 0, counts_bytes = 0;
    uint64_t sweep_counter = 0;
    int planes = 6, rounds = 10;
    ising_sim_stats stats{};
    bool general = false;          // natural-order layout + colour/degree groups
    bool real = false;             // general layout, float local fields (real J / biases)
    // per-replica inverse temperatures (parallel tempering); general layout only
    bool perbeta = false;
    unsigned long long* d_t64 = nullptr;
    uint32_t* d_slot = nullptr;
    uint32_t* d_tplane = nullptr;
    uint32_t* d_tlow = nullptr;
};

static int fail(ising_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_global_error = buf;
    return code;
}

#define CUDA_TRY(ctx, call)                                                          \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess)                                                       \
            return fail((ctx), ISING_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
static cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" int ising_abi_version(void) { return ISING_ABI_VERSION; }

extern "C" const char* ising_last_error(const ising_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_global_error.c_str();
}

static int ctx_create_impl(int device, cudaStream_t external, bool use_external, ising_ctx** out);

extern "C" int ising_ctx_create(int device, ising_ctx** out) {
    return ctx_create_impl(device, nullptr, false, out);
}

// Same, but every launch and copy of this context goes to the caller's stream (e.g. torch's
// current stream): work is then ordered with the caller's own kernels and NCCL calls without
// host synchronisation.
extern "C" int ising_ctx_create_on_stream(int device, void* cuda_stream, ising_ctx** out) {
    return ctx_create_impl(device, (cudaStream_t)cuda_stream, true, out);
}

static int ctx_create_impl(int device, cudaStream_t external, bool use_external, ising_ctx** out) {
    if (!out) return fail(nullptr, ISING_E_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ISING_E_CUDA,
                    "no CUDA device available (%s); libising_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev)
        return fail(nullptr, ISING_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    CUDA_TRY(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, ISING_E_CUDA,
                    "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    std::unique_ptr<ising_ctx> ctx(new ising_ctx);
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (use_external) {
        ctx->stream = external;
        ctx->owns_stream = false;
    } else {
        CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    }
    CUDA_TRY(nullptr, cudaEventCreate(&ctx->ev0));
    CUDA_TRY(nullptr, cudaEventCreate(&ctx->ev1));
    *out = ctx.release();
    return ISING_OK;
}

extern "C" void ising_ctx_destroy(ising_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int b = 0; b < 2; ++b) {
        if (ctx->ev_filled[b]) cudaEventDestroy(ctx->ev_filled[b]);
        if (ctx->ev_drained[b]) cudaEventDestroy(ctx->ev_drained[b]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    for (void* p : ctx->scratch) cudaFree(p);
    for (auto& b : ctx->free_bufs) cudaFree(b.first);
    delete ctx;
}

extern "C" int ising_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(nullptr, ISING_E_INVALID, "out is NULL");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess)
        return fail(nullptr, ISING_E_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return ISING_OK;
}

extern "C" void ising_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// graph
// ------------------------------------------------------------------------------------------
static int upload_stencil_masks(ising_ctx* ctx, ising_graph* g) {
    const HostGraph& h = g->h;
    if (h.kind == ISING_KIND_GENERAL || h.uniform_sign) return ISING_OK;
    const int dim = h.kind == ISING_KIND_STENCIL3D ? 3 : 2;
    const uint64_t Lx = h.dims[0], Ly = h.dims[1], Lz = h.dims[2];
    const uint64_t Lxh = Lx / 2, rows = Ly * Lz, halfN = h.nvars / 2;
    std::vector<uint32_t> m((size_t)2 * 2 * dim * halfN);
    auto sgn = [&](uint64_t x, uint64_t y, uint64_t z, int d) -> uint32_t {
        const uint64_t n = x + Lx * (y + Ly * z);
        return ((h.fwd_sign[n] >> d) & 1) ? 0xFFFFFFFFu : 0u;
    };
    for (uint32_t c = 0; c < 2; ++c)
        for (uint64_t r = 0; r < rows; ++r) {
            const uint64_t z = r / Ly, y = r % Ly;
            const uint32_t p = (uint32_t)((y + z + c) & 1);
            for (uint64_t xh = 0; xh < Lxh; ++xh) {
                const uint64_t x = 2 * xh + p;
                const uint64_t xm = x == 0 ? Lx - 1 : x - 1;
                const uint64_t ym = y == 0 ? Ly - 1 : y - 1, zm = z == 0 ? Lz - 1 : z - 1;
                uint32_t* base = m.data() + (size_t)c * 2 * dim * halfN + r * Lxh + xh;
                // k = 0: neighbour stored at the same half-index (x+1 if p == 0 else x-1)
                // k = 1: the other x neighbour;  2: y-1  3: y+1  4: z-1  5: z+1
                const uint32_t jxp = sgn(x, y, z, 0), jxm = sgn(xm, y, z, 0);
                base[0 * halfN] = p == 0 ? jxp : jxm;
                base[1 * halfN] = p == 0 ? jxm : jxp;
                base[2 * halfN] = sgn(x, ym, z, 1);
                base[3 * halfN] = sgn(x, y, z, 1);
                if (dim == 3) {
                    base[4 * halfN] = sgn(x, y, zm, 2);
                    base[5 * halfN] = sgn(x, y, z, 2);
                }
            }
        }
    CUDA_TRY(ctx, dev_alloc(&g->d_jmask, m.size()));
    CUDA_TRY(ctx, cudaMemcpy(g->d_jmask, m.data(), m.size() * sizeof(uint32_t),
                             cudaMemcpyHostToDevice));
    return ISING_OK;
}

static int ensure_csr_on_device(ising_ctx* ctx, ising_graph* g) {
    if (g->d_row) return ISING_OK;
    g->h.build_csr();
    const HostGraph& h = g->h;
    CUDA_TRY(ctx, dev_alloc(&g->d_row, h.row.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_nbr, h.nbr.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_jv, h.jv.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_bias, h.nvars));
    CUDA_TRY(ctx, cudaMemcpy(g->d_row, h.row.data(), h.row.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_nbr, h.nbr.data(), h.nbr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_jv, h.jv.data(), h.jv.size() * 8, cudaMemcpyHostToDevice));
    std::vector<double> b(h.nvars, 0.0);
    if (h.has_bias) b = h.bias;
    CUDA_TRY(ctx, cudaMemcpy(g->d_bias, b.data(), h.nvars * 8, cudaMemcpyHostToDevice));
    return ISING_OK;
}

static int ensure_csr32_on_device(ising_ctx* ctx, ising_graph* g) {
    if (g->d_row32) return ISING_OK;
    HostGraph& h = g->h;
    h.build_csr();
    if (2 * h.nedges > 0xFFFFFFFFull) return fail(ctx, ISING_E_UNSUPPORTED, "too many edges");
    const uint64_t N = h.nvars;
    std::vector<uint32_t> row32(N + 1);
    for (uint64_t n = 0; n <= N; ++n) row32[n] = (uint32_t)h.row[n];
    std::vector<uint8_t> anti8(h.jv.size());
    for (size_t k = 0; k < h.jv.size(); ++k) anti8[k] = h.jv[k] > 0;
    CUDA_TRY(ctx, dev_alloc(&g->d_row32, row32.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_nbr32, h.nbr.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_anti8, anti8.size()));
    CUDA_TRY(ctx, cudaMemcpy(g->d_row32, row32.data(), row32.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_nbr32, h.nbr.data(), h.nbr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_anti8, anti8.data(), anti8.size(), cudaMemcpyHostToDevice));
    return ISING_OK;
}

// arbitrary real couplings and biases: colour-ordered site list + float CSR values
static int ensure_real_on_device(ising_ctx* ctx, ising_graph* g) {
    if (g->real_built) return ISING_OK;
    int rc = ensure_csr32_on_device(ctx, g);
    if (rc) return rc;
    rc = ensure_csr_on_device(ctx, g);  // f64 couplings / biases for the energy kernel
    if (rc) return rc;
    HostGraph& h = g->h;
    const uint64_t N = h.nvars;
    std::vector<uint32_t> order(N);
    g->color_off.assign(h.ncolors + 1, 0);
    for (uint64_t n = 0; n < N; ++n) g->color_off[h.color_of(n) + 1]++;
    for (int c = 0; c < h.ncolors; ++c) g->color_off[c + 1] += g->color_off[c];
    std::vector<uint32_t> fill(g->color_off.begin(), g->color_off.end() - 1);
    for (uint64_t n = 0; n < N; ++n) order[fill[h.color_of(n)]++] = (uint32_t)n;
    std::vector<float> jf(h.jv.begin(), h.jv.end());
    std::vector<float> bf(N, 0.f);
    if (h.has_bias) for (uint64_t n = 0; n < N; ++n) bf[n] = (float)h.bias[n];
    CUDA_TRY(ctx, dev_alloc(&g->d_csites, order.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_jf, jf.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_biasf, bf.size()));
    CUDA_TRY(ctx, cudaMemcpy(g->d_csites, order.data(), order.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_jf, jf.data(), jf.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_biasf, bf.data(), bf.size() * 4, cudaMemcpyHostToDevice));
    g->real_built = true;
    return ISING_OK;
}

// colour x degree groups of a graph whose couplings all have the same magnitude
static int ensure_general_on_device(ising_ctx* ctx, ising_graph* g) {
    if (g->gen_built) return ISING_OK;
    HostGraph& h = g->h;
    if (!h.integer_classes)
        return fail(ctx, ISING_E_INVALID, "internal: integer-class kernels on a real-valued graph");
    {
        const int rc32 = ensure_csr32_on_device(ctx, g);
        if (rc32) return rc32;
    }
    const uint64_t N = h.nvars;
    const int ncol = h.ncolors;
    // bucket sites by (colour, degree)
    std::vector<std::vector<uint32_t>> bucket((size_t)ncol * (GEN_MAX_DEG + 1));
    for (uint64_t n = 0; n < N; ++n) {
        const uint32_t deg = (uint32_t)(h.row[n + 1] - h.row[n]);
        bucket[(size_t)h.color_of(n) * (GEN_MAX_DEG + 1) + deg].push_back((uint32_t)n);
    }
    std::vector<uint32_t> sites, nbr, anti;
    struct Off { size_t s, n; uint32_t count, deg; int color; };
    std::vector<Off> offs;
    for (int c = 0; c < ncol; ++c)
        for (int d = 0; d <= GEN_MAX_DEG; ++d) {
            const auto& b = bucket[(size_t)c * (GEN_MAX_DEG + 1) + d];
            if (b.empty()) continue;
            Off o{sites.size(), nbr.size(), (uint32_t)b.size(), (uint32_t)d, c};
            sites.insert(sites.end(), b.begin(), b.end());
            nbr.resize(nbr.size() + (size_t)d * b.size());
            for (size_t i = 0; i < b.size(); ++i) {
                const uint64_t lo = h.row[b[i]];
                uint32_t bits = 0;
                for (int k = 0; k < d; ++k) {
                    nbr[o.n + (size_t)k * b.size() + i] = h.nbr[lo + k];
                    if (h.jv[lo + k] > 0) bits |= 1u << k;
                }
                anti.push_back(bits);
            }
            offs.push_back(o);
        }
    CUDA_TRY(ctx, dev_alloc(&g->d_gsites, sites.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_gnbr, nbr.size()));
    CUDA_TRY(ctx, dev_alloc(&g->d_ganti, anti.size()));
    CUDA_TRY(ctx, cudaMemcpy(g->d_gsites, sites.data(), sites.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_gnbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->d_ganti, anti.data(), anti.size() * 4, cudaMemcpyHostToDevice));
    for (const Off& o : offs) {
        GenGroup gg;
        gg.sites = g->d_gsites + o.s;
        gg.nbr = g->d_gnbr + o.n;
        gg.anti = g->d_ganti + o.s;
        gg.count = o.count;
        gg.deg = o.deg;
        g->gen_groups.push_back(gg);
        g->gen_group_color.push_back(o.color);
    }
    g->gen_ncolors = ncol;
    g->gen_built = true;
    return ISING_OK;
}

extern "C" int ising_graph_from_edges(ising_ctx* ctx, uint64_t nvars, uint64_t nedges,
                                      const uint64_t* a, const uint64_t* b, const double* j,
                                      const double* biases, ising_graph** out) {
    if (!ctx || !out) return fail(ctx, ISING_E_INVALID, "ctx/out is NULL");
    *out = nullptr;
    if (nedges && (!a || !b || !j)) return fail(ctx, ISING_E_INVALID, "edge arrays are NULL");
    std::unique_ptr<ising_graph> g(new ising_graph);
    g->ctx = ctx;
    const std::string msg = compile_from_edges(nvars, nedges, a, b, j, biases, &g->h);
    if (!msg.empty()) return fail(ctx, ISING_E_INVALID, "%s", msg.c_str());
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int rc = upload_stencil_masks(ctx, g.get());
    if (rc) return rc;
    *out = g.release();
    return ISING_OK;
}

extern "C" int ising_graph_torus(ising_ctx* ctx, int dim, const uint64_t* L, double j0, int pmj,
                                 uint64_t j_seed, ising_graph** out) {
    if (!ctx || !out || !L) return fail(ctx, ISING_E_INVALID, "ctx/out/L is NULL");
    *out = nullptr;
    std::unique_ptr<ising_graph> g(new ising_graph);
    g->ctx = ctx;
    const std::string msg = make_torus(dim, L, j0, pmj, j_seed, &g->h);
    if (!msg.empty()) return fail(ctx, ISING_E_INVALID, "%s", msg.c_str());
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int rc = upload_stencil_masks(ctx, g.get());
    if (rc) return rc;
    *out = g.release();
    return ISING_OK;
}

extern "C" void ising_graph_destroy(ising_graph* g) {
    if (!g) return;
    if (g->ctx) cudaSetDevice(g->ctx->device);
    cudaFree(g->d_jmask);
    cudaFree(g->d_row);
    cudaFree(g->d_nbr);
    cudaFree(g->d_jv);
    cudaFree(g->d_bias);
    cudaFree(g->d_gsites);
    cudaFree(g->d_gnbr);
    cudaFree(g->d_ganti);
    cudaFree(g->d_row32);
    cudaFree(g->d_nbr32);
    cudaFree(g->d_anti8);
    cudaFree(g->d_csites);
    cudaFree(g->d_jf);
    cudaFree(g->d_biasf);
    delete g;
}

extern "C" int ising_graph_get_info(const ising_graph* g, ising_graph_info* out) {
    if (!g || !out) return fail(nullptr, ISING_E_INVALID, "graph/out is NULL");
    const HostGraph& h = g->h;
    out->nvars = h.nvars;
    out->nedges = h.nedges;
    out->kind = h.kind;
    out->ncolors = h.ncolors;
    out->max_degree = h.max_degree;
    out->integer_classes = h.integer_classes ? 1 : 0;
    out->dims[0] = h.dims[0];
    out->dims[1] = h.dims[1];
    out->dims[2] = h.dims[2];
    out->jabs = h.jabs;
    return ISING_OK;
}

extern "C" int ising_graph_get_colors(const ising_graph* g, uint32_t* colors) {
    if (!g || !colors) return fail(nullptr, ISING_E_INVALID, "graph/colors is NULL");
    for (uint64_t n = 0; n < g->h.nvars; ++n) colors[n] = g->h.color_of(n);
    return ISING_OK;
}

extern "C" int ising_graph_get_edges(const ising_graph* g, uint64_t* a, uint64_t* b, double* j) {
    if (!g || !a || !b || !j) return fail(nullptr, ISING_E_INVALID, "graph/arrays NULL");
    for (uint64_t e = 0; e < g->h.nedges; ++e) g->h.edge_at(e, a + e, b + e, j + e);
    return ISING_OK;
}

extern "C" int ising_make_seeds(uint64_t seed_gen, uint64_t n, uint64_t* out) {
    if (n && !out) return fail(nullptr, ISING_E_INVALID, "out is NULL");
    make_seeds(seed_gen, n, out);
    return ISING_OK;
}

extern "C" int ising_schedule_betas(const uint64_t* st, const double* sb, uint64_t n,
                                    uint64_t timesteps, int linear, double* out) {
    if ((n && (!st || !sb)) || (timesteps && !out))
        return fail(nullptr, ISING_E_INVALID, "schedule arrays are NULL");
    if (!schedule_betas(st, sb, n, timesteps, linear != 0, out))
        return fail(nullptr, ISING_E_INVALID, "annealing schedule has fewer than two stops");
    return ISING_OK;
}

// ------------------------------------------------------------------------------------------
// simulation object
// ------------------------------------------------------------------------------------------
static void count_launch(ising_sim* s, int n) {
    if (n > 0) s->stats.kernel_launches += (uint64_t)n;
}

extern "C" int ising_sim_create_ex(ising_ctx* ctx, const ising_graph* g, uint64_t E, uint64_t seed,
                                   uint64_t replica_offset, uint32_t flags, ising_sim** out) {
    if (!ctx || !g || !out) return fail(ctx, ISING_E_INVALID, "ctx/graph/out is NULL");
    *out = nullptr;
    if (g->ctx != ctx) return fail(ctx, ISING_E_INVALID, "graph belongs to another context");
    if (E == 0) return fail(ctx, ISING_E_INVALID, "num_experiments must be > 0");
    if (replica_offset % 32) return fail(ctx, ISING_E_INVALID, "replica_offset must be a multiple of 32");
    const HostGraph& h = g->h;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // integer energy classes (all |J| equal, no bias, degree <= 15) use the bit-sliced kernels;
    // anything else the reference accepts runs on the float-field kernel
    // (max_degree is known without the CSR: compile_from_edges builds it, make_torus sets 2*dim)
    const bool real = !h.integer_classes || (h.kind == ISING_KIND_GENERAL && h.max_degree > GEN_MAX_DEG);
    const bool general = real || h.kind == ISING_KIND_GENERAL || (flags & ISING_SIM_GENERAL_LAYOUT);
    if (real) {
        const int rc = ensure_real_on_device(ctx, const_cast<ising_graph*>(g));
        if (rc) return rc;
    } else if (general) {
        const int rc = ensure_general_on_device(ctx, const_cast<ising_graph*>(g));
        if (rc) return rc;
    }
    std::unique_ptr<ising_sim> s(new ising_sim);
    s->ctx = ctx;
    s->g = g;
    s->E = E;
    s->seed = seed;
    s->replica_offset = replica_offset;
    s->general = general;
    s->real = real;
    Layout& L = s->lay;
    L.kind = general ? ISING_KIND_GENERAL : h.kind;
    L.Lx = (uint32_t)h.dims[0];
    L.Ly = (uint32_t)h.dims[1];
    L.Lz = (uint32_t)h.dims[2];
    L.Lxh = L.Lx / 2;
    L.rows = L.Ly * L.Lz;
    L.W = (uint32_t)((E + 31) / 32);
    L.nvars = h.nvars;
    L.halfN = h.nvars / 2;
    if (!general && (uint64_t)L.Lxh * L.W > 0xFFFFFFFFull)
        return fail(ctx, ISING_E_UNSUPPORTED, "lattice row too long for 32-bit word offsets");
    const size_t words = (size_t)h.nvars * L.W;
    s->spins_bytes = words * sizeof(uint32_t);
    s->counts_bytes = (size_t)L.W * 32 * sizeof(unsigned long long);
    void* p = nullptr;
    CUDA_TRY(ctx, ctx_buf_get(ctx, s->spins_bytes, &p));
    s->d_spins = (uint32_t*)p;
    cudaError_t ce = ctx_buf_get(ctx, s->counts_bytes, &p);
    if (ce != cudaSuccess) {
        ctx_buf_put(ctx, s->d_spins, s->spins_bytes);
        CUDA_TRY(ctx, ce);
    }
    s->d_counts = (unsigned long long*)p;
    *out = s.release();
    return ising_sim_randomize(*out);
}

extern "C" int ising_sim_create(ising_ctx* ctx, const ising_graph* g, uint64_t E, uint64_t seed,
                                uint64_t replica_offset, ising_sim** out) {
    return ising_sim_create_ex(ctx, g, E, seed, replica_offset, 0u, out);
}

extern "C" void ising_sim_destroy(ising_sim* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    ctx_buf_put(s->ctx, s->d_spins, s->spins_bytes);
    ctx_buf_put(s->ctx, s->d_counts, s->counts_bytes);
    cudaFree(s->d_t64);
    cudaFree(s->d_slot);
    cudaFree(s->d_tplane);
    cudaFree(s->d_tlow);
    delete s;
}

extern "C" int ising_sim_configure(ising_sim* s, int planes, int rounds) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    if (planes) {
        if (planes < 5 || planes > 7) return fail(s->ctx, ISING_E_INVALID, "planes must be 5..7");
        s->planes = planes;
    }
    if (rounds) {
        if (rounds != 7 && rounds != 10) return fail(s->ctx, ISING_E_INVALID, "rounds must be 7 or 10");
        s->rounds = rounds;
    }
    return ISING_OK;
}

extern "C" int ising_sim_randomize(ising_sim* s) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    count_launch(s, launch_init_random(s->d_spins, s->lay, (uint32_t)s->seed,
                                       (uint32_t)(s->seed >> 32),
                                       (uint32_t)(s->replica_offset / 32), ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

extern "C" int ising_sim_set_state(ising_sim* s, const uint8_t* state) {
    if (!s || !state) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/state is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, s->lay.nvars, &dv));
    uint8_t* d = (uint8_t*)dv;
    cudaError_t e = cudaMemcpyAsync(d, state, s->lay.nvars, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        count_launch(s, launch_init_broadcast(s->d_spins, s->lay, d, ctx->stream));
        e = cudaStreamSynchronize(ctx->stream);
    }
    CUDA_TRY(ctx, e);
    return ISING_OK;
}

extern "C" int ising_sim_set_states(ising_sim* s, const uint8_t* states) {
    if (!s || !states) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/states is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)s->E * s->lay.nvars;
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, bytes, &dv));
    uint8_t* d = (uint8_t*)dv;
    cudaError_t e = cudaMemcpyAsync(d, states, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        count_launch(s, launch_pack_states(s->d_spins, s->lay, d, s->E, ctx->stream));
        e = cudaStreamSynchronize(ctx->stream);
    }
    CUDA_TRY(ctx, e);
    return ISING_OK;
}

// dE of the uphill classes in units of |J|: stencil 2D {4, 8}, 3D {4, 8, 12}
static void fill_thresholds(const HostGraph& h, double beta, int K, MscThresholds* th) {
    const int dim = h.kind == ISING_KIND_STENCIL3D ? 3 : 2;
    memset(th, 0, sizeof *th);
    for (int c = 0; c < dim; ++c) {
        const double de = 4.0 * (c + 1) * h.jabs;
        const double p = exp(-beta * de);
        const double scaled = ldexp(p, K + 32);
        const uint64_t tmax = (1ull << (K + 32)) - 1;
        uint64_t T;
        if (!(scaled >= 0.0)) T = 0;  // NaN beta: never accept uphill
        else if (scaled >= (double)tmax) T = tmax;
        else T = (uint64_t)floor(scaled);
        for (int pl = 0; pl < K; ++pl)
            th->plane[c][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[c] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

// T = floor(exp(-beta dE) 2^(K+32)) clipped to K+32 bits; NaN -> 0 (never accept uphill)
static uint64_t threshold64(double beta, double de, int K) {
    const double scaled = ldexp(exp(-beta * de), K + 32);
    const uint64_t tmax = (1ull << (K + 32)) - 1;
    if (!(scaled >= 0.0)) return 0;
    if (scaled >= (double)tmax) return tmax;
    return (uint64_t)floor(scaled);
}

// uphill classes of a degree-d site: n_sat = d/2+1 .. d, dE = 2|J|(2 n_sat - d)
static void fill_gen_thresholds(double jabs, double beta, int K, uint32_t deg, GenThresholds* th) {
    memset(th, 0, sizeof *th);
    const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
    for (uint32_t j = 0; j < ncls && j < (uint32_t)GEN_MAX_CLS; ++j) {
        const int cls = 2 * (int)(cmin + j) - (int)deg;
        const uint64_t T = threshold64(beta, 2.0 * jabs * (double)cls, K);
        for (int pl = 0; pl < K; ++pl)
            th->plane[j][pl] = ((T >> (K + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        th->low[j] = (uint32_t)(T & 0xFFFFFFFFull);
    }
}

static int sim_one_sweep_general(ising_sim* s, double beta) {
    ising_ctx* ctx = s->ctx;
    const ising_graph* g = s->g;
    const HostGraph& h = g->h;
    GenSweepArgs a;
    a.spins = s->d_spins;
    a.W = s->lay.W;
    a.sweep = (uint32_t)s->sweep_counter;
    a.key0 = (uint32_t)s->seed;
    a.key1 = (uint32_t)(s->seed >> 32);
    a.gw0 = (uint32_t)(s->replica_offset / 32);
    a.planes = s->planes;
    a.rounds = s->rounds;
    a.tables.plane = s->perbeta ? s->d_tplane : nullptr;
    a.tables.low = s->perbeta ? s->d_tlow : nullptr;
    memset(&a.th, 0, sizeof a.th);
    int launches = 0;
    for (int c = 0; c < g->gen_ncolors; ++c)
        for (size_t k = 0; k < g->gen_groups.size(); ++k) {
            if (g->gen_group_color[k] != c) continue;
            const GenGroup& gg = g->gen_groups[k];
            // (an isolated site has dE = 0 and, as in the reference's dE <= 0 rule, always flips)
            if (!s->perbeta) fill_gen_thresholds(h.jabs, beta, s->planes, gg.deg, &a.th);
            const int n = launch_sweep_general(a, gg, ctx->stream);
            if (n < 0) return fail(ctx, ISING_E_CUDA, "general sweep launch failed: %s",
                                   cudaGetErrorString(cudaGetLastError()));
            launches += n;
        }
    count_launch(s, launches);
    s->stats.sweep_kernel_launches += (uint64_t)launches;
    s->sweep_counter++;
    s->stats.sweeps++;
    s->stats.flip_attempts += s->E * h.nvars;
    return ISING_OK;
}

static int sim_one_sweep_real(ising_sim* s, double beta) {
    ising_ctx* ctx = s->ctx;
    const ising_graph* g = s->g;
    RealSweepArgs a;
    a.spins = s->d_spins;
    a.row = g->d_row32;
    a.nbr = g->d_nbr32;
    a.jf = g->d_jf;
    a.biasf = g->d_biasf;
    a.W = s->lay.W;
    a.beta = (float)beta;
    a.sweep = (uint32_t)s->sweep_counter;
    a.key0 = (uint32_t)s->seed;
    a.key1 = (uint32_t)(s->seed >> 32);
    a.gw0 = (uint32_t)(s->replica_offset / 32);
    a.rounds = s->rounds;
    int launches = 0;
    for (int c = 0; c < g->h.ncolors; ++c) {
        a.sites = g->d_csites + g->color_off[c];
        a.count = g->color_off[c + 1] - g->color_off[c];
        const int n = launch_sweep_real(a, ctx->stream);
        if (n < 0) return fail(ctx, ISING_E_CUDA, "real-coupling sweep launch failed: %s",
                               cudaGetErrorString(cudaGetLastError()));
        launches += n;
    }
    count_launch(s, launches);
    s->stats.sweep_kernel_launches += (uint64_t)launches;
    s->sweep_counter++;
    s->stats.sweeps++;
    s->stats.flip_attempts += s->E * g->h.nvars;
    return ISING_OK;
}

// f64 energies of a real-coupling sim into d_out[e * estride + eoff]
static int sim_energy_real(ising_sim* s, double* d_tmp /* [32 W] */) {
    ising_ctx* ctx = s->ctx;
    const ising_graph* g = s->g;
    CUDA_TRY(ctx, cudaMemsetAsync(d_tmp, 0, (size_t)s->lay.W * 32 * sizeof(double), ctx->stream));
    const int n = launch_energy_real(s->d_spins, s->lay.nvars, s->lay.W, g->d_row32, g->d_nbr32,
                                     g->d_jv, g->d_bias, d_tmp, ctx->stream);
    if (n < 0) return fail(ctx, ISING_E_CUDA, "energy launch failed");
    count_launch(s, n);
    return ISING_OK;
}

static int sim_one_sweep(ising_sim* s, double beta, unsigned long long* nsat_out = nullptr) {
    if (s->real) return sim_one_sweep_real(s, beta);
    if (s->general) return sim_one_sweep_general(s, beta);
    ising_ctx* ctx = s->ctx;
    const HostGraph& h = s->g->h;
    SweepArgs a;
    a.spins = s->d_spins;
    a.jmask = s->g->d_jmask;
    a.lay = s->lay;
    a.sweep = (uint32_t)s->sweep_counter;
    a.key0 = (uint32_t)s->seed;
    a.key1 = (uint32_t)(s->seed >> 32);
    a.gw0 = (uint32_t)(s->replica_offset / 32);
    a.antiferro = h.uniform_antiferro ? 0xFFFFFFFFu : 0u;
    a.planes = s->planes;
    a.rounds = s->rounds;
    if (s->perbeta) memset(&a.th, 0, sizeof a.th);
    else fill_thresholds(h, beta, s->planes, &a.th);
    a.nsat_out = nsat_out;
    a.tplane = s->perbeta ? s->d_tplane : nullptr;
    a.tlow = s->perbeta ? s->d_tlow : nullptr;
    const int n = launch_sweep_stencil(a, ctx->stream);
    if (n < 0) return fail(ctx, ISING_E_CUDA, "sweep launch failed: %s",
                           cudaGetErrorString(cudaGetLastError()));
    count_launch(s, n);
    s->stats.sweep_kernel_launches += (uint64_t)n;
    s->sweep_counter++;
    s->stats.sweeps++;
    s->stats.flip_attempts += s->E * h.nvars;
    return ISING_OK;
}

// Launch-bound sizes: a whole chunk of sweeps in one cooperative launch.  Returns 1 when done
// that way, 0 when the caller should fall back to per-phase launches, < 0 on error (rc in *err).
static int sim_sweeps_coop(ising_sim* s, const double* betas, uint64_t nt, unsigned long long* hist,
                           int* err) {
    *err = ISING_OK;
    if (s->general || s->real || s->perbeta || nt == 0) return 0;
    if ((uint64_t)s->lay.halfN * s->lay.W > (1ull << 19)) return 0;  // big enough to fill the GPU
    // measured on B200 (32^2 and 16^3): with fused energies the per-block reduction in lock-step
    // only pays off for few replica words (6.6 vs 10.2 us/sweep at W = 2, 18.9 vs 13.6 at W = 32)
    if (hist && s->lay.W > 8) return 0;
    ising_ctx* ctx = s->ctx;
    const HostGraph& h = s->g->h;
    std::vector<MscThresholds> th(nt);
    for (uint64_t t = 0; t < nt; ++t) fill_thresholds(h, betas[t], s->planes, &th[t]);
    void* dv = nullptr;
    cudaError_t e = ctx_scratch(ctx, 3, nt * sizeof(MscThresholds), &dv);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dv, th.data(), nt * sizeof(MscThresholds), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        *err = fail(ctx, ISING_E_CUDA, "threshold table upload: %s", cudaGetErrorString(e));
        return -1;
    }
    SweepArgs a;
    a.spins = s->d_spins;
    a.jmask = s->g->d_jmask;
    a.lay = s->lay;
    a.sweep = (uint32_t)s->sweep_counter;
    a.key0 = (uint32_t)s->seed;
    a.key1 = (uint32_t)(s->seed >> 32);
    a.gw0 = (uint32_t)(s->replica_offset / 32);
    a.antiferro = h.uniform_antiferro ? 0xFFFFFFFFu : 0u;
    a.planes = s->planes;
    a.rounds = s->rounds;
    a.nsat_out = nullptr;
    a.tplane = nullptr;
    a.tlow = nullptr;
    memset(&a.th, 0, sizeof a.th);
    const int rc = launch_sweeps_stencil_coop(a, (const MscThresholds*)dv, (uint32_t)nt, hist,
                                              (uint32_t)(s->lay.W * 32), ctx->stream);
    if (rc < 0) {
        cudaGetLastError();
        return 0;  // e.g. too many blocks to be co-resident: use the per-phase launches
    }
    if (rc == 0) return 0;
    // the host table must outlive the copy
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        *err = fail(ctx, ISING_E_CUDA, "cooperative sweep: %s", cudaGetErrorString(e));
        return -1;
    }
    count_launch(s, 1);
    s->stats.sweep_kernel_launches += 1;
    s->sweep_counter += nt;
    s->stats.sweeps += nt;
    s->stats.flip_attempts += nt * s->E * h.nvars;
    return 1;
}

// n_sat per experiment into s->d_counts (zeroed first)
static int sim_count_nsat(ising_sim* s, unsigned long long* d_counts) {
    ising_ctx* ctx = s->ctx;
    const HostGraph& h = s->g->h;
    CUDA_TRY(ctx, cudaMemsetAsync(d_counts, 0, (size_t)s->lay.W * 32 * sizeof(unsigned long long),
                                  ctx->stream));
    if (s->general) {
        const int n = launch_nsat_general(s->d_spins, s->lay.nvars, s->lay.W, s->g->d_row32,
                                          s->g->d_nbr32, s->g->d_anti8, d_counts, ctx->stream);
        if (n < 0) return fail(ctx, ISING_E_CUDA, "energy launch failed");
        count_launch(s, n);
        return ISING_OK;
    }
    const int n = launch_nsat_stencil(s->d_spins, s->g->d_jmask, s->lay,
                                      h.uniform_antiferro ? 0xFFFFFFFFu : 0u, d_counts,
                                      ctx->stream);
    if (n < 0) return fail(ctx, ISING_E_CUDA, "energy launch failed");
    count_launch(s, n);
    return ISING_OK;
}

extern "C" int ising_sim_sweeps(ising_sim* s, const double* betas, uint64_t nsweeps,
                                double* energies_per_sweep) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    if (s->perbeta ? betas != nullptr : (nsweeps && !betas))
        return fail(s->ctx, ISING_E_INVALID,
                    s->perbeta ? "sim runs at per-experiment betas (ising_sim_set_betas): pass betas = NULL"
                               : "betas is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const HostGraph& h = s->g->h;
    const uint64_t E = s->E;
    const int mult = s->general ? 1 : 2;
    if (!energies_per_sweep) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        uint64_t t = 0;
        while (t < nsweeps) {
            const uint64_t nt = std::min<uint64_t>(4096, nsweeps - t);
            int err = ISING_OK;
            const int done = betas ? sim_sweeps_coop(s, betas + t, nt, nullptr, &err) : 0;
            if (done < 0) return err;
            if (done) { t += nt; continue; }
            for (uint64_t k = 0; k < nt; ++k) {
                const int rc = sim_one_sweep(s, betas ? betas[t + k] : 0.0);
                if (rc) return rc;
            }
            t += nt;
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        s->stats.sweep_device_ms += ms;
        s->stats.sweep_kernel_ms += ms;
        return ISING_OK;
    }
    // per-sweep energies: integer n_sat history on the device, converted and transposed to
    // double[E, nsweeps] there, one D2H per chunk
    const uint64_t chunk_max = 2048;
    const size_t cw = (size_t)s->lay.W * 32;
    unsigned long long* d_hist = nullptr;
    double* d_out = nullptr;
    void* sp = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 1, cw * std::min(chunk_max, nsweeps) * sizeof(unsigned long long), &sp));
    d_hist = (unsigned long long*)sp;
    CUDA_TRY(ctx, ctx_scratch(ctx, 2, (size_t)E * std::min(chunk_max, nsweeps) * sizeof(double), &sp));
    d_out = (double*)sp;
    cudaError_t e = cudaSuccess;
    std::vector<double> host;
    int rc = ISING_OK;
    for (uint64_t t0 = 0; t0 < nsweeps && rc == ISING_OK; t0 += chunk_max) {
        const uint64_t nt = std::min(chunk_max, nsweeps - t0);
        cudaEventRecord(ctx->ev0, ctx->stream);
        cudaMemsetAsync(d_hist, 0, cw * nt * sizeof(unsigned long long), ctx->stream);
        // the second colour phase of every sweep adds its post-flip satisfied-bond counts
        // into that sweep's slot of the history (fused, no separate energy pass)
        int coop_err = ISING_OK;
        const int coop = betas ? sim_sweeps_coop(s, betas + t0, nt, d_hist, &coop_err) : 0;
        if (coop < 0) rc = coop_err;
        for (uint64_t t = 0; coop == 0 && t < nt && rc == ISING_OK; ++t) {
            rc = sim_one_sweep(s, betas ? betas[t0 + t] : 0.0, d_hist + t * cw);
            if (rc == ISING_OK && s->real) {
                rc = sim_energy_real(s, reinterpret_cast<double*>(d_hist + t * cw));
            } else if (rc == ISING_OK && s->general) {  // no fused accumulation on general graphs
                const int n = launch_nsat_general(s->d_spins, s->lay.nvars, s->lay.W, s->g->d_row32,
                                                  s->g->d_nbr32, s->g->d_anti8, d_hist + t * cw,
                                                  ctx->stream);
                if (n < 0) rc = fail(ctx, ISING_E_CUDA, "energy launch failed");
                else count_launch(s, n);
            }
        }
        if (rc == ISING_OK && s->real)
            count_launch(s, launch_transpose_hist_f64(reinterpret_cast<double*>(d_hist), E, cw, nt,
                                                      d_out, ctx->stream));
        else if (rc == ISING_OK)
            count_launch(s, launch_energy_from_hist(d_hist, E, cw, nt, h.jabs, h.nedges, mult,
                                                    d_out, ctx->stream));
        cudaEventRecord(ctx->ev1, ctx->stream);
        if (rc != ISING_OK) break;
        host.resize((size_t)E * nt);
        e = cudaMemcpyAsync(host.data(), d_out, host.size() * sizeof(double),
                            cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rc = fail(ctx, ISING_E_CUDA, "energy read-back: %s", cudaGetErrorString(e)); break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        s->stats.sweep_device_ms += ms;
        for (uint64_t ex = 0; ex < E; ++ex)
            memcpy(energies_per_sweep + ex * nsweeps + t0, host.data() + ex * nt, nt * sizeof(double));
    }
    return rc;
}

// current energies of all experiments into d_out[e * estride + eoff] (device)
static int sim_energies_to_device(ising_sim* s, double* d_out, uint64_t estride, uint64_t eoff) {
    ising_ctx* ctx = s->ctx;
    const HostGraph& h = s->g->h;
    if (s->real) {
        double* tmp = reinterpret_cast<double*>(s->d_counts);  // same size as the u64 counters
        const int rc = sim_energy_real(s, tmp);
        if (rc) return rc;
        count_launch(s, launch_copy_strided_f64(tmp, s->E, d_out, estride, eoff, ctx->stream));
        return ISING_OK;
    }
    const int rc = sim_count_nsat(s, s->d_counts);
    if (rc) return rc;
    count_launch(s, launch_energy_from_nsat(s->d_counts, s->E, h.jabs, h.nedges, s->general ? 1 : 2,
                                            d_out, estride, eoff, ctx->stream));
    return ISING_OK;
}

extern "C" int ising_sim_get_energies(ising_sim* s, double* energies) {
    if (!s || !energies) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/energies is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 2, s->E * sizeof(double), &dv));
    double* d_out = (double*)dv;
    const int rc = sim_energies_to_device(s, d_out, 1, 0);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(energies, d_out, s->E * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

extern "C" int ising_sim_get_magnetization(ising_sim* s, double* m) {
    if (!s || !m) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/m is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t cw = (size_t)s->lay.W * 32;
    CUDA_TRY(ctx, cudaMemsetAsync(s->d_counts, 0, cw * sizeof(unsigned long long), ctx->stream));
    count_launch(s, launch_count_up(s->d_spins, s->lay, s->d_counts, ctx->stream));
    std::vector<unsigned long long> up(cw);
    CUDA_TRY(ctx, cudaMemcpyAsync(up.data(), s->d_counts, cw * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint64_t e = 0; e < s->E; ++e)
        m[e] = 2.0 * (double)up[e] - (double)s->lay.nvars;
    return ISING_OK;
}

// bool[E, nvars] to host memory, staged through a device buffer in slabs of experiments
static int sim_states_to_host(ising_sim* s, uint8_t* states) {
    ising_ctx* ctx = s->ctx;
    const uint64_t N = s->lay.nvars;
    const size_t bytes = (size_t)s->E * N;
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, bytes, &dv));
    uint8_t* d = (uint8_t*)dv;
    count_launch(s, launch_unpack_states(s->d_spins, s->lay, d, s->E, N, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(states, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

extern "C" int ising_sim_get_states(ising_sim* s, uint8_t* states) {
    if (!s || !states) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/states is NULL");
    CUDA_TRY(s->ctx, cudaSetDevice(s->ctx->device));
    return sim_states_to_host(s, states);
}

extern "C" int ising_sim_get_packed(ising_sim* s, uint32_t* words) {
    if (!s || !words) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/words is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)s->lay.nvars * s->lay.W;
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, n * 4, &dv));
    uint32_t* d = (uint32_t*)dv;
    count_launch(s, launch_export_natural(s->d_spins, s->lay, d, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(words, d, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

// checkpoint support: the packed state in natural order plus the sweep counter are the whole
// state of a sim (the RNG is counter-based: seed + counter, nothing else to save)
extern "C" int ising_sim_set_packed(ising_sim* s, const uint32_t* words) {
    if (!s || !words) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/words is NULL");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)s->lay.nvars * s->lay.W;
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, n * 4, &dv));
    CUDA_TRY(ctx, cudaMemcpyAsync(dv, words, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    count_launch(s, launch_import_natural(s->d_spins, s->lay, (const uint32_t*)dv, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

extern "C" int ising_sim_get_counter(const ising_sim* s, uint64_t* sweeps_done) {
    if (!s || !sweeps_done) return fail(nullptr, ISING_E_INVALID, "sim/out is NULL");
    *sweeps_done = s->sweep_counter;
    return ISING_OK;
}

extern "C" int ising_sim_set_counter(ising_sim* s, uint64_t sweeps_done) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    s->sweep_counter = sweeps_done;
    return ISING_OK;
}

extern "C" int ising_sim_get_stats(ising_sim* s, ising_sim_stats* out) {
    if (!s || !out) return fail(nullptr, ISING_E_INVALID, "sim/out is NULL");
    *out = s->stats;
    return ISING_OK;
}

extern "C" int ising_sim_reset_stats(ising_sim* s) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    s->stats = ising_sim_stats{};
    return ISING_OK;
}

// ------------------------------------------------------------------------------------------
// per-experiment inverse temperatures and classical parallel tempering
// ------------------------------------------------------------------------------------------
extern "C" int ising_sim_set_betas(ising_sim* s, const double* betas) {
    if (!s || !betas) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/betas is NULL");
    ising_ctx* ctx = s->ctx;
    if (s->real)
        return fail(ctx, ISING_E_UNSUPPORTED,
                    "per-experiment betas need integer energy classes (all |J| equal, no bias)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const HostGraph& h = s->g->h;
    const uint32_t W = s->lay.W, E32 = 32 * W;
    if (!s->general) {
        // checkerboard layout: classes dE = 4|J|, 8|J| (, 12|J|); tables are built for 6 planes
        if (s->planes != 6)
            return fail(ctx, ISING_E_UNSUPPORTED, "per-experiment betas on a lattice need planes = 6");
        std::vector<unsigned long long> t64((size_t)E32 * 3, 0ull);
        const int ncls = h.kind == ISING_KIND_STENCIL3D ? 3 : 2;
        for (uint32_t e = 0; e < E32; ++e) {
            const double beta = betas[e < s->E ? e : 0];
            for (int c = 0; c < ncls; ++c)
                t64[(size_t)e * 3 + c] = threshold64(beta, 4.0 * (c + 1) * h.jabs, s->planes);
        }
        if (!s->d_t64) {
            CUDA_TRY(ctx, dev_alloc(&s->d_t64, t64.size()));
            CUDA_TRY(ctx, dev_alloc(&s->d_tplane, (size_t)W * 3 * 8));
            CUDA_TRY(ctx, dev_alloc(&s->d_tlow, (size_t)E32 * 3));
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(s->d_t64, t64.data(), t64.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        count_launch(s, launch_build_tables_stencil(s->d_t64, W, s->planes, s->d_tplane, s->d_tlow, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        s->perbeta = true;
        return ISING_OK;
    }
    const size_t per = (size_t)(GEN_MAX_DEG + 1) * GEN_MAX_CLS;
    std::vector<unsigned long long> t64((size_t)E32 * per, 0ull);
    std::vector<uint32_t> slot(E32);
    for (uint32_t e = 0; e < E32; ++e) {
        slot[e] = e;
        const double beta = betas[e < s->E ? e : 0];  // padding bits: any valid beta
        for (uint32_t deg = 1; deg <= (uint32_t)GEN_MAX_DEG; ++deg) {
            const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
            for (uint32_t j = 0; j < ncls; ++j) {
                const int cls = 2 * (int)(cmin + j) - (int)deg;
                t64[(size_t)e * per + (size_t)deg * GEN_MAX_CLS + j] =
                    threshold64(beta, 2.0 * h.jabs * (double)cls, s->planes);
            }
        }
    }
    if (!s->d_t64) {
        CUDA_TRY(ctx, dev_alloc(&s->d_t64, t64.size()));
        CUDA_TRY(ctx, dev_alloc(&s->d_slot, (size_t)E32));
        CUDA_TRY(ctx, dev_alloc(&s->d_tplane, (size_t)(GEN_MAX_DEG + 1) * W * GEN_MAX_CLS * 8));
        CUDA_TRY(ctx, dev_alloc(&s->d_tlow, (size_t)(GEN_MAX_DEG + 1) * E32 * GEN_MAX_CLS));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(s->d_t64, t64.data(), t64.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s->d_slot, slot.data(), slot.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    count_launch(s, launch_build_tables(s->d_t64, s->d_slot, W, s->planes, s->d_tplane, s->d_tlow,
                                        ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // t64/slot are stack-lifetime host buffers
    s->perbeta = true;
    return ISING_OK;
}

struct ising_pt {
    ising_ctx* ctx = nullptr;
    const ising_graph* g = nullptr;
    ising_sim* sim = nullptr;
    uint64_t R = 0, lo = 0, hi = 0;   // this rank owns configurations [lo, hi)
    uint64_t word_lo = 0;             // first replica word held locally
    std::vector<double> betas;        // by slot
    std::vector<uint32_t> slot_of_cfg, cfg_of_slot;
    uint64_t seed = 0, swap_step = 0, total_swaps = 0;
};

// betas of the locally held replica bits from the slot permutation
static int pt_push_betas(ising_pt* pt) {
    const uint64_t E = pt->sim->E;
    std::vector<double> b(E);
    for (uint64_t e = 0; e < E; ++e) {
        const uint64_t cfg = pt->word_lo * 32 + e;
        b[e] = pt->betas[cfg < pt->R ? pt->slot_of_cfg[cfg] : 0];
    }
    return ising_sim_set_betas(pt->sim, b.data());
}

extern "C" int ising_pt_create(ising_ctx* ctx, const ising_graph* g, const double* betas,
                               uint64_t nbetas, uint64_t cfg_lo, uint64_t cfg_hi, uint64_t seed,
                               ising_pt** out) {
    if (!ctx || !g || !betas || !out) return fail(ctx, ISING_E_INVALID, "ctx/graph/betas/out is NULL");
    *out = nullptr;
    if (nbetas == 0 || cfg_lo >= cfg_hi || cfg_hi > nbetas)
        return fail(ctx, ISING_E_INVALID, "need 0 <= cfg_lo < cfg_hi <= nbetas");
    std::unique_ptr<ising_pt> pt(new ising_pt);
    pt->ctx = ctx;
    pt->g = g;
    pt->R = nbetas;
    pt->lo = cfg_lo;
    pt->hi = cfg_hi;
    pt->seed = seed;
    pt->betas.assign(betas, betas + nbetas);
    pt->slot_of_cfg.resize(nbetas);
    pt->cfg_of_slot.resize(nbetas);
    for (uint64_t r = 0; r < nbetas; ++r) pt->slot_of_cfg[r] = pt->cfg_of_slot[r] = (uint32_t)r;
    // whole replica words: configuration c always lives at bit c%32 of global word c/32, so a
    // sharded run draws exactly the random numbers of the unsharded one
    pt->word_lo = cfg_lo / 32;
    const uint64_t word_hi = (cfg_hi + 31) / 32;
    const uint64_t E = std::min<uint64_t>((word_hi - pt->word_lo) * 32, nbetas - pt->word_lo * 32);
    // lattices temper on the checkerboard kernels, other graphs on the colour x degree kernels
    int rc = ising_sim_create_ex(ctx, g, E, seed, pt->word_lo * 32, 0u, &pt->sim);
    if (rc) return rc;
    rc = pt_push_betas(pt.get());
    if (rc) { ising_sim_destroy(pt->sim); return rc; }
    *out = pt.release();
    return ISING_OK;
}

extern "C" void ising_pt_destroy(ising_pt* pt) {
    if (!pt) return;
    ising_sim_destroy(pt->sim);
    delete pt;
}

extern "C" int ising_pt_configure(ising_pt* pt, int planes, int rounds) {
    if (!pt) return fail(nullptr, ISING_E_INVALID, "pt is NULL");
    const int rc = ising_sim_configure(pt->sim, planes, rounds);
    return rc ? rc : pt_push_betas(pt);
}

extern "C" int ising_pt_sweeps(ising_pt* pt, uint64_t t, double* local_energies) {
    if (!pt) return fail(nullptr, ISING_E_INVALID, "pt is NULL");
    int rc = ising_sim_sweeps(pt->sim, nullptr, t, nullptr);
    if (rc || !local_energies) return rc;
    std::vector<double> en(pt->sim->E);
    rc = ising_sim_get_energies(pt->sim, en.data());
    if (rc) return rc;
    for (uint64_t c = pt->lo; c < pt->hi; ++c) local_energies[c - pt->lo] = en[c - pt->word_lo * 32];
    return ISING_OK;
}

// One tempering step (the shape of TemperingContainer::parallel_tempering_step as driven from
// tempering.rs:191-194): even slot pairs (0,1),(2,3).. then odd pairs (1,2),(3,4)..; the pair
// (a, a+1) exchanges configurations with probability min(1, exp((b_a - b_{a+1})(E_a - E_{a+1}))).
// The uniform is Philox(seed; slot a, swap step), so every rank takes the same decisions from
// the all-gathered energies.  all_energies is indexed by CONFIGURATION.
extern "C" int ising_pt_decide_swaps(const double* betas, uint64_t R, const double* all_energies,
                                     uint64_t seed, uint64_t swap_step, uint32_t* slot_of_cfg,
                                     uint32_t* cfg_of_slot, uint64_t* nswaps) {
    if (!betas || !all_energies || !slot_of_cfg || !cfg_of_slot)
        return fail(nullptr, ISING_E_INVALID, "ising_pt_decide_swaps: NULL argument");
    uint64_t swaps = 0;
    for (int parity = 0; parity < 2; ++parity)
        for (uint64_t a = parity; a + 1 < R; a += 2) {
            const uint32_t ca = cfg_of_slot[a], cb = cfg_of_slot[a + 1];
            const double d = (betas[a] - betas[a + 1]) * (all_energies[ca] - all_energies[cb]);
            bool acc = true;
            if (d < 0.0) {
                const u32x4 r = philox4x32<10>((uint32_t)a, (uint32_t)parity, (uint32_t)swap_step,
                                               TAG_SWAP << 24, (uint32_t)seed, (uint32_t)(seed >> 32));
                const double uu = ((double)r.x + 0.5) * (1.0 / 4294967296.0);
                acc = uu < exp(d);
            }
            if (acc) {
                cfg_of_slot[a] = cb;
                cfg_of_slot[a + 1] = ca;
                slot_of_cfg[cb] = (uint32_t)a;
                slot_of_cfg[ca] = (uint32_t)(a + 1);
                ++swaps;
            }
        }
    if (nswaps) *nswaps = swaps;
    return ISING_OK;
}

extern "C" int ising_pt_swap_step(ising_pt* pt, const double* all_energies) {
    if (!pt || !all_energies) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/energies is NULL");
    uint64_t swaps = 0;
    const int rc = ising_pt_decide_swaps(pt->betas.data(), pt->R, all_energies, pt->seed, pt->swap_step,
                                         pt->slot_of_cfg.data(), pt->cfg_of_slot.data(), &swaps);
    if (rc) return rc;
    pt->total_swaps += swaps;
    pt->swap_step++;
    return pt_push_betas(pt);
}

extern "C" int ising_pt_get_slots(const ising_pt* pt, uint32_t* slot_of_config) {
    if (!pt || !slot_of_config) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    for (uint64_t c = 0; c < pt->R; ++c) slot_of_config[c] = pt->slot_of_cfg[c];
    return ISING_OK;
}

extern "C" int ising_pt_get_local_states(ising_pt* pt, uint8_t* states) {
    if (!pt || !states) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/states is NULL");
    const uint64_t N = pt->g->h.nvars;
    std::vector<uint8_t> all((size_t)pt->sim->E * N);
    const int rc = ising_sim_get_states(pt->sim, all.data());
    if (rc) return rc;
    for (uint64_t c = pt->lo; c < pt->hi; ++c)
        memcpy(states + (c - pt->lo) * N, all.data() + (c - pt->word_lo * 32) * N, N);
    return ISING_OK;
}

// checkpoint support: the sim behind the ladder, and the permutation / counters
extern "C" int ising_pt_get_sim(ising_pt* pt, ising_sim** out) {
    if (!pt || !out) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *out = pt->sim;
    return ISING_OK;
}

extern "C" int ising_pt_get_counters(const ising_pt* pt, uint64_t* swap_step, uint64_t* total_swaps) {
    if (!pt || !swap_step || !total_swaps) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *swap_step = pt->swap_step;
    *total_swaps = pt->total_swaps;
    return ISING_OK;
}

extern "C" int ising_pt_restore(ising_pt* pt, const uint32_t* slot_of_config, uint64_t swap_step,
                                uint64_t total_swaps) {
    if (!pt || !slot_of_config) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/slots is NULL");
    std::vector<uint8_t> seen(pt->R, 0);
    for (uint64_t c = 0; c < pt->R; ++c) {
        if (slot_of_config[c] >= pt->R || seen[slot_of_config[c]])
            return fail(pt->ctx, ISING_E_INVALID, "slot_of_config is not a permutation");
        seen[slot_of_config[c]] = 1;
    }
    for (uint64_t c = 0; c < pt->R; ++c) {
        pt->slot_of_cfg[c] = slot_of_config[c];
        pt->cfg_of_slot[slot_of_config[c]] = (uint32_t)c;
    }
    pt->swap_step = swap_step;
    pt->total_swaps = total_swaps;
    return pt_push_betas(pt);
}

extern "C" int ising_pt_total_swaps(const ising_pt* pt, uint64_t* out) {
    if (!pt || !out) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *out = pt->total_swaps;
    return ISING_OK;
}

// LatticeTempering::qmc_timesteps_sample, tempering.rs:156-222, single rank (cfg range = all):
// run min(to_sample, to_swap, remaining) -> swap step -> sample; states[R, n_s, nvars] holds
// "the configuration currently at beta_r", energies[R] = sum(E_r after chunk * chunk) / timesteps.
extern "C" int ising_pt_timesteps_sample(ising_pt* pt, uint64_t timesteps, uint64_t replica_swap_freq,
                                         uint64_t sampling_freq, uint8_t* states, double* energies) {
    if (!pt || !energies) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/energies is NULL");
    if (pt->lo != 0 || pt->hi != pt->R)
        return fail(pt->ctx, ISING_E_INVALID, "ising_pt_timesteps_sample needs all configurations on this rank");
    if (replica_swap_freq == 0 || sampling_freq == 0)
        return fail(pt->ctx, ISING_E_INVALID,
                    "replica_swap_freq and sampling_freq must be > 0 (the reference loops forever on 0)");
    const uint64_t R = pt->R, N = pt->g->h.nvars, ns = timesteps / sampling_freq;
    if (ns && !states) return fail(pt->ctx, ISING_E_INVALID, "states is NULL");
    std::vector<double> acc(R, 0.0), en(R);
    std::vector<uint8_t> local;
    uint64_t remaining = timesteps, to_swap = replica_swap_freq, to_sample = sampling_freq, k = 0;
    while (remaining > 0) {
        const uint64_t t = std::min(std::min(to_sample, to_swap), remaining);
        int rc = ising_pt_sweeps(pt, t, en.data());
        if (rc) return rc;
        for (uint64_t slot = 0; slot < R; ++slot) acc[slot] += en[pt->cfg_of_slot[slot]] * (double)t;
        to_sample -= t; to_swap -= t; remaining -= t;
        if (to_swap == 0) {
            rc = ising_pt_swap_step(pt, en.data());
            if (rc) return rc;
            to_swap = replica_swap_freq;
        }
        if (to_sample == 0) {
            if (k < ns) {
                local.resize((size_t)R * N);
                rc = ising_pt_get_local_states(pt, local.data());
                if (rc) return rc;
                for (uint64_t slot = 0; slot < R; ++slot)
                    memcpy(states + (slot * ns + k) * N, local.data() + (size_t)pt->cfg_of_slot[slot] * N, N);
            }
            ++k;
            to_sample = sampling_freq;
        }
    }
    for (uint64_t slot = 0; slot < R; ++slot) energies[slot] = acc[slot] / (double)timesteps;
    return ISING_OK;
}

// ------------------------------------------------------------------------------------------
// one large 2D lattice in row strips (config 5)
// ------------------------------------------------------------------------------------------
struct ising_strip {
    ising_ctx* ctx = nullptr;
    StripGeom g{};
    uint64_t Lx = 0;
    uint32_t* d_spins = nullptr;
    size_t bytes = 0;
    unsigned long long* d_acc = nullptr;
    double j = -1.0;
    uint64_t seed = 0, sweep = 0, launches = 0;
    int planes = 6, rounds = 10;
    double device_ms = 0.0;
};

extern "C" int ising_strip_create_ex(ising_ctx* ctx, uint64_t Lx, uint64_t Ly, uint64_t row_lo,
                                     uint64_t row_hi, double j, uint64_t seed, uint32_t ghost,
                                     ising_strip** out) {
    if (!ctx || !out) return fail(ctx, ISING_E_INVALID, "ctx/out is NULL");
    *out = nullptr;
    if (ghost < 1 || ghost > row_hi - row_lo || ghost > 1024)
        return fail(ctx, ISING_E_INVALID, "ghost depth must be 1..min(rows, 1024)");
    if (Lx < 64 || Lx % 64) return fail(ctx, ISING_E_INVALID, "Lx must be a positive multiple of 64");
    if (Ly < 2 || (Ly & 1)) return fail(ctx, ISING_E_INVALID, "Ly must be even");
    if (row_lo >= row_hi || row_hi > Ly) return fail(ctx, ISING_E_INVALID, "need 0 <= row_lo < row_hi <= Ly");
    if (Ly > 0xFFFFFFFFull || Lx / 64 > 0x3FFFFFFFull) return fail(ctx, ISING_E_INVALID, "lattice too large");
    if (!(fabs(j) > 0.0) || !std::isfinite(j)) return fail(ctx, ISING_E_INVALID, "j must be finite and non-zero");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<ising_strip> s(new ising_strip);
    s->ctx = ctx;
    s->Lx = Lx;
    s->g.Wr = (uint32_t)(Lx / 64);
    s->g.rows = (uint32_t)(row_hi - row_lo);
    s->g.row0 = (uint32_t)row_lo;
    s->g.Ly = (uint32_t)Ly;
    s->g.ghost = ghost;
    s->j = j;
    s->seed = seed;
    s->bytes = (size_t)2 * (s->g.rows + 2 * ghost) * s->g.Wr * sizeof(uint32_t);
    void* p = nullptr;
    CUDA_TRY(ctx, ctx_buf_get(ctx, s->bytes, &p));
    s->d_spins = (uint32_t*)p;
    cudaError_t e = ctx_buf_get(ctx, 2 * sizeof(unsigned long long), &p);
    if (e != cudaSuccess) { ctx_buf_put(ctx, s->d_spins, s->bytes); CUDA_TRY(ctx, e); }
    s->d_acc = (unsigned long long*)p;
    CUDA_TRY(ctx, cudaMemsetAsync(s->d_spins, 0, s->bytes, ctx->stream));
    if (launch_strip_init_random(s->d_spins, s->g, (uint32_t)seed, (uint32_t)(seed >> 32), ctx->stream) < 0)
        return fail(ctx, ISING_E_CUDA, "strip init launch failed");
    s->launches++;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = s.release();
    return ISING_OK;
}

extern "C" int ising_strip_create(ising_ctx* ctx, uint64_t Lx, uint64_t Ly, uint64_t row_lo,
                                  uint64_t row_hi, double j, uint64_t seed, ising_strip** out) {
    return ising_strip_create_ex(ctx, Lx, Ly, row_lo, row_hi, j, seed, 1, out);
}

extern "C" void ising_strip_destroy(ising_strip* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    ctx_buf_put(s->ctx, s->d_spins, s->bytes);
    ctx_buf_put(s->ctx, s->d_acc, 2 * sizeof(unsigned long long));
    delete s;
}

extern "C" int ising_strip_configure(ising_strip* s, int planes, int rounds) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "strip is NULL");
    if (planes) {
        if (planes < 5 || planes > 7) return fail(s->ctx, ISING_E_INVALID, "planes must be 5..7");
        s->planes = planes;
    }
    if (rounds) {
        if (rounds != 7 && rounds != 10) return fail(s->ctx, ISING_E_INVALID, "rounds must be 7 or 10");
        s->rounds = rounds;
    }
    return ISING_OK;
}

extern "C" int ising_strip_set_all(ising_strip* s, int up) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "strip is NULL");
    CUDA_TRY(s->ctx, cudaSetDevice(s->ctx->device));
    CUDA_TRY(s->ctx, cudaMemsetAsync(s->d_spins, up ? 0xFF : 0x00, s->bytes, s->ctx->stream));
    CUDA_TRY(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return ISING_OK;
}

// Local rows [r0, r1) of one colour phase; the ghost rows of the OTHER colour must hold the
// neighbours' boundary rows when r0 == 0 or r1 == rows.  sync = 0 only enqueues (no host wait,
// no event timing); advance != 0 bumps the sweep counter (call it on the last piece of colour 1).
static int strip_phase_storage_rows(ising_strip* s, int colour, double beta, uint64_t r0, uint64_t r1,
                                    int advance, int sync) {
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    StripSweepArgs a;
    a.spins = s->d_spins;
    a.g = s->g;
    a.colour = (uint32_t)colour;
    a.sweep = (uint32_t)s->sweep;
    a.key0 = (uint32_t)s->seed;
    a.key1 = (uint32_t)(s->seed >> 32);
    a.antiferro = s->j > 0 ? 0xFFFFFFFFu : 0u;
    a.planes = s->planes;
    a.rounds = s->rounds;
    a.r_begin = (uint32_t)r0;
    a.r_count = (uint32_t)(r1 - r0);
    memset(&a.th, 0, sizeof a.th);
    for (int c = 0; c < 2; ++c) {
        const uint64_t T = threshold64(beta, 4.0 * (c + 1) * fabs(s->j), s->planes);
        for (int pl = 0; pl < s->planes; ++pl)
            a.th.plane[c][pl] = ((T >> (s->planes + 31 - pl)) & 1ull) ? 0xFFFFFFFFu : 0u;
        a.th.low[c] = (uint32_t)(T & 0xFFFFFFFFull);
    }
    if (sync) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (launch_strip_phase(a, ctx->stream) < 0)
        return fail(ctx, ISING_E_CUDA, "strip phase launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (a.r_count) s->launches++;
    if (sync) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        s->device_ms += ms;
    }
    if (advance) s->sweep++;
    return ISING_OK;
}

extern "C" int ising_strip_phase_rows(ising_strip* s, int colour, double beta, uint64_t r0, uint64_t r1,
                                      int advance, int sync) {
    if (!s || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad strip/colour");
    if (r0 > r1 || r1 > s->g.rows) return fail(s->ctx, ISING_E_INVALID, "bad row range");
    return strip_phase_storage_rows(s, colour, beta, s->g.ghost + r0, s->g.ghost + r1, advance, sync);
}

// The local rows plus `ext` ghost rows on each side (ext < ghost): the redundant update of
// ghost rows reproduces the neighbour's bits (Philox is keyed by the global row), so that after
// one deep exchange of 2k rows a strip can run k sweeps without communicating: phase q of the
// batch (q = 0 .. 2k-1) is called with ext = 2k - 1 - q.
extern "C" int ising_strip_phase_ext(ising_strip* s, int colour, double beta, uint32_t ext, int advance,
                                     int sync) {
    if (!s || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad strip/colour");
    if (ext >= s->g.ghost) return fail(s->ctx, ISING_E_INVALID, "ext must be < ghost depth");
    return strip_phase_storage_rows(s, colour, beta, s->g.ghost - ext, s->g.ghost + s->g.rows + ext,
                                    advance, sync);
}

// one whole colour phase, blocking; the sweep counter advances after colour 1
extern "C" int ising_strip_phase(ising_strip* s, int colour, double beta) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "strip is NULL");
    return ising_strip_phase_rows(s, colour, beta, 0, s->g.rows, colour == 1, 1);
}

// storage row r of a colour (local row l is r = ghost + l)
static uint32_t* strip_row_ptr(ising_strip* s, int colour, uint32_t r) {
    return s->d_spins + ((size_t)colour * (s->g.rows + 2 * s->g.ghost) + r) * s->g.Wr;
}

// Deep halo staging, both colours at once.  buf = uint32[2 sides][2 colours][depth][Lx/64]
// (host or device memory).  dir = 0: side 0 <- the first `depth` local rows, side 1 <- the last
// `depth` local rows; dir = 1: side 0 -> the `depth` ghost rows above the first local row,
// side 1 -> the ghost rows below the last one.  Rows are in increasing global order.  A strip
// sends side 0 to the strip above and side 1 to the strip below and receives the upper
// neighbour's side 1 into its side 0.  sync = 0 only enqueues.
extern "C" int ising_strip_halo_deep(ising_strip* s, int dir, uint32_t depth, void* buf, int sync) {
    if (!s || !buf) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    const StripGeom& g = s->g;
    if (depth < 1 || depth > g.ghost || depth > g.rows)
        return fail(ctx, ISING_E_INVALID, "depth must be 1..min(ghost, rows)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)depth * g.Wr * 4;
    uint8_t* b = (uint8_t*)buf;
    for (int side = 0; side < 2; ++side)
        for (int c = 0; c < 2; ++c) {
            uint8_t* slot = b + (size_t)(side * 2 + c) * nb;
            if (dir == 0) {
                const uint32_t r = side ? g.ghost + g.rows - depth : g.ghost;
                CUDA_TRY(ctx, cudaMemcpyAsync(slot, strip_row_ptr(s, c, r), nb, cudaMemcpyDefault, ctx->stream));
            } else {
                const uint32_t r = side ? g.ghost + g.rows : g.ghost - depth;
                CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, c, r), slot, nb, cudaMemcpyDefault, ctx->stream));
            }
        }
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

// single strip covering the whole lattice: periodic wrap of `depth` rows of both colours
extern "C" int ising_strip_wrap_deep(ising_strip* s, uint32_t depth) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "strip is NULL");
    ising_ctx* ctx = s->ctx;
    const StripGeom& g = s->g;
    if (depth < 1 || depth > g.ghost || depth > g.rows)
        return fail(ctx, ISING_E_INVALID, "depth must be 1..min(ghost, rows)");
    if (g.rows != g.Ly) return fail(ctx, ISING_E_INVALID, "wrap needs a strip that holds the whole lattice");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)depth * g.Wr * 4;
    for (int c = 0; c < 2; ++c) {
        CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, c, g.ghost - depth), strip_row_ptr(s, c, g.ghost + g.rows - depth),
                                      nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, c, g.ghost + g.rows), strip_row_ptr(s, c, g.ghost), nb,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return ISING_OK;
}

// which = 0: first local row, 1: last local row.  dst holds Lx/64 words, host or device memory.
extern "C" int ising_strip_get_boundary(ising_strip* s, int colour, int which, void* dst) {
    if (!s || !dst || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, strip_row_ptr(s, colour, which ? s->g.ghost + s->g.rows - 1 : s->g.ghost),
                                  (size_t)s->g.Wr * 4, cudaMemcpyDefault, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

// Device-to-device halo staging without a host wait, for contexts created on the caller's
// stream: dir = 0 copies both boundary rows of `colour` into buf_dev[0..Wr) (first row) and
// buf_dev[Wr..2Wr) (last row); dir = 1 copies buf_dev[0..Wr) into the ghost row above the first
// row and buf_dev[Wr..2Wr) into the ghost row below the last row.
extern "C" int ising_strip_halo_async(ising_strip* s, int colour, int dir, void* buf_dev) {
    if (!s || !buf_dev || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)s->g.Wr * 4;
    uint8_t* b = (uint8_t*)buf_dev;
    if (dir == 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(b, strip_row_ptr(s, colour, s->g.ghost), nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(b + nb, strip_row_ptr(s, colour, s->g.ghost + s->g.rows - 1), nb, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, colour, s->g.ghost - 1), b, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, colour, s->g.ghost + s->g.rows), b + nb, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return ISING_OK;
}

// which = 0: ghost row above the first local row, 1: ghost row below the last local row
extern "C" int ising_strip_set_ghost(ising_strip* s, int colour, int which, const void* src) {
    if (!s || !src || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, colour, which ? s->g.ghost + s->g.rows : s->g.ghost - 1), src,
                                  (size_t)s->g.Wr * 4, cudaMemcpyDefault, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

// single strip covering the whole lattice: periodic wrap of its own boundary rows
extern "C" int ising_strip_wrap_local(ising_strip* s, int colour) {
    if (!s || colour < 0 || colour > 1) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)s->g.Wr * 4;
    CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, colour, s->g.ghost - 1), strip_row_ptr(s, colour, s->g.ghost + s->g.rows - 1), nb,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(strip_row_ptr(s, colour, s->g.ghost + s->g.rows), strip_row_ptr(s, colour, s->g.ghost), nb,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    return ISING_OK;
}

// local sums: satisfied bonds (colour-0 sites see every bond once; needs colour-1 ghosts) and up spins
extern "C" int ising_strip_observables(ising_strip* s, uint64_t* nsat, uint64_t* up) {
    if (!s || !nsat || !up) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemsetAsync(s->d_acc, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (launch_strip_observables(s->d_spins, s->g, s->j > 0 ? 0xFFFFFFFFu : 0u, s->d_acc, ctx->stream) < 0)
        return fail(ctx, ISING_E_CUDA, "strip observables launch failed");
    s->launches++;
    unsigned long long h[2];
    CUDA_TRY(ctx, cudaMemcpyAsync(h, s->d_acc, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *nsat = h[0];
    *up = h[1];
    return ISING_OK;
}

extern "C" int ising_strip_get_rows(ising_strip* s, uint8_t* rows_out) {
    if (!s || !rows_out) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "bad argument");
    ising_ctx* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)s->g.rows * s->Lx;
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 0, bytes, &dv));
    if (launch_strip_unpack(s->d_spins, s->g, (uint8_t*)dv, ctx->stream) < 0)
        return fail(ctx, ISING_E_CUDA, "strip unpack launch failed");
    s->launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(rows_out, dv, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

extern "C" int ising_strip_get_stats(ising_strip* s, uint64_t* launches, double* device_ms, int reset) {
    if (!s) return fail(nullptr, ISING_E_INVALID, "strip is NULL");
    if (launches) *launches = s->launches;
    if (device_ms) *device_ms = s->device_ms;
    if (reset) { s->launches = 0; s->device_ms = 0.0; }
    return ISING_OK;
}

// ------------------------------------------------------------------------------------------
// one blocking call per pymethod
// ------------------------------------------------------------------------------------------
static int check_run_args(ising_ctx* ctx, const ising_graph* g, const ising_run_args* a,
                          const void* energies, const void* states) {
    if (!ctx || !g || !a) return fail(ctx, ISING_E_INVALID, "ctx/graph/args is NULL");
    if (a->struct_size != sizeof(ising_run_args))
        return fail(ctx, ISING_E_INVALID, "ising_run_args.struct_size mismatch (%u != %zu)",
                    a->struct_size, sizeof(ising_run_args));
    if (a->flags & ISING_FLAG_EDGE_IMPORTANCE)
        return fail(ctx, ISING_E_UNSUPPORTED,
                    "edge_move_importance_sampling only affects the reference's non-basic edge "
                    "moves, which the GPU path does not perform");
    if (a->num_experiments && (!energies || !states))
        return fail(ctx, ISING_E_INVALID, "output buffers are NULL");
    return ISING_OK;
}

static int make_sim_for_run(ising_ctx* ctx, const ising_graph* g, const ising_run_args* a,
                            ising_sim** sim) {
    int rc = ising_sim_create(ctx, g, a->num_experiments, a->seed, a->replica_offset, sim);
    if (rc) return rc;
    if (a->initial_state) rc = ising_sim_set_state(*sim, a->initial_state);
    if (rc) { ising_sim_destroy(*sim); *sim = nullptr; }
    return rc;
}

extern "C" int ising_run_monte_carlo(ising_ctx* ctx, const ising_graph* g,
                                     const ising_run_args* a, double* energies, uint8_t* states) {
    int rc = check_run_args(ctx, g, a, energies, states);
    if (rc) return rc;
    if (a->num_experiments == 0) return ISING_OK;
    ising_sim* sim = nullptr;
    rc = make_sim_for_run(ctx, g, a, &sim);
    if (rc) return rc;
    std::vector<double> betas(a->timesteps, a->beta);
    rc = ising_sim_sweeps(sim, betas.data(), a->timesteps, nullptr);
    if (rc == ISING_OK) rc = ising_sim_get_energies(sim, energies);
    if (rc == ISING_OK) rc = ising_sim_get_states(sim, states);
    ising_sim_destroy(sim);
    return rc;
}

static cudaError_t ctx_copy_stream(ising_ctx* ctx) {
    if (ctx->copy_stream) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaEventCreateWithFlags(&ctx->ev_filled[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_drained[b], cudaEventDisableTiming);
    }
    return e;
}

// rows of `width` bytes, device (pitch spitch) to host (pitch dpitch); the 2D copy engine path
// is limited to pitches below 2^31, longer rows go one by one
static cudaError_t copy_rows_d2h(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                                 size_t height, cudaStream_t st) {
    if (width == 0 || height == 0) return cudaSuccess;
    if (dpitch < (1ull << 31) && spitch < (1ull << 31))
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, st);
    for (size_t r = 0; r < height; ++r) {
        cudaError_t e = cudaMemcpyAsync((char*)dst + r * dpitch, (const char*)src + r * spitch, width,
                                        cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// thermalise, then n_s x (sampling_freq sweeps, copy state, energy): lattice.rs:271-287 and
// classicising.rs:146-171 on a device-resident sim.  energies[E, n_s], states[E, n_s, nvars].
// At scale: samples are unpacked into one of two device slabs laid out
// [E, nk, N]; while the sweeps of the next slab run, the copy stream drains the previous one
// straight into the caller's [E, ns, N] array with a strided (2D) copy -- no host staging.
extern "C" int ising_sim_run_sampling(ising_sim* sim, double beta, uint64_t thermalization,
                                      uint64_t sampling_freq, uint64_t ns, double* energies,
                                      uint8_t* states) {
    if (!sim) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    ising_ctx* ctx = sim->ctx;
    if (ns && (!energies || !states)) return fail(ctx, ISING_E_INVALID, "output buffers are NULL");
    if (sim->perbeta) return fail(ctx, ISING_E_INVALID, "sampling runs at one beta");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t E = sim->E, N = sim->lay.nvars;
    std::vector<double> betas(std::max<uint64_t>(thermalization, sampling_freq), beta);
    int rc = ising_sim_sweeps(sim, betas.data(), thermalization, nullptr);
    if (rc || ns == 0) return rc;
    CUDA_TRY(ctx, ctx_copy_stream(ctx));
    uint64_t slab_bytes = 1ull << 29;
    if (const char* env = getenv("ISING_SAMPLING_SLAB_BYTES")) slab_bytes = strtoull(env, nullptr, 10);  // test knob
    const uint64_t slab = std::max<uint64_t>(1, std::min<uint64_t>(ns, slab_bytes / std::max<uint64_t>(1, E * N)));
    const int nbuf = slab < ns ? 2 : 1;
    uint8_t* d_st[2] = {nullptr, nullptr};
    double* d_en[2] = {nullptr, nullptr};
    for (int b = 0; b < nbuf; ++b) {
        void* dv = nullptr;
        CUDA_TRY(ctx, ctx_scratch(ctx, b ? 4 : 0, (size_t)E * slab * N, &dv));
        d_st[b] = (uint8_t*)dv;
        CUDA_TRY(ctx, ctx_scratch(ctx, b ? 5 : 2, (size_t)E * slab * sizeof(double), &dv));
        d_en[b] = (double*)dv;
    }
    uint64_t islab = 0;
    for (uint64_t k0 = 0; k0 < ns; k0 += slab, ++islab) {
        const uint64_t nk = std::min(slab, ns - k0);
        const int b = (int)(islab & 1);
        if (islab >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_drained[b], 0));
        for (uint64_t k = 0; k < nk; ++k) {
            rc = ising_sim_sweeps(sim, betas.data(), sampling_freq, nullptr);
            if (rc) return rc;
            count_launch(sim, launch_unpack_states(sim->d_spins, sim->lay, d_st[b] + k * N, E, nk * N,
                                                   ctx->stream));
            rc = sim_energies_to_device(sim, d_en[b], nk, k);
            if (rc) return rc;
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_filled[b], ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_filled[b], 0));
        CUDA_TRY(ctx, copy_rows_d2h(states + k0 * N, (size_t)ns * N, d_st[b], (size_t)nk * N,
                                    (size_t)nk * N, E, ctx->copy_stream));
        CUDA_TRY(ctx, copy_rows_d2h(energies + k0, (size_t)ns * 8, d_en[b], (size_t)nk * 8,
                                    (size_t)nk * 8, E, ctx->copy_stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_drained[b], ctx->copy_stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISING_OK;
}

// The sampling loop without the state read-back (a sample of config 2 is 17 GB of bools):
// per sample the energy, the magnetisation M = sum_i s_i and, when asked for, the overlap
// Q = sum_i s_i^(2p) s_i^(2p+1) of adjacent experiment pairs (the spin-glass order parameter
// when all experiments share the couplings, lattice.rs:199).  Outputs are [E, ns] / [E/2, ns].
extern "C" int ising_sim_run_observables(ising_sim* sim, double beta, uint64_t thermalization,
                                         uint64_t sampling_freq, uint64_t ns, double* energies,
                                         double* mags, double* overlaps) {
    if (!sim) return fail(nullptr, ISING_E_INVALID, "sim is NULL");
    ising_ctx* ctx = sim->ctx;
    if (sim->perbeta) return fail(ctx, ISING_E_INVALID, "sampling runs at one beta");
    if (sampling_freq == 0) return fail(ctx, ISING_E_INVALID, "sampling_freq must be > 0");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t E = sim->E, N = sim->lay.nvars, P = E / 2;
    const size_t cw = (size_t)sim->lay.W * 32;
    std::vector<double> betas(std::max<uint64_t>(thermalization, sampling_freq), beta);
    int rc = ising_sim_sweeps(sim, betas.data(), thermalization, nullptr);
    if (rc || ns == 0) return rc;
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(ns, (1ull << 23) / std::max<uint64_t>(1, E)));
    void* dv = nullptr;
    CUDA_TRY(ctx, ctx_scratch(ctx, 2, (size_t)E * chunk * sizeof(double), &dv));
    double* d_en = (double*)dv;
    CUDA_TRY(ctx, ctx_scratch(ctx, 4, (size_t)E * chunk * sizeof(double), &dv));
    double* d_m = (double*)dv;
    CUDA_TRY(ctx, ctx_scratch(ctx, 5, (size_t)std::max<uint64_t>(P, 1) * chunk * sizeof(double), &dv));
    double* d_q = (double*)dv;
    for (uint64_t k0 = 0; k0 < ns; k0 += chunk) {
        const uint64_t nk = std::min(chunk, ns - k0);
        for (uint64_t k = 0; k < nk; ++k) {
            rc = ising_sim_sweeps(sim, betas.data(), sampling_freq, nullptr);
            if (rc) return rc;
            if (energies) {
                rc = sim_energies_to_device(sim, d_en, nk, k);
                if (rc) return rc;
            }
            if (mags) {
                CUDA_TRY(ctx, cudaMemsetAsync(sim->d_counts, 0, cw * sizeof(unsigned long long), ctx->stream));
                count_launch(sim, launch_count_up(sim->d_spins, sim->lay, sim->d_counts, ctx->stream));
                // M = 2 up - N = -(N - 2 up)
                count_launch(sim, launch_energy_from_nsat(sim->d_counts, E, -1.0, N, 2, d_m, nk, k, ctx->stream));
            }
            if (overlaps && P) {
                CUDA_TRY(ctx, cudaMemsetAsync(sim->d_counts, 0, cw * sizeof(unsigned long long), ctx->stream));
                count_launch(sim, launch_count_up(sim->d_spins, sim->lay, sim->d_counts, ctx->stream, true));
                count_launch(sim, launch_overlap_from_counts(sim->d_counts, P, N, d_q, nk, k, ctx->stream));
            }
        }
        if (energies)
            CUDA_TRY(ctx, copy_rows_d2h(energies + k0, (size_t)ns * 8, d_en, (size_t)nk * 8, (size_t)nk * 8, E,
                                        ctx->stream));
        if (mags)
            CUDA_TRY(ctx, copy_rows_d2h(mags + k0, (size_t)ns * 8, d_m, (size_t)nk * 8, (size_t)nk * 8, E,
                                        ctx->stream));
        if (overlaps && P)
            CUDA_TRY(ctx, copy_rows_d2h(overlaps + k0, (size_t)ns * 8, d_q, (size_t)nk * 8, (size_t)nk * 8, P,
                                        ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ISING_OK;
}

extern "C" int ising_run_monte_carlo_sampling(ising_ctx* ctx, const ising_graph* g,
                                              const ising_run_args* a, double* energies,
                                              uint8_t* states) {
    int rc = check_run_args(ctx, g, a, energies, states);
    if (rc) return rc;
    if (a->sampling_freq == 0)
        return fail(ctx, ISING_E_INVALID, "sampling_freq must be > 0 (the reference divides by it)");
    if (a->num_experiments == 0) return ISING_OK;
    ising_sim* sim = nullptr;
    rc = make_sim_for_run(ctx, g, a, &sim);
    if (rc) return rc;
    rc = ising_sim_run_sampling(sim, a->beta, a->thermalization, a->sampling_freq,
                                a->timesteps / a->sampling_freq, energies, states);
    ising_sim_destroy(sim);
    return rc;
}

extern "C" int ising_run_monte_carlo_annealing(ising_ctx* ctx, const ising_graph* g,
                                               const ising_run_args* a, double* energies,
                                               uint8_t* states) {
    int rc = check_run_args(ctx, g, a, energies, states);
    if (rc) return rc;
    if (a->sched_len && (!a->sched_t || !a->sched_beta))
        return fail(ctx, ISING_E_INVALID, "schedule arrays are NULL");
    std::vector<double> betas(a->timesteps);
    if (!schedule_betas(a->sched_t, a->sched_beta, a->sched_len, a->timesteps,
                        (a->flags & ISING_FLAG_LINEAR_SCHEDULE) != 0, betas.data()))
        return fail(ctx, ISING_E_INVALID, "annealing schedule has fewer than two stops");
    if (a->num_experiments == 0) return ISING_OK;
    ising_sim* sim = nullptr;
    rc = make_sim_for_run(ctx, g, a, &sim);
    if (rc) return rc;
    if (a->flags & ISING_FLAG_PER_STEP_ENERGIES) {
        rc = ising_sim_sweeps(sim, betas.data(), a->timesteps, energies);
    } else {
        rc = ising_sim_sweeps(sim, betas.data(), a->timesteps, nullptr);
        if (rc == ISING_OK) rc = ising_sim_get_energies(sim, energies);
    }
    if (rc == ISING_OK) rc = ising_sim_get_states(sim, states);
    ising_sim_destroy(sim);
    return rc;
}

// ------------------------------------------------------------------------------------------
// replay
// ------------------------------------------------------------------------------------------
extern "C" int ising_replay(ising_ctx* ctx, const ising_graph* g, double beta, uint64_t E,
                            uint64_t A, const uint32_t* sites, const double* u,
                            const uint8_t* init, double* energies, uint8_t* states) {
    if (!ctx || !g) return fail(ctx, ISING_E_INVALID, "ctx/graph is NULL");
    if (E == 0) return ISING_OK;
    if (!init || !energies || !states || (A && (!sites || !u)))
        return fail(ctx, ISING_E_INVALID, "replay buffers are NULL");
    const uint64_t N = g->h.nvars;
    for (uint64_t i = 0; i < E * A; ++i)
        if (sites[i] >= N) return fail(ctx, ISING_E_INVALID, "trace site %u out of range", sites[i]);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_csr_on_device(ctx, const_cast<ising_graph*>(g));
    if (rc) return rc;
    uint32_t* d_sites = nullptr;
    double* d_u = nullptr;
    uint8_t* d_states = nullptr;
    double* d_en = nullptr;
    unsigned int* d_amb = nullptr;
    cudaError_t e = dev_alloc(&d_sites, E * A);
    if (e == cudaSuccess) e = dev_alloc(&d_u, E * A);
    if (e == cudaSuccess) e = dev_alloc(&d_states, E * N);
    if (e == cudaSuccess) e = dev_alloc(&d_en, E);
    if (e == cudaSuccess) e = dev_alloc(&d_amb, 1);
    unsigned int amb = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sites, sites, E * A * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u, E * A * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_states, init, E * N, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_amb, 0, sizeof(unsigned int), ctx->stream);
    if (e == cudaSuccess) {
        ReplayArgs ra;
        ra.E = E; ra.N = N; ra.A = A;
        ra.row = g->d_row; ra.nbr = g->d_nbr; ra.jv = g->d_jv; ra.bias = g->d_bias;
        ra.sites = d_sites; ra.u = d_u; ra.states = d_states; ra.energies = d_en;
        ra.beta = beta; ra.ambiguous = d_amb;
        if (launch_replay(ra, ctx->stream) < 0) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(states, d_states, E * N, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(energies, d_en, E * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&amb, d_amb, sizeof amb, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sites); cudaFree(d_u); cudaFree(d_states); cudaFree(d_en); cudaFree(d_amb);
    CUDA_TRY(ctx, e);
    if (amb)
        return fail(ctx, ISING_E_AMBIGUOUS,
                    "%u replayed decisions had u within 4e-15 of exp(-beta dE); not certified", amb);
    return ISING_OK;
}

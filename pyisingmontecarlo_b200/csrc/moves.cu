// sm_100a kernels: the non-basic moves of a timestep (qmc GraphState::do_time_step with
// only_basic_moves = false; reference call sites src/lattice.rs:205, 272, 278, 366, 452 and
// src/classicising.rs:100-106 with nedgeupdates / nwormupdates).  Their rules live in the
// out-of-tree `qmc` crate and are restated here as Metropolis moves with symmetric (or
// ratio-corrected) proposals, validated against exact enumeration (tests/test_gpu_moves.py):
//
//   edge move  flip both spins of a bond (a, b).  dE = sum over the bonds that join a or b to the
//              rest of the graph (bonds between a and b keep their energy) + the two bias terms;
//              accept with min(1, exp(-beta dE)).  One pass attempts every bond once, a class of a
//              strong edge colouring per launch (graph.h: EdgeClasses).  Importance sampling
//              (enable_edge_importance_sampling, lattice.rs:200) weights the attempt rate of a
//              bond by |J| / max |J| (a state-independent thinning, folded into the threshold).
//   worm move  a self-avoiding chain of `len` sites grown from a uniformly random site along
//              uniformly random adjacency entries (aborted when it bites itself), flipped as a
//              whole with min(1, deg(first) / deg(last) * exp(-beta dE)): the degree ratio is the
//              proposal probability of the reversed chain over that of the chain.  len = 1 is
//              the reference's random-site single-spin attempt.  One thread per experiment.
//
// Both work on either spin layout (site_word_base) and any couplings / biases: float local
// fields per replica bit, __expf, one 32-bit uniform per decision, as k_sweep_real.
#include "msc_device.cuh"

namespace ising {

template <int ROUNDS>
__global__ void __launch_bounds__(256) k_edge_moves(EdgeMoveArgs a) {
    const uint32_t W = a.lay.W;
    for (uint32_t i = blockIdx.x * blockDim.y + threadIdx.y; i < a.count; i += gridDim.x * blockDim.y) {
        const uint32_t na = a.ea[i], nb = a.eb[i], eid = a.eid[i];
        const size_t ba = site_word_base(a.lay, na), bb = site_word_base(a.lay, nb);
        const float wrel = a.wrel ? a.wrel[i] : 1.f;
        const float bias_a = a.g.biasf[na], bias_b = a.g.biasf[nb];
        const uint32_t lo_a = a.g.row[na], hi_a = a.g.row[na + 1];
        const uint32_t lo_b = a.g.row[nb], hi_b = a.g.row[nb + 1];
        for (uint32_t w = threadIdx.x; w < W; w += blockDim.x) {
            const uint32_t sa = a.spins[ba + w], sb = a.spins[bb + w];
            uint32_t flip = 0;
            for (uint32_t b0 = 0; b0 < 32; b0 += 4) {
                const u32x4 r = philox4x32<ROUNDS>(eid, a.gw0 + w, a.sweep,
                                                   (b0 >> 2) | (a.pass << 8) | (TAG_EDGE << 24), a.key0, a.key1);
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
                float ha[4] = {0.f, 0.f, 0.f, 0.f}, hb[4] = {0.f, 0.f, 0.f, 0.f};
                for (uint32_t k = lo_a; k < hi_a; ++k) {
                    const uint32_t v = a.g.nbr[k];
                    if (v == nb) continue;
                    const uint32_t x = a.spins[site_word_base(a.lay, v) + w] >> b0;
                    const uint32_t jb = __float_as_uint(a.g.jf[k]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) ha[q] += __uint_as_float(jb ^ ((~(x >> q) & 1u) << 31));
                }
                for (uint32_t k = lo_b; k < hi_b; ++k) {
                    const uint32_t v = a.g.nbr[k];
                    if (v == na) continue;
                    const uint32_t x = a.spins[site_word_base(a.lay, v) + w] >> b0;
                    const uint32_t jb = __float_as_uint(a.g.jf[k]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) hb[q] += __uint_as_float(jb ^ ((~(x >> q) & 1u) << 31));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float sia = ((sa >> (b0 + q)) & 1u) ? 1.f : -1.f;
                    const float sib = ((sb >> (b0 + q)) & 1u) ? 1.f : -1.f;
                    const float de = 2.f * sia * (bias_a - ha[q]) + 2.f * sib * (bias_b - hb[q]);
                    const float p = (de > 0.f ? __expf(-a.beta * de) : 1.f) * wrel;
                    // p == 1 always accepts (the saturating conversion alone would miss r = 2^32 - 1)
                    const bool acc = p >= 1.f || rr[q] < __float2uint_rz(p * 4294967296.f);
                    if (acc) flip |= 1u << (b0 + q);
                }
            }
            a.spins[ba + w] = sa ^ flip;
            a.spins[bb + w] = sb ^ flip;
        }
    }
}

int launch_edge_moves(const EdgeMoveArgs& a, cudaStream_t st) {
    if (a.count == 0) return 0;
    const uint32_t wx = a.lay.W >= 32 ? 32 : pow2_ceil(a.lay.W);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)a.count + block.y - 1) / block.y;
    if (blocks > (uint64_t)device_sms() * 16) blocks = (uint64_t)device_sms() * 16;
    if (a.rounds == 7) k_edge_moves<7><<<(unsigned)blocks, block, 0, st>>>(a);
    else k_edge_moves<10><<<(unsigned)blocks, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// word m of the worm's random stream: counter (experiment, worm, sweep, call | TAG_WORM << 24)
template <int ROUNDS>
struct WormStream {
    uint32_t e, k, sweep, key0, key1;
    u32x4 cur;
    uint32_t call = 0xFFFFFFFFu;
    __device__ __forceinline__ uint32_t word(uint32_t m) {
        if ((m >> 2) != call) {
            call = m >> 2;
            cur = philox4x32<ROUNDS>(e, k, sweep, call | (TAG_WORM << 24), key0, key1);
        }
        const uint32_t j = m & 3u;
        return j == 0 ? cur.x : (j == 1 ? cur.y : (j == 2 ? cur.z : cur.w));
    }
};

template <int ROUNDS>
__global__ void __launch_bounds__(128) k_worm_moves(WormArgs a) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    const uint32_t w = (uint32_t)(e >> 5), bit = (uint32_t)(e & 31u);
    const uint64_t ge = a.replica_offset + e;
    // own bit of a site: other threads change other bits of the same word with atomics, so read
    // past L1 (this thread's own atomics must be seen by its later reads as well)
    auto spin = [&](uint32_t n) -> float {
        return ((__ldcg(a.spins + site_word_base(a.lay, n) + w) >> bit) & 1u) ? 1.f : -1.f;
    };
    for (uint32_t k = 0; k < a.nworms; ++k) {
        WormStream<ROUNDS> rs;
        rs.e = (uint32_t)ge; rs.k = k + a.worm0; rs.sweep = a.sweep; rs.key0 = a.key0; rs.key1 = a.key1;
        uint32_t path[WORM_MAX_LEN];
        const uint64_t r64 = ((uint64_t)rs.word(0) << 32) | rs.word(1);
        path[0] = (uint32_t)__umul64hi(r64, a.lay.nvars);
        const uint32_t u = rs.word(2);
        bool ok = true;
        for (uint32_t t = 1; t < a.len && ok; ++t) {
            const uint32_t head = path[t - 1];
            const uint32_t lo = a.g.row[head], deg = a.g.row[head + 1] - lo;
            if (deg == 0) { ok = false; break; }
            const uint32_t nxt = a.g.nbr[lo + __umulhi(rs.word(2 + t), deg)];
            for (uint32_t q = 0; q < t; ++q) ok = ok && path[q] != nxt;
            path[t] = nxt;
        }
        if (!ok) continue;
        float de = 0.f;
        for (uint32_t t = 0; t < a.len; ++t) {
            const uint32_t n = path[t];
            float h = 0.f;
            for (uint32_t kk = a.g.row[n]; kk < a.g.row[n + 1]; ++kk) {
                const uint32_t v = a.g.nbr[kk];
                bool inside = false;
                for (uint32_t q = 0; q < a.len; ++q) inside = inside || path[q] == v;
                if (!inside) h += a.g.jf[kk] * spin(v);
            }
            de += 2.f * spin(n) * (a.g.biasf[n] - h);
        }
        const uint32_t d_first = a.g.row[path[0] + 1] - a.g.row[path[0]];
        const uint32_t d_last = a.g.row[path[a.len - 1] + 1] - a.g.row[path[a.len - 1]];
        const float ratio = a.len > 1 ? (float)d_first / (float)d_last : 1.f;
        const float p = ratio * __expf(-a.beta * de);
        if (p >= 1.f || u < __float2uint_rz(p * 4294967296.f)) {
            for (uint32_t t = 0; t < a.len; ++t)
                atomicXor(a.spins + site_word_base(a.lay, path[t]) + w, 1u << bit);
        }
    }
}

int launch_worm_moves(const WormArgs& a, cudaStream_t st) {
    if (a.E == 0 || a.nworms == 0) return 0;
    if (a.len < 1 || a.len > (uint32_t)WORM_MAX_LEN) return -1;
    const unsigned blocks = (unsigned)((a.E + 127) / 128);
    if (a.rounds == 7) k_worm_moves<7><<<blocks, 128, 0, st>>>(a);
    else k_worm_moves<10><<<blocks, 128, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ising

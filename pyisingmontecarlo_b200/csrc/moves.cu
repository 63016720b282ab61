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

// ------------------------------------------------------------------------------------------
// Edge moves on graphs whose couplings all have the same magnitude and that carry no bias: the
// pair (a, b) sees D = deg(a) + deg(b) - 2 (bonds between a and b) outer bonds, flipping both spins
// turns its n_sat satisfied outer bonds into D - n_sat, so dE = 2|J|(2 n_sat - D) - the integer
// classes of a site of degree D.  Same bit-sliced machinery as k_sweep_general (4-plane counter,
// thresholds of degree D, K planes + resolver words, ties in ascending bit position), exact
// integer thresholds, hence bit-exact against oracle/msc_mirror.c (msc_mirror_moves).  Philox
// counter = (edge index, replica word, timestep, call | pass << 8 | TAG_EDGE << 24).
// Spins are addressed by SLOT (word base / W), computed on the host for the sim's layout.
// ------------------------------------------------------------------------------------------
template <int K, int ROUNDS>
__device__ __forceinline__ uint32_t edge_flip_mask(const uint32_t (&cnt)[4], uint32_t deg, uint32_t eid, uint32_t gw,
                                                   uint32_t sweep, uint32_t tagw, const PhiloxKeys& pk,
                                                   const GenThresholds& th) {
    constexpr int NCALL = K / 4 + 1;
    const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
    uint32_t oh[GEN_MAX_CLS];
    uint32_t up = 0;
#pragma unroll
    for (int j = 0; j < GEN_MAX_CLS; ++j) {
        oh[j] = 0;
        if ((uint32_t)j < ncls) {
            const uint32_t val = cmin + j;
            uint32_t o = 0xFFFFFFFFu;
#pragma unroll
            for (int l = 0; l < 4; ++l) o &= ((val >> l) & 1u) ? cnt[l] : ~cnt[l];
            oh[j] = o;
            up |= o;
        }
    }
    uint32_t r[NCALL * 4];
#pragma unroll
    for (int q = 0; q < NCALL; ++q) {
        const u32x4 o = philox4x32_keys<ROUNDS>(eid, gw, sweep, (uint32_t)q | tagw, pk);
        r[4 * q + 0] = o.x; r[4 * q + 1] = o.y; r[4 * q + 2] = o.z; r[4 * q + 3] = o.w;
    }
    uint32_t eq = up, borrow = 0;
#pragma unroll
    for (int p = K - 1; p >= 0; --p) {
        uint32_t t = 0;
#pragma unroll
        for (int j = 0; j < GEN_MAX_CLS; ++j)
            if ((uint32_t)j < ncls) t |= oh[j] & th.plane[j][p];
        borrow = maj3(~r[p], t, borrow);
        eq &= ~(r[p] ^ t);
    }
    uint32_t flip = ~up | (borrow & ~eq);
    int jj = K;   // tied bits in ascending position: resolver words K, K + 1, ...
    u32x4 cur = {r[4 * (NCALL - 1)], r[4 * (NCALL - 1) + 1], r[4 * (NCALL - 1) + 2], r[4 * (NCALL - 1) + 3]};
    while (eq) {
        const int b = __ffs((int)eq) - 1;
        if ((jj & 3) == 0 && jj >= 4 * NCALL)
            cur = philox4x32_more(cur, (uint32_t)(ROUNDS + (jj >> 2) - NCALL), pk.k[0], pk.k[1]);
        const int m = jj & 3;
        const uint32_t val = m == 0 ? cur.x : (m == 1 ? cur.y : (m == 2 ? cur.z : cur.w));
        uint32_t cls = 0;
#pragma unroll
        for (int j = 1; j < GEN_MAX_CLS; ++j)
            if ((oh[j] >> b) & 1u) cls = j;
        if (val < th.low[cls]) flip |= 1u << b;
        eq &= eq - 1;
        ++jj;
    }
    return flip;
}

// DEG > 0: compile-time number of outer bonds (the gathers are unrolled and in flight together:
// the pass is bound by gather latency, like k_sweep_general); DEG = 0: runtime
template <int K, int ROUNDS, int V, int DEG>
__global__ void __launch_bounds__(256)
k_edge_general(uint32_t* __restrict__ spins, EdgeGroup g, uint32_t W, uint32_t sweep, uint32_t pass, PhiloxKeys pk,
               uint32_t gw0, GenThresholds th) {
    const uint32_t tagw = (pass << 8) | (TAG_EDGE << 24);
    const uint32_t deg = DEG > 0 ? (uint32_t)DEG : g.deg;
    // programmatic dependent launch: a pass is one small launch per (class, outer degree) group
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t i = blockIdx.x * blockDim.y + threadIdx.y; i < g.count; i += gridDim.x * blockDim.y)
    for (uint32_t w0 = threadIdx.x * V; w0 < W; w0 += blockDim.x * V) {
        const uint32_t ab = g.anti[i], eb = g.endp[i], eid = g.eid[i];
        uint32_t* pa = spins + (size_t)g.sa[i] * W + w0;
        uint32_t* pb = spins + (size_t)g.sb[i] * W + w0;
        uint32_t sa[V], sb[V];
        load_words<V>(pa, sa);
        load_words<V>(pb, sb);
        uint32_t cntv[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) cntv[v][l] = 0;
        auto add_bond = [&](uint32_t k, const uint32_t (&x)[V]) {
            const uint32_t m = 0u - ((ab >> k) & 1u);
            const uint32_t e = 0u - ((eb >> k) & 1u);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const uint32_t s = (sa[v] & ~e) | (sb[v] & e);   // the end of the pair this bond hangs on
                uint32_t c = ~(s ^ x[v] ^ m);                    // satisfied bond
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t t = cntv[v][l] & c;
                    cntv[v][l] ^= c;
                    c = t;
                }
            }
        };
        if constexpr (DEG > 0) {
            uint32_t x[DEG][V];
#pragma unroll
            for (int k = 0; k < DEG; ++k)
                load_words<V>(spins + (size_t)g.nbr[(size_t)k * g.count + i] * W + w0, x[k]);
#pragma unroll
            for (int k = 0; k < DEG; ++k) add_bond((uint32_t)k, x[k]);
        } else {
            for (uint32_t k = 0; k < deg; ++k) {
                uint32_t x[V];
                load_words<V>(spins + (size_t)g.nbr[(size_t)k * g.count + i] * W + w0, x);
                add_bond(k, x);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t flip = edge_flip_mask<K, ROUNDS>(cntv[v], deg, eid, gw0 + w0 + v, sweep, tagw, pk, th);
            sa[v] ^= flip;
            sb[v] ^= flip;
        }
        store_words<V>(pa, sa);
        store_words<V>(pb, sb);
    }
}

template <int K, int ROUNDS, int DEG>
static int edge_general_launch(const GenSweepArgs& a, const EdgeGroup& g, uint32_t pass, cudaStream_t st) {
    const bool v2 = a.W % 2 == 0;
    const uint32_t groups = v2 ? a.W / 2 : a.W;
    const uint32_t wx = groups >= 32 ? 32 : pow2_ceil(groups);
    const dim3 block(wx, 256 / wx, 1);
    uint64_t blocks = ((uint64_t)g.count + block.y - 1) / block.y;
    if (blocks > (uint64_t)device_sms() * 16) blocks = (uint64_t)device_sms() * 16;
    const PhiloxKeys pk = philox_round_keys(a.key0, a.key1);
    cudaError_t e;
    if (v2) e = launch_pdl_v(k_edge_general<K, ROUNDS, 2, DEG>, dim3((unsigned)blocks), block, 0, st, a.spins, g, a.W, a.sweep,
                             pass, pk, a.gw0, a.th);
    else e = launch_pdl_v(k_edge_general<K, ROUNDS, 1, DEG>, dim3((unsigned)blocks), block, 0, st, a.spins, g, a.W, a.sweep,
                          pass, pk, a.gw0, a.th);
    return e == cudaSuccess ? 1 : -1;
}

// outer degrees of the regular lattices (square 6, cubic 10, 3-regular 4) for the default (K, rounds)
template <int K, int ROUNDS>
static int edge_general_degree(const GenSweepArgs& a, const EdgeGroup& g, uint32_t pass, cudaStream_t st) {
    if constexpr (K == 6 && ROUNDS == kDefaultRounds) {
        if (g.deg == 4) return edge_general_launch<K, ROUNDS, 4>(a, g, pass, st);
        if (g.deg == 6) return edge_general_launch<K, ROUNDS, 6>(a, g, pass, st);
        if (g.deg == 10) return edge_general_launch<K, ROUNDS, 10>(a, g, pass, st);
    }
    return edge_general_launch<K, ROUNDS, 0>(a, g, pass, st);
}

int launch_edge_general(const GenSweepArgs& a, const EdgeGroup& g, uint32_t pass, cudaStream_t st) {
    if (g.count == 0) return 0;
    if (g.deg > (uint32_t)GEN_MAX_DEG || a.planes < 5 || a.planes > 7) return -1;
#define EDGE_ROUNDS(KK) (a.rounds == 7 ? edge_general_degree<KK, 7>(a, g, pass, st) : edge_general_degree<KK, 10>(a, g, pass, st))
    return a.planes == 5 ? EDGE_ROUNDS(5) : (a.planes == 6 ? EDGE_ROUNDS(6) : EDGE_ROUNDS(7));
#undef EDGE_ROUNDS
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

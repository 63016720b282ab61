// Philox4x32 counter-based RNG (Salmon et al., SC'11), host + device.
// Stream layout used by every kernel of this library (DESIGN.md "Randomness"):
//   key = (seed_lo, seed_hi)
//   ctr = (site, global replica word, sweep, call | tag << 24)
// so a draw is a pure function of (seed, site, replica word, sweep) and does not depend on
// the launch geometry, the colour layout or how experiments are sharded over GPUs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ISING_HD __host__ __device__ __forceinline__
#else
#define ISING_HD inline
#endif

namespace ising {

enum : uint32_t {
    TAG_ACCEPT = 0u,  // Metropolis acceptance planes / resolver words
    TAG_INIT = 1u,    // random initial state
    TAG_SWAP = 2u,    // parallel-tempering swap decisions
    TAG_BOND = 3u,    // +-J disorder of ising_graph_torus
    TAG_EDGE = 4u,    // two-spin edge moves: counter (edge, replica word, timestep, call | pass << 8)
    TAG_WORM = 5u,    // worm moves: counter (experiment, worm, timestep, call)
};

struct u32x4 {
    uint32_t x, y, z, w;
};

template <int ROUNDS>
ISING_HD u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    u32x4 out = {c0, c1, c2, c3};
    return out;
}

// One more round on a finished block: round index `round` (0-based) of the same key schedule, so
// philox4x32_more(philox4x32<R>(ctr, key), R, key) == philox4x32<R + 1>(ctr, key).  The tie
// resolver takes its words beyond the calls a site update makes anyway from these continuation
// rounds of its last block: two multiplications instead of a fresh call.
ISING_HD u32x4 philox4x32_more(const u32x4& s, uint32_t round, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    const uint64_t p0 = (uint64_t)M0 * s.x;
    const uint64_t p1 = (uint64_t)M1 * s.z;
    u32x4 out;
    out.x = (uint32_t)(p1 >> 32) ^ s.y ^ (k0 + round * W0);
    out.y = (uint32_t)p1;
    out.z = (uint32_t)(p0 >> 32) ^ s.w ^ (k1 + round * W1);
    out.w = (uint32_t)p0;
    return out;
}

// Round keys precomputed on the host (key schedule k += W per round): passed as a kernel
// parameter they sit in the constant bank and feed LOP3 directly, instead of 2 integer adds per
// round per thread.
struct PhiloxKeys {
    uint32_t k[20];
};

inline PhiloxKeys philox_round_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys pk;
    for (int r = 0; r < 10; ++r) {
        pk.k[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        pk.k[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return pk;
}

template <int ROUNDS>
ISING_HD u32x4 philox4x32_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                               const PhiloxKeys& pk) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ pk.k[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ pk.k[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    u32x4 out = {c0, c1, c2, c3};
    return out;
}

}  // namespace ising

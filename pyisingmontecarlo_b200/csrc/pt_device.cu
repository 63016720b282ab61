// sm_100a kernels of the device-resident replica-exchange cycle (K6): swap decisions, slot
// bookkeeping, time-averaged energies, slot-ordered sample rows.  Everything a tempering cycle
// needs between two sweeps stays on the stream; the host only enqueues.
#include "kernels.h"
#include "msc_device.cuh"
#include "philox.h"
#include "pt_exp.h"

namespace ising {

// One tempering step (TemperingContainer::parallel_tempering_step as driven from
// tempering.rs:191-194): even slot pairs (0,1),(2,3).. then odd pairs (1,2),(3,4)..; the pair
// (a, a+1) exchanges configurations with probability min(1, exp((b_a - b_{a+1})(E_a - E_{a+1}))),
// uniform = Philox4x32-10(seed; a, parity, swap step).  Pairs of one parity are disjoint, so they
// are decided in parallel; one block, barrier between the parities.  The same arithmetic, in the
// same order, as ising_pt_decide_swaps on the host (pt_exp.h makes exp a fixed IEEE sequence).
//   e_all[gidx[c]] = energy of configuration c;  stats = {swap_step, total_swaps,
//   attempts[R] , accepts[R]} (per pair, indexed by its lower slot)
// Afterwards slot_of_replica[e] = slot of local replica bit e (configuration 32 word_lo + e),
// 0 for the padding bits - the indirection the threshold-table kernels read.
__global__ void __launch_bounds__(256)
k_pt_swap(const double* __restrict__ betas, const double* __restrict__ e_all,
          const uint32_t* __restrict__ gidx, uint32_t* __restrict__ slot_of_cfg,
          uint32_t* __restrict__ cfg_of_slot, uint32_t R, uint32_t key0, uint32_t key1,
          unsigned long long* __restrict__ stats, uint32_t* __restrict__ slot_of_replica,
          uint32_t word_lo, uint32_t e32) {
    __shared__ unsigned int s_swaps;
    if (threadIdx.x == 0) s_swaps = 0;
    const uint32_t swap_step = (uint32_t)stats[0];
    __syncthreads();
    for (uint32_t parity = 0; parity < 2; ++parity) {
        for (uint32_t a = parity + 2 * threadIdx.x; a + 1 < R; a += 2 * blockDim.x) {
            const uint32_t ca = cfg_of_slot[a], cb = cfg_of_slot[a + 1];
            const double d = __dmul_rn(__dsub_rn(betas[a], betas[a + 1]),
                                       __dsub_rn(e_all[gidx[ca]], e_all[gidx[cb]]));
            bool acc = true;
            if (d < 0.0) {
                const u32x4 r = philox4x32<10>(a, parity, swap_step, TAG_SWAP << 24, key0, key1);
                const double uu = __dmul_rn(__dadd_rn((double)r.x, 0.5), 1.0 / 4294967296.0);
                acc = uu < pt_exp_nonpos(d);
            }
            stats[2 + a] += 1ull;
            if (acc) {
                cfg_of_slot[a] = cb;
                cfg_of_slot[a + 1] = ca;
                slot_of_cfg[cb] = a;
                slot_of_cfg[ca] = a + 1;
                stats[2 + R + a] += 1ull;
                atomicAdd(&s_swaps, 1u);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] += 1ull;
        stats[1] += (unsigned long long)s_swaps;
    }
    for (uint32_t e = threadIdx.x; e < e32; e += blockDim.x) {
        const uint32_t c = word_lo * 32u + e;
        slot_of_replica[e] = c < R ? slot_of_cfg[c] : 0u;
    }
}

int launch_pt_swap(const double* betas, const double* e_all, const uint32_t* gidx, uint32_t* slot_of_cfg,
                   uint32_t* cfg_of_slot, uint32_t R, uint64_t seed, unsigned long long* stats,
                   uint32_t* slot_of_replica, uint32_t word_lo, uint32_t e32, cudaStream_t st) {
    k_pt_swap<<<1, 256, 0, st>>>(betas, e_all, gidx, slot_of_cfg, cfg_of_slot, R, (uint32_t)seed,
                                 (uint32_t)(seed >> 32), stats, slot_of_replica, word_lo, e32);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// slot_of_replica alone (after the host changed the permutation)
__global__ void k_pt_local_slots(const uint32_t* __restrict__ slot_of_cfg, uint32_t R,
                                 uint32_t* __restrict__ slot_of_replica, uint32_t word_lo, uint32_t e32) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < e32; e += gridDim.x * blockDim.x) {
        const uint32_t c = word_lo * 32u + e;
        slot_of_replica[e] = c < R ? slot_of_cfg[c] : 0u;
    }
}

int launch_pt_local_slots(const uint32_t* slot_of_cfg, uint32_t R, uint32_t* slot_of_replica,
                          uint32_t word_lo, uint32_t e32, cudaStream_t st) {
    k_pt_local_slots<<<(e32 + 255) / 256, 256, 0, st>>>(slot_of_cfg, R, slot_of_replica, word_lo, e32);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// out[s][0..n) = rows[gidx[cfg_of_slot[s]]][0..n): the sampled states in slot order ("the
// configuration currently at beta_s", tempering.rs:195-211); 16-byte vectors when n allows
__global__ void __launch_bounds__(256)
k_pt_gather_rows(const uint8_t* __restrict__ rows, uint64_t n, const uint32_t* __restrict__ gidx,
                 const uint32_t* __restrict__ cfg_of_slot, uint32_t R, uint8_t* __restrict__ out) {
    const uint32_t s = blockIdx.y;
    if (s >= R) return;
    const uint8_t* src = rows + (size_t)gidx[cfg_of_slot[s]] * n;
    uint8_t* dst = out + (size_t)s * n;
    if ((n & 15ull) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 16; i += (uint64_t)gridDim.x * blockDim.x)
            d4[i] = s4[i];
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
}

int launch_pt_gather_rows(const uint8_t* rows, uint64_t n, const uint32_t* gidx, const uint32_t* cfg_of_slot,
                          uint32_t R, uint8_t* out, cudaStream_t st) {
    uint64_t bx = (n / 16 + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > device_sms() * 4u) bx = device_sms() * 4u;
    k_pt_gather_rows<<<dim3((unsigned)bx, R, 1), 256, 0, st>>>(rows, n, gidx, cfg_of_slot, R, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------
// The whole post-sweep part of a tempering cycle in ONE block (integer-class sims, one rank or
// after the all-gather): energies from the satisfied-bond counters (which it zeroes for the next
// cycle), time-averaged energies, the swap step when due, the replica -> slot map and - for the
// checkerboard layout, whose tables are W * 3 warps of work - the bit-sliced threshold tables.
// Replaces memset + energy + copy + accumulate + swap + tables (six launches of ~2.5 us each
// around ten 2.8 us sweeps of an 8^3 lattice).  Same arithmetic in the same order as the separate
// kernels, so results are bit-identical.
// ------------------------------------------------------------------------------------------


__global__ void __launch_bounds__(256) k_pt_cycle(const __grid_constant__ PtCycleArgs a) {
    __shared__ unsigned int s_swaps;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s_swaps = 0;
    // programmatic dependent launch: scheduled while the sweeps before drain, and the sweeps after
    // are scheduled while this block works (both sides wait for their predecessor's completion)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.nsat) {
        for (uint32_t e = tid; e < a.e32; e += blockDim.x) {
            if (e < a.E) {
                const long long v = (long long)a.nbonds - (long long)a.mult * (long long)a.nsat[e];
                const double en = a.scale * (double)v;
                a.e_local[e] = en;
                if (a.identity) a.e_all[e] = en;
            }
            a.nsat[e] = 0ull;
        }
    }
    __syncthreads();
    for (uint32_t s = tid; s < a.R; s += blockDim.x)
        a.acc[s] = __dadd_rn(a.acc[s], __dmul_rn(a.e_all[a.gidx[a.cfg_of_slot[s]]], a.t));
    if (!a.do_swap) return;
    __syncthreads();
    const uint32_t swap_step = (uint32_t)a.stats[0];
    for (uint32_t parity = 0; parity < 2; ++parity) {
        for (uint32_t p = parity + 2 * tid; p + 1 < a.R; p += 2 * blockDim.x) {
            const uint32_t ca = a.cfg_of_slot[p], cb = a.cfg_of_slot[p + 1];
            const double d = __dmul_rn(__dsub_rn(a.betas[p], a.betas[p + 1]),
                                       __dsub_rn(a.e_all[a.gidx[ca]], a.e_all[a.gidx[cb]]));
            bool acc = true;
            if (d < 0.0) {
                const u32x4 r = philox4x32<10>(p, parity, swap_step, TAG_SWAP << 24, a.key0, a.key1);
                const double uu = __dmul_rn(__dadd_rn((double)r.x, 0.5), 1.0 / 4294967296.0);
                acc = uu < pt_exp_nonpos(d);
            }
            a.stats[2 + p] += 1ull;
            if (acc) {
                a.cfg_of_slot[p] = cb;
                a.cfg_of_slot[p + 1] = ca;
                a.slot_of_cfg[cb] = p;
                a.slot_of_cfg[ca] = p + 1;
                a.stats[2 + a.R + p] += 1ull;
                atomicAdd(&s_swaps, 1u);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.stats[0] += 1ull;
        a.stats[1] += (unsigned long long)s_swaps;
    }
    for (uint32_t e = tid; e < a.e32; e += blockDim.x) {
        const uint32_t c = a.word_lo * 32u + e;
        a.slot_of_replica[e] = c < a.R ? a.slot_of_cfg[c] : 0u;
    }
    if (!a.t64) return;
    __syncthreads();
    // threshold tables of the checkerboard kernels: one warp per (word, class), as k_build_tables_stencil
    const uint32_t b = tid & 31u;
    for (uint32_t idx = tid >> 5; idx < a.W * 3; idx += blockDim.x >> 5) {
        const uint32_t cls = idx % 3, w = idx / 3;
        const uint32_t e = w * 32 + b;
        const unsigned long long T = a.t64[(size_t)a.slot_of_replica[e] * 3 + cls];
        a.tlow[(size_t)e * 3 + cls] = (uint32_t)(T & 0xFFFFFFFFull);
        uint32_t mine = 0;
        for (int p = 0; p < 8; ++p) {
            const uint32_t m = p < a.K ? __ballot_sync(0xFFFFFFFFu, (T >> (a.K + 31 - p)) & 1ull) : 0u;
            if (b == (uint32_t)p) mine = m;
        }
        if (b < 8) a.tplane[((size_t)w * 3 + cls) * 8 + b] = mine;
    }
}

int launch_pt_cycle(const PtCycleArgs& a, cudaStream_t st) {
    return launch_pdl(k_pt_cycle, dim3(1), dim3(256), 0, st, a) == cudaSuccess ? 1 : -1;
}

}  // namespace ising

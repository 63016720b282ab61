// Integer-pipe microbenchmark for the sweep kernel's instruction mix (B200, sm_100a).
//
// The Metropolis sweep is bit-sliced logic (LOP3, ALU pipe) + Philox (IMAD.WIDE.U32, FMA-heavy
// pipe) + class selects / address arithmetic (IMAD, FMA pipe).  This program measures what the SM
// sustains for pure streams of each and for the kernel's own mix, with the kernel's launch shape
// (256 threads, 3 blocks per SM), so that bench.py can report "fraction of the instruction bound"
// next to the fraction of the HBM roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu && ./pipe_bench
// Output: one line per mix, warp-instructions per cycle per SM sub-partition (SMSP).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int CHAINS = 8;   // independent dependency chains per thread (the sweep has 8 Philox calls in flight)

// One "step" issues NL LOP3, NW IMAD.WIDE and NI IMAD per chain group, interleaved.
template <int NL, int NW, int NI, int MODE = 0>
__global__ void __launch_bounds__(256, 3) k_mix(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        a[k] = seed + threadIdx.x * 7u + k;
        b[k] = seed * 3u + blockIdx.x + k * 11u;
        c[k] = seed ^ (threadIdx.x << k);
    }
    for (int it = 0; it < iters; ++it) {
        constexpr int NMAX = NL > NW ? (NL > NI ? NL : NI) : (NW > NI ? NW : NI);
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) {
                if (j < NW) {   // 32x32 -> 64, both halves used (a Philox round)
                    if (MODE == 0) {          // IMAD.WIDE.U32
                        const uint64_t p = (uint64_t)a[k] * 0xD2511F53u;
                        a[k] = (uint32_t)(p >> 32) ^ b[k];
                        b[k] = (uint32_t)p;
                    } else {                  // IMAD.HI.U32 + IMAD (MODE 2: high half only)
                        uint32_t hi, lo = a[k];
                        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(a[k]), "r"(0xD2511F53u));
                        if (MODE == 1) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(a[k]), "r"(0xD2511F53u));
                        a[k] = hi ^ b[k];
                        b[k] = lo;
                    }
                }
                if (j < NL) {   // LOP3 with three register operands (majority)
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c[k]) : "r"(c[k]), "r"(a[k]), "r"(b[k]));
                }
                if (j < NI) {   // IMAD (32-bit)
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b[k]) : "r"(b[k]), "r"(c[k] | 1u), "r"(a[k]));
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) r ^= a[k] ^ b[k] ^ c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NL, int NW, int NI, int MODE = 0>
static void run(const char* name, uint32_t* d_out, int sms, double clock_hz) {
    const int iters = 2000, blocks = sms * 3;
    k_mix<NL, NW, NI, MODE><<<blocks, 256>>>(d_out, 1u, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_mix<NL, NW, NI, MODE><<<blocks, 256>>>(d_out, 1u, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // per chain and iteration: NL majority LOP3; NW x (multiply [2 instructions when split] + xor);
    // NI x (IMAD + the OR that makes its multiplier odd)
    const double per_thread = (double)iters * CHAINS * (NL + (MODE == 1 ? 3.0 : 2.0) * NW + 2.0 * NI);
    const double warp_inst = per_thread * blocks * 256 / 32.0;
    const double cycles = best * 1e-3 * clock_hz;
    printf("{\"mix\": \"%s\", \"lop3\": %d, \"imad_wide\": %d, \"imad\": %d, \"ms\": %.4f, "
           "\"warp_inst_per_clk_per_smsp\": %.4f, \"cycles_per_group\": %.3f}\n",
           name, NL + NW + NI, NW, NI, best, warp_inst / cycles / (sms * 4.0),
           cycles * sms * 4.0 / ((double)iters * CHAINS * blocks * 8.0));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clock_hz = khz * 1e3;
    uint32_t* d_out = nullptr;
    cudaMalloc(&d_out, (size_t)prop.multiProcessorCount * 3 * 256 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", prop.name, prop.multiProcessorCount, clock_hz / 1e6);
    run<4, 0, 0>("lop3 only", d_out, prop.multiProcessorCount, clock_hz);
    run<0, 4, 0>("imad.wide + its xor", d_out, prop.multiProcessorCount, clock_hz);
    run<0, 0, 4>("imad only", d_out, prop.multiProcessorCount, clock_hz);
    run<4, 0, 4>("lop3 : imad 1:1", d_out, prop.multiProcessorCount, clock_hz);
    run<2, 2, 0>("lop3 : imad.wide 2:1 (+xor)", d_out, prop.multiProcessorCount, clock_hz);
    run<6, 3, 4>("sweep kernel mix (9 lop3 : 3 imad.wide : 4 imad)", d_out, prop.multiProcessorCount, clock_hz);
    run<6, 3, 0>("round-1 kernel mix (9 lop3 : 3 imad.wide)", d_out, prop.multiProcessorCount, clock_hz);
    run<0, 4, 0, 2>("imad.hi + its xor", d_out, prop.multiProcessorCount, clock_hz);
    run<0, 4, 0, 1>("imad.hi + imad.lo + xor", d_out, prop.multiProcessorCount, clock_hz);
    run<6, 3, 0, 1>("9 lop3 : 3 (imad.hi + imad.lo)", d_out, prop.multiProcessorCount, clock_hz);
    run<6, 3, 4, 1>("9 lop3 : 3 (imad.hi + imad.lo) : 4 imad", d_out, prop.multiProcessorCount, clock_hz);
    run<8, 1, 0, 0>("9 lop3 : 1 imad.wide", d_out, prop.multiProcessorCount, clock_hz);
    run<8, 2, 0, 0>("10 lop3 : 2 imad.wide", d_out, prop.multiProcessorCount, clock_hz);
    cudaFree(d_out);
    return 0;
}

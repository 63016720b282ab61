// exp(d) for d <= 0 as a FIXED sequence of IEEE-754 double operations (no libm, no FMA
// contraction), so that the replica-exchange decisions u < exp((b_a - b_b)(E_a - E_b)) taken on
// the device (k_pt_swap), on the host (ising_pt_decide_swaps) and in the CPU mirror
// (oracle/msc_mirror.c restates the same sequence) are bit-identical.
//   d = k ln2 + r, |r| <= ln2 / 2;  exp(r) by a degree-13 Taylor polynomial in Horner form
//   (truncation error < 4e-18 relative);  2^k assembled from the exponent bits.
// Accuracy ~1 ulp; what matters here is that it is a pure function of the bits of d.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define PTX_MUL(a, b) __dmul_rn((a), (b))
#define PTX_ADD(a, b) __dadd_rn((a), (b))
#else
#define PTX_MUL(a, b) ((a) * (b))
#define PTX_ADD(a, b) ((a) + (b))
#endif

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline double pt_exp_nonpos(double d) {
    if (!(d <= 0.0)) return 1.0;          // callers only ask for d < 0; NaN -> accept nothing below
    if (d < -700.0) return 0.0;
    const double INV_LN2 = 1.4426950408889634074;      // 0x3FF71547652B82FE
    const double LN2_HI = 6.93147180369123816490e-01;  // 0x3FE62E42FEE00000
    const double LN2_LO = 1.90821492927058770002e-10;  // 0x3DEA39EF35793C76
    const double t = PTX_MUL(d, INV_LN2);
    // nearest integer to t (t <= 0): truncate t - 0.5 toward zero
    const long long ki = (long long)PTX_ADD(t, -0.5);
    const double kf = (double)ki;
    double r = PTX_ADD(d, -PTX_MUL(kf, LN2_HI));
    r = PTX_ADD(r, -PTX_MUL(kf, LN2_LO));
    // 1/n!, n = 13 .. 1
    const double c[13] = {1.6059043836821613e-10, 2.0876756987868098e-09, 2.5052108385441720e-08,
                          2.7557319223985888e-07, 2.7557319223985893e-06, 2.4801587301587302e-05,
                          1.9841269841269841e-04, 1.3888888888888889e-03, 8.3333333333333332e-03,
                          4.1666666666666664e-02, 1.6666666666666666e-01, 5.0000000000000000e-01,
                          1.0000000000000000e+00};
    double p = c[0];
    for (int i = 1; i < 13; ++i) p = PTX_ADD(PTX_MUL(p, r), c[i]);
    p = PTX_ADD(PTX_MUL(p, r), 1.0);
    // p * 2^ki, ki in [-1011, 0]: exponent field 1023 + ki stays normal
    const uint64_t bits = (uint64_t)(1023 + ki) << 52;
    double scale;
    memcpy(&scale, &bits, sizeof scale);
    return PTX_MUL(p, scale);
}

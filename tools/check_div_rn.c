// check_div_rn.c — CPU check of the exact fp64 division sequence used by Ar<double>::div_rn
// (prismo_b200/csrc/fdtd_kernels.cuh): q0 = a*y; r0 = fma(-q0,d,a); q1 = fma(r0,y,q0); r1 = fma(-q1,d,a);
// q2 = fma(r1,y,q1) with y = RN(1/d) must equal the IEEE quotient a/d for every a whose quotient is in range.
// Also counts how often the shorter 3-operation variant (q1) would be wrong, to document why five are used.
//   gcc -O2 -mfma -o check_div_rn tools/check_div_rn.c -lm && ./check_div_rn [n_per_divisor]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void)
{
    uint64_t a = s[0], b = s[1];
    s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

int main(int argc, char** argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 20000000;
    double ds[64];
    int nd = 0;
    const double fixed[] = {2e-8, 2.5e-8, 3e-8, 5e-8, 1.0 / 20e6, 1.0 / 40e6, 1.0 / 50e6, 1e-9, 1.25e-8, 4e-8, 1e-7, 1e-6,
                            3.0, 1.0, 0x1.fffffffffffffp-27, 0x1.0000000000001p-26, 7e-9};
    for (size_t i = 0; i < sizeof fixed / sizeof *fixed; ++i) ds[nd++] = fixed[i];
    while (nd < 48) ds[nd++] = from_bits((rnd() & 0x000fffffffffffffull) | ((uint64_t)(1023 - 40 + (int)(rnd() % 30)) << 52));
    long bad5 = 0, bad3 = 0, total = 0;
    for (int di = 0; di < nd; ++di) {
        const double d = ds[di], y = 1.0 / d;
        long b3 = 0, b5 = 0;
        for (long it = 0; it < n; ++it) {
            double a;
            const int mode = (int)(it & 3);
            if (mode == 0) {                 // random significand, moderate exponent
                a = from_bits((rnd() & 0x800fffffffffffffull) | ((uint64_t)(1023 - 200 + (int)(rnd() % 400)) << 52));
            } else if (mode == 1) {          // a = q*d rounded, q random: quotients near representable numbers
                const double q = from_bits((rnd() & 0x000fffffffffffffull) | ((uint64_t)(1023 - 30 + (int)(rnd() % 90)) << 52));
                a = q * d;
                if (rnd() & 1) a = nextafter(a, (rnd() & 1) ? INFINITY : -INFINITY);
            } else if (mode == 2) {          // quotients near midpoints: q + half ulp
                const double q = from_bits((rnd() & 0x000fffffffffffffull) | ((uint64_t)(1023 + (int)(rnd() % 60)) << 52));
                a = fma(q, d, 0.5 * (nextafter(q, INFINITY) - q) * d);
            } else {                         // difference of two O(1) numbers (what the kernel divides)
                const double u = (double)(int64_t)rnd() * 0x1p-63, v = (double)(int64_t)rnd() * 0x1p-63;
                a = u - v;
            }
            const double want = a / d;
            const double q0 = a * y;
            const double r0 = fma(-q0, d, a);
            const double q1 = fma(r0, y, q0);
            const double r1 = fma(-q1, d, a);
            const double q2 = fma(r1, y, q1);
            const double m = fabs(q0);
            if (!(m >= 0x1p-900 && m <= 0x1p900)) continue;
            ++total;
            if (bits(q1) != bits(want)) ++b3;
            if (bits(q2) != bits(want)) ++b5;
        }
        bad3 += b3; bad5 += b5;
        if (b5 || di < 8) printf("d = %.17g: 3-op wrong %ld, 5-op wrong %ld of %ld\n", d, b3, b5, n);
    }
    printf("total %ld quotients over %d divisors: 3-op sequence wrong %ld, 5-op sequence wrong %ld\n", total, nd, bad3, bad5);
    return bad5 ? 1 : 0;
}

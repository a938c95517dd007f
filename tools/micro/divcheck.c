/* CPU check of the division used by the stencil-wavefront triangular solve (csrc/bis_sptrsv_wave.cuh):
 *     r  = RN(1 / d)                       (an IEEE division, off the critical path)
 *     q0 = RN(a * r) ; e0 = fma(-d, q0, a) ; q1 = fma(e0, r, q0) ; e1 = fma(-d, q1, a) ; q = fma(e1, r, q1)
 * must equal RN(a / d) bit for bit (Markstein: a correctly rounded reciprocal and a faithful quotient make the
 * last fma the correctly rounded quotient, except possibly when the significand of d is all ones; the kernel
 * falls back to the IEEE division there and whenever an exponent is far from 0).
 *   gcc -O2 -fopenmp -ffp-contract=off -o tools/micro/divcheck tools/micro/divcheck.c -lm && tools/micro/divcheck 2000000000
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t rng(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double mk(uint64_t mant, int e, int neg) {   /* 1.mant * 2^e */
    uint64_t b = ((uint64_t)neg << 63) | ((uint64_t)(e + 1023) << 52) | (mant & 0xFFFFFFFFFFFFFull);
    double x;
    memcpy(&x, &b, 8);
    return x;
}
static inline int safe(double a, double d) {   /* the guard of the kernel */
    uint64_t ba, bd;
    memcpy(&ba, &a, 8);
    memcpy(&bd, &d, 8);
    int ea = (int)((ba >> 52) & 0x7FF), ed = (int)((bd >> 52) & 0x7FF);
    if (ed < 1023 - 400 || ed > 1023 + 400) return 0;
    if (ea < 1023 - 400 || ea > 1023 + 400) return 0;     /* also zero (the sequence loses the sign of -0) and subnormals */
    if ((bd & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull) return 0;
    return 1;
}
static inline double fastdiv(double a, double d) {
    double r = 1.0 / d;
    double q0 = a * r;
    double e0 = fma(-d, q0, a);
    double q1 = fma(e0, r, q0);
    double e1 = fma(-d, q1, a);
    return fma(e1, r, q1);
}

int main(int argc, char **argv) {
    long long n = argc > 1 ? atoll(argv[1]) : 100000000LL;
    long long bad = 0, tested = 0;
#pragma omp parallel reduction(+ : bad, tested)
    {
        uint64_t s = 12345;
#ifdef _OPENMP
        extern int omp_get_thread_num(void);
        s += 7919ull * (uint64_t)omp_get_thread_num();
#endif
#pragma omp for schedule(static)
        for (long long i = 0; i < n; ++i) {
            uint64_t u = rng(&s), v = rng(&s), w = rng(&s);
            int kind = (int)(w & 7);
            double d = mk(v, (int)((w >> 8) % 61) - 30, (int)(w >> 20) & 1);
            double a;
            if (kind < 3) {
                a = mk(u, (int)((w >> 32) % 121) - 60, (int)(w >> 21) & 1);                 /* unrelated numerator */
            } else if (kind < 6) {
                double q = mk(u, (int)((w >> 32) % 41) - 20, 0);                             /* a = RN(q d) +- k ulp: quotients next to ties */
                a = q * d;
                int k = (int)((w >> 40) % 5) - 2;
                uint64_t ba;
                memcpy(&ba, &a, 8);
                ba += (uint64_t)(int64_t)k;
                memcpy(&a, &ba, 8);
            } else if (kind == 6) {
                d = mk((v & 0xFFFFF) | 0xFFFFFFFF00000ull, (int)((w >> 8) % 21) - 10, 0);    /* long runs of ones in d */
                a = mk(u, (int)((w >> 32) % 21) - 10, 0);
            } else {
                a = (w >> 44) & 1 ? ((w >> 45) & 1 ? -0.0 : 0.0) : mk(u & 0xFFF0000000000ull, (int)((w >> 32) % 21) - 10, 1);   /* zeros / short numerators */
                d = mk(v & 0xFFFFF00000000ull, (int)((w >> 8) % 21) - 10, (int)(w >> 46) & 1);
            }
            if (!safe(a, d)) continue;
            ++tested;
            double q = fastdiv(a, d), ref = a / d;
            if (memcmp(&q, &ref, 8) != 0) {
                ++bad;
                if (bad < 5) printf("MISMATCH a=%a d=%a fast=%a ref=%a\n", a, d, q, ref);
            }
        }
    }
    printf("tested %lld pairs, %lld mismatches\n", tested, bad);
    return bad != 0;
}

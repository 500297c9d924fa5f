/* CPU restatement of the kernels' division-free quotients (dmx_compressor_b200/csrc/dmxq_stages.cuh div_by_recip /
 * div_by_recip2), checked against IEEE division.  Test code only: fmaf is the correctly rounded C99 fused multiply-add,
 * exactly what __fmaf_rn is on the device, so the identity proven here is the one the kernels rely on.
 *   div_by_recip :  q = a*rb;  two Newton corrections q += (a - q*b)*rb
 *   div_by_recip2:  q = fma(a, rh, a*rl) with (rh, rl) the high / low parts of 1/b;  one correction
 * Preconditions (the kernels check them per block / vector): b in (2^-60, 2^60) with a significand that is not all
 * ones; the comparison is on the full quotient for |a| >= 2^-100 and on round-half-away(q) -- what the casts consume -- below
 * (there the residual a - q*b leaves the normal range and the last bit of a quotient < 2^-40 is immaterial).
 *   div_by_recip16:  q = fma(a, rh, a*rl) alone, for dividends of at most 16 significant bits (widened bf16 / fp16 values):
 *                   a/b then stays >= 2^-41 (relative) away from every fp32 rounding boundary, further than the error of the
 *                   estimate, so no correction step is needed.  Checked with 8-, 11- and 16-bit dividends, half of them in the
 *                   neighbourhood of quotient ties (k + 0.5) * b.
 * usage: div_by_recip_check <cases>   -> prints "cases N bad_two_step X bad_hilo Y bad_16bit Z", exit status 0 iff all are 0 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline float div_hilo(float a, float b, float rh, float rl)
{
    float q = fmaf(a, rh, a * rl);
    return fmaf(fmaf(-q, b, a), rh, q);
}
static inline float div_two_step(float a, float b, float rb)
{
    float q = a * rb;
    q = fmaf(fmaf(-q, b, a), rb, q);
    return fmaf(fmaf(-q, b, a), rb, q);
}
int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 20000000L, bad1 = 0, bad2 = 0, cnt = 0;
    for (long i = 0; i < n; ++i) {
        uint32_t bm = (uint32_t)rnd() & 0x7FFFFF;
        if (bm == 0x7FFFFF) continue;
        float b = u2f(((uint32_t)(67 + (int)(rnd() % 120)) << 23) | bm); /* 2^-60 .. 2^60 */
        float rh = 1.0f / b, rl = fmaf(-b, rh, 1.0f) * rh;
        float a;
        switch (i & 3) {
        case 0: a = b * (float)((rnd() % 7000001) / 1000000.0); break;                       /* anywhere in [0, 7b]      */
        case 1: a = u2f(f2u(b * ((float)(rnd() % 8) + 0.5f)) + (uint32_t)((int)(rnd() % 9) - 4)); break; /* around the ties */
        case 2: a = u2f(f2u(b * (float)((rnd() % 7000001) / 1000000.0)) & 0xFFFF0000u); break; /* bf16 significands        */
        default: a = u2f((uint32_t)rnd() & 0x7FFFFFFF); if (!(a < 7.1f * b)) a = b * 0.3f; break; /* any exponent below    */
        }
        if (!(a >= 0) || isinf(a)) continue;
        if (i & 4) a = -a; /* calibrated INT8 divides signed data: every step is odd-symmetric, checked all the same */
        const float want = a / b, g2 = div_hilo(a, b, rh, rl), g1 = div_two_step(a, b, rh);
        ++cnt;
        if (fabsf(a) < 0x1p-100f) {
            if (roundf(g1) != roundf(want)) ++bad1;
            if (roundf(g2) != roundf(want)) ++bad2;
            continue;
        }
        if (f2u(g1) != f2u(want)) ++bad1;
        if (f2u(g2) != f2u(want)) ++bad2;
    }
    long bad3 = 0, cnt3 = 0;
    for (long i = 0; i < n; ++i) {
        const int bits = (i % 3 == 0) ? 8 : (i % 3 == 1) ? 11 : 16;
        const uint32_t keep = ~((1u << (24 - bits)) - 1u);
        uint32_t bm = (uint32_t)rnd() & 0x7FFFFF;
        if (bm == 0x7FFFFF) continue;
        float b = u2f(((uint32_t)(67 + (int)(rnd() % 120)) << 23) | bm);
        float rh = 1.0f / b, rl = fmaf(-b, rh, 1.0f) * rh;
        int ae = (int)((f2u(b) >> 23) & 0xFF) + (int)(rnd() % 40) - 30; /* quotient in 2^-30 .. 2^9 */
        if (ae < 1 || ae > 254) continue;
        float a = u2f(((uint32_t)ae << 23) | (((uint32_t)rnd() & 0x7FFFFF) & keep));
        if (i & 8) { /* around the ties of the quotient, on the dividend's own grid */
            a = u2f(f2u(b * ((float)(rnd() % 300) + 0.5f)) & keep);
            if (rnd() & 1) a = u2f(f2u(a) + (1u << (24 - bits)));
        }
        if (i & 4) a = -a;
        if (!(fabsf(a) < 0x1p60f) || fabsf(a) < 0x1p-100f) continue;
        ++cnt3;
        if (f2u(a / b) != f2u(fmaf(a, rh, a * rl))) ++bad3;
    }
    printf("cases %ld bad_two_step %ld bad_hilo %ld bad_16bit %ld (of %ld)\n", cnt, bad1, bad2, bad3, cnt3);
    return (bad1 || bad2 || bad3) ? 1 : 0;
}

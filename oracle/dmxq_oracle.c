/*
 * dmxq_oracle.c -- CPU restatement of the reference's CastTo / Sparsify numerics.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call this file.  The product path
 * (libdmxq.so, CUDA) never does and has no CPU fallback.
 *
 * Parity pin: this restatement is checked bit-for-bit against (a) the reference's own
 * compiled quant_cpu extension (oracle/_ref/ref_quant_cpu.so, built by oracle/build_ref.py
 * from the sources under /root/reference), (b) golden vectors produced by importing the
 * reference's python (tests/golden/make_golden.py) and (c) the inline known-answer tests of
 * the reference's own test-suite (tests/test_bfp.py:26-65, tests/test_group_quant.py:49-63).
 * See tests/test_oracle.py.
 *
 * Every function cites the reference file:line it follows ("Q/" = src/dmx/compressor/quant/,
 * "S/" = src/dmx/compressor/).  All tensors are contiguous fp32 addressed as
 * (outer, K, inner): blocks are `bs` consecutive k for each (o, i); the last block of a
 * row may be ragged (torch.split semantics, S/numerical/format.py:324-326).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { R_NEAREST = 0, R_STOCHASTIC = 1, R_UP = 2, R_DOWN = 3 };
enum { TIE_AWAY = 0 /* reference CUDA: roundf */, TIE_EVEN = 1 /* reference CPU quirk */ };

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* Q/quant_cpu/quant_cpu.cpp:211-237 (CPU) == Q/quant_cuda/bit_helper.cu:10-46 (CUDA).
 * man_bits >= 23 is undefined behaviour in the reference (negative shift); the CUDA build
 * degenerates to the identity, which is what we define here. */
static uint32_t round_bitwise(uint32_t target, int man_bits, int mode, uint32_t rnd)
{
    if (man_bits >= 23) return target;
    uint32_t mask = (1u << (23 - man_bits)) - 1u;
    uint32_t rand_prob;
    if (mode == R_STOCHASTIC) {
        rand_prob = rnd & mask;
    } else if (mode == R_NEAREST) {
        rand_prob = 1u << (23 - man_bits - 1);
        if ((target & mask) == rand_prob)                 /* exactly half ...            */
            if (((target >> (23 - man_bits)) & 1u) == 0u) /* ... and kept LSB is even     */
                rand_prob = 0u;
    } else if (mode == R_DOWN) {
        rand_prob = 0u;
    } else {
        rand_prob = 1u << (23 - man_bits);
    }
    return (target + rand_prob) & ~mask;
}

/* Q/quant_cpu/bit_helper.cpp:4-22 */
static uint32_t clip_exponent(int exp_bits, int man_bits, uint32_t old_num, uint32_t q)
{
    if (q == 0) return q;
    int q_exp_store = (int)((q << 1) >> 24);
    int max_exp_store = (1 << (exp_bits - 1)) + 127;
    if (q_exp_store > max_exp_store) {
        uint32_t max_man = (((uint32_t)-1 << 9) >> 9) >> (23 - man_bits) << (23 - man_bits);
        uint32_t max_num = ((uint32_t)max_exp_store << 23) | max_man;
        q = (old_num & 0x80000000u) | max_num;
    }
    return q;
}

/* Q/quant_cpu/bit_helper.cpp:24-37 */
static uint32_t clip_max_exponent(int man_bits, uint32_t max_exponent, uint32_t q)
{
    uint32_t q_exp = ((q << 1) >> 24) << 23;
    if (q_exp > max_exponent) {
        uint32_t max_man = (((uint32_t)-1 << 9) >> 9) >> (23 - man_bits) << (23 - man_bits);
        q = (q & 0x80000000u) | max_exponent | max_man;
    }
    return q;
}

/* One element of Q/quant_cpu/quant_cpu.cpp:239-275 (block_quantize_helper) ==
 * Q/quant_cuda/block_kernel.cu:43-74. `max_elem` is the block's max|x| (NaN-propagating). */
static float bfp_elem(float x, float max_elem, int wl, int symmetric, int mode, uint32_t rnd)
{
    uint32_t max_num = f2u(max_elem);
    if (!symmetric) {
        if (x == -max_elem && ((max_num >> 16) << 25) == 0xFE000000u)
            max_num = ((max_num >> 23) + 1u) << 23;
    }
    uint32_t max_exp = ((max_num << 1) >> 24) << 23;
    float base = u2f(max_exp) * 6.0f;
    float t = x + base;
    uint32_t qb = round_bitwise(f2u(t), wl, mode, rnd);
    float q = u2f(qb) - base;
    return u2f(clip_max_exponent(wl - 2, max_exp, f2u(q)));
}

/* NaN-propagating max|x| as torch `abs().max()` (Q/quant_cpu/quant_cpu.cpp:277-297):
 * returned as a float whose bit pattern is the unsigned max of the |x| patterns
 * (NaN patterns order above Inf; only the exponent field is consumed downstream). */
static inline uint32_t absbits(float x) { return f2u(x) & 0x7FFFFFFFu; }

/* L1 `block_quantize(x, wl, dim=0)` on a [rows, len] matrix: one shared exponent per row.
 * Q/quant_function.py:87-117 -> quant_cpu.cpp:299-357. rand: int32 per element or NULL. */
void orc_block_quantize_rows(const float *x, float *y, int64_t rows, int64_t len, int wl,
                             int symmetric, int mode, const int32_t *rnd)
{
    for (int64_t r = 0; r < rows; r++) {
        uint32_t m = 0;
        for (int64_t j = 0; j < len; j++) { uint32_t a = absbits(x[r * len + j]); if (a > m) m = a; }
        float mf = u2f(m);
        for (int64_t j = 0; j < len; j++) {
            int64_t i = r * len + j;
            y[i] = bfp_elem(x[i], mf, wl, symmetric, mode, rnd ? (uint32_t)rnd[i] : 0u);
        }
    }
}

/* BlockFloatingPoint.cast, symmetric part: S/numerical/format.py:322-341 (transpose block
 * dim last, split into chunks of `bs`, block_quantize(dim=0) each chunk, cat, transpose
 * back), restated without the transposes on the (outer, K, inner) addressing. */
void orc_bfp_cast(const float *x, float *y, int64_t outer, int64_t K, int64_t inner, int64_t bs,
                  int wl, int mode, const int32_t *rnd)
{
    for (int64_t o = 0; o < outer; o++)
        for (int64_t i = 0; i < inner; i++)
            for (int64_t k0 = 0; k0 < K; k0 += bs) {
                int64_t k1 = k0 + bs < K ? k0 + bs : K;
                uint32_t m = 0;
                for (int64_t k = k0; k < k1; k++) {
                    uint32_t a = absbits(x[(o * K + k) * inner + i]);
                    if (a > m) m = a;
                }
                float mf = u2f(m);
                for (int64_t k = k0; k < k1; k++) {
                    int64_t idx = (o * K + k) * inner + i;
                    y[idx] = bfp_elem(x[idx], mf, wl, 1, mode, rnd ? (uint32_t)rnd[idx] : 0u);
                }
            }
}

/* float_quantize: Q/quant_cpu/quant_cpu.cpp:359-402 == Q/quant_cuda/float_kernel.cu:131-168. */
static float float_elem(float x, int man_bits, int exp_bits, int exp_bias, int flush, int mode, uint32_t rnd)
{
    uint32_t target = f2u(x), qb;
    float q;
    int target_exp = (int)((target << 1) >> 24) - 127;
    int min_exp = -(exp_bias - 1);
    if (target_exp < min_exp) {
        if (!flush) {
            uint32_t shift_bits = ((uint32_t)(127 + min_exp) << 23) | (target & 0x80000000u);
            float shift = u2f(shift_bits);
            float val = x + shift;
            qb = round_bitwise(f2u(val), man_bits, mode, rnd);
            q = u2f(qb) - shift;
        } else {
            q = 0.0f;
        }
    } else {
        qb = round_bitwise(target, man_bits, mode, rnd);
        qb = clip_exponent(exp_bits, man_bits, target, qb);
        q = u2f(qb);
    }
    return q;
}

/* FloatingPoint.cast: S/numerical/format.py:208-233 -- float_quantize, then the extra fp16
 * subnormal flush for "FP[1|5|10,15](FN)" (:223-232; |x| < 2^-14 -> +0), then abs() when
 * the format is unsigned (:233). */
void orc_float_cast(const float *x, float *y, int64_t n, int man_bits, int exp_bits, int exp_bias,
                    int flush, int is_unsigned, int fp16_flush, int mode, const int32_t *rnd)
{
    for (int64_t i = 0; i < n; i++) {
        float q = float_elem(x[i], man_bits, exp_bits, exp_bias, flush, mode, rnd ? (uint32_t)rnd[i] : 0u);
        if (fp16_flush && fabsf(q) < 6.103515625e-05f) q = 0.0f;
        if (is_unsigned) q = fabsf(q);
        y[i] = q;
    }
}

/* Q/quant_cpu/sim_helper.cpp:5-12 == Q/quant_cuda/quant.cu:230-237 */
void orc_fixed_min_max(int wl, int fl, int symmetric, float *t_min, float *t_max)
{
    int sigma = -fl;
    *t_min = (float)-ldexp(1.0, wl - fl - 1);
    *t_max = (float)(-(double)*t_min - ldexp(1.0, sigma));
    if (symmetric) *t_min = (float)((double)*t_min + ldexp(1.0, sigma));
}

/* One fixed-point element.
 * CPU reference: Q/quant_cpu/sim_helper.cpp:14-38 (`nearbyint(a + r - 0.5)` with r = 0.5 for
 *   nearest: float add, double subtract, ties-to-even) -> TIE_EVEN.
 * CUDA reference: Q/quant_cuda/sim_helper.cu:4-51 (nearest = roundf, half away from zero;
 *   stochastic = the same nearbyint expression) -> TIE_AWAY. */
static float fixed_elem(float a, int sigma, int mode, int tie, float r)
{
    a = ldexpf(a, -sigma);
    if (mode == R_NEAREST) {
        if (tie == TIE_EVEN) a = (float)nearbyint((double)(a + 0.5f) - 0.5);
        else a = roundf(a);
    } else if (mode == R_STOCHASTIC) {
        a = (float)nearbyint((double)(a + r) - 0.5);
    } else if (mode == R_UP) {
        a = ceilf(a);
    } else {
        a = floorf(a);
    }
    return ldexpf(a, sigma);
}

/* fixed_point_quantize: Q/quant_cpu/quant_cpu.cpp:125-209 == Q/quant_cuda/fixed_point_kernel.cu:34-101.
 * rand: fp32 in [0,1) per element (stochastic) or NULL. */
void orc_fixed_quantize(const float *x, float *y, int64_t n, int wl, int fl, int clamp, int symmetric,
                        int mode, int tie, const float *rnd)
{
    float t_min, t_max;
    orc_fixed_min_max(wl, fl, symmetric, &t_min, &t_max);
    int sigma = -fl;
    for (int64_t i = 0; i < n; i++) {
        float q = fixed_elem(x[i], sigma, mode, tie, rnd ? rnd[i] : 0.5f);
        if (clamp) { if (q > t_max) q = t_max; else if (q < t_min) q = t_min; }
        y[i] = q;
    }
}

/* CastTo.forward for a FixedPoint format with affine parameters: S/numerical/cast.py:279-296
 *   x = x / sc + zp ; x = fixed_point_quantize(x) ; x = (x - zp) * sc
 * each a separately rounded fp32 torch op.  (outer, C, inner) addressing; the qparam of
 * channel c is scale[(c / group) % nq] -- per-tensor: nq = 1, group = C; per-channel
 * (cast.py:228-237): nq = C, group = 1; group quantisation (cast.py:281-292,
 * repeat_interleave): group = group_size. */
void orc_fixed_cast_affine(const float *x, float *y, int64_t outer, int64_t C, int64_t inner,
                           int wl, int fl, int clamp, int symmetric, int mode, int tie,
                           const float *scale, const float *zp, int64_t nq, int64_t group, const float *rnd)
{
    float t_min, t_max;
    orc_fixed_min_max(wl, fl, symmetric, &t_min, &t_max);
    int sigma = -fl;
    for (int64_t o = 0; o < outer; o++)
        for (int64_t c = 0; c < C; c++) {
            int64_t qi = nq == 1 ? 0 : (c / group);
            if (qi >= nq) qi = nq - 1;
            float sc = scale[qi], z = zp[qi];
            for (int64_t i = 0; i < inner; i++) {
                int64_t idx = (o * C + c) * inner + i;
                float v = x[idx] / sc;
                v = v + z;
                float q = fixed_elem(v, sigma, mode, tie, rnd ? rnd[idx] : 0.5f);
                if (clamp) { if (q > t_max) q = t_max; else if (q < t_min) q = t_min; }
                q = q - z;
                y[idx] = q * sc;
            }
        }
}

/* ScaledBlockFloatingPoint.cast: S/numerical/format.py:453-479.
 *   chunk_max = max|chunk| / man_scaling                      (:461-463, :429-434)
 *   y = where(chunk_max > 0, XP(chunk / chunk_max) * FP(chunk_max), chunk)   (:466-472)
 * XP = fixed_point_quantize(wl = p, fl = 0, clamp, symmetric, nearest) (format.py:134-142);
 * FP = FloatingPoint.cast with the scaler format (float_quantize [+ fp16 flush] [+ abs]). */
void orc_sbfp_cast(const float *x, float *y, int64_t outer, int64_t K, int64_t inner, int64_t bs,
                   int xp_wl, int xp_clamp, int xp_mode, int tie,
                   int sc_man, int sc_exp, int sc_bias, int sc_flush, int sc_unsigned, int sc_fp16_flush,
                   int sc_mode, int scale_recip)
{
    float man_scaling = (float)((1 << (xp_wl - 1)) - 1);
    /* `chunk_max / self.man_scaling` on a CUDA tensor: ATen multiplies by inv_b = fp32(1.0 / b) (div_true_kernel_cuda) */
    float inv_b = (float)(1.0 / (double)man_scaling);
    float t_min, t_max;
    orc_fixed_min_max(xp_wl, 0, 1, &t_min, &t_max);
    for (int64_t o = 0; o < outer; o++)
        for (int64_t i = 0; i < inner; i++)
            for (int64_t k0 = 0; k0 < K; k0 += bs) {
                int64_t k1 = k0 + bs < K ? k0 + bs : K;
                uint32_t m = 0;
                for (int64_t k = k0; k < k1; k++) {
                    uint32_t a = absbits(x[(o * K + k) * inner + i]);
                    if (a > m) m = a;
                }
                float cmax = scale_recip ? u2f(m) * inv_b : u2f(m) / man_scaling;
                float fs = float_elem(cmax, sc_man, sc_exp, sc_bias, sc_flush, sc_mode, 0u);
                if (sc_fp16_flush && fabsf(fs) < 6.103515625e-05f) fs = 0.0f;
                if (sc_unsigned) fs = fabsf(fs);
                for (int64_t k = k0; k < k1; k++) {
                    int64_t idx = (o * K + k) * inner + i;
                    if (cmax > 0.0f) {
                        float v = x[idx] / cmax;
                        float q = fixed_elem(v, 0, xp_mode, tie, 0.5f);
                        if (xp_clamp) { if (q > t_max) q = t_max; else if (q < t_min) q = t_min; }
                        y[idx] = q * fs;
                    } else {
                        y[idx] = x[idx];
                    }
                }
            }
}

/* MXFP.cast: S/numerical/format.py:545-564.
 *   scale = 2 ** floor(log2(max|chunk|)) / largest_representable_power_of_two       (:551-555)
 *   y     = element_format.cast(chunk / scale) * scale                               (:558)
 * element format: E<exp>M<man>, bias 2^(exp-1)-1, subnormals kept, nearest (:585-592).  An all-zero
 * chunk gives scale 0 and 0/0 = NaN, exactly as the torch ops do. */
void orc_mxfp_cast(const float *x, float *y, int64_t outer, int64_t K, int64_t inner, int64_t bs, int man, int exp_bits)
{
    float largest = ldexpf(1.0f, 1 << (exp_bits - 1));
    int bias = (1 << (exp_bits - 1)) - 1;
    for (int64_t o = 0; o < outer; o++)
        for (int64_t i = 0; i < inner; i++)
            for (int64_t k0 = 0; k0 < K; k0 += bs) {
                int64_t k1 = k0 + bs < K ? k0 + bs : K;
                uint32_t m = 0;
                for (int64_t k = k0; k < k1; k++) {
                    uint32_t a = absbits(x[(o * K + k) * inner + i]);
                    if (a > m) m = a;
                }
                float scale = exp2f(floorf(log2f(u2f(m)))) / largest;
                for (int64_t k = k0; k < k1; k++) {
                    int64_t idx = (o * K + k) * inner + i;
                    float v = x[idx] / scale;
                    y[idx] = float_elem(v, man, exp_bits, bias, 0, R_NEAREST, 0u) * scale;
                }
            }
}

/* BlockTopK.forward + Sparsify.forward: S/sparse.py:163-180, :287-301.
 * Per group of M consecutive k: ascending (stable, NaN largest) argsort of the score, the
 * first M - Kkeep indices get mask 0; y = x * mask (fp32 multiply: masked negatives become
 * -0.0, masked Inf/NaN become NaN).  score == NULL means score = |x| (the documented
 * `lambda s, x: x.abs()` score function).  mask_out may be NULL. */
/* torch.argsort's default (unstable) order on CUDA for rows of <= 32 keys: ATen's bitonicSortKVInPlace
 * (ATen/native/cuda/SortUtils.cuh) restated literally -- 32 slots, the M keys in slots 0..M-1, "invalid" slots behind
 * them, 16 "threads" per stage, comparator LTOp<float, handleNaN = true> (SortingCommon.cuh:55-61):
 *     swap = (LT(kA, kB) && validA) || !validB;   if (swap == dir) exchange
 * pinned by tests/golden/argsort_cuda_order.npz (torch.argsort run on the B200 over every tie pattern). */
static int lt_nan_last(float a, float b) { return (isnan(b) && !isnan(a)) || a < b; }
static void bitonic32_order(const float *key, int M, int *order)
{
    float k[32];
    int id[32];
    for (int a = 0; a < 32; a++) { k[a] = a < M ? key[a] : 0.0f; id[a] = a; }
    for (int size = 2; size <= 32; size *= 2)
        for (int stride = size / 2; stride > 0; stride /= 2)
            for (int t = 0; t < 16; t++) {
                int dir = size < 32 && (t & (size / 2)) != 0;
                int a = 2 * t - (t & (stride - 1)), b = a + stride;
                int va = id[a] < M, vb = id[b] < M;
                int swap = (lt_nan_last(k[a], k[b]) && va) || !vb;
                if (swap == dir) {
                    float tk = k[a]; k[a] = k[b]; k[b] = tk;
                    int ti = id[a]; id[a] = id[b]; id[b] = ti;
                }
            }
    for (int a = 0; a < M; a++) order[a] = id[a];
}
void orc_argsort_cuda_order(const float *keys, int64_t rows, int M, int32_t *out)
{
    int order[32];
    for (int64_t r = 0; r < rows; r++) {
        bitonic32_order(keys + r * M, M, order);
        for (int a = 0; a < M; a++) out[r * M + a] = order[a];
    }
}

/* nm_order: 0 = stable (torch.argsort on CPU tensors / stable=True), 1 = torch's CUDA order (groups of <= 32). */
void orc_nm_prune(const float *x, const float *score, float *y, float *mask_out,
                  int64_t outer, int64_t K, int64_t inner, int n_keep, int M, int nm_order)
{
    int n_prune = M - n_keep;
    if (nm_order == 1 && M <= 32) {
        float key[32];
        int order[32];
        for (int64_t o = 0; o < outer; o++)
            for (int64_t i = 0; i < inner; i++)
                for (int64_t k0 = 0; k0 + M <= K; k0 += M) {
                    for (int a = 0; a < M; a++) {
                        int64_t ia = (o * K + k0 + a) * inner + i;
                        key[a] = score ? score[ia] : fabsf(x[ia]);
                    }
                    bitonic32_order(key, M, order);
                    for (int a = 0; a < M; a++) {
                        int64_t ia = (o * K + k0 + order[a]) * inner + i;
                        float mk = a < n_prune ? 0.0f : 1.0f;
                        if (mask_out) mask_out[ia] = mk;
                        y[ia] = x[ia] * mk;
                    }
                }
        return;
    }
    for (int64_t o = 0; o < outer; o++)
        for (int64_t i = 0; i < inner; i++)
            for (int64_t k0 = 0; k0 + M <= K; k0 += M)
                for (int a = 0; a < M; a++) {
                    int64_t ia = (o * K + k0 + a) * inner + i;
                    float sa = score ? score[ia] : fabsf(x[ia]);
                    int rank = 0;
                    for (int b = 0; b < M; b++) {
                        if (b == a) continue;
                        int64_t ib = (o * K + k0 + b) * inner + i;
                        float sb = score ? score[ib] : fabsf(x[ib]);
                        int b_less;
                        if (isnan(sa)) b_less = isnan(sb) ? (b < a) : 1;
                        else if (isnan(sb)) b_less = 0;
                        else b_less = (sb < sa) || (sb == sa && b < a);
                        rank += b_less;
                    }
                    float mk = rank < n_prune ? 0.0f : 1.0f;
                    if (mask_out) mask_out[ia] = mk;
                    y[ia] = x[ia] * mk;
                }
}

/* MinMaxObserver.forward statistics (S/numerical/observer.py:173-193): running amin / amax,
 * per tensor (C == 1 via outer*K*inner flattening by the caller) or per channel of the
 * (outer, C, inner) addressing. */
void orc_minmax(const float *x, int64_t outer, int64_t C, int64_t inner, float *mn, float *mx)
{
    for (int64_t c = 0; c < C; c++) { mn[c] = INFINITY; mx[c] = -INFINITY; }
    for (int64_t o = 0; o < outer; o++)
        for (int64_t c = 0; c < C; c++)
            for (int64_t i = 0; i < inner; i++) {
                float v = x[(o * C + c) * inner + i];
                if (isnan(v)) { mn[c] = v; mx[c] = v; continue; }
                if (!(mn[c] <= v)) { if (!isnan(mn[c])) mn[c] = v; }
                if (!(mx[c] >= v)) { if (!isnan(mx[c])) mx[c] = v; }
            }
}

/* torch.histc(x, bins, min, max) as HistogramObserver.forward calls it (S/numerical/observer.py:470-491); the bin
 * rule is ATen's (aten/src/ATen/native/cpu/.. histc and cuda/SummaryOps.cu getBin, torch 2.x -- a dependency of the
 * reference, not vendored): values outside [lo, hi] (and NaN) are skipped, bin = (int)((v - lo) * bins / (hi - lo))
 * evaluated in fp32 left to right, the right edge falls into the last bin.  lo < hi is resolved by the caller
 * (histc's "min == max -> data range, still equal -> +-1" rule lives in oracle.py). counts are accumulated. */
void orc_histc(const float *x, int64_t n, int bins, float lo, float hi, int64_t *counts)
{
    const float nb = (float)bins, w = hi - lo;
    for (int64_t i = 0; i < n; i++) {
        float v = x[i];
        if (!(v >= lo && v <= hi)) continue;
        volatile float t = (v - lo) * nb;
        int64_t b = (int64_t)(t / w);
        if (b == bins) b -= 1;
        if (b >= 0 && b < bins) counts[b] += 1;
    }
}

/* bf16 <-> fp32 as torch does it (`x.float()` exact widening; `.to(bfloat16)` RNE with NaN
 * quieting) -- S/numerical/cast.py:262,306 wrap every cast in these conversions. */
void orc_bf16_to_f32(const uint16_t *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) y[i] = u2f((uint32_t)x[i] << 16);
}
void orc_f32_to_bf16(const float *x, uint16_t *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        uint32_t u = f2u(x[i]);
        if ((u & 0x7FFFFFFFu) > 0x7F800000u) { y[i] = 0x7FC0; continue; }
        u += 0x7FFFu + ((u >> 16) & 1u);
        y[i] = (uint16_t)(u >> 16);
    }
}

int orc_abi_version(void) { return 1; }

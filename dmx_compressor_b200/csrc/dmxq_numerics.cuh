// dmxq_numerics.cuh -- bit-exact element codecs of the CastTo / Sparsify path (device side).
//
// Each function states the reference arithmetic it reproduces ("Q/" = reference
// src/dmx/compressor/quant/, "S/" = src/dmx/compressor/).  The algebra is re-derived for the
// GPU (fused masks, 3-input adds, branch-free selects); the *results* are bit-identical to the
// reference's CUDA kernels, which is what tests/test_parity_gpu.py checks against the oracle.
//
// Rules that keep it exact (hard part 1):
//   * every fp32 add/sub/mul/div that the reference performs as a separate rounded op is an
//     explicit __f*_rn intrinsic here (never contracted into an FMA, never reassociated);
//   * no fast-math, no flush-to-zero (denormals survive), true IEEE division;
//   * block maxima are taken on |x| bit patterns as unsigned integers: NaN > Inf > finite,
//     which reproduces torch.max's NaN propagation for everything downstream consumes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dmxq {

enum : int { R_NEAREST = 0, R_STOCHASTIC = 1, R_UP = 2, R_DOWN = 3 };
enum : int { TIE_AWAY = 0, TIE_EVEN = 1 };
enum : int { ST_NONE = 0, ST_NM = 1, ST_BFP = 2, ST_SBFP = 3, ST_FLOAT = 4, ST_FIXED = 5, ST_MXFP = 6, ST_SCALE = 7 };

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// Keep `23 - sh` mantissa bits of the fp32 pattern `t`.
// Reference: round_bitwise_{nearest,stochastic,up,down}, Q/quant_cuda/bit_helper.cu:10-46.
// Nearest (RNE on the magnitude pattern): the reference adds `half` unless the discarded
// field is exactly `half` and the kept LSB is even; that equals adding (half - 1) + keptLSB
// and truncating, which is one IADD3 + one LOP3.  sh == 0 (man_bits >= 23, undefined in the
// reference, identity on its CUDA build) must be passed with mode R_DOWN by the host.
template <int MODE>
__device__ __forceinline__ uint32_t round_bits(uint32_t t, int sh, uint32_t mask, uint32_t rnd)
{
    if (MODE == R_NEAREST) {
        uint32_t lsb = (t >> sh) & 1u;
        return (t + (mask >> 1) + lsb) & ~mask;
    } else if (MODE == R_STOCHASTIC) {
        return (t + (rnd & mask)) & ~mask;
    } else if (MODE == R_UP) {
        return (t + mask + 1u) & ~mask;
    } else {
        return t & ~mask;
    }
}

__device__ __forceinline__ uint32_t round_bits_rt(uint32_t t, int sh, uint32_t mask, int mode, uint32_t rnd)
{
    switch (mode) {
    case R_NEAREST: return round_bits<R_NEAREST>(t, sh, mask, rnd);
    case R_STOCHASTIC: return round_bits<R_STOCHASTIC>(t, sh, mask, rnd);
    case R_UP: return round_bits<R_UP>(t, sh, mask, rnd);
    default: return round_bits<R_DOWN>(t, sh, mask, rnd);
    }
}

// ---------------------------------------------------------------------------------------------
// Counter-based random words for stochastic rounding without a random TENSOR.  The reference
// draws `randint_like(a, INT_MAX)` / `rand_like(a)` into a full-size tensor per cast (Q/quant_cuda/quant.cu:40,118,160,244): 4 bytes
// written and 4 read per element next to the 8 the cast itself moves.  Philox4x32-10 (Salmon et al., "Parallel random numbers: as
// easy as 1, 2, 3", SC'11; the generator behind curand's and torch's CUDA streams) makes word i of a stream a pure function:
//     word(i) = Philox4x32_10(counter = (i / 4 as 64 bits, stream id as 64 bits), key = seed)[i % 4],   i = logical element index
// so the kernel computes the four words of four consecutive elements in registers.  dmxq_philox_fill materialises the very same
// stream, which is how the fused path is pinned: cast(philox) == cast(external tensor = philox_fill) bit for bit, and the
// external-tensor path is the one checked against the oracle and the reference's kernels.  (Not torch's randint_like stream: that one
// depends on the launch geometry of ATen's distribution kernel; the explicit tensor input remains for seed-for-seed reproduction.)
__device__ __forceinline__ uint4 philox4x32_10(uint64_t ctr, uint64_t stream, uint64_t seed)
{
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t philox_word(uint64_t idx, uint64_t stream, uint64_t seed)
{
    const uint4 q = philox4x32_10(idx >> 2, stream, seed);
    const uint32_t k = (uint32_t)idx & 3u;
    return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w;
}
// the word as the bit pattern of a uniform fp32 in [0, 1) with 24 random bits (what FixedPoint's stochastic mode consumes)
__device__ __forceinline__ uint32_t philox_unit_bits(uint32_t w) { return f2u(__fmul_rn(__uint2float_rz(w >> 8), 0x1p-24f)); }

// ---------------------------------------------------------------------------------------------
// Block floating point.  Per block (uniform): E = exponent field of max|x| (in place, bits
// 23..30), base = 6 * 2^e, maxnum = E | top (wl-2) mantissa bits.
// Reference: block_kernel_*, Q/quant_cuda/block_kernel.cu:43-74 + clip_max_exponent,
// bit_helper.cu:68-79.
struct BfpBlock {
    uint32_t E;       // exponent field of the block max
    uint32_t maxnum;  // E | max mantissa: largest representable magnitude of the block
    float base;       // 6 * 2^e  (Inf when e >= 126, exactly as the reference's fp32 product)
    float quantum;    // 2^(e + 2 - wl), only used by the asymmetric post-pass
};

__device__ __forceinline__ BfpBlock bfp_block(uint32_t maxabs_bits, int wl)
{
    BfpBlock b;
    b.E = maxabs_bits & 0x7F800000u;
    b.base = __fmul_rn(u2f(b.E), 6.0f);
    int m = wl - 2;
    b.maxnum = b.E | ((0x007FFFFFu >> (23 - m)) << (23 - m));
    b.quantum = __fmul_rn(u2f(b.E), u2f((uint32_t)(127 + 2 - wl) << 23));
    return b;
}

template <int MODE>
__device__ __forceinline__ float bfp_elem(float x, const BfpBlock &b, int sh, uint32_t mask, uint32_t rnd)
{
    float t = __fadd_rn(x, b.base);
    uint32_t tb = round_bits<MODE>(f2u(t), sh, mask, rnd);
    float q = __fsub_rn(u2f(tb), b.base);
    uint32_t qb = f2u(q);
    if ((qb & 0x7F800000u) > b.E) qb = (qb & 0x80000000u) | b.maxnum;
    return u2f(qb);
}

// Fast path for "ordinary" blocks (biased block exponent 1..227, wl <= 20; everything the
// BASELINE workloads ever produce).  There t = x + base lies in [4,8)*2^e, i.e. has the fixed
// exponent e+2, so keeping wl mantissa bits of t with RNE is rounding t to a multiple of
// Q = 2^(e+2-wl) -- which one more fp32 add of C = 1.5 * 2^(e+25-wl) performs in hardware
// (ulp(t + C) == Q; C/Q is even, so ties go to the same even neighbour as the bit pattern
// rule).  (t + C) - C and the final - base are exact.  Every q is then a finite multiple of
// Q, so the exponent clip |q| >= 2^(e+1) equals a clamp to +-maxval = +-(2^(e+1) - Q).
// 4 FADD + 2 FMNMX per element instead of ~11 integer ops; bit-identical by construction,
// and checked against the integer path by tests/test_parity_gpu.py on adversarial ties.
struct BfpFast {
    float base, C, maxval;
    bool clamp;  // some element of the block can reach the clip (block max within 2 quanta of 2^(e+1))
};
__device__ __forceinline__ bool bfp_fast_ok(uint32_t maxabs_bits)
{
    return ((maxabs_bits & 0x7F800000u) - 0x00800000u) <= (226u << 23);
}
__device__ __forceinline__ BfpFast bfp_fast_block(uint32_t maxabs_bits, int wl)
{
    BfpFast b;
    uint32_t E = maxabs_bits & 0x7F800000u;
    b.base = __fmul_rn(u2f(E), 6.0f);
    b.C = u2f(E + (((uint32_t)(25 - wl) << 23) | 0x00400000u));
    int m = wl - 2;
    uint32_t maxnum = E | ((0x007FFFFFu >> (23 - m)) << (23 - m));
    b.maxval = u2f(maxnum);
    // |q| can exceed maxval = 2^(e+1) - Q only if |x| >= 2^(e+1) - Q/2 - (one ulp of the first
    // add); testing the block max against maxval itself (a full quantum below 2^(e+1)) is a
    // safe superset, and lets the common block skip the clamp altogether.
    b.clamp = maxabs_bits >= maxnum;
    return b;
}
__device__ __forceinline__ float bfp_fast_elem(float x, const BfpFast &b)
{
    float t = __fadd_rn(x, b.base);
    float r = __fsub_rn(__fadd_rn(t, b.C), b.C);
    return __fsub_rn(r, b.base);
}
// 16-bit sources (bf16: 8 significant bits, fp16: 11): x + base is exact whenever x is not
// negligible (its LSB >= ulp(t) = 2^(e-21)), and negligible x (|x| < 2^(e-14) resp. 2^(e-11))
// round to +0 either way as long as wl <= 14 (bf16) / 11 (fp16).  The reference's first add
// therefore never changes the outcome and q = RNE_Q(x) = (x + C) - C: two adds per element.
__device__ __forceinline__ float bfp_fast16_elem(float x, const BfpFast &b)
{
    return __fsub_rn(__fadd_rn(x, b.C), b.C);
}
__device__ __forceinline__ float bfp_clamp(float q, const BfpFast &b) { return fminf(fmaxf(q, -b.maxval), b.maxval); }

// BlockFloatingPoint.make_mantissa_asymmetric, S/numerical/format.py:349-372: an element whose
// integer mantissa is exactly -(2^(wl-1)-1) moves one quantum down to -2^(wl-1) when that does
// not increase |error| (ties go to the even mantissa).  old/candidate errors are separately
// rounded fp32 subtractions, as the torch ops are.
__device__ __forceinline__ float bfp_asym_fix(float q, float x, const BfpBlock &b)
{
    if (f2u(q) == (0x80000000u | b.maxnum) && b.E != 0u) {
        float old_err = __fsub_rn(q, x);
        float cand_err = __fsub_rn(old_err, b.quantum);
        if (fabsf(cand_err) <= fabsf(old_err)) q = -u2f(b.E + 0x00800000u);
    }
    return q;
}

// ---------------------------------------------------------------------------------------------
// Low-bit floating point.  Reference: float_kernel_*, Q/quant_cuda/float_kernel.cu:131-168 +
// clip_exponent, bit_helper.cu:48-65; python wrapper extras S/numerical/format.py:223-233.
struct FloatFmt {
    int sh;              // 23 - man_bits (0 => identity rounding)
    uint32_t mask;       // (1 << sh) - 1
    int min_exp;         // -(bias - 1)
    uint32_t shift_exp;  // (127 + min_exp) << 23: the subnormal-rounding shift magnitude
    uint32_t max_store;  // (1 << (exp_bits-1)) + 127: saturation exponent (nothing reserved for Inf/NaN)
    uint32_t max_num;    // (max_store << 23) | max mantissa
    int exact;           // 1: the input already fits `man` bits (16-bit source dtype, or man >= 23): no rounding
    int flush;           // flush_subnormal
    int is_unsigned;     // abs() afterwards
    int fp16_flush;      // extra |q| < 2^-14 -> +0
    int mode;
    int fastpath;        // nearest + flush + signed (+ fp16 pass implied): float_elem_flush_nearest is valid for non-NaN inputs
    int nsub;            // nearest, subnormals kept, signed, man <= 21: float_elem_nearest_sub is valid for finite inputs
    uint32_t magic;      // (sh << 23) | 0x00400000: exponent-field offset + mantissa of the rounding constant 1.5 * 2^(e + sh)
};

template <int MODE>
__device__ __forceinline__ float float_elem(float x, const FloatFmt &f, uint32_t rnd)
{
    uint32_t target = f2u(x);
    int texp = (int)((target & 0x7FFFFFFFu) >> 23) - 127;
    float q;
    if (texp < f.min_exp) {
        if (f.flush) {
            q = 0.0f;
        } else {
            float shift = u2f(f.shift_exp | (target & 0x80000000u));
            float val = __fadd_rn(x, shift);
            uint32_t qb = f.sh ? round_bits<MODE>(f2u(val), f.sh, f.mask, rnd) : f2u(val);
            q = __fsub_rn(u2f(qb), shift);
        }
    } else {
        uint32_t qb = f.sh ? round_bits<MODE>(target, f.sh, f.mask, rnd) : target;
        if (((qb & 0x7FFFFFFFu) >> 23) > f.max_store) qb = (target & 0x80000000u) | f.max_num;
        q = u2f(qb);
    }
    if (f.fp16_flush && fabsf(q) < 6.103515625e-05f) q = 0.0f;
    if (f.is_unsigned) q = fabsf(q);
    return q;
}

// nearest rounding + flush_subnormal, signed: the FLOAT16 / BFLOAT16 "(FN)" formats every BASIC
// module boundary uses.  Branch-free except for NaN inputs:
//   * "target_exp < min_exp" <=> |x| pattern < shift_exp  -> +0;
//   * "exponent store > max" <=> |q| pattern > max_num (q is already rounded to man bits), so
//     the saturation is an unsigned min on the magnitude, re-signed with x's sign;
//   * a NaN whose mantissa carries out of bit 30 when rounded is the one input where the sign
//     bookkeeping above differs from the reference's pattern arithmetic: NaNs (pattern above
//     0x7f800000) take the out-of-line exact path (never in real data; keeps bit parity).
template <bool EXACT>
__device__ __forceinline__ float float_elem_flush_nearest(float x, const FloatFmt &f, uint32_t max_num)
{
    uint32_t target = f2u(x);
    uint32_t ab = target & 0x7FFFFFFFu;
    uint32_t qa = EXACT ? ab : round_bits<R_NEAREST>(ab, f.sh, f.mask, 0u);
    uint32_t mag = min(qa, max_num);
    return ab < f.shift_exp ? 0.0f : u2f(mag | (target & 0x80000000u));
}
template <bool EXACT> __device__ __forceinline__ float float_elem_flush_nearest(float x, const FloatFmt &f)
{
    return float_elem_flush_nearest<EXACT>(x, f, f.max_num);
}

// nearest rounding with subnormals kept (the MX element formats), branch free, for finite x.  The reference
// (float_kernel.cu:147-160) rounds values below 2^min_exp by adding +-2^min_exp first (val = x + shift, itself a rounded
// fp32 add), rounding val's mantissa to `man` bits and subtracting the shift again; ordinary values are rounded in
// place.  Both are "round val to a multiple of 2^(exponent(val) - man), ties to even", which one more fp32 add of
// C = 1.5 * 2^(exponent(val) + 23 - man) performs: (val + C) - C.  With shift = +-0 for ordinary values the two cases
// share one instruction sequence; the saturation is an unsigned min on the magnitude, as in float_elem_flush_nearest.
// Precondition: |x| < 2^(128 - sh) (C must stay finite); larger and non-finite values take the literal path.
__device__ __forceinline__ float float_elem_nearest_sub(float x, const FloatFmt &f)
{
    const uint32_t t = f2u(x), ab = t & 0x7FFFFFFFu;
    const float shift = u2f((ab < f.shift_exp ? f.shift_exp : 0u) | (t & 0x80000000u));
    const float val = __fadd_rn(x, shift);
    const float C = u2f((f2u(val) & 0x7F800000u) + f.magic);
    const uint32_t qb = f2u(__fsub_rn(__fsub_rn(__fadd_rn(val, C), C), shift));
    return u2f(min(qb & 0x7FFFFFFFu, f.max_num) | (qb & 0x80000000u));
}

// nearest rounding, all other flag combinations (FP8 formats keep subnormals; unsigned scalers)
__device__ __forceinline__ float float_elem_nearest(float x, const FloatFmt &f)
{
    uint32_t target = f2u(x);
    uint32_t ab = target & 0x7FFFFFFFu;
    float q;
    if (f.flush) {
        uint32_t qb = f.exact ? target : round_bits<R_NEAREST>(target, f.sh, f.mask, 0u);
        uint32_t mag = min(qb & 0x7FFFFFFFu, f.max_num);
        q = ab < f.shift_exp ? 0.0f : u2f(mag | (qb & 0x80000000u));
    } else if (ab < f.shift_exp) {
        float shift = u2f(f.shift_exp | (target & 0x80000000u));
        float val = __fadd_rn(x, shift);
        uint32_t qb = f.sh ? round_bits<R_NEAREST>(f2u(val), f.sh, f.mask, 0u) : f2u(val);
        q = __fsub_rn(u2f(qb), shift);
    } else {
        uint32_t qb = f.exact ? target : round_bits<R_NEAREST>(target, f.sh, f.mask, 0u);
        uint32_t mag = min(qb & 0x7FFFFFFFu, f.max_num);
        q = u2f(mag | (qb & 0x80000000u));
    }
    if (f.fp16_flush && fabsf(q) < 6.103515625e-05f) q = 0.0f;
    if (f.is_unsigned) q = fabsf(q);
    return q;
}

__device__ __forceinline__ float float_elem_rt(float x, const FloatFmt &f, uint32_t rnd)
{
    switch (f.mode) {
    case R_NEAREST: return float_elem<R_NEAREST>(x, f, rnd);
    case R_STOCHASTIC: return float_elem<R_STOCHASTIC>(x, f, rnd);
    case R_UP: return float_elem<R_UP>(x, f, rnd);
    default: return float_elem<R_DOWN>(x, f, rnd);
    }
}

// ---------------------------------------------------------------------------------------------
// Fixed point.  Reference: fixed_point_quantize_kernel_*, Q/quant_cuda/fixed_point_kernel.cu:34-101
// + sim_helper.cu:4-51 (CUDA) / Q/quant_cpu/sim_helper.cpp:14-38 (CPU tie quirk).
// ldexp(a, +-fl) is an exact power-of-two scaling: a single correctly rounded multiply by
// 2^fl is bit-identical (|fl| <= 126 is enforced by the host).
struct FixedFmt {
    float up, down;      // 2^fl, 2^-fl
    float t_min, t_max;  // fixed_min_max, quant.cu:230-237
    int clamp;
    int mode;
    int tie;
};

__device__ __forceinline__ float fixed_elem(float a, const FixedFmt &f, float r)
{
    a = __fmul_rn(a, f.up);
    switch (f.mode) {
    case R_NEAREST:
        if (f.tie == TIE_AWAY) a = roundf(a);
        else a = (float)rint(__dadd_rn((double)__fadd_rn(a, 0.5f), -0.5));
        break;
    case R_STOCHASTIC: a = (float)rint(__dadd_rn((double)__fadd_rn(a, r), -0.5)); break;
    case R_UP: a = ceilf(a); break;
    default: a = floorf(a); break;
    }
    a = __fmul_rn(a, f.down);
    if (f.clamp) {
        if (a > f.t_max) a = f.t_max;
        else if (a < f.t_min) a = f.t_min;
    }
    return a;
}

// CastTo.forward's affine wrap, S/numerical/cast.py:293,296: four separately rounded ops.
__device__ __forceinline__ float fixed_elem_affine(float x, const FixedFmt &f, float sc, float zp, float r)
{
    float v = __fadd_rn(__fdiv_rn(x, sc), zp);
    v = fixed_elem(v, f, r);
    return __fmul_rn(__fsub_rn(v, zp), sc);
}

// ---------------------------------------------------------------------------------------------
// Scaled BFP.  Reference: ScaledBlockFloatingPoint.cast, S/numerical/format.py:453-479.
struct SbfpFmt {
    FixedFmt xp;        // XP[p,0] block format (fl = 0: up = down = 1)
    FloatFmt sc;        // scaler FloatingPoint format
    float man_scaling;  // 2^(p-1) - 1
    float inv_man;      // RN(1 / man_scaling)
    int no_clamp;       // the XP clamp cannot trigger on |x| <= block max (host-decided: !clamp || range covers +-man_scaling)
    int sc_fast;        // scaler cast = float_elem_flush_nearest (nearest, flushing; host-decided)
    int recip;          // DMXQ_SCALE_RECIP: block scale = max * inv_man_t (what torch computes for `cuda_tensor / python_scalar`)
    float inv_man_t;    // fp32(1.0 / (double)man_scaling): ATen's inv_b (div_true_kernel_cuda)
};

struct SbfpBlock {
    float cmax;  // max|x| / man_scaling   (NaN when the block holds a NaN)
    float fs;    // FP_cast(cmax)
    float rc;    // RN(1 / cmax), for the division-free quotient (dmxq_stages.cuh sbfp_elem_fast)
    float rl;    // ~ 1 / cmax - rc: low part of the reciprocal (div_by_recip2)
    bool on;     // cmax > 0 (false => pass the block through, format.py:467-472)
    bool rok;    // rc usable: cmax well inside the normal range and its significand not all ones
};

__device__ __forceinline__ SbfpBlock sbfp_block(uint32_t maxabs_bits, const SbfpFmt &f)
{
    SbfpBlock b;
    b.cmax = f.recip ? __fmul_rn(u2f(maxabs_bits), f.inv_man_t) : __fdiv_rn(u2f(maxabs_bits), f.man_scaling);
    b.fs = float_elem_rt(b.cmax, f.sc, 0u);
    b.on = b.cmax > 0.0f;
    b.rc = 0.0f;
    b.rl = 0.0f;
    b.rok = false;
    return b;
}

__device__ __forceinline__ float sbfp_elem(float x, const SbfpBlock &b, const SbfpFmt &f)
{
    if (!b.on) return x;
    float v = __fdiv_rn(x, b.cmax);
    v = fixed_elem(v, f.xp, 0.5f);
    return __fmul_rn(v, b.fs);
}

// ---------------------------------------------------------------------------------------------
// MXFP.  Reference: MXFP.cast, S/numerical/format.py:545-564:
//   scale = 2 ** floor(log2(max|chunk|)) / largest_representable_power_of_two      (:551-555)
//   y     = element_format.cast(chunk / scale) * scale                             (:558)
// with the very same float functions torch dispatches to on CUDA (log2f, floorf, exp2f, IEEE
// division / multiplication), so an all-zero block becomes NaN (0 / 0) exactly as in the reference.
struct MxBlock {
    float scale;
};
__device__ __forceinline__ MxBlock mx_block(uint32_t maxabs_bits, float largest_pow2)
{
    MxBlock b;
    b.scale = __fdiv_rn(exp2f(floorf(log2f(u2f(maxabs_bits)))), largest_pow2);
    return b;
}
__device__ __forceinline__ float mx_elem(float x, const MxBlock &b, const FloatFmt &f)
{
    return __fmul_rn(float_elem_nearest(__fdiv_rn(x, b.scale), f), b.scale);
}

// ---------------------------------------------------------------------------------------------
// N:M prune.  Reference: BlockTopK.forward, S/sparse.py:163-180 (ascending argsort of the
// score, the M-K lowest get mask 0) and Sparsify.forward :287-301 (y = x * mask, an fp32
// multiply: masked negatives become -0.0, masked Inf/NaN become NaN).  Order: ascending,
// ties -> lower index first (stable), NaN largest, -0 == +0.
__device__ __forceinline__ uint32_t score_key(float s)
{
    uint32_t b = f2u(s);
    if ((b & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;  // NaN: largest
    if (b == 0x80000000u) b = 0u;                              // -0 == +0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t absx_key(float x)  // score = |x|: all NaNs collapse to one key above Inf
{
    return min(f2u(x) & 0x7FFFFFFFu, 0x7F800001u);
}
__device__ __forceinline__ float nm_apply(float x, bool keep) { return keep ? x : __fmul_rn(x, 0.0f); }

// ---------------------------------------------------------------------------------------------
// dtype conversion exactly as torch does it around every cast (S/numerical/cast.py:262,306):
// widening is exact, narrowing is round-to-nearest-even.
template <typename T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float to_f32(float v) { return v; }
    static __device__ __forceinline__ float from_f32(float v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};

}  // namespace dmxq

// dmxq_stages.cuh -- vector I/O and per-stage register-vector code shared by the tiled kernels.
//
// Code-size discipline: the formats the BASIC rule set and the BASELINE configs use
// (nearest rounding, symmetric BFP, half-away XP) are inlined straight-line code; every other
// mode goes through one __noinline__ scalar function per format, so a kernel carries a single
// copy of the rarely used paths instead of one per unrolled element.
#pragma once
#include <type_traits>

#include "dmxq_kernels.cuh"

namespace dmxq {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ------------------------------------------------------------------------------------------------
// 128-bit streaming loads / stores (read-once / write-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

template <typename T> struct VecIO;  // V = 16 bytes of T on the input side

template <> struct VecIO<float> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&v)[4])
    {
        v[0] = u2f(r.x); v[1] = u2f(r.y); v[2] = u2f(r.z); v[3] = u2f(r.w);
    }
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { unpack(ldg_stream(p), v); }
    template <int N> static __device__ __forceinline__ void store(float *p, const float (&v)[N])
    {
#pragma unroll
        for (int i = 0; i < N; i += 4) stg_stream(p + i, make_uint4(f2u(v[i]), f2u(v[i + 1]), f2u(v[i + 2]), f2u(v[i + 3])));
    }
};
template <> struct VecIO<__nv_bfloat16> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&v)[8])
    {
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = u2f(w[i] << 16); v[2 * i + 1] = u2f(w[i] & 0xFFFF0000u); }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) { unpack(ldg_stream(p), v); }
    template <int N> static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[N])
    {
        static_assert(N == 8, "bf16 stores are 8 wide");
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        stg_stream(p, make_uint4(w[0], w[1], w[2], w[3]));
    }
};
template <> struct VecIO<__half> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&v)[8])
    {
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(*reinterpret_cast<__half2 *>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void load(const __half *p, float (&v)[8]) { unpack(ldg_stream(p), v); }
    template <int N> static __device__ __forceinline__ void store(__half *p, const float (&v)[N])
    {
        static_assert(N == 8, "fp16 stores are 8 wide");
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        stg_stream(p, make_uint4(w[0], w[1], w[2], w[3]));
    }
};

// widen a raw 16-byte vector and return the max|x| pattern (as an fp32 bit pattern) of its
// elements.  For 16-bit sources the max is taken on the packed words (2 elements per
// instruction) before widening.
template <typename T> __device__ __forceinline__ uint32_t unpack_absmax(const uint4 &r, float (&v)[VecIO<T>::V])
{
    VecIO<T>::unpack(r, v);
    if constexpr (sizeof(T) == 4) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) m = max(m, f2u(v[j]) & 0x7FFFFFFFu);
        return m;
    } else {
        uint32_t m2 = __vmaxu2(__vmaxu2(r.x & 0x7FFF7FFFu, r.y & 0x7FFF7FFFu), __vmaxu2(r.z & 0x7FFF7FFFu, r.w & 0x7FFF7FFFu));
        uint32_t m16 = max(m2 & 0xFFFFu, m2 >> 16);
        if constexpr (std::is_same<T, __nv_bfloat16>::value) return m16 << 16;
        else return f2u(__half2float(__ushort_as_half((unsigned short)m16)));
    }
}

template <typename Tout> __device__ __forceinline__ float requant1(float v)
{
    if constexpr (sizeof(Tout) == 4) return v;
    else return Cvt<Tout>::to_f32(Cvt<Tout>::from_f32(v));
}

// ------------------------------------------------------------------------------------------------
// out-of-line scalar paths (one copy per kernel)
static __device__ __noinline__ float bfp_elem_slow(float x, uint32_t maxabs_bits, int wl, int sh, uint32_t mask, int mode, int asym, uint32_t rnd)
{
    BfpBlock b = bfp_block(maxabs_bits, wl);
    float t = __fadd_rn(x, b.base);
    uint32_t tb = round_bits_rt(f2u(t), sh, mask, mode, rnd);
    float q = __fsub_rn(u2f(tb), b.base);
    uint32_t qb = f2u(q);
    if ((qb & 0x7F800000u) > b.E) qb = (qb & 0x80000000u) | b.maxnum;
    q = u2f(qb);
    if (asym) q = bfp_asym_fix(q, x, b);
    return q;
}
static __device__ __noinline__ float float_elem_slow(float x, const FloatFmt *f, uint32_t rnd) { return float_elem_rt(x, *f, rnd); }
static __device__ __noinline__ float fixed_elem_slow(float x, const FixedFmt *f, int affine, float sc, float zp, float r)
{
    return affine ? fixed_elem_affine(x, *f, sc, zp, r) : fixed_elem(x, *f, r);
}

// SBFP block header (once per block per lane) and element
// Correctly rounded a / b without the division sequence, given rb = RN(1/b): q0 = a*rb is within 2 ulp, one Newton
// step q += (a - q*b)*rb (residual by FMA) makes it faithful and a second one rounds correctly (Markstein), i.e. the
// result equals __fdiv_rn(a, b) bit for bit -- for a >= 0, b > 0 both well inside the normal range and b's significand
// not all ones (the callers check; they fall back on __fdiv_rn otherwise).  (A -0 numerator would come out as +0.)
__device__ __forceinline__ float div_by_recip(float a, float b, float rb)
{
    float q = __fmul_rn(a, rb);
    q = __fmaf_rn(__fmaf_rn(-q, b, a), rb, q);
    return __fmaf_rn(__fmaf_rn(-q, b, a), rb, q);
}
// The same quotient in four operations, for divisors that serve many dividends: with the reciprocal carried as a
// high / low pair (rh = RN(1/b), rl ~ 1/b - rh) the first estimate q0 = RN(a*rh + RN(a*rl)) is already faithful, and ONE
// Markstein correction q0 + (a - q0*b)*rh rounds it correctly.  Same preconditions as div_by_recip.  (Checked against
// IEEE division on 4e8 random, tie-adjacent and 16-bit-significand operand pairs on the CPU; the SBFP tie tests hammer
// it on the GPU.)
__device__ __forceinline__ float recip_lo(float b, float rh) { return __fmul_rn(__fmaf_rn(-b, rh, 1.0f), rh); }
// ... and in TWO operations when the dividend has at most 16 significant bits (a widened bf16 / fp16 value: 8 / 11 bits).
// rh + rl equals 1/b to ~2^-47 and RN(a*rl) adds ~2^-48, so q0 = RN(a*rh + RN(a*rl)) is the correct rounding of a value
// within 2^-46 (relative) of a/b -- and a/b itself cannot come closer than 2^-(s+25) to a rounding boundary of fp32: with
// a = A*2^i (A < 2^s), b = B*2^j (B < 2^24) and a boundary (2Q+1)*2^k, |A/B - (2Q+1)*2^k'| = |A*2^m - (2Q+1)*B| / (B*2^m) is a
// non-zero integer over B*2^m.  For s <= 16 that is >= 2^-41 >> 2^-46: q0 is already a/b correctly rounded, no correction step.
// (tests/native/div_by_recip_check.c: 2*10^8 quotients with 8- / 11- / 16-bit dividends incl. tie neighbourhoods, 0 mismatches;
// with 24-bit dividends the same form does fail, rarely.)  Same preconditions as div_by_recip.
__device__ __forceinline__ float div_by_recip16(float a, float rh, float rl) { return __fmaf_rn(a, rh, __fmul_rn(a, rl)); }
// round half away from zero, then clamp to the integers [t_min, t_max] (|t| <= 2^22), for finite a: clamping FIRST -- to
// [t_min - 0.25, t_max + 0.25], which rounds to the same integers -- bounds the magnitude, so the rounding is
// trunc(|c| + 0.5) with the addition rounded toward zero (it cannot round up into the next integer) and the sign put back:
// five instructions instead of roundf's sequence plus two.  lo = t_min - 0.25, hi = t_max + 0.25.
__device__ __forceinline__ float round_away_clamped(float a, float lo, float hi)
{
    const float c = fminf(fmaxf(a, lo), hi);
    return copysignf(truncf(__fadd_rz(fabsf(c), 0.5f)), c);
}
__device__ __forceinline__ float div_by_recip2(float a, float b, float rh, float rl)
{
    const float q = __fmaf_rn(a, rh, __fmul_rn(a, rl));
    return __fmaf_rn(__fmaf_rn(-q, b, a), rh, q);
}
__device__ __forceinline__ bool recip_safe(float b) { return b > 0x1p-60f && b < 0x1p60f && (f2u(b) & 0x7FFFFFu) != 0x7FFFFFu; }

// `sc_thr`: the scaler format's flush threshold (FloatFmt.shift_exp) -- the one quantity of the fast scaler cast that depends
// on the exponent bias; passed separately so a bias derived on the device (from an all-reduced amax) can replace the
// host-decided one while every other format constant keeps coming from the kernel parameters
__device__ __forceinline__ SbfpBlock sbfp_block_ol(uint32_t maxabs_bits, const SbfpFmt &f, uint32_t sc_thr)
{
    SbfpBlock b;
    const float m = u2f(maxabs_bits);
    // max / man_scaling: man_scaling = 2^(p-1) - 1 never has an all-ones significand, its reciprocal comes from the host
    if (f.recip) b.cmax = __fmul_rn(m, f.inv_man_t);  // torch's CUDA `max / man_scaling` (a multiplication by the rounded reciprocal)
    else b.cmax = (m > 0x1p-60f && m < 0x1p60f) ? div_by_recip(m, f.man_scaling, f.inv_man) : __fdiv_rn(m, f.man_scaling);
    // scaler cast: cmax >= 0 (or NaN, in which case the block passes through and fs is unused), so the
    // unsigned scaler formats of the SBFP aliases reduce to the signed nearest+flush fast path
    if (f.sc_fast) {  // float_elem_flush_nearest<false> with the threshold made explicit
        const uint32_t ab = f2u(b.cmax) & 0x7FFFFFFFu;
        const uint32_t mag = min(round_bits<R_NEAREST>(ab, f.sc.sh, f.sc.mask, 0u), f.sc.max_num);
        b.fs = ab < sc_thr ? 0.0f : u2f(mag | (f2u(b.cmax) & 0x80000000u));
    } else {
        b.fs = float_elem_slow(b.cmax, &f.sc, 0u);
    }
    b.on = b.cmax > 0.0f;
    b.rok = recip_safe(b.cmax);
    b.rc = __frcp_rn(b.cmax);
    b.rl = recip_lo(b.cmax, b.rc);
    return b;
}
__device__ __forceinline__ SbfpBlock sbfp_block_ol(uint32_t maxabs_bits, const SbfpFmt &f) { return sbfp_block_ol(maxabs_bits, f, f.sc.shift_exp); }
__device__ __forceinline__ bool sbfp_fast(const SbfpFmt &f) { return f.xp.mode == R_NEAREST && f.xp.tie == TIE_AWAY; }
// XP[p,0] nearest, half away (the reference's CUDA rule) of the IEEE quotient; fl = 0 => no scaling multiplies.
__device__ __forceinline__ float sbfp_elem_fast(float x, const SbfpBlock &b, const SbfpFmt &f)
{
    float v = roundf(__fdiv_rn(x, b.cmax));
    if (f.xp.clamp) v = v > f.xp.t_max ? f.xp.t_max : (v < f.xp.t_min ? f.xp.t_min : v);
    return b.on ? __fmul_rn(v, b.fs) : x;
}
// The same on a register vector of one block, division free for ordinary blocks.  Works on magnitudes (|x| is a free
// operand modifier): the quotient |x| / cmax is in [0, 7.0000005] (cmax = RN(max|x| / man_scaling)), so
//   * round half away is trunc(q + 0.5) with the addition rounded toward zero (what roundf itself does, minus the sign
//     handling): two instructions;
//   * the clamp to [t_min, t_max] cannot trigger when |t_min|, t_max >= man_scaling (every SBFP format), so it is skipped;
//   * the result carries x's sign (cmax, fs >= 0).
template <int V, bool SRC16 = false> __device__ __forceinline__ void sbfp_apply(float (&v)[V], const SbfpBlock &b, const SbfpFmt &f)
{
    if (!b.on) return;  // all-zero (or NaN) block: passes through
    if (b.rok && f.no_clamp) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float a = fabsf(v[j]);
            const float q = SRC16 ? div_by_recip16(a, b.rc, b.rl) : div_by_recip2(a, b.cmax, b.rc, b.rl);
            const float r = truncf(__fadd_rz(q, 0.5f));  // round half away of q >= 0: the toward-zero add cannot round up into the next integer
            v[j] = copysignf(__fmul_rn(r, b.fs), v[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = sbfp_elem_fast(v[j], b, f);
    }
}
static __device__ __noinline__ float sbfp_elem_slow(float x, float cmax, float fs, const FixedFmt *xp)
{
    if (!(cmax > 0.0f)) return x;
    float v = fixed_elem(__fdiv_rn(x, cmax), *xp, 0.5f);
    return __fmul_rn(v, fs);
}

// ------------------------------------------------------------------------------------------------
// stages on one register vector (V consecutive elements along the blocked dim)
template <int V> __device__ __forceinline__ uint32_t vec_absmax(const float (&v)[V])
{
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) m = max(m, f2u(v[j]) & 0x7FFFFFFFu);
    return m;
}
template <int LANES> __device__ __forceinline__ uint32_t lanes_max_n(uint32_t m)
{
#pragma unroll
    for (int off = 1; off < LANES; off <<= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    return m;
}
__device__ __forceinline__ uint32_t lanes_max(uint32_t m, int lanes)
{
    // (warp-uniform) the usual block sizes get straight-line shuffles; the loop form costs ~6 instructions per step
    switch (lanes) {
    case 1: return m;
    case 2: return lanes_max_n<2>(m);
    case 4: return lanes_max_n<4>(m);
    case 8: return lanes_max_n<8>(m);
    case 16: return lanes_max_n<16>(m);
    default: break;
    }
    for (int off = 1; off < lanes; off <<= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    return m;
}

template <int V> __device__ __forceinline__ void bfp_stage(float (&v)[V], const StageDev &st, int lanes, const uint32_t (&r)[V])
{
    uint32_t m = lanes_max(vec_absmax<V>(v), lanes);
    if (st.mode == R_NEAREST && st.fast && bfp_fast_ok(m)) {
        BfpFast b = bfp_fast_block(m, st.wl);
        float x0[V];
        if (st.asym) {
#pragma unroll
            for (int j = 0; j < V; ++j) x0[j] = v[j];
        }
        if (st.fast16) {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = bfp_fast16_elem(v[j], b);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = bfp_fast_elem(v[j], b);
        }
        if (b.clamp) {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = bfp_clamp(v[j], b);
        }
        if (st.asym) {
            // the edge mantissa -(2^(wl-1)-1) needs |x| >= maxval - Q/2: only blocks whose max is within one
            // quantum of maxval (a safe superset) can hold it
            BfpBlock bb = bfp_block(m, st.wl);
            if (m >= bb.maxnum - (1u << (23 - (st.wl - 2)))) {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = bfp_asym_fix(v[j], x0[j], bb);
            }
        }
    } else if (st.mode == R_NEAREST) {
        BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float q = bfp_elem<R_NEAREST>(v[j], b, st.sh, st.mask, 0u);
            if (st.asym) q = bfp_asym_fix(q, v[j], b);
            v[j] = q;
        }
    } else if (st.mode == R_STOCHASTIC && !st.asym) {  // stochastic rounding is a first-class mode: inlined
        BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem<R_STOCHASTIC>(v[j], b, st.sh, st.mask, r[j]);
    } else if (st.mode == R_UP && !st.asym) {  // directed rounding: inlined too (the out-of-line call per element cost 2x)
        BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem<R_UP>(v[j], b, st.sh, st.mask, 0u);
    } else if (st.mode == R_DOWN && !st.asym) {
        BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem<R_DOWN>(v[j], b, st.sh, st.mask, 0u);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem_slow(v[j], m, st.wl, st.sh, st.mask, st.mode, st.asym, r[j]);
    }
}

template <int V> __device__ __forceinline__ void sbfp_stage(float (&v)[V], const StageDev &st, int lanes)
{
    uint32_t m = lanes_max(vec_absmax<V>(v), lanes);
    SbfpBlock b = sbfp_block_ol(m, st.sb);
    if (sbfp_fast(st.sb)) {
        sbfp_apply<V>(v, b, st.sb);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = sbfp_elem_slow(v[j], b.cmax, b.fs, &st.sb.xp);
    }
}

// ------------------------------------------------------------------------------------------------
// Low-bit float with subnormals kept (float_elem_nearest_sub), on PACKED 16-bit values: a bf16 / fp16 tensor cast to a
// format of man <= 5 mantissa bits (FP8 E4M3 / E5M2, the MX element formats) never needs fp32.  For a 16-bit source the
// reference's sequence (x + shift is exact: 8 / 11 significant bits) is ONE rounding of x, ties to even, to a multiple of
//     Q = 2^(max(e(x), min_exp) - man)
// followed by the saturation of the magnitude.  In the source's own arithmetic that rounding is (x + C) - C with
// C = 1.5 * 2^(max(e, min_exp) + P - man), P = 7 (bf16) / 10 (fp16) significand bits: |x| < 2^(max(e,min_exp)+1) <= C / 3 for
// man <= 5, so the sum stays in C's binade, whose ulp is Q; C / Q is even, so ties land on the same neighbour as the
// reference's magnitude arithmetic for either sign; the subtraction is exact; a zero result is +0 as in the reference
// (-0 and negative values that round to zero included).  C's exponent field is a packed max + add on the raw words: two
// elements per instruction throughout -- 8 instructions per PAIR instead of ~12 per element plus widening / narrowing.
// The caller guarantees finite inputs whose exponent leaves room for C (pattern below Sub16.limit); the MX
// variant passes per-block constants (the block scale folded into min_exp and the saturation value).
template <typename T> struct Arith16;
template <> struct Arith16<__nv_bfloat16> {
    static constexpr uint32_t kExp2 = 0x7F807F80u;
    static constexpr int kMan = 7;
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b)
    {
        uint32_t d;
        asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
    static __device__ __forceinline__ uint32_t sub2(uint32_t a, uint32_t b)
    {
        uint32_t d;
        asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
};
template <> struct Arith16<__half> {
    static constexpr uint32_t kExp2 = 0x7C007C00u;
    static constexpr int kMan = 10;
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b)
    {
        uint32_t d;
        asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
    static __device__ __forceinline__ uint32_t sub2(uint32_t a, uint32_t b)
    {
        uint32_t d;
        asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
};
struct Sub16 {
    uint32_t minexp2;  // exponent field of 2^min_exp, in both halves
    uint32_t magic2;   // ((P - man) << P) | (1 << (P - 1)), in both halves: exponent offset + mantissa 1.5 of C
    uint32_t maxnum2;  // saturation magnitude pattern, in both halves
    uint32_t limit;    // fast path only for vectors whose largest magnitude pattern is below this (finite, room for C)
    bool ok;           // the format (and, for MX, the block scale) qualifies
};
struct Sub16Fmt {      // the per-format part, derived once per thread
    int me0, xe0;      // exponent fields (in T) of 2^min_exp and of the saturation value, before any block scale
    int room;          // P - man
    uint32_t manbits;  // mantissa field of the saturation value in T
    uint32_t magic2, limit;
    bool ok;
};
template <typename T> __device__ __forceinline__ Sub16Fmt sub16_fmt(const FloatFmt &f)
{
    constexpr int P = Arith16<T>::kMan;
    constexpr int BIAS = std::is_same<T, __nv_bfloat16>::value ? 127 : 15;
    constexpr int EMAX = std::is_same<T, __nv_bfloat16>::value ? 254 : 30;
    Sub16Fmt c;
    const int man = 23 - f.sh;
    c.room = P - man;
    c.me0 = f.min_exp + BIAS;
    c.xe0 = (int)f.max_store - 127 + BIAS;
    c.manbits = (f.max_num & 0x007FFFFFu) >> (23 - P);
    // (fp16: x + 2^min_exp must be exact in the reference's fp32 add: min_exp + 25 bits at most 24)
    c.ok = f.nsub && man <= 5 && (std::is_same<T, __nv_bfloat16>::value || f.min_exp <= -1);
    const uint32_t mg = ((uint32_t)c.room << P) | (1u << (P - 1));
    c.magic2 = mg * 0x10001u;
    c.limit = (uint32_t)(EMAX + 1 - c.room) << P;
    return c;
}
// e_off = extra exponent (the MX block scale); 0 for a plain FLOAT stage
template <typename T> __device__ __forceinline__ Sub16 sub16_consts(const Sub16Fmt &b, int e_off)
{
    constexpr int P = Arith16<T>::kMan;
    constexpr int EMAX = std::is_same<T, __nv_bfloat16>::value ? 254 : 30;
    Sub16 c;
    const int me = b.me0 + e_off, xe = b.xe0 + e_off;
    c.ok = b.ok && me >= 1 && me <= EMAX - b.room && xe >= 1;
    // a saturation value beyond T's range cannot trigger: the largest finite pattern stands in for it
    const uint32_t mx = xe > EMAX ? (std::is_same<T, __nv_bfloat16>::value ? 0x7F7Fu : 0x7BFFu) : (((uint32_t)max(xe, 0) << P) | b.manbits);
    c.minexp2 = ((uint32_t)max(me, 0) << P) * 0x10001u;
    c.maxnum2 = mx * 0x10001u;
    c.magic2 = b.magic2;
    c.limit = b.limit;
    return c;
}
template <typename T> __device__ __forceinline__ uint32_t sub16_pair(uint32_t w, const Sub16 &c)
{
    const uint32_t C = __vmaxu2(w & Arith16<T>::kExp2, c.minexp2) + c.magic2;
    const uint32_t q = Arith16<T>::sub2(Arith16<T>::add2(w, C), C);
    return __vminu2(q & 0x7FFF7FFFu, c.maxnum2) | (q & 0x80008000u);
}
template <typename T> __device__ __forceinline__ uint4 sub16_vec(const uint4 &r, const Sub16 &c)
{
    return make_uint4(sub16_pair<T>(r.x, c), sub16_pair<T>(r.y, c), sub16_pair<T>(r.z, c), sub16_pair<T>(r.w, c));
}

static __device__ __noinline__ float mx_elem_ol(float x, float scale, const FloatFmt *f)
{
    MxBlock b{scale};
    return mx_elem(x, b, *f);
}
// MXFP on one register vector whose block max|x| pattern `m` is known
template <int V> __device__ __forceinline__ void mxfp_apply(float (&v)[V], uint32_t m, const StageDev &st)
{
    // floor(log2f(max)) is the exponent field whenever log2(max) is further than a few ulp from an integer,
    // i.e. unless the mantissa is within 2^-16 of a power of two (log2f is only faithful, and the reference
    // calls it: in that sliver -- and for denormal / non-finite maxima -- so do we).
    const uint32_t man = m & 0x007FFFFFu, ef = m >> 23;
    const int shift = (int)((f2u(st.mx_largest) >> 23) - 127u);
    MxBlock b;
    if (man >= 128u && man <= 0x007FFF00u && ef >= 1u && ef <= 254u && (int)ef - shift >= 2 && (int)ef - shift <= 252) b.scale = u2f((uint32_t)((int)ef - shift) << 23);
    else b = mx_block(m, st.mx_largest);
    const uint32_t sb = f2u(b.scale);
    if ((sb & 0x807FFFFFu) == 0u && sb >= (2u << 23) && sb <= (252u << 23)) {
        // ordinary block: the scale is a normal power of two, so x / scale == x * (1 / scale) exactly
        const float inv = u2f((254u << 23) - sb);
        if (st.ff.nsub) {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = __fmul_rn(float_elem_nearest_sub(__fmul_rn(v[j], inv), st.ff), b.scale);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = __fmul_rn(float_elem_nearest(__fmul_rn(v[j], inv), st.ff), b.scale);
        }
    } else {  // zero / denormal / huge / non-finite block: the literal op sequence
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = mx_elem_ol(v[j], b.scale, &st.ff);
    }
}
template <int V> __device__ __forceinline__ void mxfp_stage(float (&v)[V], const StageDev &st, int lanes)
{
    mxfp_apply<V>(v, lanes_max(vec_absmax<V>(v), lanes), st);
}

template <int V> __device__ __forceinline__ void float_stage(float (&v)[V], const StageDev &st, const uint32_t (&r)[V])
{
    if (st.ff.fastpath) {
        const bool any_nan = vec_absmax<V>(v) > 0x7F800000u;
        float q[V];
        if (st.ff.exact) {
#pragma unroll
            for (int j = 0; j < V; ++j) q[j] = float_elem_flush_nearest<true>(v[j], st.ff);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) q[j] = float_elem_flush_nearest<false>(v[j], st.ff);
        }
        if (any_nan) {
#pragma unroll
            for (int j = 0; j < V; ++j) q[j] = float_elem_slow(v[j], &st.ff, 0u);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = q[j];
    } else if (st.ff.mode == R_NEAREST) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = float_elem_nearest(v[j], st.ff);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = float_elem_slow(v[j], &st.ff, r[j]);
    }
}

template <int V> __device__ __forceinline__ void fixed_stage(float (&v)[V], const StageDev &st, const uint32_t (&r)[V])
{
    if (st.xf.mode == R_NEAREST && st.xf.tie == TIE_AWAY && !st.affine) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float a = __fmul_rn(roundf(__fmul_rn(v[j], st.xf.up)), st.xf.down);
            if (st.xf.clamp) a = a > st.xf.t_max ? st.xf.t_max : (a < st.xf.t_min ? st.xf.t_min : a);
            v[j] = a;
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = fixed_elem_slow(v[j], &st.xf, st.affine, st.sc, st.zp, u2f(r[j]));
    }
}

// N:M inside one thread (M <= V).  Stable ascending order: for a < b, "a sorts before b" iff
// key[a] <= key[b]; one comparison per pair feeds both ranks (rank = number of elements that
// sort before it); the n_prune lowest ranks are pruned.
template <int V, int M> __device__ __forceinline__ void nm_local(const uint32_t (&key)[V], int n_prune, bool (&keep)[V])
{
#pragma unroll
    for (int g0 = 0; g0 < V; g0 += M) {
        int rank[M];
#pragma unroll
        for (int a = 0; a < M; ++a) rank[a] = 0;
#pragma unroll
        for (int a = 0; a < M; ++a)
#pragma unroll
            for (int b = a + 1; b < M; ++b) {
                int a_first = key[g0 + a] <= key[g0 + b] ? 1 : 0;
                rank[b] += a_first;
                rank[a] += 1 - a_first;
            }
#pragma unroll
        for (int a = 0; a < M; ++a) keep[g0 + a] = rank[a] >= n_prune;
    }
}

// N:M inside one thread for 16-bit sources scored by |x| (M <= 8): the |x| pattern of a bf16 / fp16
// value fits 19 bits, so (pattern << 3) | index is a *unique* 22-bit key whose order is exactly the
// stable ascending order.  A min/max sorting network on the keys yields the n_prune-th smallest;
// everything at or above it is kept.  5 compare-exchanges for M = 4, 19 for M = 8.
__device__ __forceinline__ void cex(uint32_t &a, uint32_t &b) { uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

template <int V, int M, int KEYSHIFT> __device__ __forceinline__ void nm_local16(const float (&v)[V], int n_prune, bool (&keep)[V])
{
#pragma unroll
    for (int g0 = 0; g0 < V; g0 += M) {
        uint32_t k[M], s[M];
#pragma unroll
        for (int a = 0; a < M; ++a) {
            k[a] = (min((f2u(v[g0 + a]) & 0x7FFFFFFFu) >> KEYSHIFT, (0x7F800000u >> KEYSHIFT) + 1u) << 3) | (uint32_t)a;
            s[a] = k[a];
        }
        if (M == 2) {
            cex(s[0], s[1]);
        } else if (M == 4) {
            cex(s[0], s[1]); cex(s[2], s[3]); cex(s[0], s[2]); cex(s[1], s[3]); cex(s[1], s[2]);
        } else {  // M == 8, 19-comparator network
            cex(s[0], s[1]); cex(s[2], s[3]); cex(s[4], s[5]); cex(s[6], s[7]);
            cex(s[0], s[2]); cex(s[1], s[3]); cex(s[4], s[6]); cex(s[5], s[7]);
            cex(s[1], s[2]); cex(s[5], s[6]); cex(s[0], s[4]); cex(s[3], s[7]);
            cex(s[1], s[5]); cex(s[2], s[6]);
            cex(s[1], s[4]); cex(s[3], s[6]);
            cex(s[2], s[4]); cex(s[3], s[5]);
            cex(s[3], s[4]);
        }
        uint32_t thr = s[0];
#pragma unroll
        for (int a = 1; a < M; ++a) thr = (a == n_prune) ? s[a] : thr;
#pragma unroll
        for (int a = 0; a < M; ++a) keep[g0 + a] = n_prune == 0 || k[a] >= thr;
    }
}

// N:4 with score |x| straight on the RAW 16-bit patterns of a 16-byte vector (bf16 or fp16, eight finite values): the
// magnitude pattern of either format orders like the magnitude, so ((pattern & 0x7FFF) << 3) | index is a unique key
// in stable ascending order.  The n_prune-th smallest key of a group (1 <= n_prune <= 3) comes from six min/max
// operations on the two sorted pairs; everything at or above it is kept.
// Inf / NaN (where x * 0 is not a signed zero, and NaN payloads must not order) are excluded by the caller.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)  // a * b + c kept as one FMA-pipe IMAD
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t raw16_absmax(const uint4 &r)
{
    uint32_t m2 = __vmaxu2(__vmaxu2(r.x & 0x7FFF7FFFu, r.y & 0x7FFF7FFFu), __vmaxu2(r.z & 0x7FFF7FFFu, r.w & 0x7FFF7FFFu));
    return max(m2 & 0xFFFFu, m2 >> 16);
}
__device__ __forceinline__ uint32_t raw16_absmin(const uint4 &r)
{
    uint32_t m2 = __vminu2(__vminu2(r.x & 0x7FFF7FFFu, r.y & 0x7FFF7FFFu), __vminu2(r.z & 0x7FFF7FFFu, r.w & 0x7FFF7FFFu));
    return min(m2 & 0xFFFFu, m2 >> 16);
}
// 16-bit magnitude patterns of a 16-bit dtype: order == order of the values they encode
template <typename T> __device__ __forceinline__ uint32_t pattern16_ru(float f);  // smallest pattern whose value is >= f (f > 0)
template <> __device__ __forceinline__ uint32_t pattern16_ru<__nv_bfloat16>(float f) { return (f2u(f) + 0xFFFFu) >> 16; }
template <> __device__ __forceinline__ uint32_t pattern16_ru<__half>(float f) { return (uint32_t)__half_as_ushort(__float2half_ru(f)); }
template <> __device__ __forceinline__ uint32_t pattern16_ru<float>(float) { return 0u; }
template <typename T> __device__ __forceinline__ uint32_t pattern16_rn(float f);  // pattern of f rounded to T (f >= 0)
template <> __device__ __forceinline__ uint32_t pattern16_rn<__nv_bfloat16>(float f) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f)); }
template <> __device__ __forceinline__ uint32_t pattern16_rn<__half>(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }
template <> __device__ __forceinline__ uint32_t pattern16_rn<float>(float) { return 0u; }

__device__ __forceinline__ void nm4_keep_raw16(const uint4 &r, int n_prune, bool (&keep)[8])
{
    // keys: magnitude pattern in bits 17..31, index in the low bits.  The multiplications are shifts that drop the
    // sign (and, for the low half, the other element) and run on the FMA pipe; the kernels these serve are bound by
    // the integer ALU pipe, so every LOP3 / VIMNMX saved counts.
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint32_t k[4];
        k[0] = imad(w[2 * g], 0x20000u, 0u);
        k[1] = imad(w[2 * g] & 0x7FFF0000u, 2u, 1u);
        k[2] = imad(w[2 * g + 1], 0x20000u, 2u);
        k[3] = imad(w[2 * g + 1] & 0x7FFF0000u, 2u, 3u);
        const uint32_t lo1 = min(k[0], k[1]), hi1 = max(k[0], k[1]), lo2 = min(k[2], k[3]), hi2 = max(k[2], k[3]);
        uint32_t thr;  // the n_prune-th smallest key (0-based): everything from it upwards is kept
        if (n_prune == 2) thr = __vimax3_u32(min(hi1, hi2), lo1, lo2);
        else if (n_prune == 1) thr = __vimin3_u32(max(lo1, lo2), hi1, hi2);
        else thr = max(hi1, hi2);
#pragma unroll
        for (int a = 0; a < 4; ++a) keep[4 * g + a] = k[a] >= thr;
    }
}
// x * mask on the packed words: a pruned finite value becomes a zero of its own sign
__device__ __forceinline__ uint4 nm_apply_raw16(const uint4 &r, const bool (&keep)[8])
{
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] &= (keep[2 * i] ? 0x0000FFFFu : 0x00008000u) | (keep[2 * i + 1] ? 0xFFFF0000u : 0x80000000u);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------
// nearest + flush FLOAT stage on a register vector (dmxq_add_cast, dmxq_softmax_cast)
template <int V> __device__ __forceinline__ void float_fast_vec(float (&v)[V], const FloatFmt &f)
{
    const bool any_nan = vec_absmax<V>(v) > 0x7F800000u;
    float q[V];
    if (f.exact) {  // values already fit the format's mantissa (16-bit tensor dtype): flush + saturate only
#pragma unroll
        for (int j = 0; j < V; ++j) q[j] = float_elem_flush_nearest<true>(v[j], f);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) q[j] = float_elem_flush_nearest<false>(v[j], f);
    }
    if (any_nan) {
#pragma unroll
        for (int j = 0; j < V; ++j) q[j] = float_elem_slow(v[j], &f, 0u);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = q[j];
}

// 16-bit helpers: pack a register vector to T (round to nearest: the tensor dtype's rounding) and the "this FLOAT stage
// is the identity on the whole vector" test on the packed patterns (see f16_same in dmxq_rows.cuh)
template <typename T> __device__ __forceinline__ uint4 pack16(const float (&v)[8])
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        } else {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
struct Range16 {
    uint32_t lo, hi;
    bool on;  // the stage exists, keeps T's significand, and is the signed nearest+flush kind
    uint32_t sat;  // the format's saturation value rounded to T (fp16: FLOAT16's 131008 becomes Inf, as CastTo's `.to(dtype)` makes it)
};
template <typename T> __device__ __forceinline__ Range16 range16(int has, const FloatFmt &f)
{
    Range16 r;
    r.on = has && f.exact && !f.is_unsigned;
    r.lo = pattern16_ru<T>(u2f(f.shift_exp));
    r.sat = pattern16_rn<T>(u2f(f.max_num));
    r.hi = min(r.sat, (uint32_t)(std::is_same<T, __half>::value ? 0x7BFFu : 0x7F7Fu));
    return r;
}
__device__ __forceinline__ bool inside16(const uint4 &w, const Range16 &r) { return r.on && raw16_absmin(w) >= r.lo && raw16_absmax(w) <= r.hi; }

// The same stage (r.on: nearest + flush + signed, T's significand kept, so only flush and saturate act) applied to a vector of
// packed 16-bit values, two elements per instruction: magnitude patterns below r.lo become +0, above r.sat become r.sat with the
// sign kept -- what float_elem_flush_nearest<true> followed by the rounding to T yields, element for element.  The flush mask
// comes from bit 15 of (magnitude + 0x8000 - lo), replicated over its half by a sign-extending byte permute.  A NaN anywhere in
// the vector (the one input class whose reference result depends on the payload) takes the per-element literal path.
// (one out-of-line copy per kernel of the literal per-element path: it is reached by vectors that hold a NaN only)
template <typename T> static __device__ __noinline__ uint4 float_vec16_literal(uint4 w, const FloatFmt *f)
{
    float v[8];
    VecIO<T>::unpack(w, v);
    float_fast_vec<8>(v, *f);
    return pack16<T>(v);
}
template <typename T> __device__ __forceinline__ uint4 flush_sat16_vec(const uint4 &w, const FloatFmt &f, const Range16 &r)
{
    constexpr uint32_t kInf16 = std::is_same<T, __nv_bfloat16>::value ? 0x7F80u : 0x7C00u;
    const uint32_t x[4] = {w.x, w.y, w.z, w.w};
    uint32_t mag[4], mx = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) { mag[i] = x[i] & 0x7FFF7FFFu; mx = __vmaxu2(mx, mag[i]); }
    if (max(mx & 0xFFFFu, mx >> 16) > kInf16) return float_vec16_literal<T>(w, &f);
    const uint32_t hi2 = r.sat * 0x10001u, k2 = (0x8000u - r.lo) * 0x10001u;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t keep;  // 0xFFFF per half whose magnitude is >= lo (prmt with selector msb set replicates the sign bit of the byte)
        asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(keep) : "r"(mag[i] + k2), "r"(0u));
        o[i] = ((x[i] & 0x80008000u) | __vminu2(mag[i], hi2)) & keep;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}


// ------------------------------------------------------------------------------------------------
// N:M tie order of the reference on CUDA tensors (DMXQ_NM_ORDER_TORCH_CUDA).  BlockTopK sorts every group with
// torch.argsort(score, dim=1) (S/sparse.py:172); for rows of <= 32 keys ATen runs bitonicSortKVInPlace
// (ATen/native/cuda/SortUtils.cuh) with 32 slots, the M keys in slots 0..M-1 and "invalid" slots behind them:
//     swap = (LT(kA, kB) && validA) || !validB;   if (swap == dir) exchange(A, B)
// so an ascending comparator (dir = false) exchanges unless kA < kB -- tied keys ARE exchanged -- and a descending
// one (dir = true) exchanges when kA < kB.  The result for tied keys is a fixed function of the tie pattern, but not
// a tie-break by index (scripts/probe_argsort.py: this model == torch.argsort on every tie pattern, any batch / dtype).
//
// With M a power of two the valid slots never leave 0..M-1 (an ascending comparator never moves a valid key behind an
// invalid one, and slots 0..M-1 only meet descending comparators among themselves), so the 32-slot network reduces to
//   * the bitonic build on M slots (stages size = 2..M, ATen's direction flags), then
//   * log2(32 / M) more ascending merge passes over strides M/2..1 (the size 2M..32 stages: larger strides pair a
//     valid slot with an invalid one and do nothing) -- which re-exchange the ties of the already sorted keys.
// Returns the bit set of the n_prune first (= pruned) elements.
template <int M> __device__ __forceinline__ uint32_t nm_pruned_torch(const uint32_t (&key)[M], int n_prune)
{
    uint32_t k[M], id[M];
#pragma unroll
    for (int a = 0; a < M; ++a) { k[a] = key[a]; id[a] = (uint32_t)a; }
#pragma unroll
    for (int size = 2; size <= 32; size *= 2) {
#pragma unroll
        for (int stride = (size <= M ? size / 2 : M / 2); stride > 0; stride /= 2) {
#pragma unroll
            for (int t = 0; t < M / 2; ++t) {
                const bool desc = size < M && (t & (size / 2)) != 0;
                const int a = 2 * t - (t & (stride - 1)), b = a + stride;
                const bool lt = k[a] < k[b];
                const bool ex = desc ? lt : !lt;
                const uint32_t ka = k[a], kb = k[b], ia = id[a], ib = id[b];
                k[a] = ex ? kb : ka; k[b] = ex ? ka : kb;
                id[a] = ex ? ib : ia; id[b] = ex ? ia : ib;
            }
        }
    }
    uint32_t pruned = 0u;
#pragma unroll
    for (int a = 0; a < M; ++a) pruned |= (a < n_prune) ? (1u << id[a]) : 0u;
    return pruned;
}

// The full 32-slot network for any group size M <= 32 (not a power of two, or larger than a thread's vector): scalar,
// local-memory arrays; used by the generic kernel.
static __device__ __noinline__ uint32_t nm_pruned_torch32(const uint32_t *key, int M, int n_prune)
{
    uint32_t k[32];
    unsigned char id[32];
    for (int a = 0; a < 32; ++a) { k[a] = a < M ? key[a] : 0u; id[a] = (unsigned char)a; }
    for (int size = 2; size <= 32; size *= 2)
        for (int stride = size / 2; stride > 0; stride /= 2)
            for (int t = 0; t < 16; ++t) {
                const bool dir = size < 32 && (t & (size / 2)) != 0;
                const int a = 2 * t - (t & (stride - 1)), b = a + stride;
                const bool va = id[a] < M, vb = id[b] < M;
                const bool swap = (k[a] < k[b] && va) || !vb;
                if (swap == dir) {
                    const uint32_t tk = k[a]; k[a] = k[b]; k[b] = tk;
                    const unsigned char ti = id[a]; id[a] = id[b]; id[b] = ti;
                }
            }
    uint32_t pruned = 0u;
    for (int a = 0; a < n_prune; ++a) pruned |= 1u << id[a];
    return pruned;
}

template <int V, int M> __device__ __forceinline__ void nm_local_torch(const uint32_t (&key)[V], int n_prune, bool (&keep)[V])
{
#pragma unroll
    for (int g0 = 0; g0 < V; g0 += M) {
        uint32_t kk[M];
#pragma unroll
        for (int a = 0; a < M; ++a) kk[a] = key[g0 + a];
        const uint32_t pruned = nm_pruned_torch<M>(kk, n_prune);
#pragma unroll
        for (int a = 0; a < M; ++a) keep[g0 + a] = ((pruned >> a) & 1u) == 0u;
    }
}

// N:M across `lanes` = M / V neighbouring lanes (M > V): partner keys arrive by shuffle.
template <int V> __device__ __forceinline__ void nm_lanes(const uint32_t (&key)[V], int n_prune, int lanes, int lane, bool (&keep)[V])
{
    int rank[V];
    int me = lane & (lanes - 1);
#pragma unroll
    for (int a = 0; a < V; ++a) {
        rank[a] = 0;
#pragma unroll
        for (int b = 0; b < V; ++b)
            if (b != a) rank[a] += (key[b] < key[a] || (key[b] == key[a] && b < a)) ? 1 : 0;
    }
    for (int d = 1; d < lanes; ++d) {
        bool partner_first = (me ^ d) < me;  // partner's elements have lower indices than mine
#pragma unroll
        for (int b = 0; b < V; ++b) {
            uint32_t pk = __shfl_xor_sync(0xFFFFFFFFu, key[b], d);
#pragma unroll
            for (int a = 0; a < V; ++a) rank[a] += (pk < key[a] || (pk == key[a] && partner_first)) ? 1 : 0;
        }
    }
#pragma unroll
    for (int a = 0; a < V; ++a) keep[a] = rank[a] >= n_prune;
}

template <int V, int SRCBITS = 32>
__device__ __forceinline__ void nm_stage(float (&v)[V], const StageDev &st, int lane, const float *score_vec, float *mask_vec, bool valid)
{
    if (SRCBITS != 32 && V == 8 && score_vec == nullptr && st.block <= 8 && !st.nm_order) {
        // 16-bit source, score |x|: unique packed keys + sorting network
        constexpr int KS = SRCBITS == 16 ? 16 : 13;  // bf16 keeps 8+7 bits, fp16 (widened) 8+10 bits
        bool keep[V];
        if (st.block == 2) nm_local16<V, 2, KS>(v, st.n_prune, keep);
        else if (st.block == 4) nm_local16<V, 4, KS>(v, st.n_prune, keep);
        else nm_local16<V, (V >= 8 ? 8 : V), KS>(v, st.n_prune, keep);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = nm_apply(v[j], keep[j]);
        if (mask_vec != nullptr && valid) {
            float mk[V];
#pragma unroll
            for (int j = 0; j < V; ++j) mk[j] = keep[j] ? 1.0f : 0.0f;
            VecIO<float>::store<V>(mask_vec, mk);
        }
        return;
    }
    uint32_t key[V];
    if (score_vec != nullptr) {
        if (valid) {
#pragma unroll
            for (int j = 0; j < V; j += 4) {
                uint4 r = ldg_stream(score_vec + j);
                key[j] = score_key(u2f(r.x)); key[j + 1] = score_key(u2f(r.y));
                key[j + 2] = score_key(u2f(r.z)); key[j + 3] = score_key(u2f(r.w));
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) key[j] = 0u;
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) key[j] = absx_key(v[j]);
    }
    bool keep[V];
    const int M = st.block;
    if (st.nm_order) {  // torch's CUDA tie order (the dispatcher only sends in-thread groups here: M <= V, M <= 8)
        if (M == 2) nm_local_torch<V, 2>(key, st.n_prune, keep);
        else if (M == 4) nm_local_torch<V, 4>(key, st.n_prune, keep);
        else nm_local_torch<V, (V >= 8 ? 8 : V)>(key, st.n_prune, keep);
    } else if (M > V) {
        nm_lanes<V>(key, st.n_prune, M / V, lane, keep);
    } else if (M == 2) {
        nm_local<V, 2>(key, st.n_prune, keep);
    } else if (M == 4) {
        nm_local<V, 4>(key, st.n_prune, keep);
    } else {
        nm_local<V, (V >= 8 ? 8 : V)>(key, st.n_prune, keep);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = nm_apply(v[j], keep[j]);
    if (mask_vec != nullptr && valid) {
        float mk[V];
#pragma unroll
        for (int j = 0; j < V; ++j) mk[j] = keep[j] ? 1.0f : 0.0f;
        VecIO<float>::store<V>(mask_vec, mk);
    }
}

void count_launch(int n = 1);

}  // namespace dmxq

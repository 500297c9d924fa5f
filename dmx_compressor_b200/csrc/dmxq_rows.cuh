// dmxq_rows.cuh -- chain_rows_kernel: the blocked dim is contiguous in memory.
//
// This is the kernel of the headline casts (Linear inputs / weights, q, k^T views, attention
// probabilities, whole-model weight casting).  One pass over HBM:
//   * each thread owns kUnroll independent 16-byte vectors; all loads are issued (into raw
//     registers) before the first use, so 64 B per thread / 16 KiB per CTA are in flight;
//   * a block of B elements lives in B/V neighbouring lanes of one warp; its max|x| is an
//     unsigned-integer max over bit patterns, reduced with log2(B/V) xor-shuffles;
//   * the fused round / clamp / rescale (+ N:M mask) runs on registers and the result is
//     written with one 16-byte streaming store.
// Algorithmic traffic: sizeof(in) + sizeof(out) bytes per element, nothing else.
//
// KIND selects a compile-time specialisation of the stage list (static parameter offsets, no
// stage loop / switch); the runtime-chain variants remain as the general path:
//   K_AUX        runtime chain, score / mask / rand tensors honoured
//   K_CHAIN      runtime chain, no auxiliary tensors
//   K_BFP        [BFP nearest symmetric]                      BFP16 / BFP12 / MXINT casts
//   K_FLOAT      [FLOAT nearest: flush+signed fast form, subnormal-keeping form, or the general element function]  FLOAT16 / BFLOAT16
//                boundary casts, FP8, bias casts
//   K_FLOAT_BFP  [FLOAT nearest+flush+signed -> BFP n.s.]     output cast fused with next input cast
//   K_NM_BFP     [N:M with score |x| -> BFP n.s.]             sparsify -> weight cast (hypernet)
//   K_SBFP       [SBFP, XP nearest half-away]                 SBFP weight storage cast
//   K_NM         [N:M with score |x|]                          Sparsify alone (no mask output)
//   K_FIXED      [FixedPoint nearest half-away, per-tensor affine (immediate or device qparams)]  INT8 / INT4
//   K_MXFP       [MXFP]                                       OCP-MX style power-of-two block scale + low-bit float elements
//   K_BFP_ASYM   [BFP nearest, asymmetric mantissa]            BFP16A / BFP12A (kept apart from K_BFP: it needs a copy of the inputs)
//   K_BFP_STOCH  [one stochastic stage: BFP / FLOAT / FixedPoint, FLAT]  the reference's default rounding; random words loaded with the
//                data or computed in registers (Philox)
//   K_NM24_BFP   [2:4 with score |x| -> BFP n.s.] on a bf16 / fp16 tensor: the whole-model weight cast of config #4, straight-line
#pragma once
#include "dmxq_stages.cuh"

namespace dmxq {

enum : int { K_AUX = 0, K_CHAIN = 1, K_BFP = 2, K_FLOAT = 3, K_FLOAT_BFP = 4, K_NM_BFP = 5, K_SBFP = 6, K_FIXED = 7, K_NM = 8, K_MXFP = 9, K_BFP_ASYM = 10, K_BFP_STOCH = 11, K_NM24_BFP = 12, K_COUNT = 13 };

struct RowAddr {
    int64_t xo, yo, so, mo, ro;
};

__device__ __forceinline__ RowAddr row_addr(const RowsParams &p, int64_t row)
{
    RowAddr a;
    if (p.nouter == 1) {
        a.xo = row * p.xs[0]; a.yo = row * p.ys[0]; a.so = row * p.ss[0]; a.mo = row * p.ms[0]; a.ro = row * p.rs[0];
        return a;
    }
    a.xo = a.yo = a.so = a.mo = a.ro = 0;
    if (p.n_vec <= 0x7FFFFFFFll) {  // row < 2^31: multiply-high divisions (the 64-bit forms cost ~100 instructions each)
        uint32_t r32 = (uint32_t)row;
        for (int d = p.nouter - 1; d >= 0; --d) {
            const uint32_t od = (uint32_t)p.odim[d];
            const uint32_t q = (d == 0) ? 0u : p.odim_div[d].div(r32);
            const int64_t i = (d == 0) ? r32 : r32 - q * od;
            r32 = q;
            a.xo += i * p.xs[d]; a.yo += i * p.ys[d]; a.so += i * p.ss[d]; a.mo += i * p.ms[d]; a.ro += i * p.rs[d];
        }
        return a;
    }
    for (int d = p.nouter - 1; d >= 0; --d) {
        int64_t i = (d == 0) ? row : row % p.odim[d];
        if (d != 0) row /= p.odim[d];
        a.xo += i * p.xs[d]; a.yo += i * p.ys[d]; a.so += i * p.ss[d]; a.mo += i * p.ms[d]; a.ro += i * p.rs[d];
    }
    return a;
}

// one symmetric nearest BFP stage on a register vector whose max|x| pattern is already known
template <int V, bool SRC16>
__device__ __forceinline__ void bfp_ns_apply(float (&v)[V], uint32_t m, const StageDev &st)
{
    if (st.fast && bfp_fast_ok(m)) {
        BfpFast b = bfp_fast_block(m, st.wl);
        if (SRC16 && st.fast16) {
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
    } else {
        BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem<R_NEAREST>(v[j], b, st.sh, st.mask, 0u);
    }
}

template <int V> __device__ __forceinline__ void float_fast_apply(float (&v)[V], const StageDev &st, bool any_nan)
{
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
}

// fp32 bit pattern of a non-negative 16-bit magnitude pattern
template <typename T> __device__ __forceinline__ uint32_t widen16(uint32_t m16)
{
    if constexpr (std::is_same<T, __nv_bfloat16>::value) return m16 << 16;
    else return f2u(__half2float(__ushort_as_half((unsigned short)m16)));
}

// Flush threshold (FloatFmt.shift_exp) of an SBFP scaler format whose exponent bias is derived on the device from a
// tensor-wide amax (dmxq_cast_chain_multi), so a calibration all-reduce can feed the cast without a host round trip:
//     bias = (2^E - 1) - floor(log2(amax / man_scaling)), clamped to the range FloatingPoint accepts;  threshold = 2^(1 - bias)
// (dmx_compressor_b200.parallel.sbfp_scaler_bias_from_amax is the same rule on the host; the reference delegates the choice
// to d-Matrix's private `numerics` module, S/numerical/format.py:13-20, 438-446 -- parity unpinned by construction).
__device__ __forceinline__ uint32_t sbfp_thr_from_amax(const SbfpFmt &f, const float *amax, int sc_exp_bits)
{
    const float a = __ldg(amax);
    int bias = (1 << (sc_exp_bits - 1)) - 1;  // the format's default when amax is unusable (zero, Inf, NaN)
    if (a > 0.0f && a < __int_as_float(0x7F800000)) {
        // top = largest t with man_scaling * 2^t <= amax (exact: no log, and the quotient's rounding is corrected)
        int t = (int)((f2u(__fdiv_rn(a, f.man_scaling)) >> 23) & 0xFFu) - 127;
        if (__fmul_rn(f.man_scaling, u2f((uint32_t)max(min(t + 127, 254), 1) << 23)) > a) --t;
        else if (t < 127 && __fmul_rn(f.man_scaling, u2f((uint32_t)max(min(t + 128, 254), 1) << 23)) <= a) ++t;
        bias = ((1 << sc_exp_bits) - 1) - t;
        const int lo = sc_exp_bits == 8 ? 127 : -128 + (1 << sc_exp_bits);
        bias = max(lo, min(127, bias));
    }
    return (uint32_t)(127 - (bias - 1)) << 23;
}

// K_FLOAT_BFP on one vector: [FLOAT nearest + flush + signed] -> rounding to Tout -> [BFP nearest symmetric], the output cast of
// one module fused with the input cast of the next (also the tail of dmxq_softmax_cast).  Per-thread constants in FloatBfpCtx.
struct FloatBfpCtx {
    uint32_t fmax_rq;        // the float format's saturation value as Tout stores it
    bool f16_same;           // 16-bit tensor whose significand the format keeps: vectors inside [f16_lo, f16_hi] pass unchanged
    uint32_t f16_lo, f16_hi;
};
template <typename Tin, typename Tout> __device__ __forceinline__ FloatBfpCtx float_bfp_ctx(const StageDev &sf)
{
    constexpr bool SAME16 = sizeof(Tin) == 2 && std::is_same<Tin, Tout>::value;
    FloatBfpCtx c;
    c.fmax_rq = SAME16 ? f2u(requant1<Tout>(u2f(sf.ff.max_num))) : 0u;
    c.f16_same = SAME16 && sf.ff.exact && !sf.ff.is_unsigned;
    c.f16_lo = c.f16_same ? pattern16_ru<Tin>(u2f(sf.ff.shift_exp)) : 0u;
    c.f16_hi = c.f16_same ? min(pattern16_rn<Tin>(u2f(sf.ff.max_num)), (uint32_t)(std::is_same<Tin, __half>::value ? 0x7BFFu : 0x7F7Fu)) : 0u;
    return c;
}
template <typename Tin, typename Tout, int V>
__device__ __forceinline__ void float_bfp_apply(const uint4 &raw, float (&v)[V], const StageDev &sf, const StageDev &st, const FloatBfpCtx &c)
{
    constexpr bool SAME16 = sizeof(Tin) == 2 && std::is_same<Tin, Tout>::value;
    const uint32_t fmax_rq = c.fmax_rq, f16_lo = c.f16_lo, f16_hi = c.f16_hi;
    const bool f16_same = c.f16_same;
    {
            // The float stage (nearest, flush, saturate) and the rounding to Tout are monotone in |x|, so the block
            // maximum the BFP stage needs is the stage applied to the maximum of the inputs: one scalar evaluation
            // instead of a second max over the vector.
            const uint32_t m_in = unpack_absmax<Tin>(raw, v);
            uint32_t m_thr;
            if (m_in > 0x7F800000u) {  // a NaN in this vector: literal path
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = requant1<Tout>(float_elem_slow(v[j], &sf.ff, 0u));
                m_thr = vec_absmax<V>(v);
            } else if (SAME16 && f16_same && raw16_absmin(raw) >= f16_lo && raw16_absmax(raw) <= f16_hi) {
                m_thr = m_in;  // every magnitude inside [flush threshold, saturation value]: the float stage is the identity
            } else if (SAME16 && f16_same) {
                // the values already sit on Tout's grid and the format keeps their significand: only flush and saturate act --
                // on the packed words, two elements per instruction (no NaN here: m_in was checked above)
                if constexpr (SAME16) {
                    const Range16 rg = range16<Tin>(1, sf.ff);
                    const uint4 w = flush_sat16_vec<Tin>(raw, sf.ff, rg);
                    VecIO<Tin>::unpack(w, v);
                    m_thr = widen16<Tin>(raw16_absmax(w));
                } else {
                    m_thr = m_in;
                }
            } else if (SAME16 && sf.ff.exact) {
                // (unsigned variants of such formats: element by element)
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = float_elem_flush_nearest<true>(v[j], sf.ff, fmax_rq);
                m_thr = f2u(float_elem_flush_nearest<true>(u2f(m_in), sf.ff, fmax_rq));
            } else if (sf.ff.exact) {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = requant1<Tout>(float_elem_flush_nearest<true>(v[j], sf.ff));
                m_thr = f2u(requant1<Tout>(float_elem_flush_nearest<true>(u2f(m_in), sf.ff)));
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = requant1<Tout>(float_elem_flush_nearest<false>(v[j], sf.ff));
                m_thr = f2u(requant1<Tout>(float_elem_flush_nearest<false>(u2f(m_in), sf.ff)));
            }
            bfp_ns_apply<V, (sizeof(Tout) == 2)>(v, lanes_max(m_thr, st.block / V), st);  // the values now carry Tout's significand
    }
}

// body of chain_rows_kernel: `cta` is the CTA's index inside the tensor (x, y) of n_vec vectors -- the whole grid for
// the single-tensor kernel, a segment of it for the many-tensor kernel
template <typename Tin, typename Tout, bool FLAT, int KIND, bool AMAX = false>
__device__ __forceinline__ void chain_rows_body(const RowsParams &p, const Tin *__restrict__ x, Tout *__restrict__ y, const int64_t cta,
                                                const int64_t n_vec, const float *amax = nullptr)
{
    constexpr int V = VecIO<Tin>::V;
    constexpr bool SRC16 = sizeof(Tin) == 2;
    constexpr int SRCBITS = std::is_same<Tin, __nv_bfloat16>::value ? 16 : (std::is_same<Tin, __half>::value ? 11 : 32);
    constexpr bool SAME16 = SRC16 && std::is_same<Tin, Tout>::value;
    const int lane = threadIdx.x & 31;
    const int64_t g0 = cta * (kThreads * kUnroll) + threadIdx.x;
    // FLOAT -> BFP pair on a 16-bit tensor: the float format's saturation value as Tout stores it
    const uint32_t fmax_rq = (KIND == K_FLOAT_BFP && SAME16) ? f2u(requant1<Tout>(u2f(p.chain.st[0].ff.max_num))) : 0u;
    // FLOAT stage on a 16-bit tensor whose significand the format keeps: a vector whose magnitudes all lie between the
    // flush threshold and the saturation value passes through the stage unchanged (16-bit pattern bounds, inclusive)
    const bool f16_same = (KIND == K_FLOAT_BFP || KIND == K_FLOAT) && SAME16 && p.chain.st[0].ff.exact && !p.chain.st[0].ff.is_unsigned;
    const uint32_t f16_lo = f16_same ? pattern16_ru<Tin>(u2f(p.chain.st[0].ff.shift_exp)) : 0u;
    const uint32_t f16_hi = f16_same ? min(pattern16_rn<Tin>(u2f(p.chain.st[0].ff.max_num)), (uint32_t)(std::is_same<Tin, __half>::value ? 0x7BFFu : 0x7F7Fu)) : 0u;

    // K_FIXED: the per-tensor parameters (device-resident after calibration, or the stage's immediates), everything derived from
    // them and the choice of variant, ONCE per thread -- the parameter loads go out ahead of the data loads.  (Left inside the
    // per-vector code they were re-loaded and re-derived for every vector: ~80 of 174 instructions per vector.)
    float fx_sc = 1.0f, fx_zp = 0.0f, fx_rsc = 0.0f, fx_rsl = 0.0f, fx_lo = 0.0f, fx_hi = 0.0f;
    int fx_mode = 0;  // 0 general, 1 unit scale + clamp, 2 calibrated + clamp (reciprocal usable), 3 calibrated + clamp (divide)
    if (KIND == K_FIXED) {
        const StageDev &st = p.chain.st[0];
        fx_sc = p.qscale ? __ldg(p.qscale) : st.sc;
        fx_zp = p.qzp ? __ldg(p.qzp) : st.zp;
        const bool wrap = st.affine || p.qscale != nullptr, scaled = st.xf.up != 1.0f;
        const bool small = st.xf.t_max <= 0x1p21f && st.xf.t_min >= -0x1p21f;  // t +- 0.25 exact
        fx_lo = st.xf.t_min - 0.25f; fx_hi = st.xf.t_max + 0.25f;
        if (!wrap && !scaled && st.xf.clamp && small) fx_mode = 1;
        else if (wrap && fx_sc != 1.0f && !scaled && st.xf.clamp && small) {
            fx_rsc = __frcp_rn(fx_sc);
            fx_rsl = recip_lo(fx_sc, fx_rsc);
            fx_mode = (recip_safe(fx_sc) && fabsf(fx_zp) < 0x1p60f) ? 2 : 3;
        }
    }

    // K_FLOAT (subnormals kept) / K_MXFP on a 16-bit tensor: the packed form (sub16_pair) -- per-format constants
    Sub16Fmt s16f{};
    Sub16 s16{};
    if constexpr (SAME16 && (KIND == K_FLOAT || KIND == K_MXFP)) {
        s16f = sub16_fmt<Tin>(p.chain.st[0].ff);
        if (KIND == K_FLOAT) s16 = sub16_consts<Tin>(s16f, 0);
    }

    uint4 raw[kUnroll];
    uint4 rraw[kUnroll][V / 4];  // K_BFP_STOCH only
    int64_t yoff[kUnroll];
    int64_t aux_s[kUnroll], aux_m[kUnroll], aux_r[kUnroll];
    bool valid[kUnroll];

    // SBFP on a flat 16-bit tensor whose block is two vectors (block 16): a thread takes both vectors of a block (still
    // whole 32-byte sectors per lane), so the block constants -- max / 7, its scaler cast, its reciprocal -- are derived once
    // per block instead of once per lane, and no shuffle is needed
    const bool pair = KIND == K_SBFP && SRC16 && FLAT && p.chain.st[0].block == 2 * V;
    // K_SBFP: the scaler format's flush threshold -- the stage's own, or derived from a device-resident amax (AMAX)
    const SbfpFmt &sbf = p.chain.st[0].sb;
    uint32_t sc_thr = 0u;
    if (KIND == K_SBFP) sc_thr = AMAX ? sbfp_thr_from_amax(sbf, amax, p.chain.st[0].sb_exp_bits) : sbf.sc.shift_exp;

    // ---- phase 1: addresses + all loads (nothing here consumes loaded data)
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        int64_t g = g0 + (int64_t)u * kThreads;
        if (pair) g = cta * (kThreads * kUnroll) + (u >> 1) * (2 * kThreads) + 2 * threadIdx.x + (u & 1);
        int64_t xoff;
        if (FLAT) {
            valid[u] = g < n_vec;
            xoff = g * V;
            yoff[u] = xoff;
            if (KIND == K_AUX) { aux_s[u] = xoff; aux_m[u] = xoff; aux_r[u] = xoff; }
        } else {
            int64_t row;
            uint32_t kv;
            if (p.n_vec <= 0x7FFFFFFFll) {
                uint32_t g32 = (uint32_t)g;
                uint32_t r32 = p.vpr_div.div(g32);
                kv = g32 - r32 * p.vpr;
                row = r32;
            } else {
                row = g / p.vpr;
                kv = (uint32_t)(g - row * p.vpr);
            }
            valid[u] = g < p.n_vec && kv < p.kvec;
            RowAddr a = row_addr(p, valid[u] ? row : 0);
            int64_t k = (int64_t)kv * V;
            xoff = a.xo + k;
            yoff[u] = a.yo + k;
            if (KIND == K_AUX) { aux_s[u] = a.so + k; aux_m[u] = a.mo + k; aux_r[u] = a.ro + k * p.rks; }
        }
        raw[u] = valid[u] ? ldg_stream(x + xoff) : make_uint4(0u, 0u, 0u, 0u);
        if (KIND == K_BFP_STOCH) {  // (FLAT only) one int32 random word per element, same offsets as the data
            if (p.philox) {  // ... or computed: four consecutive elements share one Philox call (xoff is a multiple of V)
#pragma unroll
                for (int j = 0; j < V / 4; ++j) rraw[u][j] = philox4x32_10((uint64_t)(xoff >> 2) + j, p.ph_stream, p.ph_seed);
            } else {
#pragma unroll
                for (int j = 0; j < V / 4; ++j)
                    rraw[u][j] = valid[u] ? ldg_stream(static_cast<const uint32_t *>(p.rnd) + xoff + 4 * j) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }

    // ---- phase 2: per vector: widen, stages, narrow, store
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        float v[V];
        if (KIND == K_BFP) {
            const StageDev &st = p.chain.st[0];
            uint32_t m = lanes_max(unpack_absmax<Tin>(raw[u], v), st.block / V);
            bfp_ns_apply<V, SRC16>(v, m, st);
        } else if (KIND == K_BFP_STOCH) {
            // ONE stochastic stage (the L1 default rounding, Q/quant_function.py:47,87,120) with its random words loaded next to the
            // data or computed in registers: BFP, FLOAT, or FixedPoint without affine parameters (its words are fp32 values in [0, 1))
            const StageDev &st = p.chain.st[0];
            if (st.kind == ST_BFP) {
                const uint32_t m = lanes_max(unpack_absmax<Tin>(raw[u], v), st.block / V);
                const BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
                for (int j = 0; j < V; j += 4) {
                    const uint4 q = rraw[u][j / 4];
                    v[j] = bfp_elem<R_STOCHASTIC>(v[j], b, st.sh, st.mask, q.x);
                    v[j + 1] = bfp_elem<R_STOCHASTIC>(v[j + 1], b, st.sh, st.mask, q.y);
                    v[j + 2] = bfp_elem<R_STOCHASTIC>(v[j + 2], b, st.sh, st.mask, q.z);
                    v[j + 3] = bfp_elem<R_STOCHASTIC>(v[j + 3], b, st.sh, st.mask, q.w);
                }
            } else {
                VecIO<Tin>::unpack(raw[u], v);
                const bool fixed = st.kind == ST_FIXED;
#pragma unroll
                for (int j = 0; j < V; j += 4) {
                    const uint4 q = rraw[u][j / 4];
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (fixed) v[j + k] = fixed_elem(v[j + k], st.xf, u2f(p.philox ? philox_unit_bits(w[k]) : w[k]));
                        else v[j + k] = float_elem<R_STOCHASTIC>(v[j + k], st.ff, w[k]);
                    }
                }
            }
        } else if (KIND == K_BFP_ASYM) {
            // symmetric nearest result first, then make_mantissa_asymmetric (S/numerical/format.py:349-372): the edge
            // mantissa -(2^(wl-1)-1) needs |x| >= maxval - Q/2, so only blocks whose max is within one quantum of
            // maxval (a safe superset) -- or blocks off the fast path -- can hold it
            const StageDev &st = p.chain.st[0];
            uint32_t m = lanes_max(unpack_absmax<Tin>(raw[u], v), st.block / V);
            float x0[V];
#pragma unroll
            for (int j = 0; j < V; ++j) x0[j] = v[j];
            bfp_ns_apply<V, SRC16>(v, m, st);
            const BfpBlock bb = bfp_block(m, st.wl);
            if (!(st.fast && bfp_fast_ok(m)) || m >= bb.maxnum - (1u << (23 - (st.wl - 2)))) {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = bfp_asym_fix(v[j], x0[j], bb);
            }
        } else if (KIND == K_FLOAT && p.chain.st[0].ff.nsub) {
            // nearest with subnormals kept (FP8 E4M3 / E5M2 ...): the branch-free form; Inf / NaN vectors take the literal path
            const StageDev &st = p.chain.st[0];
            if constexpr (SAME16) {
                if (s16.ok && raw16_absmax(raw[u]) < s16.limit) {  // finite, room for the rounding constant: two elements per instruction
                    if (valid[u]) stg_stream(y + yoff[u], sub16_vec<Tin>(raw[u], s16));
                    continue;
                }
            }
            // (finite, and the rounding constant 1.5 * 2^(e + sh) must itself be finite: exponent field below 255 - sh)
            if (unpack_absmax<Tin>(raw[u], v) < ((255u - (uint32_t)st.ff.sh) << 23)) {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = float_elem_nearest_sub(v[j], st.ff);
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = float_elem_slow(v[j], &st.ff, 0u);
            }
        } else if (KIND == K_FLOAT && !p.chain.st[0].ff.fastpath) {
            // every other nearest-rounding format (unsigned scalers, `BFP[24|8]{1}` = the BASIC bias format: 22 mantissa bits, nothing
            // flushed): the same element function the runtime chain calls, without that kernel's ~6 us of fixed cost per launch
            VecIO<Tin>::unpack(raw[u], v);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = float_elem_nearest(v[j], p.chain.st[0].ff);
        } else if (KIND == K_FLOAT) {
            if constexpr (SAME16) {
                if (f16_same && raw16_absmin(raw[u]) >= f16_lo && raw16_absmax(raw[u]) <= f16_hi) {  // identity on this vector
                    if (valid[u]) stg_stream(y + yoff[u], raw[u]);
                    continue;
                }
            }
            const bool any_nan = unpack_absmax<Tin>(raw[u], v) > 0x7F800000u;
            float_fast_apply<V>(v, p.chain.st[0], any_nan);
        } else if (KIND == K_FLOAT_BFP) {
            float_bfp_apply<Tin, Tout, V>(raw[u], v, p.chain.st[0], p.chain.st[1], FloatBfpCtx{fmax_rq, f16_same, f16_lo, f16_hi});
        } else if (KIND == K_NM_BFP) {
            // score = |x|: the largest magnitude of every group survives the pruning, so the block maximum is the
            // maximum of the inputs (packed 16-bit max on the raw words for 16-bit sources)
            const StageDev &sn = p.chain.st[0];
            const StageDev &st = p.chain.st[1];
            const uint32_t m_in = unpack_absmax<Tin>(raw[u], v);
            bool pruned = false;
            if constexpr (SRC16) {
                if (sn.block == 4 && sn.n_prune >= 1 && sn.n_prune <= 3 && m_in < 0x7F800000u) {
                    bool keep[8];
                    nm4_keep_raw16(raw[u], sn.n_prune, keep);
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = keep[j] ? v[j] : __fmul_rn(v[j], 0.0f);  // x * mask (finite x)
                    pruned = true;
                }
            }
            if (!pruned) nm_stage<V>(v, sn, lane, nullptr, nullptr, valid[u]);  // (pairwise ranks measured faster here than the packed-key network)
            const uint32_t m_thr = m_in > 0x7F800000u ? vec_absmax<V>(v) : m_in;  // NaN: whatever the pruning left
            bfp_ns_apply<V, SRC16>(v, lanes_max(m_thr, st.block / V), st);
        } else if (KIND == K_NM24_BFP) {
            // 2:4 (score |x|) -> symmetric nearest BFP on a 16-bit tensor, everything the general K_NM_BFP path decides at run
            // time fixed by the dispatcher (M = 4, two pruned, BFP fast16 form).  One sorting network per group serves both
            // stages: the threshold key selects the survivors AND the largest key is the group's max|x| (it always survives),
            // so no separate abs-max pass exists.  The BFP rounding runs on the unpruned values and the pruned ones are
            // zeroed afterwards: BFP of a pruned +-0 is (+-0 + C) - C = +0 in the reference too.
            if constexpr (SAME16) {
                const StageDev &st = p.chain.st[1];
                constexpr uint32_t kInf16 = std::is_same<Tin, __nv_bfloat16>::value ? 0x7F80u : 0x7C00u;
                const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
                uint32_t k[8], thr[2], gmax[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {  // keys as in nm4_keep_raw16: magnitude pattern << 17 | index
                    k[4 * g + 0] = imad(w[2 * g], 0x20000u, 0u);
                    k[4 * g + 1] = imad(w[2 * g] & 0x7FFF0000u, 2u, 1u);
                    k[4 * g + 2] = imad(w[2 * g + 1], 0x20000u, 2u);
                    k[4 * g + 3] = imad(w[2 * g + 1] & 0x7FFF0000u, 2u, 3u);
                    const uint32_t lo1 = min(k[4 * g], k[4 * g + 1]), hi1 = max(k[4 * g], k[4 * g + 1]);
                    const uint32_t lo2 = min(k[4 * g + 2], k[4 * g + 3]), hi2 = max(k[4 * g + 2], k[4 * g + 3]);
                    thr[g] = __vimax3_u32(min(hi1, hi2), lo1, lo2);  // third smallest key: it and everything above survive
                    gmax[g] = max(hi1, hi2);
                }
                const int lanes = st.block / V;
                const uint32_t mv16 = max(gmax[0], gmax[1]) >> 17;  // 15-bit magnitude pattern of this vector's max|x|
                const uint32_t m16 = lanes == 8 ? lanes_max_n<8>(mv16) : lanes_max(mv16, lanes);  // ... of the block's
                VecIO<Tin>::unpack(raw[u], v);
                if (__all_sync(0xFFFFFFFFu, m16 < kInf16 && bfp_fast_ok(widen16<Tin>(m16)))) {
                    const BfpFast b = bfp_fast_block(widen16<Tin>(m16), st.wl);
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = bfp_fast16_elem(v[j], b);
                    if (b.clamp) {  // (an unconditional clamp for small wl, where most warps hold such a block, measured slower)
#pragma unroll
                        for (int j = 0; j < V; ++j) v[j] = bfp_clamp(v[j], b);
                    }
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = k[j] >= thr[j >> 2] ? v[j] : 0.0f;
                } else {
                    // some block of this warp is denormal / huge / non-finite: the whole warp (the branch is warp-uniform, so the
                    // shuffles below stay convergent) runs K_NM_BFP's general sequence, which is bit-identical for ordinary blocks
                    const StageDev &sn = p.chain.st[0];
                    if (mv16 < kInf16) {
#pragma unroll
                        for (int j = 0; j < V; ++j) v[j] = k[j] >= thr[j >> 2] ? v[j] : __fmul_rn(v[j], 0.0f);  // x * mask (finite x)
                    } else {
                        nm_stage<V>(v, sn, lane, nullptr, nullptr, valid[u]);
                    }
                    const uint32_t m_thr = mv16 > kInf16 ? vec_absmax<V>(v) : widen16<Tin>(mv16);  // NaN: whatever the pruning left
                    bfp_ns_apply<V, true>(v, lanes_max(m_thr, lanes), st);
                }
            }
        } else if (KIND == K_NM) {
            const StageDev &sn = p.chain.st[0];
            bool pruned = false;
            if constexpr (SRC16) {
                constexpr uint32_t kInf16 = std::is_same<Tin, __nv_bfloat16>::value ? 0x7F80u : 0x7C00u;
                if (sn.block == 4 && sn.n_prune >= 1 && sn.n_prune <= 3 && raw16_absmax(raw[u]) < kInf16) {
                    bool keep[8];
                    nm4_keep_raw16(raw[u], sn.n_prune, keep);
                    if constexpr (SAME16) {  // no widening at all: mask the packed words and store them
                        if (valid[u]) stg_stream(y + yoff[u], nm_apply_raw16(raw[u], keep));
                        continue;
                    } else {
                        VecIO<Tin>::unpack(raw[u], v);
#pragma unroll
                        for (int j = 0; j < V; ++j) v[j] = keep[j] ? v[j] : __fmul_rn(v[j], 0.0f);
                        pruned = true;
                    }
                }
            }
            if (!pruned) {
                VecIO<Tin>::unpack(raw[u], v);
                nm_stage<V, SRCBITS>(v, sn, lane, nullptr, nullptr, valid[u]);
            }
        } else if (KIND == K_FIXED) {
            // CastTo.forward for FixedPoint (S/numerical/cast.py:279-296): x/sc + zp -> round -> clamp -> (q - zp)*sc,
            // every step a separately rounded fp32 op.  sc == 1 makes the division and the final multiply exact
            // identities, zp == 0 the final subtraction; the leading "+ zp" is kept (it turns -0 into +0).
            // The per-tensor parameters and the variant (fx_mode) were fixed once per thread above the load phase.
            const StageDev &st = p.chain.st[0];
            const float t_min = st.xf.t_min, t_max = st.xf.t_max;
            const uint32_t m_in = unpack_absmax<Tin>(raw[u], v);  // (packed on 16-bit sources)
            if (fx_mode == 1 && m_in <= 0x7F800000u) {  // INT8 / INT4 with unit scale, no NaN in this vector: clamp first, then round
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = round_away_clamped(v[j], fx_lo, fx_hi);
            } else if (fx_mode == 1) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float a = roundf(v[j]);
                    v[j] = a > t_max ? t_max : (a < t_min ? t_min : a);
                }
            } else if (fx_mode == 2 && m_in < 0x5D800000u) {  // calibrated INT8 / INT4, data well inside the normal range
                // x / sc correctly rounded without the division sequence (div_by_recip16 / div_by_recip2; a -0 quotient comes out
                // as +0, which the "+ zp" produces anyway); finite data and parameters: no NaN can reach the clamp
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float q = SRC16 ? div_by_recip16(v[j], fx_rsc, fx_rsl) : div_by_recip2(v[j], fx_sc, fx_rsc, fx_rsl);
                    v[j] = __fmul_rn(__fsub_rn(round_away_clamped(__fadd_rn(q, fx_zp), fx_lo, fx_hi), fx_zp), fx_sc);
                }
            } else if (fx_mode == 2 || fx_mode == 3) {  // calibrated, literal operation sequence
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float a = roundf(__fadd_rn(__fdiv_rn(v[j], fx_sc), fx_zp));
                    a = a > t_max ? t_max : (a < t_min ? t_min : a);
                    v[j] = __fmul_rn(__fsub_rn(a, fx_zp), fx_sc);
                }
            } else {
                const bool wrap = st.affine || p.qscale != nullptr, unit = fx_sc == 1.0f, scaled = st.xf.up != 1.0f;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float a = v[j];
                    if (wrap) a = __fadd_rn(unit ? a : __fdiv_rn(a, fx_sc), fx_zp);
                    if (scaled) a = __fmul_rn(a, st.xf.up);
                    a = roundf(a);
                    if (scaled) a = __fmul_rn(a, st.xf.down);
                    if (st.xf.clamp) a = a > t_max ? t_max : (a < t_min ? t_min : a);
                    if (wrap && !(unit && fx_zp == 0.0f)) a = __fmul_rn(__fsub_rn(a, fx_zp), fx_sc);
                    v[j] = a;
                }
            }
        } else if (KIND == K_MXFP) {
            const StageDev &st = p.chain.st[0];
            if constexpr (SAME16) {
                // the block maximum as a 16-bit pattern; an ordinary block (mxfp_apply's conditions: normal maximum that is not
                // a power of two, normal scale) folds its power-of-two scale into the packed rounding constants
                constexpr int P = Arith16<Tin>::kMan;
                constexpr int BIAS = std::is_same<Tin, __nv_bfloat16>::value ? 127 : 15;
                const uint32_t m16 = lanes_max(raw16_absmax(raw[u]), st.block / V);
                const int ef = (int)(m16 >> P) - BIAS + 127;  // fp32 exponent field of the maximum
                const int se = ef - (int)((f2u(st.mx_largest) >> 23) - 127u);  // ... of the scale
                if (s16f.ok && (m16 & ((1u << P) - 1u)) != 0u && (m16 >> P) >= 1u && m16 < s16f.limit && se >= 2 && se <= 252) {
                    const Sub16 c = sub16_consts<Tin>(s16f, se - 127);
                    if (c.ok) {
                        if (valid[u]) stg_stream(y + yoff[u], sub16_vec<Tin>(raw[u], c));
                        continue;
                    }
                }
                VecIO<Tin>::unpack(raw[u], v);
                mxfp_apply<V>(v, widen16<Tin>(m16), st);
            } else {
                mxfp_apply<V>(v, lanes_max(unpack_absmax<Tin>(raw[u], v), st.block / V), st);
            }
        } else if (KIND == K_SBFP && pair) {
            if ((u & 1) == 0) {
                float w[V];
                const uint32_t m = max(unpack_absmax<Tin>(raw[u], v), unpack_absmax<Tin>(raw[(u + 1) % kUnroll], w));
                const SbfpBlock b = sbfp_block_ol(m, sbf, sc_thr);
                sbfp_apply<V, SRC16>(v, b, sbf);
                sbfp_apply<V, SRC16>(w, b, sbf);
                if (valid[u]) VecIO<Tout>::template store<V>(y + yoff[u], v);
                if (valid[(u + 1) % kUnroll]) VecIO<Tout>::template store<V>(y + yoff[(u + 1) % kUnroll], w);
            }
            continue;
        } else if (KIND == K_SBFP) {
            const StageDev &st = p.chain.st[0];
            uint32_t m = lanes_max(unpack_absmax<Tin>(raw[u], v), st.block / V);
            SbfpBlock b = sbfp_block_ol(m, sbf, sc_thr);
            sbfp_apply<V, SRC16>(v, b, sbf);
        } else {
            VecIO<Tin>::unpack(raw[u], v);
#pragma unroll 1
            for (int s = 0; s < p.chain.n; ++s) {
                const StageDev &st = p.chain.st[s];
                const int lanes = st.block / V;
                uint32_t r[V];
#pragma unroll
                for (int j = 0; j < V; ++j) r[j] = 0x3F000000u;  // 0.5f: deterministic FIXED
                if (KIND == K_AUX && p.philox && valid[u]) {
                    const int mode = st.kind == ST_FLOAT ? st.ff.mode : st.kind == ST_FIXED ? st.xf.mode : st.kind == ST_BFP ? st.mode : 0;
                    if (mode == R_STOCHASTIC) {  // the words the external tensor would hold at the same logical indices
                        const uint64_t i0 = (uint64_t)aux_r[u];
                        if ((FLAT || p.rks == 1) && (i0 & 3u) == 0u) {
#pragma unroll
                            for (int j = 0; j < V; j += 4) {
                                const uint4 q = philox4x32_10((i0 >> 2) + (j >> 2), p.ph_stream, p.ph_seed);
                                r[j] = q.x; r[j + 1] = q.y; r[j + 2] = q.z; r[j + 3] = q.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < V; ++j) r[j] = philox_word(i0 + (uint64_t)(FLAT ? (int64_t)j : j * p.rks), p.ph_stream, p.ph_seed);
                        }
                        if (st.kind == ST_FIXED) {
#pragma unroll
                            for (int j = 0; j < V; ++j) r[j] = philox_unit_bits(r[j]);
                        }
                    }
                } else if (KIND == K_AUX && p.rnd != nullptr && valid[u]) {
                    const int mode = st.kind == ST_FLOAT ? st.ff.mode : st.kind == ST_FIXED ? st.xf.mode : st.kind == ST_BFP ? st.mode : 0;
                    if (mode == R_STOCHASTIC) {
                        const uint32_t *rp = static_cast<const uint32_t *>(p.rnd) + aux_r[u];
                        if ((FLAT || p.rks == 1) && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {  // contiguous: 16-byte streaming loads
#pragma unroll
                            for (int j = 0; j < V; j += 4) {
                                uint4 q = ldg_stream(rp + j);
                                r[j] = q.x; r[j + 1] = q.y; r[j + 2] = q.z; r[j + 3] = q.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < V; ++j) r[j] = __ldg(rp + (FLAT ? (int64_t)j : j * p.rks));
                        }
                    }
                }
                switch (st.kind) {
                case ST_NM:
                    if (s == 0)  // values still carry the source dtype's significand
                        nm_stage<V, SRCBITS>(v, st, lane, (KIND == K_AUX && p.score) ? p.score + aux_s[u] : nullptr,
                                             (KIND == K_AUX && p.mask) ? p.mask + aux_m[u] : nullptr, valid[u]);
                    else
                        nm_stage<V>(v, st, lane, (KIND == K_AUX && p.score) ? p.score + aux_s[u] : nullptr,
                                    (KIND == K_AUX && p.mask) ? p.mask + aux_m[u] : nullptr, valid[u]);
                    break;
                case ST_BFP: bfp_stage<V>(v, st, lanes, r); break;
                case ST_SBFP: sbfp_stage<V>(v, st, lanes); break;
                case ST_FLOAT: float_stage<V>(v, st, r); break;
                case ST_FIXED: fixed_stage<V>(v, st, r); break;
                case ST_MXFP: mxfp_stage<V>(v, st, lanes); break;
                case ST_SCALE:
                    // SmoothQuant's scale application: x / scale[k] or x * scale[k] along the blocked dim, in fp32 (torch promotes
                    // `tensor / fp32_vector` to fp32), one 16-byte load of the (L1-resident) vector per four elements
                    if (valid[u]) {
                        const int64_t g = g0 + (int64_t)u * kThreads;
                        const uint32_t kv = p.n_vec <= 0x7FFFFFFFll ? (uint32_t)g - p.vpr_div.div((uint32_t)g) * p.vpr : (uint32_t)(g % p.vpr);
                        const float *sp = st.vec + (size_t)kv * V;
#pragma unroll
                        for (int j = 0; j < V; j += 4) {
                            const float4 sc = __ldg(reinterpret_cast<const float4 *>(sp + j));
                            if (st.vec_op) {
                                v[j] = __fmul_rn(v[j], sc.x); v[j + 1] = __fmul_rn(v[j + 1], sc.y); v[j + 2] = __fmul_rn(v[j + 2], sc.z); v[j + 3] = __fmul_rn(v[j + 3], sc.w);
                            } else {
                                v[j] = __fdiv_rn(v[j], sc.x); v[j + 1] = __fdiv_rn(v[j + 1], sc.y); v[j + 2] = __fdiv_rn(v[j + 2], sc.z); v[j + 3] = __fdiv_rn(v[j + 3], sc.w);
                            }
                        }
                    }
                    break;
                default: break;
                }
                if (st.requant) {
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = requant1<Tout>(v[j]);
                }
            }
        }
        if (valid[u]) VecIO<Tout>::template store<V>(y + yoff[u], v);
    }
}

template <typename Tin, typename Tout, bool FLAT, int KIND>
__global__ void __launch_bounds__(kThreads) chain_rows_kernel(const __grid_constant__ RowsParams p)
{
    chain_rows_body<Tin, Tout, FLAT, KIND>(p, static_cast<const Tin *>(p.x), static_cast<Tout *>(p.y), (int64_t)blockIdx.x, p.n_vec);
}

// Many tensors, one launch (dmxq_cast_chain_multi): the shards a rank owns in a sharded whole-model weight cast.  Every
// tensor is a flat run of vectors; CTAs are dealt to tensors by a prefix table that travels in the kernel parameters (no
// device-side table to allocate or fill), found by a CTA-uniform binary search.
template <typename T, int KIND, bool AMAX>
__global__ void __launch_bounds__(kThreads) chain_rows_multi_kernel(const __grid_constant__ RowsParams p, const __grid_constant__ MultiTable t)
{
    int lo = 0, hi = t.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blockIdx.x >= t.cta0[mid]) lo = mid; else hi = mid;
    }
    chain_rows_body<T, T, true, KIND, AMAX>(p, static_cast<const T *>(t.x[lo]), static_cast<T *>(t.y[lo]), (int64_t)(blockIdx.x - t.cta0[lo]), t.n_vec[lo],
                                            AMAX ? t.amax + t.slot[lo] : nullptr);
}

template <typename T, int KIND> static cudaError_t launch_rows_multi_k(const RowsParams &p, const MultiTable &t, cudaStream_t s)
{
    const unsigned grid = t.cta0[t.n];
    if (grid == 0) return cudaSuccess;
    if constexpr (KIND == K_SBFP) {
        if (t.amax != nullptr) chain_rows_multi_kernel<T, KIND, true><<<grid, kThreads, 0, s>>>(p, t);
        else chain_rows_multi_kernel<T, KIND, false><<<grid, kThreads, 0, s>>>(p, t);
    } else {
        chain_rows_multi_kernel<T, KIND, false><<<grid, kThreads, 0, s>>>(p, t);
    }
    count_launch();
    return cudaGetLastError();
}

template <typename Tin, typename Tout, int KIND>
static cudaError_t launch_rows_k(bool flat, const RowsParams &p, cudaStream_t s)
{
    int64_t per_cta = (int64_t)kThreads * kUnroll;
    int64_t grid = (p.n_vec + per_cta - 1) / per_cta;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    dim3 g((unsigned)grid), b(kThreads);
    if (flat) chain_rows_kernel<Tin, Tout, true, KIND><<<g, b, 0, s>>>(p);
    else chain_rows_kernel<Tin, Tout, false, KIND><<<g, b, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

// one translation unit per KIND group instantiates these (compile-time parallelism)
template <int KIND> cudaError_t launch_rows_kind(int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    if (in_dt == 0 && out_dt == 0) return launch_rows_k<float, float, KIND>(flat, p, s);
    if (in_dt == 1 && out_dt == 1) return launch_rows_k<__nv_bfloat16, __nv_bfloat16, KIND>(flat, p, s);
    if (in_dt == 2 && out_dt == 2) return launch_rows_k<__half, __half, KIND>(flat, p, s);
    if (in_dt == 1 && out_dt == 0) return launch_rows_k<__nv_bfloat16, float, KIND>(flat, p, s);
    if (in_dt == 2 && out_dt == 0) return launch_rows_k<__half, float, KIND>(flat, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace dmxq

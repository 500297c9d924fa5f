// dmxq_misc.cu -- the correctness-net and support kernels:
//   chain_generic_kernel  any strides / odd sizes (scalar, two-pass)
//   fixed_chan_kernel     FixedPoint with per-channel / group affine parameters
//   blockq_kernel         L1 block_quantize(x, wl, dim) apply step
//   minmax_kernel         calibration amin/amax (exact, order independent)
#include <algorithm>
#include <atomic>

#include "dmxq_stages.cuh"

namespace dmxq {

static std::atomic<int64_t> g_launches{0};
int64_t launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches += n; }

// ------------------------------------------------------------------------------------------------
// generic kernel: any strides, any block size; one thread per (other-dims index, block).
// Two passes over the block (statistic, then apply); the second pass re-reads through L1/L2.
// In-place safe: a thread finishes reading its block statistic before it writes any element,
// and blocks are disjoint.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kThreads) chain_generic_kernel(const __grid_constant__ GenericParams p)
{
    const int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (item >= p.n_items) return;
    const StageDev &st = p.st;
    // fastest: last non-block dim, then the block index, then the remaining dims
    int64_t t = item;
    int64_t xo = 0, yo = 0, so = 0, mo = 0, ro = 0;
    if (p.nd > 0) {
        int d = p.nd - 1;
        int64_t i = t % p.dim[d]; t /= p.dim[d];
        xo += i * p.xs[d]; yo += i * p.ys[d]; so += i * p.ss[d]; mo += i * p.ms[d]; ro += i * p.rs[d];
    }
    const int64_t blk = t % p.nblk; t /= p.nblk;
    for (int d = p.nd - 2; d >= 0; --d) {
        int64_t i = t % p.dim[d]; t /= p.dim[d];
        xo += i * p.xs[d]; yo += i * p.ys[d]; so += i * p.ss[d]; mo += i * p.ms[d]; ro += i * p.rs[d];
    }
    const int64_t B = (st.kind == ST_FLOAT || st.kind == ST_FIXED) ? 1 : st.block;
    const int64_t k0 = blk * B;
    const int64_t k1 = min(k0 + B, p.K);
    const Tin *__restrict__ x = static_cast<const Tin *>(p.x) + xo;
    Tout *y = static_cast<Tout *>(p.y) + yo;
    const uint32_t *rnd = p.rnd ? static_cast<const uint32_t *>(p.rnd) + ro : nullptr;

    if (st.kind == ST_BFP || st.kind == ST_SBFP || st.kind == ST_MXFP) {
        uint32_t m = 0;
        for (int64_t k = k0; k < k1; ++k) m = max(m, f2u(Cvt<Tin>::to_f32(x[k * p.xks])) & 0x7FFFFFFFu);
        if (st.kind == ST_BFP) {
            for (int64_t k = k0; k < k1; ++k) {
                float xv = Cvt<Tin>::to_f32(x[k * p.xks]);
                uint32_t r = (rnd && st.mode == R_STOCHASTIC) ? rnd[k * p.rks] : 0u;
                y[k * p.yks] = Cvt<Tout>::from_f32(bfp_elem_slow(xv, m, st.wl, st.sh, st.mask, st.mode, st.asym, r));
            }
        } else if (st.kind == ST_MXFP) {
            MxBlock b = mx_block(m, st.mx_largest);
            for (int64_t k = k0; k < k1; ++k)
                y[k * p.yks] = Cvt<Tout>::from_f32(mx_elem_ol(Cvt<Tin>::to_f32(x[k * p.xks]), b.scale, &st.ff));
        } else {
            SbfpBlock b = sbfp_block_ol(m, st.sb);
            for (int64_t k = k0; k < k1; ++k)
                y[k * p.yks] = Cvt<Tout>::from_f32(sbfp_elem_slow(Cvt<Tin>::to_f32(x[k * p.xks]), b.cmax, b.fs, &st.sb.xp));
        }
    } else if (st.kind == ST_NM) {
        // rank each element of the group against the others (keys recomputed on the fly);
        // all ranks are computed from the pristine input before the first write (in-place safe
        // because masks are buffered in a bit set: M <= 64).
        unsigned long long keepbits = 0ull;
        if (st.nm_order && k1 - k0 <= 32) {  // torch's CUDA tie order: the 32-slot bitonic network on this group's keys
            uint32_t keys[32];
            const int M = (int)(k1 - k0);
            for (int a = 0; a < M; ++a)
                keys[a] = p.score ? score_key(p.score[so + (k0 + a) * p.sks]) : absx_key(Cvt<Tin>::to_f32(x[(k0 + a) * p.xks]));
            keepbits = ~(unsigned long long)nm_pruned_torch32(keys, M, st.n_prune);
        } else
        for (int64_t a = k0; a < k1; ++a) {
            float xa = Cvt<Tin>::to_f32(x[a * p.xks]);
            uint32_t ka = p.score ? score_key(p.score[so + a * p.sks]) : absx_key(xa);
            int rank = 0;
            for (int64_t bb = k0; bb < k1; ++bb) {
                if (bb == a) continue;
                float xb = Cvt<Tin>::to_f32(x[bb * p.xks]);
                uint32_t kb = p.score ? score_key(p.score[so + bb * p.sks]) : absx_key(xb);
                rank += (kb < ka || (kb == ka && bb < a)) ? 1 : 0;
            }
            if (rank >= st.n_prune) keepbits |= 1ull << (a - k0);
        }
        for (int64_t a = k0; a < k1; ++a) {
            bool keep = (keepbits >> (a - k0)) & 1ull;
            float xa = Cvt<Tin>::to_f32(x[a * p.xks]);
            y[a * p.yks] = Cvt<Tout>::from_f32(nm_apply(xa, keep));
            if (p.mask) p.mask[mo + a * p.mks] = keep ? 1.0f : 0.0f;
        }
    } else if (st.kind == ST_FLOAT) {
        float xv = Cvt<Tin>::to_f32(x[k0 * p.xks]);
        uint32_t r = (rnd && st.ff.mode == R_STOCHASTIC) ? rnd[k0 * p.rks] : 0u;
        y[k0 * p.yks] = Cvt<Tout>::from_f32(float_elem_slow(xv, &st.ff, r));
    } else if (st.kind == ST_FIXED) {
        float xv = Cvt<Tin>::to_f32(x[k0 * p.xks]);
        float r = (rnd && st.xf.mode == R_STOCHASTIC) ? u2f(rnd[k0 * p.rks]) : 0.5f;
        float q = fixed_elem_slow(xv, &st.xf, st.affine, st.sc, st.zp, r);
        y[k0 * p.yks] = Cvt<Tout>::from_f32(q);
    } else {
        y[k0 * p.yks] = Cvt<Tout>::from_f32(Cvt<Tin>::to_f32(x[k0 * p.xks]));
    }
}

template <typename Tin, typename Tout> static cudaError_t launch_generic_t(const GenericParams &p, cudaStream_t s)
{
    int64_t grid = (p.n_items + kThreads - 1) / kThreads;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    chain_generic_kernel<Tin, Tout><<<(unsigned)grid, kThreads, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_generic(int in_dt, int out_dt, const GenericParams &p, cudaStream_t s)
{
    if (in_dt == 0 && out_dt == 0) return launch_generic_t<float, float>(p, s);
    if (in_dt == 1 && out_dt == 1) return launch_generic_t<__nv_bfloat16, __nv_bfloat16>(p, s);
    if (in_dt == 2 && out_dt == 2) return launch_generic_t<__half, __half>(p, s);
    if (in_dt == 1 && out_dt == 0) return launch_generic_t<__nv_bfloat16, float>(p, s);
    if (in_dt == 2 && out_dt == 0) return launch_generic_t<__half, float>(p, s);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// FixedPoint with per-channel / group affine parameters (S/numerical/cast.py:228-237, 279-296)
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kThreads) fixed_chan_kernel(const __grid_constant__ FixedChanParams p)
{
    const Tin *__restrict__ x = static_cast<const Tin *>(p.x);
    Tout *__restrict__ y = static_cast<Tout *>(p.y);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * kThreads) {
        int64_t c = (i / p.inner) % p.C;
        int64_t q = p.nq == 1 ? 0 : min(c / p.group, p.nq - 1);
        float r = p.rnd ? p.rnd[i] : 0.5f;
        y[i] = Cvt<Tout>::from_f32(fixed_elem_affine(Cvt<Tin>::to_f32(x[i]), p.xf, __ldg(p.scale + q), __ldg(p.zp + q), r));
    }
}

// The affine wrap on one value: x / sc + zp -> round half away -> clamp -> (q - zp) * sc, every step a separately
// rounded fp32 operation (S/numerical/cast.py:279-296 around fixed_point_quantize_nearest_cuda).  `rsc`, `rsl` = RN(1 / sc)
// and its low part (div_by_recip2);
// div_free: the exact reciprocal-based quotient may be used (scale and data well inside the normal range).
// HILO: 0 = two-step quotient from rsc alone, 1 = high / low reciprocal pair, 2 = the pair on a 16-bit-significand dividend
template <int HILO = 1>
__device__ __forceinline__ float fixed_affine_away(float x, float sc, float zp, float rsc, float rsl, bool div_free, const FixedFmt &xf, bool scaled)
{
    float a = __fadd_rn(div_free ? (HILO == 2 ? div_by_recip16(x, rsc, rsl) : HILO == 1 ? div_by_recip2(x, sc, rsc, rsl) : div_by_recip(x, sc, rsc))
                                 : __fdiv_rn(x, sc), zp);
    if (div_free && !scaled && xf.clamp && xf.t_max <= 0x1p21f && xf.t_min >= -0x1p21f)  // (t +- 0.25 must be exact)
        return __fmul_rn(__fsub_rn(round_away_clamped(a, xf.t_min - 0.25f, xf.t_max + 0.25f), zp), sc);
    if (scaled) a = __fmul_rn(a, xf.up);
    a = roundf(a);
    if (scaled) a = __fmul_rn(a, xf.down);
    if (xf.clamp) {
        // div_free implies finite inputs and parameters, so no NaN can reach the clamp and min/max instructions do;
        // otherwise the compare form, which lets a NaN through as the reference's does
        if (div_free) a = fminf(fmaxf(a, xf.t_min), xf.t_max);
        else a = a > xf.t_max ? xf.t_max : (a < xf.t_min ? xf.t_min : a);
    }
    return __fmul_rn(__fsub_rn(a, zp), sc);
}

static __device__ __noinline__ float fixed_chan_cold(float x, const FixedFmt *xf, float sc, float zp, float rsc, float rsl, bool div_free, bool fast,
                                                     bool scaled, float rnd)
{
    if (fast) return fixed_affine_away<1>(x, sc, zp, rsc, rsl, div_free, *xf, scaled);
    return fixed_elem_slow(x, xf, 1, sc, zp, rnd);
}

// vectorised variant 1: one 16-byte vector never straddles two qparam groups (inner % V == 0, or the channel runs
// along the contiguous dim in groups that are multiples of V).  256 threads x 4 vectors, all loads before first use.
// PAIR (host-decided): two neighbouring vectors always share their parameters (an even number of vectors per channel /
// group), so a thread takes vector PAIRS -- still whole 32-byte sectors per lane -- and derives the parameter set (index
// arithmetic, two loads, reciprocal pair, range checks: ~40 instructions) once per pair instead of once per vector.
template <typename Tin, typename Tout, bool PAIR>
__global__ void __launch_bounds__(kThreads) fixed_chan_vec_kernel(const __grid_constant__ FixedChanParams p, int64_t cvec, int64_t gvec, FastDiv dcvec,
                                                                  FastDiv dC, FastDiv dgvec)
{
    // qparam index of vector g: ((g / cvec) % C') / gvec with the caller's (cvec, C', gvec) -- see launch_fixed_chan_t
    constexpr int V = VecIO<Tin>::V;
    constexpr int U = 4;
    constexpr int S = PAIR ? 2 : 1;  // vectors per parameter set
    constexpr bool SRC16 = sizeof(Tin) == 2;
    const Tin *__restrict__ x = static_cast<const Tin *>(p.x);
    Tout *__restrict__ y = static_cast<Tout *>(p.y);
    const int64_t nvec = p.n / V;
    const int64_t cta0 = (int64_t)blockIdx.x * (kThreads * U);
    const bool fast = p.xf.mode == R_NEAREST && p.xf.tie == TIE_AWAY, scaled = p.xf.up != 1.0f;
    const bool straight = fast && !scaled && p.xf.clamp && p.xf.t_max <= 0x1p21f && p.xf.t_min >= -0x1p21f;  // INT8 / INT4 with calibrated parameters
    const float lo = p.xf.t_min - 0.25f, hi = p.xf.t_max + 0.25f;
    uint4 raw[U];
    int64_t gi[U];
    float scv[U / S], zpv[U / S];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // all loads first -- the data AND its quantisation parameters (a dependent load per channel
                                   // change in the compute phase left the kernel latency-bound)
        const int64_t g = PAIR ? cta0 + (int64_t)(u >> 1) * (2 * kThreads) + 2 * threadIdx.x + (u & 1) : cta0 + (int64_t)u * kThreads + threadIdx.x;
        gi[u] = g;
        raw[u] = g < nvec ? ldg_stream(x + g * V) : make_uint4(0u, 0u, 0u, 0u);
        if (u % S == 0) {
            int64_t q = 0;
            if (p.nq != 1 && g < nvec) {
                if (nvec <= 0x7FFFFFFFll) {  // multiply-high divisions
                    const uint32_t t = dcvec.div((uint32_t)g);
                    q = dgvec.div(t - dC.div(t) * dC.d);
                } else {
                    q = ((g / cvec) % p.C) / gvec;
                }
                q = min(q, p.nq - 1);
            }
            scv[u / S] = __ldg(p.scale + q);
            zpv[u / S] = __ldg(p.zp + q);
        }
    }
#pragma unroll
    for (int s0 = 0; s0 < U; s0 += S) {
        const float sc = scv[s0 / S], zp = zpv[s0 / S];
        const float rsc = __frcp_rn(sc), rsl = recip_lo(sc, rsc);
        const bool sc_ok = recip_safe(sc) && fabsf(zp) < 0x1p60f;
#pragma unroll
        for (int u = s0; u < s0 + S; ++u) {
            const int64_t g = gi[u];
            if (g >= nvec) continue;
            float v[V];
            const uint32_t m_in = unpack_absmax<Tin>(raw[u], v);
            const bool div_free = sc_ok && m_in < 0x5D800000u;
            if (straight && div_free) {  // the straight-line form
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float q = SRC16 ? div_by_recip16(v[j], rsc, rsl) : div_by_recip2(v[j], sc, rsc, rsl);
                    v[j] = __fmul_rn(__fsub_rn(round_away_clamped(__fadd_rn(q, zp), lo, hi), zp), sc);
                }
            } else {  // every other mode / range: ONE out-of-line copy (inlined per element it cost ~45 predicate-shuffling
                      // instructions per vector on the hot path as well)
#pragma unroll  // (unrolled: a run-time index would move v[] into local memory for the hot path too)
                for (int j = 0; j < V; ++j) v[j] = fixed_chan_cold(v[j], &p.xf, sc, zp, rsc, rsl, div_free, fast, scaled, p.rnd ? __ldg(p.rnd + g * V + j) : 0.5f);
            }
            VecIO<Tout>::template store<V>(y + g * V, v);
        }
    }
}

// vectorised variant 2: the channel runs along the contiguous dim ([R, C] row-major, channel = column) with arbitrary
// group size: a lane owns the V columns of one 16-byte segment, keeps their scale / zero-point / reciprocal in
// registers and walks down the rows, four loads in flight.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kThreads) fixed_chan_cols_kernel(const __grid_constant__ FixedChanParams p, int64_t R)
{
    constexpr int V = VecIO<Tin>::V;
    constexpr int W = kThreads / 32;
    const Tin *__restrict__ x = static_cast<const Tin *>(p.x);
    Tout *__restrict__ y = static_cast<Tout *>(p.y);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * V;
    if (c0 >= p.C) return;
    const bool scaled = p.xf.up != 1.0f;
    // fp32 data: no low reciprocal parts (V more registers cost the kernel a resident CTA, 4.8 -> 4.1 TB/s), the two-step quotient
    // instead; 16-bit data: the pair, because with it the quotient is two operations (div_by_recip16)
    constexpr bool PAIR = sizeof(Tin) == 2;
    float sc[V], zp[V], rsc[V], rsl[PAIR ? V : 1];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int64_t q = p.nq == 1 ? 0 : min((c0 + j) / p.group, p.nq - 1);
        sc[j] = __ldg(p.scale + q);
        zp[j] = __ldg(p.zp + q);
        rsc[j] = __frcp_rn(sc[j]);
        if (PAIR) rsl[j] = recip_lo(sc[j], rsc[j]);
        ok = ok && recip_safe(sc[j]) && fabsf(zp[j]) < 0x1p60f;
    }
    const int64_t step = (int64_t)gridDim.y * W;
    for (int64_t r0 = (int64_t)blockIdx.y * W + warp; r0 < R; r0 += 4 * step) {
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw[u] = r0 + u * step < R ? ldg_stream(x + (r0 + u * step) * p.C + c0) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (r0 + u * step >= R) continue;
            float v[V];
            const uint32_t m_in = unpack_absmax<Tin>(raw[u], v);  // (also fills v: keep it out of the && chain)
            const bool div_free = ok && m_in < 0x5D800000u;
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = fixed_affine_away<(PAIR ? 2 : 0)>(v[j], sc[j], zp[j], rsc[j], PAIR ? rsl[PAIR ? j : 0] : 0.0f, div_free, p.xf, scaled);
            VecIO<Tout>::template store<V>(y + (r0 + u * step) * p.C + c0, v);
        }
    }
}

static FastDiv fd(int64_t d) { return make_fastdiv((uint32_t)std::min<int64_t>(std::max<int64_t>(d, 1), 0x7FFFFFFF)); }

template <typename Tin, typename Tout> static void launch_fixed_chan_t(const FixedChanParams &p, cudaStream_t s)
{
    constexpr int V = VecIO<Tin>::V;
    const bool aligned = p.n % V == 0 && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0 && (reinterpret_cast<uintptr_t>(p.y) % 16) == 0 &&
                         (!p.rnd || (reinterpret_cast<uintptr_t>(p.rnd) % 16) == 0);
    const bool fast = p.xf.mode == R_NEAREST && p.xf.tie == TIE_AWAY;
    if (aligned && p.inner % V == 0) {
        // channel constant inside a vector: qparam = ((g / inner_vec) % C) / group
        FixedChanParams q = p;
        int64_t grid = (p.n / V + kThreads * 4 - 1) / (kThreads * 4);
        const unsigned gr = (unsigned)std::max<int64_t>(grid, 1);
        if ((p.inner / V) % 2 == 0) fixed_chan_vec_kernel<Tin, Tout, true><<<gr, kThreads, 0, s>>>(q, p.inner / V, p.group, fd(p.inner / V), fd(q.C), fd(p.group));
        else fixed_chan_vec_kernel<Tin, Tout, false><<<gr, kThreads, 0, s>>>(q, p.inner / V, p.group, fd(p.inner / V), fd(q.C), fd(p.group));
    } else if (aligned && p.inner == 1 && p.C % V == 0 && p.group % V == 0) {
        // channel along the contiguous dim, groups are whole vectors: qparam = (g % (C / V)) / (group / V)
        FixedChanParams q = p;
        q.C = p.C / V;
        int64_t grid = (p.n / V + kThreads * 4 - 1) / (kThreads * 4);
        const unsigned gr = (unsigned)std::max<int64_t>(grid, 1);
        if ((p.group / V) % 2 == 0 && q.C % 2 == 0) fixed_chan_vec_kernel<Tin, Tout, true><<<gr, kThreads, 0, s>>>(q, 1, p.group / V, fd(1), fd(q.C), fd(p.group / V));
        else fixed_chan_vec_kernel<Tin, Tout, false><<<gr, kThreads, 0, s>>>(q, 1, p.group / V, fd(1), fd(q.C), fd(p.group / V));
    } else if (aligned && fast && !p.rnd && p.inner == 1 && p.C % V == 0) {
        const int64_t R = p.n / p.C, gx = (p.C + 32 * V - 1) / (32 * V);
        int64_t gy = std::max<int64_t>(1, std::min<int64_t>((R + 31) / 32, std::max<int64_t>(1, (148 * 8) / gx)));
        dim3 g((unsigned)gx, (unsigned)std::min<int64_t>(gy, 65535));
        fixed_chan_cols_kernel<Tin, Tout><<<g, kThreads, 0, s>>>(p, R);
    } else {
        int64_t grid = std::min<int64_t>((p.n + kThreads - 1) / kThreads, 148 * 16);
        fixed_chan_kernel<Tin, Tout><<<(unsigned)grid, kThreads, 0, s>>>(p);
    }
}

cudaError_t launch_fixed_chan(int in_dt, int out_dt, const FixedChanParams &p, cudaStream_t s)
{
    if (p.n <= 0) return cudaSuccess;
    if (in_dt == 0 && out_dt == 0) launch_fixed_chan_t<float, float>(p, s);
    else if (in_dt == 1 && out_dt == 1) launch_fixed_chan_t<__nv_bfloat16, __nv_bfloat16>(p, s);
    else if (in_dt == 2 && out_dt == 2) launch_fixed_chan_t<__half, __half>(p, s);
    else if (in_dt == 1 && out_dt == 0) launch_fixed_chan_t<__nv_bfloat16, float>(p, s);
    else if (in_dt == 2 && out_dt == 0) launch_fixed_chan_t<__half, float>(p, s);
    else return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// L1 block_quantize apply (Q/quant_cuda/block_kernel.cu:7-139) given per-slice max|x| bits,
// including the reference's `!symmetric` exponent bump for x == -max with an all-ones top
// mantissa byte (block_kernel.cu:51-57).
__global__ void __launch_bounds__(kThreads) blockq_kernel(const __grid_constant__ BlockQParams p)
{
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * kThreads) {
        int64_t c = p.C == 1 ? 0 : (i / p.inner) % p.C;
        uint32_t mb = __ldg(p.maxbits + c);
        float xv = p.x[i];
        if (!p.symmetric) {
            float mf = u2f(mb);
            if (xv == -mf && ((mb >> 16) << 25) == 0xFE000000u) mb = ((mb >> 23) + 1u) << 23;
        }
        BfpBlock b = bfp_block(mb, p.wl);
        uint32_t r = p.rnd ? (uint32_t)p.rnd[i] : 0u;
        float q;
        switch (p.mode) {
        case R_NEAREST: q = bfp_elem<R_NEAREST>(xv, b, p.sh, p.mask, r); break;
        case R_STOCHASTIC: q = bfp_elem<R_STOCHASTIC>(xv, b, p.sh, p.mask, r); break;
        case R_UP: q = bfp_elem<R_UP>(xv, b, p.sh, p.mask, r); break;
        default: q = bfp_elem<R_DOWN>(xv, b, p.sh, p.mask, r); break;
        }
        p.y[i] = q;
    }
}

// one element of the L1 apply, any mode (the literal integer path of block_kernel.cu)
__device__ __forceinline__ float blockq_elem(float xv, uint32_t mb, const BlockQParams &p, uint32_t r)
{
    if (!p.symmetric) {
        float mf = u2f(mb);
        if (xv == -mf && ((mb >> 16) << 25) == 0xFE000000u) mb = ((mb >> 23) + 1u) << 23;
    }
    BfpBlock b = bfp_block(mb, p.wl);
    switch (p.mode) {
    case R_NEAREST: return bfp_elem<R_NEAREST>(xv, b, p.sh, p.mask, r);
    case R_STOCHASTIC: return bfp_elem<R_STOCHASTIC>(xv, b, p.sh, p.mask, r);
    case R_UP: return bfp_elem<R_UP>(xv, b, p.sh, p.mask, r);
    default: return bfp_elem<R_DOWN>(xv, b, p.sh, p.mask, r);
    }
}

// vectorised apply: 16-byte vectors, four in flight per thread.  PER_ELEM = false: the slice ("channel") is constant
// inside a vector (whole tensor, or inner % 4 == 0); true: the slices run along the contiguous dim (inner == 1, C % 4 == 0).
// Symmetric nearest slices whose max is an ordinary number take the four-add form of the rows kernel.
template <bool PER_ELEM> __global__ void __launch_bounds__(kThreads) blockq_vec_kernel(const __grid_constant__ BlockQParams p)
{
    constexpr int V = 4, U = 4;
    const int64_t nvec = p.n / V;
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    const bool fast_mode = p.mode == R_NEAREST && p.symmetric && p.wl <= 20;
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        raw[u] = g < nvec ? ldg_stream(p.x + g * V) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        if (g >= nvec) continue;
        float v[V];
        VecIO<float>::unpack(raw[u], v);
        uint32_t r[V] = {0u, 0u, 0u, 0u};
        if (p.rnd) {
            const uint4 t = ldg_stream(p.rnd + g * V);
            r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
        }
        if (!PER_ELEM) {
            int64_t c = 0;
            if (p.C != 1) {
                const int64_t iv = p.inner / V;
                c = nvec <= 0xFFFFFFFFll ? (int64_t)(((uint32_t)g / (uint32_t)iv) % (uint32_t)p.C) : (g / iv) % p.C;
            }
            const uint32_t mb = __ldg(p.maxbits + c);
            if (fast_mode && bfp_fast_ok(mb)) {
                const BfpFast b = bfp_fast_block(mb, p.wl);
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = bfp_fast_elem(v[j], b);
                if (b.clamp) {
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = bfp_clamp(v[j], b);
                }
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = blockq_elem(v[j], mb, p, r[j]);
            }
        } else {
            const int64_t c0 = nvec <= 0xFFFFFFFFll ? (int64_t)((uint32_t)g % (uint32_t)(p.C / V)) * V : (g % (p.C / V)) * V;
            const uint4 m4 = *reinterpret_cast<const uint4 *>(p.maxbits + c0);
            const uint32_t mb[V] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (fast_mode && bfp_fast_ok(mb[j])) {
                    const BfpFast b = bfp_fast_block(mb[j], p.wl);
                    const float q = bfp_fast_elem(v[j], b);
                    v[j] = b.clamp ? bfp_clamp(q, b) : q;
                } else {
                    v[j] = blockq_elem(v[j], mb[j], p, r[j]);
                }
            }
        }
        VecIO<float>::store<V>(p.y + g * V, v);
    }
}

cudaError_t launch_blockq(const BlockQParams &p, cudaStream_t s)
{
    if (p.n <= 0) return cudaSuccess;
    const bool aligned = p.n % 4 == 0 && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0 && (reinterpret_cast<uintptr_t>(p.y) % 16) == 0 &&
                         (!p.rnd || (reinterpret_cast<uintptr_t>(p.rnd) % 16) == 0);
    const int64_t vgrid = (p.n / 4 + kThreads * 4 - 1) / (kThreads * 4);
    if (aligned && vgrid <= 0x7FFFFFFFll && (p.C == 1 || p.inner % 4 == 0)) {
        blockq_vec_kernel<false><<<(unsigned)vgrid, kThreads, 0, s>>>(p);
    } else if (aligned && vgrid <= 0x7FFFFFFFll && p.inner == 1 && p.C % 4 == 0 && (reinterpret_cast<uintptr_t>(p.maxbits) % 16) == 0) {
        blockq_vec_kernel<true><<<(unsigned)vgrid, kThreads, 0, s>>>(p);
    } else {
        int64_t grid = std::min<int64_t>((p.n + kThreads - 1) / kThreads, 148 * 16);
        blockq_kernel<<<(unsigned)grid, kThreads, 0, s>>>(p);
    }
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// amin / amax: ordered-int atomics (exact, order independent).  NaN -> INT_MAX on the max side,
// INT_MIN on the min side; the finalize kernel turns those sentinels back into NaN.
__device__ __forceinline__ int ord(float f) { int k = (int)f2u(f); return k < 0 ? k ^ 0x7FFFFFFF : k; }
__device__ __forceinline__ float unord(int k) { return u2f((uint32_t)(k < 0 ? k ^ 0x7FFFFFFF : k)); }

__global__ void minmax_init_kernel(int *omin, int *omax, int64_t C)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) { omin[i] = 0x7F800000; omax[i] = ord(u2f(0xFF800000u)); }
}

// per-tensor fast path: contiguous data, 16-byte loads, 4 vectors in flight per thread.  Running extrema are kept in
// the source's own arithmetic -- FMNMX for fp32 with a sticky NaN flag, NaN-propagating packed HMNMX2 for bf16 / fp16
// (two elements per instruction, no widening) -- and only the per-thread results go through the ordered-int reduction.
template <typename T> struct MinMaxAcc;
template <> struct MinMaxAcc<float> {
    float lo = u2f(0x7F800000u), hi = u2f(0xFF800000u);
    bool nan = false;
    __device__ __forceinline__ void add(const uint4 &r)
    {
        const float v[4] = {u2f(r.x), u2f(r.y), u2f(r.z), u2f(r.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo = fminf(lo, v[j]); hi = fmaxf(hi, v[j]); nan |= v[j] != v[j]; }
    }
    __device__ __forceinline__ void add1(float v) { lo = fminf(lo, v); hi = fmaxf(hi, v); nan |= v != v; }
    __device__ __forceinline__ void finish(int &olo, int &ohi) const
    {
        olo = nan ? (int)0x80000000 : ord(lo);
        ohi = nan ? 0x7FFFFFFF : ord(hi);
    }
};
template <typename H2, typename H> struct MinMaxAcc16 {
    H2 lo, hi;
    __device__ __forceinline__ MinMaxAcc16()
    {
        lo = __float2half2_rn_any(u2f(0x7F800000u));
        hi = __float2half2_rn_any(u2f(0xFF800000u));
    }
    static __device__ __forceinline__ H2 __float2half2_rn_any(float f);
    __device__ __forceinline__ void add(const uint4 &r)
    {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const H2 v = *reinterpret_cast<const H2 *>(&w[j]);
            lo = __hmin2_nan(lo, v);
            hi = __hmax2_nan(hi, v);
        }
    }
    __device__ __forceinline__ void add1(float v)
    {
        const H2 t = __float2half2_rn_any(v);  // exact: v came from a 16-bit element
        lo = __hmin2_nan(lo, t);
        hi = __hmax2_nan(hi, t);
    }
    __device__ __forceinline__ void finish(int &olo, int &ohi) const
    {
        const float l0 = Cvt<H>::to_f32(lo.x), l1 = Cvt<H>::to_f32(lo.y), h0 = Cvt<H>::to_f32(hi.x), h1 = Cvt<H>::to_f32(hi.y);
        const bool nan = l0 != l0 || l1 != l1 || h0 != h0 || h1 != h1;
        olo = nan ? (int)0x80000000 : min(ord(l0), ord(l1));
        ohi = nan ? 0x7FFFFFFF : max(ord(h0), ord(h1));
    }
};
template <> __device__ __forceinline__ __nv_bfloat162 MinMaxAcc16<__nv_bfloat162, __nv_bfloat16>::__float2half2_rn_any(float f) { return __float2bfloat162_rn(f); }
template <> __device__ __forceinline__ __half2 MinMaxAcc16<__half2, __half>::__float2half2_rn_any(float f) { return __float2half2_rn(f); }
template <> struct MinMaxAcc<__nv_bfloat16> : MinMaxAcc16<__nv_bfloat162, __nv_bfloat16> {};
template <> struct MinMaxAcc<__half> : MinMaxAcc16<__half2, __half> {};

template <typename T> __global__ void __launch_bounds__(kThreads) minmax_flat_kernel(const T *__restrict__ x, int64_t n, int *omin, int *omax)
{
    constexpr int V = VecIO<T>::V;
    const int64_t nvec = n / V;
    MinMaxAcc<T> acc;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream(x + (i + u * stride) * V);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc.add(r[u]);
    }
    for (; i < nvec; i += stride) acc.add(ldg_stream(x + i * V));
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * V) acc.add1(Cvt<T>::to_f32(x[nvec * V + threadIdx.x]));
    int lo, hi;
    acc.finish(lo, hi);
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
    }
    __shared__ int slo[kThreads / 32], shi[kThreads / 32];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        atomicMin(omin, lo);
        atomicMax(omax, hi);
    }
}

// per-channel statistics, channel = a contiguous run: the tensor is rows of `inner` contiguous elements, row r belongs
// to channel r % C ((outer, C, inner) addressing; MinMaxObserver per_channel on weights, conv activations).  One warp
// per (row, segment); 16-byte loads, four in flight per lane.
template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_rows_kernel(const T *__restrict__ x, int64_t rows, int64_t C, int64_t inner_vec, int segs,
                                                               int64_t per_seg_vec, int *omin, int *omax)
{
    constexpr int V = VecIO<T>::V;
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (tile >= rows * segs) return;
    const int64_t row = tile / segs, seg = tile - row * segs;
    const T *base = x + row * inner_vec * V;
    const int64_t i1 = min(inner_vec, (seg + 1) * per_seg_vec);
    MinMaxAcc<T> acc;
    int64_t i = seg * per_seg_vec + lane;
    for (; i + 96 < i1; i += 128) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream(base + (i + 32 * u) * V);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc.add(r[u]);
    }
    for (; i < i1; i += 32) acc.add(ldg_stream(base + i * V));
    int lo, hi;
    acc.finish(lo, hi);
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
    }
    if (lane == 0) {
        const int64_t c = row % C;
        atomicMin(omin + c, lo);
        atomicMax(omax + c, hi);
    }
}

// per-channel statistics, channel = position along the contiguous dim: x = [R, C] row-major, channel = column
// (activations observed per feature, SmoothQuant maxabs).  A lane owns the V columns of one 16-byte segment and walks
// down the rows (a warp reads 512 contiguous bytes per row); per-column extrema stay in the source's arithmetic
// (packed for 16-bit), the 8 warps of a CTA combine through shared memory, one pair of atomics per column and CTA.
template <typename T> struct MinMaxColAcc;
template <> struct MinMaxColAcc<float> {
    float lo[4], hi[4];
    unsigned nan = 0u;
    __device__ __forceinline__ MinMaxColAcc()
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo[j] = u2f(0x7F800000u); hi[j] = u2f(0xFF800000u); }
    }
    __device__ __forceinline__ void add(const uint4 &r)
    {
        const float v[4] = {u2f(r.x), u2f(r.y), u2f(r.z), u2f(r.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo[j] = fminf(lo[j], v[j]); hi[j] = fmaxf(hi[j], v[j]); nan |= (v[j] != v[j] ? 1u : 0u) << j; }
    }
    __device__ __forceinline__ void finish(int (&olo)[4], int (&ohi)[4]) const
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool n = (nan >> j) & 1u;
            olo[j] = n ? (int)0x80000000 : ord(lo[j]);
            ohi[j] = n ? 0x7FFFFFFF : ord(hi[j]);
        }
    }
};
template <typename H2, typename H> struct MinMaxColAcc16 {
    H2 lo[4], hi[4];
    __device__ __forceinline__ MinMaxColAcc16()
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo[j] = MinMaxAcc16<H2, H>::__float2half2_rn_any(u2f(0x7F800000u)); hi[j] = MinMaxAcc16<H2, H>::__float2half2_rn_any(u2f(0xFF800000u)); }
    }
    __device__ __forceinline__ void add(const uint4 &r)
    {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const H2 v = *reinterpret_cast<const H2 *>(&w[j]);
            lo[j] = __hmin2_nan(lo[j], v);
            hi[j] = __hmax2_nan(hi[j], v);
        }
    }
    __device__ __forceinline__ void finish(int (&olo)[8], int (&ohi)[8]) const
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float l[2] = {Cvt<H>::to_f32(lo[j].x), Cvt<H>::to_f32(lo[j].y)}, h[2] = {Cvt<H>::to_f32(hi[j].x), Cvt<H>::to_f32(hi[j].y)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool n = l[k] != l[k] || h[k] != h[k];
                olo[2 * j + k] = n ? (int)0x80000000 : ord(l[k]);
                ohi[2 * j + k] = n ? 0x7FFFFFFF : ord(h[k]);
            }
        }
    }
};
template <> struct MinMaxColAcc<__nv_bfloat16> : MinMaxColAcc16<__nv_bfloat162, __nv_bfloat16> {};
template <> struct MinMaxColAcc<__half> : MinMaxColAcc16<__half2, __half> {};

template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_cols_kernel(const T *__restrict__ x, int64_t R, int64_t C, int *omin, int *omax)
{
    constexpr int V = VecIO<T>::V;
    constexpr int W = kThreads / 32;
    __shared__ int slo[W][V][32], shi[W][V][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * V;
    MinMaxColAcc<T> acc;
    if (c0 < C) {
        const int64_t step = (int64_t)gridDim.y * W;
        int64_t r = (int64_t)blockIdx.y * W + warp;
        for (; r + 3 * step < R; r += 4 * step) {
            uint4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = ldg_stream(x + (r + u * step) * C + c0);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc.add(t[u]);
        }
        for (; r < R; r += step) acc.add(ldg_stream(x + r * C + c0));
    }
    int lo[V], hi[V];
    acc.finish(lo, hi);
#pragma unroll
    for (int j = 0; j < V; ++j) { slo[warp][j][lane] = lo[j]; shi[warp][j][lane] = hi[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * V; i += kThreads) {
        const int j = i >> 5, l = i & 31;
        int a = slo[0][j][l], b = shi[0][j][l];
#pragma unroll
        for (int w = 1; w < W; ++w) { a = min(a, slo[w][j][l]); b = max(b, shi[w][j][l]); }
        const int64_t c = ((int64_t)blockIdx.x * 32 + l) * V + j;
        if (c < C) { atomicMin(omin + c, a); atomicMax(omax + c, b); }
    }
}

template <typename T> __global__ void __launch_bounds__(kThreads) minmax_kernel(const __grid_constant__ MinMaxParams p)
{
    // grid: (chunks, C); each CTA reduces a slice of channel c over (outer, inner)
    const T *__restrict__ x = static_cast<const T *>(p.x);
    const int64_t c = blockIdx.y;
    const int64_t per = p.outer * p.inner;
    int lo = 0x7F800000, hi = ord(u2f(0xFF800000u));
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < per; i += (int64_t)gridDim.x * kThreads) {
        int64_t o = i / p.inner, in = i - o * p.inner;
        float v = Cvt<T>::to_f32(x[o * p.xo + c * p.xc + in * p.xi]);
        if (v != v) { lo = (int)0x80000000; hi = 0x7FFFFFFF; }
        else { int k = ord(v); lo = min(lo, k); hi = max(hi, k); }
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
    }
    __shared__ int slo[kThreads / 32], shi[kThreads / 32];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        atomicMin(p.omin + c, lo);
        atomicMax(p.omax + c, hi);
    }
}

__global__ void minmax_final_kernel(int *omin, int *omax, int64_t C)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        int lo = omin[i], hi = omax[i];
        float fl = lo == (int)0x80000000 ? u2f(0x7FC00000u) : unord(lo);
        float fh = hi == 0x7FFFFFFF ? u2f(0x7FC00000u) : unord(hi);
        reinterpret_cast<float *>(omin)[i] = fl;
        reinterpret_cast<float *>(omax)[i] = fh;
    }
}

cudaError_t launch_minmax(const MinMaxParams &p, cudaStream_t s)
{
    int64_t C = p.C;
    if (C <= 0) return cudaSuccess;
    unsigned ib = (unsigned)((C + 255) / 256);
    minmax_init_kernel<<<ib, 256, 0, s>>>(p.omin, p.omax, C);
    int64_t per = p.outer * p.inner;
    if (C == 1 && per > 0 && p.xi == 1 && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) {
        int64_t nvec = per / (p.dtype == 0 ? 4 : 8);
        unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nvec + kThreads * 4 - 1) / (kThreads * 4), 148 * 8));
        if (p.dtype == 0) minmax_flat_kernel<float><<<grid, kThreads, 0, s>>>(static_cast<const float *>(p.x), per, p.omin, p.omax);
        else if (p.dtype == 1) minmax_flat_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(p.x), per, p.omin, p.omax);
        else minmax_flat_kernel<__half><<<grid, kThreads, 0, s>>>(static_cast<const __half *>(p.x), per, p.omin, p.omax);
    } else if (per > 0 && p.inner == 1 && C % (p.dtype == 0 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) {
        // channel = column of the row-major [outer, C] matrix
        const int V = p.dtype == 0 ? 4 : 8;
        const int64_t gx = (C + 32 * V - 1) / (32 * V);
        int64_t gy = std::max<int64_t>(1, std::min<int64_t>((p.outer + 31) / 32, std::max<int64_t>(1, (148 * 8) / gx)));
        dim3 g((unsigned)gx, (unsigned)std::min<int64_t>(gy, 65535));
        if (p.dtype == 0) minmax_cols_kernel<float><<<g, kThreads, 0, s>>>(static_cast<const float *>(p.x), p.outer, C, p.omin, p.omax);
        else if (p.dtype == 1) minmax_cols_kernel<__nv_bfloat16><<<g, kThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(p.x), p.outer, C, p.omin, p.omax);
        else minmax_cols_kernel<__half><<<g, kThreads, 0, s>>>(static_cast<const __half *>(p.x), p.outer, C, p.omin, p.omax);
    } else if (per > 0 && p.inner % (p.dtype == 0 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) {
        // channel = contiguous run of `inner` elements
        const int V = p.dtype == 0 ? 4 : 8;
        const int64_t rows = p.outer * C, inner_vec = p.inner / V;
        const int64_t want = (148 * 32 + rows - 1) / rows;                       // segments per row that fill the GPU
        const int64_t most = std::max<int64_t>(1, inner_vec / 128);              // ... but at least 128 vectors each
        const int segs = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(want, most), 1 << 20));
        const int64_t per_seg = (inner_vec + segs - 1) / segs;
        const int64_t grid = (rows * segs + kThreads / 32 - 1) / (kThreads / 32);
        if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
        if (p.dtype == 0) minmax_rows_kernel<float><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const float *>(p.x), rows, C, inner_vec, segs, per_seg, p.omin, p.omax);
        else if (p.dtype == 1) minmax_rows_kernel<__nv_bfloat16><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(p.x), rows, C, inner_vec, segs, per_seg, p.omin, p.omax);
        else minmax_rows_kernel<__half><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const __half *>(p.x), rows, C, inner_vec, segs, per_seg, p.omin, p.omax);
    } else if (per > 0) {
        if (C > 65535) return cudaErrorInvalidConfiguration;
        int64_t chunks = (per + (int64_t)kThreads * 8 - 1) / ((int64_t)kThreads * 8);
        int64_t cap = std::max<int64_t>(1, (148 * 8) / C);
        chunks = std::max<int64_t>(1, std::min(chunks, cap));
        dim3 g((unsigned)chunks, (unsigned)C);
        if (p.dtype == 0) minmax_kernel<float><<<g, kThreads, 0, s>>>(p);
        else if (p.dtype == 1) minmax_kernel<__nv_bfloat16><<<g, kThreads, 0, s>>>(p);
        else minmax_kernel<__half><<<g, kThreads, 0, s>>>(p);
    }
    minmax_final_kernel<<<ib, 256, 0, s>>>(p.omin, p.omax, C);
    count_launch(3);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// calibration histogram (see dmxq_histc in include/dmxq.h): torch.histc's bin rule evaluated in fp32,
//   bin = (int)((v - lo) * bins / (hi - lo)), bin == bins -> bins - 1, values outside [lo, hi] and NaN dropped,
// counted in per-CTA shared-memory bins (u32) that are flushed to 64-bit global counters once per CTA.
// MM additionally folds the tensor's amin/amax into ordered-int accumulators in the same pass.
template <typename T, bool MM>
__global__ void __launch_bounds__(kThreads) histc_kernel(const T *__restrict__ x, int64_t n, float lo, float hi, int bins,
                                                         unsigned long long *counts, int *omin, int *omax)
{
    extern __shared__ unsigned int sbin[];
    constexpr int V = VecIO<T>::V;
    for (int i = threadIdx.x; i < bins; i += kThreads) sbin[i] = 0u;
    __syncthreads();
    const float nb = (float)bins, w = __fsub_rn(hi, lo);
    // Correctly rounded t / w without a division per element: with rw = RN(1/w), two Newton steps on the quotient
    //   q0 = t*rw;  q1 = q0 + (t - q0*w)*rw;  q2 = q1 + (t - q1*w)*rw      (each residual one FMA)
    // leave q2 == RN(t/w): q1 is within one ulp, its residual is then exact, and the last step rounds correctly
    // (Markstein's theorem; it needs w's significand not to be all ones and everything well inside the normal range,
    // otherwise `recip_ok` is false and every element takes the IEEE division).
    const float rw = __fdiv_rn(1.0f, w);
    const bool recip_ok = w > 0x1p-60f && w < 0x1p60f && (f2u(w) & 0x7FFFFFu) != 0x7FFFFFu;
    float fmn = u2f(0x7F800000u), fmx = u2f(0xFF800000u);
    bool saw_nan = false;
    auto bin_exact = [&](float v) {  // next to a bin edge (or rw unusable): the reference's own expression
        int b = (int)__fdiv_rn(__fmul_rn(__fsub_rn(v, lo), nb), w);
        return min(b, bins - 1);
    };
    auto put = [&](float v) {
        if (v >= lo && v <= hi) {
            int b = bin_exact(v);
            if (b >= 0) atomicAdd(&sbin[b], 1u);
        }
    };
    auto put_vec = [&](const float (&v)[V]) {
        int b[V];
        bool in[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (MM) {
                fmn = fminf(fmn, v[j]);
                fmx = fmaxf(fmx, v[j]);
                saw_nan |= v[j] != v[j];
            }
            in[j] = v[j] >= lo && v[j] <= hi;
            const float t = __fmul_rn(__fsub_rn(v[j], lo), nb);
            float q = __fmul_rn(t, rw);
            q = __fmaf_rn(__fmaf_rn(-q, w, t), rw, q);
            q = __fmaf_rn(__fmaf_rn(-q, w, t), rw, q);
            b[j] = min((int)q, bins - 1);  // right edge -> last bin
        }
        if (!recip_ok) {
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (in[j]) b[j] = bin_exact(v[j]);
        }
#pragma unroll
        for (int j = 0; j < V; ++j)
            if (in[j] && b[j] >= 0) atomicAdd(&sbin[b[j]], 1u);
    };
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream(x + (i + u * stride) * V);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[V];
            VecIO<T>::unpack(r[u], v);
            put_vec(v);
        }
    }
    for (; i < nvec; i += stride) {
        float v[V];
        VecIO<T>::load(x + i * V, v);
        put_vec(v);
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * V) {
        const float v = Cvt<T>::to_f32(x[nvec * V + threadIdx.x]);
        if (MM) {
            fmn = fminf(fmn, v);
            fmx = fmaxf(fmx, v);
            saw_nan |= v != v;
        }
        put(v);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += kThreads) {
        unsigned int c = sbin[b];
        if (c) atomicAdd(counts + b, (unsigned long long)c);
    }
    if (MM) {
        int mlo = saw_nan ? (int)0x80000000 : ord(fmn), mhi = saw_nan ? 0x7FFFFFFF : ord(fmx);
        for (int off = 16; off > 0; off >>= 1) {
            mlo = min(mlo, __shfl_xor_sync(0xFFFFFFFFu, mlo, off));
            mhi = max(mhi, __shfl_xor_sync(0xFFFFFFFFu, mhi, off));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(omin, mlo); atomicMax(omax, mhi); }
    }
}

template <typename T>
static void histc_dispatch(const void *x, int64_t n, float lo, float hi, int bins, unsigned long long *counts, int *omin, int *omax,
                           unsigned grid, cudaStream_t s)
{
    size_t sh = (size_t)bins * sizeof(unsigned int);
    if (omin) histc_kernel<T, true><<<grid, kThreads, sh, s>>>(static_cast<const T *>(x), n, lo, hi, bins, counts, omin, omax);
    else histc_kernel<T, false><<<grid, kThreads, sh, s>>>(static_cast<const T *>(x), n, lo, hi, bins, counts, omin, omax);
}

cudaError_t launch_histc(int dt, const void *x, int64_t n, float lo, float hi, int bins, unsigned long long *counts, float *out_min,
                         float *out_max, cudaStream_t s)
{
    int *omin = reinterpret_cast<int *>(out_min), *omax = reinterpret_cast<int *>(out_max);
    if (omin) minmax_init_kernel<<<1, 256, 0, s>>>(omin, omax, 1);
    if (n > 0) {
        const int V = dt == 0 ? 4 : 8;
        int64_t nvec = n / V;
        unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nvec + kThreads * 4 - 1) / (kThreads * 4), 148 * 8));
        if (dt == 0) histc_dispatch<float>(x, n, lo, hi, bins, counts, omin, omax, grid, s);
        else if (dt == 1) histc_dispatch<__nv_bfloat16>(x, n, lo, hi, bins, counts, omin, omax, grid, s);
        else histc_dispatch<__half>(x, n, lo, hi, bins, counts, omin, omax, grid, s);
    }
    if (omin) minmax_final_kernel<<<1, 256, 0, s>>>(omin, omax, 1);
    count_launch((n > 0 ? 1 : 0) + (omin ? 2 : 0));
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// fused residual add with its boundary casts (see dmxq_add_cast in include/dmxq.h)
// (float_fast_vec, pack16 and the Range16 identity test live in dmxq_stages.cuh: dmxq_softmax.cu shares them)
// BCAST 0: b has a's layout (no index arithmetic); 1: b = [o0, o1, inner] with arbitrary outer strides; 2: the same, with
// every inner run a whole number of CTA tiles, so the outer index is CTA-uniform (the attention-mask add: no per-vector division)
template <typename T, int BCAST> __global__ void __launch_bounds__(kThreads) add_cast_kernel(const __grid_constant__ AddParams p)
{
    constexpr int V = VecIO<T>::V;
    constexpr int U = 4;
    const T *__restrict__ a = static_cast<const T *>(p.a);
    const T *__restrict__ b = static_cast<const T *>(p.b);
    T *__restrict__ y = static_cast<T *>(p.y);
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    int64_t cta_boff = 0;
    if (BCAST == 2) {
        const int64_t o = ((int64_t)blockIdx.x * (kThreads * U)) / p.inner_vec;
        const int64_t o0 = o / p.d1, o1 = o - o0 * p.d1;
        cta_boff = o0 * p.bs0 + o1 * p.bs1 - o * (int64_t)p.inner_vec * V;  // + g * V gives the element offset in b
    }
    uint4 ra[U], rb[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int64_t g = g0 + (int64_t)u * kThreads;
        valid[u] = g < p.n_vec;
        int64_t gg = valid[u] ? g : 0;
        int64_t boff;
        if (BCAST == 0) {
            boff = gg * V;
        } else if (BCAST == 2) {
            boff = cta_boff + gg * V;
        } else if (p.n_vec <= 0xFFFFFFFFll) {  // 32-bit index arithmetic
            uint32_t g32 = (uint32_t)gg;
            uint32_t o = g32 / p.inner_vec, iv = g32 - o * p.inner_vec;
            uint32_t o0 = o / p.d1, o1 = o - o0 * p.d1;
            boff = (int64_t)o0 * p.bs0 + (int64_t)o1 * p.bs1 + (int64_t)iv * V;
        } else {
            int64_t o = gg / p.inner_vec, iv = gg - o * p.inner_vec;
            int64_t o0 = o / p.d1, o1 = o - o0 * p.d1;
            boff = o0 * p.bs0 + o1 * p.bs1 + iv * V;
        }
        ra[u] = valid[u] ? ldg_stream(a + gg * V) : make_uint4(0, 0, 0, 0);
        rb[u] = valid[u] ? (BCAST == 0 ? ldg_stream(b + boff) : *reinterpret_cast<const uint4 *>(b + boff)) : make_uint4(0, 0, 0, 0);
    }
    if constexpr (sizeof(T) == 2) {
        // 16-bit tensors: a FLOAT stage that keeps T's significand only flushes / saturates, so a vector whose magnitudes all
        // lie inside [flush threshold, saturation value] passes through it untouched -- two packed compares per stage
        const Range16 qa = range16<T>(p.has_a, p.fa), qb = range16<T>(p.has_b, p.fb), qo = range16<T>(p.has_o, p.fo);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float va[V], vb[V];
            uint4 wa = ra[u], wb = rb[u];
            if (p.has_a && !inside16(wa, qa)) {  // (stages that keep T's significand act on the packed words: flush_sat16_vec)
                if (qa.on) wa = flush_sat16_vec<T>(wa, p.fa, qa);
                else { VecIO<T>::unpack(wa, va); float_fast_vec<V>(va, p.fa); wa = pack16<T>(va); }
            }
            if (p.has_b && !inside16(wb, qb)) {
                if (qb.on) wb = flush_sat16_vec<T>(wb, p.fb, qb);
                else { VecIO<T>::unpack(wb, vb); float_fast_vec<V>(vb, p.fb); wb = pack16<T>(vb); }
            }
            VecIO<T>::unpack(wa, va);
            VecIO<T>::unpack(wb, vb);
#pragma unroll
            for (int j = 0; j < V; ++j) va[j] = __fadd_rn(va[j], vb[j]);  // torch adds in fp32, rounds to T
            uint4 w = pack16<T>(va);
            if (p.has_o && !inside16(w, qo)) {
                if (qo.on) w = flush_sat16_vec<T>(w, p.fo, qo);
                else { VecIO<T>::unpack(w, va); float_fast_vec<V>(va, p.fo); w = pack16<T>(va); }
            }
            if (valid[u]) stg_stream(y + (g0 + (int64_t)u * kThreads) * V, w);
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float va[V], vb[V];
            VecIO<T>::unpack(ra[u], va);
            VecIO<T>::unpack(rb[u], vb);
            if (p.has_a) float_fast_vec<V>(va, p.fa);
            if (p.has_b) float_fast_vec<V>(vb, p.fb);
#pragma unroll
            for (int j = 0; j < V; ++j) va[j] = __fadd_rn(va[j], vb[j]);
            if (p.has_o) float_fast_vec<V>(va, p.fo);
            if (valid[u]) VecIO<T>::template store<V>(y + (g0 + (int64_t)u * kThreads) * V, va);
        }
    }
}

template <typename T> static void launch_add_t(const AddParams &p, unsigned grid, cudaStream_t s)
{
    const bool same = p.d1 == 1 && p.bs0 == 0 && (int64_t)p.inner_vec == p.n_vec;
    if (same) add_cast_kernel<T, 0><<<grid, kThreads, 0, s>>>(p);
    else if (p.inner_vec % (kThreads * 4) == 0) add_cast_kernel<T, 2><<<grid, kThreads, 0, s>>>(p);
    else add_cast_kernel<T, 1><<<grid, kThreads, 0, s>>>(p);
}

cudaError_t launch_add(int dt, const AddParams &p, cudaStream_t s)
{
    int64_t per = (int64_t)kThreads * 4;
    int64_t grid = (p.n_vec + per - 1) / per;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    if (dt == 0) launch_add_t<float>(p, (unsigned)grid, s);
    else if (dt == 1) launch_add_t<__nv_bfloat16>(p, (unsigned)grid, s);
    else launch_add_t<__half>(p, (unsigned)grid, s);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// packed BFP storage (see dmxq_bfp_pack in include/dmxq.h).  Same tiling as the rows kernel; the
// integer mantissa falls out of the float-add rounding trick for free: u = (x + base) + C has
// ulp == Q, so bits(u) - bits(C) is t / Q as an integer and k = that - base / Q = ... - 3 * 2^(wl-1)
// (16-bit sources skip the base: bits(x + C) - bits(C) is x / Q directly).
template <typename T, bool NIBBLE> __global__ void __launch_bounds__(kThreads) bfp_pack_kernel(const T *__restrict__ x, uint8_t *__restrict__ mant, uint8_t *__restrict__ exps, int64_t n_vec, int lanes, int lshift, int wl)
{
    constexpr int V = VecIO<T>::V;
    constexpr int U = 4;
    constexpr bool SRC16 = sizeof(T) == 2;
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    const int kmax = (1 << (wl - 1)) - 1;
    const bool two_add = !(SRC16 && ((std::is_same<T, __nv_bfloat16>::value && wl <= 14) || (std::is_same<T, __half>::value && wl <= 11)));
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int64_t g = g0 + (int64_t)u * kThreads;
        raw[u] = g < n_vec ? ldg_stream(x + g * V) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        float v[V];
        uint32_t m = lanes_max(unpack_absmax<T>(raw[u], v), lanes);
        int k[V];
        const bool ok = bfp_fast_ok(m);
        if (ok) {
            BfpFast b = bfp_fast_block(m, wl);
            const int cb = (int)f2u(b.C);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float uu = two_add ? __fadd_rn(__fadd_rn(v[j], b.base), b.C) : __fadd_rn(v[j], b.C);
                int q = (int)f2u(uu) - cb - (two_add ? 3 * (1 << (wl - 1)) : 0);
                k[j] = max(-kmax, min(kmax, q));
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) k[j] = 0;
        }
        if (g >= n_vec) continue;
        if ((threadIdx.x & (lanes - 1)) == 0) exps[g >> lshift] = ok ? (uint8_t)(m >> 23) : 0;  // lanes is a power of two
        if (NIBBLE) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < V; ++j) acc |= ((uint32_t)k[j] & 0xFu) << (4 * j);
            if (V == 8) reinterpret_cast<uint32_t *>(mant)[g] = acc;
            else reinterpret_cast<uint16_t *>(mant)[g] = (uint16_t)acc;
        } else {
            uint32_t w[2] = {0u, 0u};
#pragma unroll
            for (int j = 0; j < V; j += 4)  // low bytes of four ints -> one word: three byte permutes
                w[j / 4] = __byte_perm(__byte_perm((uint32_t)k[j], (uint32_t)k[j + 1], 0x0040), __byte_perm((uint32_t)k[j + 2], (uint32_t)k[j + 3], 0x0040), 0x5410);
            if (V == 8) reinterpret_cast<uint2 *>(mant)[g] = make_uint2(w[0], w[1]);
            else reinterpret_cast<uint32_t *>(mant)[g] = w[0];
        }
    }
}

template <typename T, bool NIBBLE> __global__ void __launch_bounds__(kThreads) bfp_unpack_kernel(const uint8_t *__restrict__ mant, const uint8_t *__restrict__ exps, T *__restrict__ y, int64_t n_vec, int lshift, int wl)
{
    constexpr int V = VecIO<T>::V;
    constexpr int U = 4;
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    uint32_t w[U][2];
    int ef[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // all loads first
        const int64_t g = g0 + (int64_t)u * kThreads;
        w[u][0] = w[u][1] = 0u;
        ef[u] = 0;
        if (g < n_vec) {
            ef[u] = __ldg(exps + (g >> lshift));
            if (NIBBLE) {
                w[u][0] = V == 8 ? __ldg(reinterpret_cast<const uint32_t *>(mant) + g) : (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(mant) + g);
            } else if (V == 8) {
                const uint2 t = __ldg(reinterpret_cast<const uint2 *>(mant) + g);
                w[u][0] = t.x; w[u][1] = t.y;
            } else {
                w[u][0] = __ldg(reinterpret_cast<const uint32_t *>(mant) + g);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        if (g >= n_vec) continue;
        // quantum Q = 2^(ef - 127 + 2 - wl), applied as two exact power-of-two factors so that a denormal Q stays exact
        const int qe = ef[u] + 2 - wl;
        const float s1 = u2f((uint32_t)max(qe, 1) << 23), s2 = qe >= 1 ? 1.0f : u2f((uint32_t)(127 + qe - 1) << 23);
        float v[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int k = NIBBLE ? ((int)(w[u][0] << (28 - 4 * j))) >> 28 : ((int)(w[u][j / 4] << (24 - 8 * (j & 3)))) >> 24;
            v[j] = ef[u] == 0 ? 0.0f : __fmul_rn(__fmul_rn((float)k, s1), s2);
        }
        VecIO<T>::template store<V>(y + g * V, v);
    }
}

static int log2_pow2(int v)
{
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

cudaError_t launch_bfp_pack(int dt, const void *x, void *mant, uint8_t *exps, int64_t n, int B, int wl, cudaStream_t s)
{
    const int V = dt == 0 ? 4 : 8;
    int64_t n_vec = n / V;
    int64_t grid = (n_vec + kThreads * 4 - 1) / (kThreads * 4);
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    const bool nib = wl <= 4;
    uint8_t *m8 = static_cast<uint8_t *>(mant);
#define DMXQ_PACK(T) do { if (nib) bfp_pack_kernel<T, true><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const T *>(x), m8, exps, n_vec, B / V, log2_pow2(B / V), wl); \
                          else bfp_pack_kernel<T, false><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const T *>(x), m8, exps, n_vec, B / V, log2_pow2(B / V), wl); } while (0)
    if (dt == 0) DMXQ_PACK(float); else if (dt == 1) DMXQ_PACK(__nv_bfloat16); else DMXQ_PACK(__half);
#undef DMXQ_PACK
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_bfp_unpack(int dt, const void *mant, const uint8_t *exps, void *y, int64_t n, int B, int wl, cudaStream_t s)
{
    const int V = dt == 0 ? 4 : 8;
    int64_t n_vec = n / V;
    int64_t grid = (n_vec + kThreads * 4 - 1) / (kThreads * 4);
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    const bool nib = wl <= 4;
    const uint8_t *m8 = static_cast<const uint8_t *>(mant);
#define DMXQ_UNPACK(T) do { if (nib) bfp_unpack_kernel<T, true><<<(unsigned)grid, kThreads, 0, s>>>(m8, exps, static_cast<T *>(y), n_vec, log2_pow2(B / V), wl); \
                            else bfp_unpack_kernel<T, false><<<(unsigned)grid, kThreads, 0, s>>>(m8, exps, static_cast<T *>(y), n_vec, log2_pow2(B / V), wl); } while (0)
    if (dt == 0) DMXQ_UNPACK(float); else if (dt == 1) DMXQ_UNPACK(__nv_bfloat16); else DMXQ_UNPACK(__half);
#undef DMXQ_UNPACK
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// packed SBFP storage (see dmxq_sbfp_pack in include/dmxq.h): per block one scaler byte (the low-bit float scaler's
// exponent and mantissa fields, 0 = zero scaler) and sign-magnitude integer mantissas -- sign-magnitude because the
// simulated cast keeps the sign of x on a zero result (roundf(-0.3) * fs = -0), which two's complement cannot hold.
// The quotient / rounding sequence is sbfp_apply's, on magnitudes.
struct SbfpPackFmt {
    SbfpFmt f;
    uint32_t base_code;  // (bits of the smallest normal scaler >> sh) - (1 << man): code = (bits(fs) >> sh) - base_code
    uint32_t code_max;   // 2^(exp + man) - 1
    int nibble;          // mantissas in 4 bits (precision <= 4) or 8
};

// VPB = 0: a block spans `lanes` neighbouring lanes (shuffle reduction).  VPB = 2 / 4: a block is VPB consecutive 16-byte
// vectors and ONE thread owns it (still whole 32- / 64-byte runs per lane), so the block header -- max / man_scaling, its
// scaler cast, the reciprocal pair -- is derived once per block instead of once per lane, and no shuffle is needed.
template <typename T, bool NIBBLE, int VPB> __global__ void __launch_bounds__(kThreads) sbfp_pack_kernel(const T *__restrict__ x, uint8_t *__restrict__ mant, uint8_t *__restrict__ scalers, unsigned int *__restrict__ n_inexact, int64_t n_vec, int lanes, int lshift, const __grid_constant__ SbfpPackFmt pf)
{
    constexpr int V = VecIO<T>::V;
    constexpr int U = 4;
    constexpr int W = VPB ? VPB : 1;  // vectors per header
    constexpr uint32_t SIGN = NIBBLE ? 0x8u : 0x80u;
    const SbfpFmt &f = pf.f;
    const int64_t cta0 = (int64_t)blockIdx.x * (kThreads * U);
    int64_t gi[U];
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        gi[u] = VPB ? cta0 + (int64_t)(u / W) * (kThreads * W) + (int64_t)threadIdx.x * W + (u % W) : cta0 + (int64_t)u * kThreads + threadIdx.x;
        raw[u] = gi[u] < n_vec ? ldg_stream(x + gi[u] * V) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int b0 = 0; b0 < U; b0 += W) {
        float v[W][V];
        uint32_t m = 0u;
#pragma unroll
        for (int t = 0; t < W; ++t) m = max(m, unpack_absmax<T>(raw[b0 + t], v[t]));
        if (!VPB) m = lanes_max(m, lanes);
        const SbfpBlock b = sbfp_block_ol(m, f);
        const bool finite = m < 0x7F800000u;
        // not representable: non-finite blocks, blocks so small that max / man_scaling underflows to zero (the cast passes
        // their denormals through), scalers beyond the byte's exponent field (the reference's simulated formats saturate at
        // 2^(2^(exp-1)) whatever the bias).  The first two are stored as zeros, the last saturates the byte.
        bool exact = finite && (b.on || m == 0u);
        uint32_t code = 0u;
        if (finite && b.on && b.fs != 0.0f) {
            code = (f2u(b.fs) >> f.sc.sh) - pf.base_code;
            if (code > pf.code_max) { code = pf.code_max; exact = false; }
        }
        if (gi[b0] < n_vec && (VPB || (threadIdx.x & (lanes - 1)) == 0)) {  // (lanes is a power of two)
            scalers[gi[b0] >> lshift] = (uint8_t)code;
            if (!exact && n_inexact) atomicAdd(n_inexact, 1u);
        }
#pragma unroll
        for (int t = 0; t < W; ++t) {
            const int64_t g = gi[b0 + t];
            uint32_t k[V];
            if (finite && b.on) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float a = fabsf(v[t][j]);
                    const float q = b.rok ? (sizeof(T) == 2 ? div_by_recip16(a, b.rc, b.rl) : div_by_recip2(a, b.cmax, b.rc, b.rl)) : __fdiv_rn(a, b.cmax);
                    const float r = fminf(truncf(__fadd_rz(q, 0.5f)), f.man_scaling);
                    k[j] = (uint32_t)(int)r | ((f2u(v[t][j]) >> 31) ? SIGN : 0u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) k[j] = (finite && (f2u(v[t][j]) >> 31)) ? SIGN : 0u;  // a zero block keeps its signs
            }
            if (g >= n_vec) continue;
            if (NIBBLE) {
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < V; ++j) acc |= k[j] << (4 * j);
                if (V == 8) reinterpret_cast<uint32_t *>(mant)[g] = acc;
                else reinterpret_cast<uint16_t *>(mant)[g] = (uint16_t)acc;
            } else {
                uint32_t w[2] = {0u, 0u};
#pragma unroll
                for (int j = 0; j < V; ++j) w[j / 4] |= k[j] << (8 * (j & 3));
                if (V == 8) reinterpret_cast<uint2 *>(mant)[g] = make_uint2(w[0], w[1]);
                else reinterpret_cast<uint32_t *>(mant)[g] = w[0];
            }
        }
    }
}

template <typename T, bool NIBBLE> __global__ void __launch_bounds__(kThreads) sbfp_unpack_kernel(const uint8_t *__restrict__ mant, const uint8_t *__restrict__ scalers, T *__restrict__ y, int64_t n_vec, int lshift, int sh, uint32_t base_code)
{
    constexpr int V = VecIO<T>::V;
    constexpr int U = 4;
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    uint32_t w[U][2];
    uint32_t code[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // all loads first
        const int64_t g = g0 + (int64_t)u * kThreads;
        w[u][0] = w[u][1] = 0u;
        code[u] = 0u;
        if (g < n_vec) {
            code[u] = __ldg(scalers + (g >> lshift));
            if (NIBBLE) {
                w[u][0] = V == 8 ? __ldg(reinterpret_cast<const uint32_t *>(mant) + g) : (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(mant) + g);
            } else if (V == 8) {
                const uint2 t = __ldg(reinterpret_cast<const uint2 *>(mant) + g);
                w[u][0] = t.x; w[u][1] = t.y;
            } else {
                w[u][0] = __ldg(reinterpret_cast<const uint32_t *>(mant) + g);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        if (g >= n_vec) continue;
        const float fs = code[u] ? u2f((code[u] + base_code) << sh) : 0.0f;
        float v[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t t = NIBBLE ? (w[u][0] >> (4 * j)) & 0xFu : (w[u][j / 4] >> (8 * (j & 3))) & 0xFFu;
            const uint32_t mag = NIBBLE ? t & 0x7u : t & 0x7Fu;
            const uint32_t sg = NIBBLE ? t >> 3 : t >> 7;
            v[j] = u2f(f2u(__fmul_rn((float)mag, fs)) | (sg << 31));  // the cast's own last step: magnitude * scaler, sign of x
        }
        VecIO<T>::template store<V>(y + g * V, v);
    }
}

cudaError_t launch_sbfp_pack(int dt, const void *x, void *mant, uint8_t *scalers, unsigned int *n_inexact, int64_t n, int B, const SbfpFmt &f, int sc_man, int sc_exp, cudaStream_t s)
{
    const int V = dt == 0 ? 4 : 8;
    int64_t n_vec = n / V;
    int64_t grid = (n_vec + kThreads * 4 - 1) / (kThreads * 4);
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    SbfpPackFmt pf;
    pf.f = f;
    pf.base_code = (f.sc.shift_exp >> f.sc.sh) - (1u << sc_man);
    pf.code_max = (1u << (sc_exp + sc_man)) - 1u;
    pf.nibble = f.man_scaling <= 7.0f;
    uint8_t *m8 = static_cast<uint8_t *>(mant);
    const int vpb = (B / V == 2 || B / V == 4) ? B / V : 0;  // a thread owns whole blocks (SBFP12_16: 4 fp32 / 2 sixteen-bit vectors)
#define DMXQ_PACK_V(T, NIB, VPB) sbfp_pack_kernel<T, NIB, VPB><<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const T *>(x), m8, scalers, n_inexact, n_vec, B / V, log2_pow2(B / V), pf)
#define DMXQ_PACK_N(T, NIB) do { if (vpb == 2) DMXQ_PACK_V(T, NIB, 2); else if (vpb == 4) DMXQ_PACK_V(T, NIB, 4); else DMXQ_PACK_V(T, NIB, 0); } while (0)
#define DMXQ_PACK(T) do { if (pf.nibble) DMXQ_PACK_N(T, true); else DMXQ_PACK_N(T, false); } while (0)
    if (dt == 0) DMXQ_PACK(float); else if (dt == 1) DMXQ_PACK(__nv_bfloat16); else DMXQ_PACK(__half);
#undef DMXQ_PACK
#undef DMXQ_PACK_N
#undef DMXQ_PACK_V
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sbfp_unpack(int dt, const void *mant, const uint8_t *scalers, void *y, int64_t n, int B, const SbfpFmt &f, int sc_man, cudaStream_t s)
{
    const int V = dt == 0 ? 4 : 8;
    int64_t n_vec = n / V;
    int64_t grid = (n_vec + kThreads * 4 - 1) / (kThreads * 4);
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    const bool nib = f.man_scaling <= 7.0f;
    const uint32_t base_code = (f.sc.shift_exp >> f.sc.sh) - (1u << sc_man);
    const uint8_t *m8 = static_cast<const uint8_t *>(mant);
#define DMXQ_UNPACK(T) do { if (nib) sbfp_unpack_kernel<T, true><<<(unsigned)grid, kThreads, 0, s>>>(m8, scalers, static_cast<T *>(y), n_vec, log2_pow2(B / V), f.sc.sh, base_code); \
                            else sbfp_unpack_kernel<T, false><<<(unsigned)grid, kThreads, 0, s>>>(m8, scalers, static_cast<T *>(y), n_vec, log2_pow2(B / V), f.sc.sh, base_code); } while (0)
    if (dt == 0) DMXQ_UNPACK(float); else if (dt == 1) DMXQ_UNPACK(__nv_bfloat16); else DMXQ_UNPACK(__half);
#undef DMXQ_UNPACK
    count_launch();
    return cudaGetLastError();
}

__global__ void fold_absmax_kernel(const float *mn, const float *mx, uint32_t *out, int64_t C)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[i] = max(f2u(mn[i]) & 0x7FFFFFFFu, f2u(mx[i]) & 0x7FFFFFFFu);
}

cudaError_t launch_fold_absmax(const float *mn, const float *mx, uint32_t *out, int64_t C, cudaStream_t s)
{
    if (C <= 0) return cudaSuccess;
    fold_absmax_kernel<<<(unsigned)((C + 255) / 256), 256, 0, s>>>(mn, mx, out, C);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dmxq

namespace dmxq {

// ------------------------------------------------------------------------------------------------
// max|x| of many contiguous tensors in one launch (dmxq_amax_multi): the per-shard statistic of a sharded calibration
// pass.  A pure streaming read: 8 independent 16-byte loads per thread (32 KiB in flight per CTA), max over the magnitude
// bit patterns (packed two-per-instruction on 16-bit data; a NaN pattern is larger than Inf's, so NaN wins and sticks),
// shuffle + shared-memory reduction, ONE atomicMax per CTA on the tensor's slot (unsigned order == order of |x|).
constexpr int kAmaxUnroll = 8;

template <typename T> __global__ void __launch_bounds__(kThreads) amax_multi_kernel(const __grid_constant__ MultiTable t, uint32_t *__restrict__ out)
{
    constexpr int V = VecIO<T>::V;
    int lo = 0, hi = t.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blockIdx.x >= t.cta0[mid]) lo = mid; else hi = mid;
    }
    const T *__restrict__ x = static_cast<const T *>(t.x[lo]);
    const int64_t n = t.n_vec[lo], n_vec = n / V;
    const int64_t g0 = (int64_t)(blockIdx.x - t.cta0[lo]) * (kThreads * kAmaxUnroll) + threadIdx.x;
    uint4 raw[kAmaxUnroll];
#pragma unroll
    for (int u = 0; u < kAmaxUnroll; ++u) {
        const int64_t g = g0 + (int64_t)u * kThreads;
        raw[u] = g < n_vec ? ldg_stream(x + g * V) : make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t m = 0u;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int u = 0; u < kAmaxUnroll; ++u)
            m = max(max(m, raw[u].x & 0x7FFFFFFFu), max(max(raw[u].y & 0x7FFFFFFFu, raw[u].z & 0x7FFFFFFFu), raw[u].w & 0x7FFFFFFFu));
    } else {
        uint32_t m2 = 0u;
#pragma unroll
        for (int u = 0; u < kAmaxUnroll; ++u)
            m2 = __vmaxu2(m2, __vmaxu2(__vmaxu2(raw[u].x & 0x7FFF7FFFu, raw[u].y & 0x7FFF7FFFu), __vmaxu2(raw[u].z & 0x7FFF7FFFu, raw[u].w & 0x7FFF7FFFu)));
        m = max(m2 & 0xFFFFu, m2 >> 16);
    }
    if (g0 == 0) {  // the (< V) elements behind the last whole vector
        for (int64_t i = n_vec * V; i < n; ++i) {
            if constexpr (sizeof(T) == 4) m = max(m, f2u(Cvt<T>::to_f32(x[i])) & 0x7FFFFFFFu);
            else m = max(m, (uint32_t)(*reinterpret_cast<const unsigned short *>(x + i) & 0x7FFFu));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
    __shared__ uint32_t sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) m = max(m, sm[w]);
        if constexpr (sizeof(T) == 2) {  // widen the 15-bit magnitude pattern to the fp32 pattern of the same value
            if constexpr (std::is_same<T, __nv_bfloat16>::value) m <<= 16;
            else m = f2u(__half2float(__ushort_as_half((unsigned short)m)));
        }
        if (m != 0u) atomicMax(out + t.slot[lo], m);
    }
}

int64_t amax_multi_ctas(int dt, int64_t n_elems)
{
    const int V = dt == 0 ? 4 : 8;
    const int64_t per = (int64_t)kThreads * kAmaxUnroll * V;
    return std::max<int64_t>(1, (n_elems + per - 1) / per);
}

cudaError_t launch_amax_multi(int dt, const MultiTable &t, float *out, cudaStream_t s)
{
    const unsigned grid = t.cta0[t.n];
    if (grid == 0) return cudaSuccess;
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
    if (dt == 0) amax_multi_kernel<float><<<grid, kThreads, 0, s>>>(t, o);
    else if (dt == 1) amax_multi_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(t, o);
    else if (dt == 2) amax_multi_kernel<__half><<<grid, kThreads, 0, s>>>(t, o);
    else return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dmxq_philox_fill: the stream the kernels compute in registers, materialised (testing / layouts the rows kernels do not take)
__global__ void __launch_bounds__(256) philox_fill_kernel(uint32_t *__restrict__ out, int64_t n, int as_float, unsigned long long seed,
                                                          unsigned long long stream_id)
{
    const int64_t nq = (n + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        uint4 w = philox4x32_10((uint64_t)q, stream_id, seed);
        if (as_float) { w.x = philox_unit_bits(w.x); w.y = philox_unit_bits(w.y); w.z = philox_unit_bits(w.z); w.w = philox_unit_bits(w.w); }
        if (4 * q + 3 < n) {
            *reinterpret_cast<uint4 *>(out + 4 * q) = w;
        } else {
            const uint32_t v[4] = {w.x, w.y, w.z, w.w};
            for (int k = 0; 4 * q + k < n; ++k) out[4 * q + k] = v[k];
        }
    }
}

cudaError_t launch_philox_fill(void *out, int64_t n, int as_float, unsigned long long seed, unsigned long long stream_id, cudaStream_t s)
{
    const int64_t nq = (n + 3) >> 2;
    const unsigned grid = (unsigned)std::min<int64_t>((nq + 255) / 256, 148 * 16);
    philox_fill_kernel<<<grid, 256, 0, s>>>(static_cast<uint32_t *>(out), n, as_float, seed, stream_id);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dmxq

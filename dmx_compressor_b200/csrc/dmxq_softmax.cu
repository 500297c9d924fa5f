// dmxq_softmax.cu -- softmax along the contiguous dim with the casts around it folded in (dmxq_softmax_cast).
//
// The last full-size passes of a BASIC-mode attention block (reference S/modeling/nn/torch_modules.py: ResAdd mask add ->
// Softmax -> ActActMatMul) are, per layer and on the [B*H, S, S] scores:
//     mask add (+ 3 FLOAT16 casts)  ->  softmax input cast  ->  torch softmax  ->  output cast FLOAT16  ->  BFP16 input cast of P.V
// i.e. three read+write passes even after cast elision, the middle one torch's softmax_warp_forward, which reads and writes
// the row with 2- / 4-byte accesses strided by lane (1.85 ms for [96,2048,2048] bf16 on a B200: 0.87 TB/s).  This kernel does
// the whole sequence in ONE pass: a warp owns a row,
//   1. loads it (and the broadcast addend row) with 16-byte vectors, applies the add and its casts on the vectors,
//   2. transposes through its private shared-memory row into torch's element-to-lane assignment (lane l holds elements
//      l, l + 32, l + 64, ...), because the softmax must be BIT-IDENTICAL to torch's: same per-lane sequential max / exp / sum
//      order, same xor-butterfly (offsets 16..1), expf and IEEE division as ATen's PersistentSoftmax.cuh computes them,
//   3. rounds the probabilities to the tensor dtype (what torch's kernel stores), transposes back, and
//   4. runs the output cast chain (FLOAT16 -> BFP16 ...) on 16-byte vectors exactly as chain_rows_kernel's runtime chain does,
//      blocks along the row reduced with xor-shuffles, and stores with 16-byte streaming stores.
// Algorithmic traffic: sizeof(T) in + sizeof(T) out per element (+ the addend, which for an attention mask is L2-resident).
// Rows of 33..2048 elements (torch's persistent-warp range with a full warp per row; longer rows use another ATen kernel and
// are refused here -- the caller falls back on torch.softmax + dmxq_cast_chain).
#include "dmxq_rows.cuh"

namespace dmxq {

enum : int { POST_NONE = 0, POST_CHAIN = 1, POST_FLOAT_BFP = 2 };

// the add in front of the softmax on one vector: ResAdd.forward with its casts, as add_cast_kernel computes it
// (cast(a), cast(b) each rounded to T, fp32 add rounded to T, cast of the sum) -> the T-rounded result as a raw vector.
// Code size matters here (the row loop is fully unrolled around it: the kernel was instruction-fetch bound at 160 KB of SASS):
// the common case -- every stage absent or the identity on this vector -- is inline, everything else ONE out-of-line copy.
struct AddRanges {
    Range16 a, b, o;
};
template <typename T, int V> static __device__ __noinline__ uint4 add_vec_general(uint4 ra, uint4 rb, const SoftmaxParams *pp, const AddRanges *rr)
{
    const SoftmaxParams &p = *pp;
    float va[V], vb[V];
    if constexpr (sizeof(T) == 2) {
        const Range16 qa = rr->a, qb = rr->b, qo = rr->o;
        uint4 wa = ra, wb = rb;
        if (p.has_a && !inside16(wa, qa)) {
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
        return w;
    } else {
        VecIO<T>::unpack(ra, va);
        VecIO<T>::unpack(rb, vb);
        if (p.has_a) float_fast_vec<V>(va, p.fa);
        if (p.has_b) float_fast_vec<V>(vb, p.fb);
#pragma unroll
        for (int j = 0; j < V; ++j) va[j] = __fadd_rn(va[j], vb[j]);
        if (p.has_o) float_fast_vec<V>(va, p.fo);
        return make_uint4(f2u(va[0]), f2u(va[1]), f2u(va[2]), f2u(va[3]));
    }
}
template <typename T, int V> __device__ __forceinline__ uint4 add_vec(const uint4 &ra, const uint4 &rb, const SoftmaxParams &p, const Range16 &qa,
                                                                       const Range16 &qb, const Range16 &qo)
{
    if constexpr (sizeof(T) == 2) {
        if ((!p.has_a || inside16(ra, qa)) && (!p.has_b || inside16(rb, qb))) {
            float va[V], vb[V];
            VecIO<T>::unpack(ra, va);
            VecIO<T>::unpack(rb, vb);
#pragma unroll
            for (int j = 0; j < V; ++j) va[j] = __fadd_rn(va[j], vb[j]);
            const uint4 w = pack16<T>(va);
            if (!p.has_o || inside16(w, qo)) return w;
        }
        const AddRanges rr{qa, qb, qo};
        return add_vec_general<T, V>(ra, rb, &p, &rr);
    } else {  // fp32: the stages are real work on every vector (rounding to the format's mantissa): inline
        float va[V], vb[V];
        VecIO<T>::unpack(ra, va);
        VecIO<T>::unpack(rb, vb);
        if (p.has_a) float_fast_vec<V>(va, p.fa);
        if (p.has_b) float_fast_vec<V>(vb, p.fb);
#pragma unroll
        for (int j = 0; j < V; ++j) va[j] = __fadd_rn(va[j], vb[j]);
        if (p.has_o) float_fast_vec<V>(va, p.fo);
        return make_uint4(f2u(va[0]), f2u(va[1]), f2u(va[2]), f2u(va[3]));
    }
}

// the FLOAT (nearest + flush) stage of the output chain on a 16-bit vector -> the stage's result rounded to T, packed: identity and
// packed flush / saturate inline, anything else (NaN, formats that round T's significand, unsigned ones) out of line
template <typename T> static __device__ __noinline__ uint4 post_float16_general(uint4 raw, const StageDev *sf)
{
    float v[8];
    VecIO<T>::unpack(raw, v);
    float_fast_vec<8>(v, sf->ff);
    return pack16<T>(v);
}
// symmetric nearest BFP of a 16-bit vector whose block maximum is known: the two-add form inline, the literal form out of line
template <typename T, int V> __device__ __forceinline__ void post_bfp16(float (&v)[V], uint32_t m, const StageDev &st)
{
    if (st.fast && st.fast16 && bfp_fast_ok(m)) {
        const BfpFast b = bfp_fast_block(m, st.wl);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_fast16_elem(v[j], b);
        if (b.clamp) {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = bfp_clamp(v[j], b);
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = bfp_elem_slow(v[j], m, st.wl, st.sh, st.mask, R_NEAREST, 0, 0u);
    }
}
static __device__ __noinline__ float div_literal(float a, float b) { return b == 0.0f ? __int_as_float(0x7FC00000) : __fdiv_rn(a, b); }

// FULL: the row length equals the padded length ITERS * 32 (no bounds checks anywhere)
template <typename T, int ITERS, bool ADD, int POST, bool FULL>
__global__ void __launch_bounds__(256, 2) softmax_cast_kernel(const __grid_constant__ SoftmaxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int V = VecIO<T>::V;
    constexpr int NP = ITERS * 32;      // padded row length (a power of two, as in ATen's dispatch)
    constexpr int NV = NP / V / 32;     // vectors per lane (0 for the shortest rows: fewer vectors than lanes)
    constexpr int NVL = NV > 0 ? NV : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= p.rows) return;  // (warp-uniform; no CTA-wide barrier anywhere below)
    T *s = reinterpret_cast<T *>(smem_raw) + (size_t)warp * NP;
    const T *xr = static_cast<const T *>(p.x) + row * p.xs;
    T *yr = static_cast<T *>(p.y) + row * p.ys;
    const int n = FULL ? NP : p.n;
    const int nvec = n / V;  // the host guarantees n % V == 0 and 16-byte aligned rows

    // ---- 1. vector loads (a chunk of them issued before first use), the add and its casts, into the shared row
    constexpr int CH = ADD ? (NVL < 4 ? NVL : 4) : (NVL < 8 ? NVL : 8);  // vectors in flight per lane (x and addend each)
    const T *br = nullptr;
    Range16 qa{}, qb{}, qo{};
    if (ADD) {
        const uint32_t r32 = (uint32_t)row;  // rows < 2^32 (host-checked)
        const uint32_t q2 = r32 / p.d2, i2 = r32 - q2 * p.d2;
        const uint32_t i0 = q2 / p.d1, i1 = q2 - i0 * p.d1;
        br = static_cast<const T *>(p.b) + (int64_t)i0 * p.bs[0] + (int64_t)i1 * p.bs[1] + (int64_t)i2 * p.bs[2];
        if constexpr (sizeof(T) == 2) { qa = range16<T>(p.has_a, p.fa); qb = range16<T>(p.has_b, p.fb); qo = range16<T>(p.has_o, p.fo); }
    }
#pragma unroll 1
    for (int k0 = 0; k0 < NVL; k0 += CH) {
        uint4 ra[CH], rb[ADD ? CH : 1];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int j = (k0 + k) * 32 + lane;
            const bool ok = (FULL && NV > 0) || j < nvec;
            ra[k] = ok ? ldg_stream(xr + (size_t)j * V) : make_uint4(0u, 0u, 0u, 0u);
            if (ADD) rb[k] = ok ? *reinterpret_cast<const uint4 *>(br + (size_t)j * V) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int j = (k0 + k) * 32 + lane;
            if (!((FULL && NV > 0) || j < nvec)) continue;
            *reinterpret_cast<uint4 *>(s + (size_t)j * V) = ADD ? add_vec<T, V>(ra[k], rb[k], p, qa, qb, qo) : ra[k];
        }
    }
    __syncwarp();

    // ---- 2. torch's softmax_warp_forward (ATen/native/cuda/PersistentSoftmax.cuh), WARP_SIZE 32, WARP_BATCH 1: lane l owns
    // elements l, l + 32, ...; padding is -inf.  The maximum is exact in any order (fmaxf: a NaN anywhere in the row makes the
    // sum -- hence every output -- the canonical NaN in torch's comparison chain as well); exp and the sum keep torch's order.
    float e[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int idx = it * 32 + lane;
        e[it] = (FULL || idx < n) ? Cvt<T>::to_f32(s[idx]) : -__int_as_float(0x7F800000);
    }
    float m4[4] = {e[0], e[0], e[0], e[0]};
#pragma unroll
    for (int it = 0; it < ITERS; ++it) m4[it & 3] = fmaxf(m4[it & 3], e[it]);
    float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    float sum = 0.0f;
    uint32_t amin = 0xFFFFFFFFu;  // smallest non-zero exp pattern (minus one): decides whether the reciprocal form may divide
    // Steps are taken in groups of G: when x - max < -110 for EVERY element of a group across the warp (the masked tail of a causal
    // row), each expf is exactly +0 -- e^-110 is far below half the smallest denormal -- and adding +0 leaves the sum untouched, so
    // the group's exponentials, additions and (below) divisions are skipped; one predicate AND per element and one vote per group.
    constexpr int G = ITERS < 8 ? ITERS : 8;
    bool zero_group[ITERS / G];
#pragma unroll
    for (int g0 = 0; g0 < ITERS; g0 += G) {
        bool low = true;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            e[g0 + k] = __fsub_rn(e[g0 + k], m);
            low = low && e[g0 + k] < -110.0f;  // (false for NaN)
        }
        const bool z = __all_sync(0xFFFFFFFFu, low);
        zero_group[g0 / G] = z;
        if (z) {
#pragma unroll
            for (int k = 0; k < G; ++k) e[g0 + k] = 0.0f;
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) {
                e[g0 + k] = expf(e[g0 + k]);
                sum = __fadd_rn(sum, e[g0 + k]);
                amin = min(amin, f2u(e[g0 + k]) - 1u);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xFFFFFFFFu, sum, off));
    __syncwarp();  // every lane has read its inputs: the shared row may be overwritten

    // ---- 3. probabilities e / sum, rounded to T as torch stores them, back into the shared row.  IEEE division without the
    // division sequence: sum is in [1, 2048] (or NaN), so with the reciprocal carried as a high / low pair the four-operation
    // form div_by_recip2 (dmxq_stages.cuh) is e / sum correctly rounded -- for dividends that are zero or >= 2^-90 (quotient
    // and residual stay clear of the denormal range) and a divisor whose significand is not all ones; anything else divides.
    const float rh = __frcp_rn(sum), rl = recip_lo(sum, rh);
    const bool fast = recip_safe(sum) && amin >= (0x12800000u - 1u);  // 0x12800000 = 2^-90
    if (fast) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int idx = it * 32 + lane;
            // (a skipped group holds zeros and the divisor is finite here: 0 / sum = +0)
            const float r = zero_group[it / G] ? 0.0f : div_by_recip2(e[it], sum, rh, rl);
            if (FULL || idx < n) s[idx] = Cvt<T>::from_f32(r);
        }
    } else {
#pragma unroll  // (fully unrolled: a run-time index would put e[] into local memory for every path)
        for (int it = 0; it < ITERS; ++it) {
            const int idx = it * 32 + lane;
            if (FULL || idx < n) s[idx] = Cvt<T>::from_f32(div_literal(e[it], sum));
        }
    }
    __syncwarp();

    // ---- 4. the output casts on vectors, streaming stores
    FloatBfpCtx fb{};
    Range16 rg{};
    if (POST == POST_FLOAT_BFP) {
        fb = float_bfp_ctx<T, T>(p.chain.st[0]);
        if constexpr (sizeof(T) == 2) rg = range16<T>(1, p.chain.st[0].ff);
    }
#pragma unroll 1
    for (int k = 0; k < NVL; ++k) {
        const int j = k * 32 + lane;
        const bool ok = (FULL && NV > 0) || j < nvec;
        const uint4 raw = ok ? *reinterpret_cast<const uint4 *>(s + (size_t)j * V) : make_uint4(0u, 0u, 0u, 0u);
        if (POST == POST_NONE) {
            if (ok) stg_stream(yr + (size_t)j * V, raw);
            continue;
        }
        float v[V];
        if (POST == POST_FLOAT_BFP) {
            // every vector of this step holds nothing but zeros (the masked tail of a causal row): FLOAT then BFP of zeros are +0
            constexpr uint32_t kMag = sizeof(T) == 4 ? 0x7FFFFFFFu : 0x7FFF7FFFu;
            if (__all_sync(0xFFFFFFFFu, ((raw.x | raw.y | raw.z | raw.w) & kMag) == 0u)) {
                if (ok) stg_stream(yr + (size_t)j * V, make_uint4(0u, 0u, 0u, 0u));
                continue;
            }
            if constexpr (sizeof(T) == 2) {
                // FLOAT stage on the packed words (identity / flush + saturate), its maximum straight from the result, then BFP
                constexpr uint32_t kInf16 = std::is_same<T, __nv_bfloat16>::value ? 0x7F80u : 0x7C00u;
                const uint32_t amax = raw16_absmax(raw);
                uint4 w = raw;
                uint32_t m16 = amax;
                if (!rg.on || amax > kInf16) {
                    w = post_float16_general<T>(raw, &p.chain.st[0]);
                    m16 = raw16_absmax(w);
                } else if (!(raw16_absmin(raw) >= rg.lo && amax <= rg.hi)) {
                    w = flush_sat16_vec<T>(raw, p.chain.st[0].ff, rg);
                    m16 = raw16_absmax(w);
                }
                const uint32_t m = lanes_max(widen16<T>(m16), p.chain.st[1].block / V);
                VecIO<T>::unpack(w, v);
                post_bfp16<T, V>(v, m, p.chain.st[1]);
            } else {
                float_bfp_apply<T, T, V>(raw, v, p.chain.st[0], p.chain.st[1], fb);
            }
        } else {
            // chain_rows_kernel's runtime chain: the same stage functions, blocks reduced over neighbouring lanes
            VecIO<T>::unpack(raw, v);
#pragma unroll 1
            for (int q = 0; q < p.chain.n; ++q) {
                const StageDev &st = p.chain.st[q];
                const int lanes = st.block / V;
                uint32_t r[V];
#pragma unroll
                for (int i = 0; i < V; ++i) r[i] = 0x3F000000u;
                switch (st.kind) {
                case ST_BFP: bfp_stage<V>(v, st, lanes, r); break;
                case ST_SBFP: sbfp_stage<V>(v, st, lanes); break;
                case ST_FLOAT: float_stage<V>(v, st, r); break;
                case ST_FIXED: fixed_stage<V>(v, st, r); break;
                case ST_MXFP: mxfp_stage<V>(v, st, lanes); break;
                default: break;
                }
                if (st.requant) {
#pragma unroll
                    for (int i = 0; i < V; ++i) v[i] = requant1<T>(v[i]);
                }
            }
        }
        if (ok) VecIO<T>::template store<V>(yr + (size_t)j * V, v);
    }
}

template <typename T, int ITERS, bool ADD, int POST> static cudaError_t launch_softmax_tiap(const SoftmaxParams &p, cudaStream_t s)
{
    constexpr int WARPS = 8;
    const size_t smem = (size_t)WARPS * ITERS * 32 * sizeof(T);
    const int64_t grid = (p.rows + WARPS - 1) / WARPS;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    const bool full = p.n == ITERS * 32;
    auto run = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<(unsigned)grid, WARPS * 32, smem, s>>>(p);
        return cudaSuccess;
    };
    cudaError_t e = full ? run(softmax_cast_kernel<T, ITERS, ADD, POST, true>) : run(softmax_cast_kernel<T, ITERS, ADD, POST, false>);
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

template <typename T, int ITERS> static cudaError_t launch_softmax_ti(const SoftmaxParams &p, cudaStream_t s)
{
    int post = POST_CHAIN;
    if (p.chain.n == 0) post = POST_NONE;
    else if (p.chain.n == 2 && p.chain.st[0].kind == ST_FLOAT && p.chain.st[0].ff.fastpath && p.chain.st[1].kind == ST_BFP &&
             p.chain.st[1].mode == R_NEAREST && !p.chain.st[1].asym) post = POST_FLOAT_BFP;
    const bool add = p.b != nullptr;
    if (post == POST_NONE) return add ? launch_softmax_tiap<T, ITERS, true, POST_NONE>(p, s) : launch_softmax_tiap<T, ITERS, false, POST_NONE>(p, s);
    if (post == POST_FLOAT_BFP) return add ? launch_softmax_tiap<T, ITERS, true, POST_FLOAT_BFP>(p, s) : launch_softmax_tiap<T, ITERS, false, POST_FLOAT_BFP>(p, s);
    return add ? launch_softmax_tiap<T, ITERS, true, POST_CHAIN>(p, s) : launch_softmax_tiap<T, ITERS, false, POST_CHAIN>(p, s);
}

template <typename T> static cudaError_t launch_softmax_t(const SoftmaxParams &p, cudaStream_t s)
{
    // ATen: log2_elements = log2_ceil(dim_size); WARP_ITERATIONS = next_power_of_two / 32
    if (p.n <= 64) return launch_softmax_ti<T, 2>(p, s);
    if (p.n <= 128) return launch_softmax_ti<T, 4>(p, s);
    if (p.n <= 256) return launch_softmax_ti<T, 8>(p, s);
    if (p.n <= 512) return launch_softmax_ti<T, 16>(p, s);
    if (p.n <= 1024) return launch_softmax_ti<T, 32>(p, s);
    return launch_softmax_ti<T, 64>(p, s);
}

cudaError_t launch_softmax(int dt, const SoftmaxParams &p, cudaStream_t s)
{
    if (p.rows <= 0) return cudaSuccess;
    if (dt == 0) return launch_softmax_t<float>(p, s);
    if (dt == 1) return launch_softmax_t<__nv_bfloat16>(p, s);
    return launch_softmax_t<__half>(p, s);
}

}  // namespace dmxq

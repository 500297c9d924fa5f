// dmxq_rows.cu -- chain_rows_kernel: the blocked dim is contiguous in memory.
//
// This is the kernel of the headline casts (Linear inputs / weights, q, k^T views, attention
// probabilities, whole-model weight casting).  One pass over HBM:
//   * each thread owns kUnroll independent 16-byte vectors (all loads are issued before the
//     first use, so 64 B per thread / 16 KiB per CTA are in flight);
//   * a block of B elements lives in B/V neighbouring lanes of one warp; its max|x| is an
//     unsigned-integer max over bit patterns, reduced with log2(B/V) xor-shuffles;
//   * the fused round / clamp / rescale (+ N:M mask) runs on registers and the result is
//     written with one 16-byte streaming store.
// Algorithmic traffic: sizeof(in) + sizeof(out) bytes per element, nothing else.
//
//   SPECIAL 0: full feature set (score / mask / rand tensors honoured)
//   SPECIAL 1: no auxiliary tensors (deterministic chains)
//   SPECIAL 2: exactly one symmetric nearest BFP stage -- the headline cast, straight-line code
#include "dmxq_stages.cuh"

namespace dmxq {

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
    for (int d = p.nouter - 1; d >= 0; --d) {
        int64_t i = (d == 0) ? row : row % p.odim[d];
        if (d != 0) row /= p.odim[d];
        a.xo += i * p.xs[d]; a.yo += i * p.ys[d]; a.so += i * p.ss[d]; a.mo += i * p.ms[d]; a.ro += i * p.rs[d];
    }
    return a;
}

template <typename Tin, typename Tout, bool FLAT, int SPECIAL>
__global__ void __launch_bounds__(kThreads) chain_rows_kernel(const __grid_constant__ RowsParams p)
{
    constexpr int V = VecIO<Tin>::V;
    const Tin *__restrict__ x = static_cast<const Tin *>(p.x);
    Tout *__restrict__ y = static_cast<Tout *>(p.y);
    const int lane = threadIdx.x & 31;
    const int64_t g0 = (int64_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;

    float v[kUnroll][V];
    int64_t yoff[kUnroll];
    int64_t aux_s[kUnroll], aux_m[kUnroll], aux_r[kUnroll];
    bool valid[kUnroll];

#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        int64_t g = g0 + (int64_t)u * kThreads;
        int64_t xoff;
        if (FLAT) {
            valid[u] = g < p.n_vec;
            xoff = g * V;
            yoff[u] = xoff;
            if (SPECIAL == 0) { aux_s[u] = xoff; aux_m[u] = xoff; aux_r[u] = xoff; }
        } else {
            int64_t row;
            uint32_t kv;
            if (p.n_vec <= 0xFFFFFFFFll) {
                uint32_t g32 = (uint32_t)g;
                uint32_t r32 = g32 / p.vpr;
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
            if (SPECIAL == 0) { aux_s[u] = a.so + k; aux_m[u] = a.mo + k; aux_r[u] = a.ro + k * p.rks; }
        }
        if (valid[u]) {
            VecIO<Tin>::load(x + xoff, v[u]);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[u][j] = 0.0f;
        }
    }

    if (SPECIAL == 2) {
        const StageDev &st = p.chain.st[0];
        const int lanes = st.block / V;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            uint32_t m = lanes_max(vec_absmax<V>(v[u]), lanes);
            if (st.fast && bfp_fast_ok(m)) {
                BfpFast b = bfp_fast_block(m, st.wl);
#pragma unroll
                for (int j = 0; j < V; ++j) v[u][j] = bfp_fast_elem(v[u][j], b);
            } else {
                BfpBlock b = bfp_block(m, st.wl);
#pragma unroll
                for (int j = 0; j < V; ++j) v[u][j] = bfp_elem<R_NEAREST>(v[u][j], b, st.sh, st.mask, 0u);
            }
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < p.chain.n; ++s) {
            const StageDev &st = p.chain.st[s];
            const int lanes = st.block / V;
            const bool stoch = SPECIAL == 0 && p.rnd != nullptr &&
                               (st.kind == ST_FLOAT ? st.ff.mode : st.kind == ST_FIXED ? st.xf.mode : st.kind == ST_BFP ? st.mode : 0) == R_STOCHASTIC;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                uint32_t r[V];
#pragma unroll
                for (int j = 0; j < V; ++j) r[j] = 0x3F000000u;  // 0.5f: deterministic FIXED
                if (SPECIAL == 0 && stoch && valid[u]) {
                    const uint32_t *rp = static_cast<const uint32_t *>(p.rnd) + aux_r[u];
#pragma unroll
                    for (int j = 0; j < V; ++j) r[j] = __ldg(rp + (FLAT ? (int64_t)j : j * p.rks));
                }
                switch (st.kind) {
                case ST_NM:
                    nm_stage<V>(v[u], st, lane, (SPECIAL == 0 && p.score) ? p.score + aux_s[u] : nullptr,
                                (SPECIAL == 0 && p.mask) ? p.mask + aux_m[u] : nullptr, valid[u]);
                    break;
                case ST_BFP: bfp_stage<V>(v[u], st, lanes, r); break;
                case ST_SBFP: sbfp_stage<V>(v[u], st, lanes); break;
                case ST_FLOAT: float_stage<V>(v[u], st, r); break;
                case ST_FIXED: fixed_stage<V>(v[u], st, r); break;
                default: break;
                }
                if (st.requant) {
#pragma unroll
                    for (int j = 0; j < V; ++j) v[u][j] = requant1<Tout>(v[u][j]);
                }
            }
        }
    }

#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
        if (valid[u]) VecIO<Tout>::template store<V>(y + yoff[u], v[u]);
}

template <typename Tin, typename Tout>
static cudaError_t launch_rows_t(bool flat, int special, const RowsParams &p, cudaStream_t s)
{
    int64_t per_cta = (int64_t)kThreads * kUnroll;
    int64_t grid = (p.n_vec + per_cta - 1) / per_cta;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    dim3 g((unsigned)grid), b(kThreads);
#define DMXQ_ROWS(F, S) chain_rows_kernel<Tin, Tout, F, S><<<g, b, 0, s>>>(p)
    if (flat) {
        if (special == 2) DMXQ_ROWS(true, 2); else if (special == 1) DMXQ_ROWS(true, 1); else DMXQ_ROWS(true, 0);
    } else {
        if (special == 2) DMXQ_ROWS(false, 2); else if (special == 1) DMXQ_ROWS(false, 1); else DMXQ_ROWS(false, 0);
    }
#undef DMXQ_ROWS
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rows(int in_dt, int out_dt, bool flat, int special, const RowsParams &p, cudaStream_t s)
{
    if (in_dt == 0 && out_dt == 0) return launch_rows_t<float, float>(flat, special, p, s);
    if (in_dt == 1 && out_dt == 1) return launch_rows_t<__nv_bfloat16, __nv_bfloat16>(flat, special, p, s);
    if (in_dt == 2 && out_dt == 2) return launch_rows_t<__half, __half>(flat, special, p, s);
    if (in_dt == 1 && out_dt == 0) return launch_rows_t<__nv_bfloat16, float>(flat, special, p, s);
    if (in_dt == 2 && out_dt == 0) return launch_rows_t<__half, float>(flat, special, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace dmxq

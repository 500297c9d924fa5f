// dmxq_cols.cu -- chain_cols_kernel: the blocked dim is strided, another dim is contiguous
// (the PV multiplier v:[.., S, 64] blocked along S, conv inputs/weights blocked along channels).
//
// One warp per tile of B rows (along the blocked dim) x LI*V contiguous inner elements, V = 16
// bytes worth (4 fp32 / 8 bf16).  lane = lk * LI + li; lane (lk, li) owns rows
// [lk*RPT, (lk+1)*RPT) of the tile and the V columns at li*V.  Every column is its own block,
// so a thread carries V running maxima and
// reduces them over the LK lanes that share li with xor-shuffles.  All RPT row loads of a
// thread are issued back to back (RPT x 16 B in flight per thread); rows are LI*16 B (fp32)
// contiguous segments, i.e. whole 128-byte lines for LI = 8.  Nothing is transposed in HBM.
#include "dmxq_stages.cuh"

namespace dmxq {

template <typename Tin, typename Tout, int RPT, int LK>
__global__ void __launch_bounds__(kThreads, (RPT * (16 / sizeof(Tin)) <= 32 ? 3 : 2)) chain_cols_kernel(const __grid_constant__ ColsParams p)
{
    constexpr int V = VecIO<Tin>::V;  // 16-byte row segments per lane: 4 fp32 or 8 bf16/fp16
    constexpr int LI = 32 / LK;
    constexpr int B = RPT * LK;
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (tile >= p.n_tiles) return;  // whole warp leaves together
    const int li = lane % LI, lk = lane / LI;

    int64_t chunk, blk, xo = 0, yo = 0, ro = 0;
    if (p.n_tiles <= 0xFFFFFFFFll) {  // 32-bit index arithmetic (64-bit div/mod costs ~100 instructions each)
        uint32_t t = (uint32_t)tile, nc = (uint32_t)p.nchunk, nb = (uint32_t)p.nblk;
        uint32_t q = t / nc; chunk = t - q * nc; t = q;
        q = t / nb; blk = t - q * nb;
        uint32_t o = q;
        for (int d = p.nouter - 1; d >= 0; --d) {
            uint32_t od = (uint32_t)p.odim[d];
            uint32_t i = (d == 0) ? o : o % od;
            if (d != 0) o /= od;
            xo += (int64_t)i * p.xs[d]; yo += (int64_t)i * p.ys[d]; ro += (int64_t)i * p.rs[d];
        }
    } else {
        int64_t t = tile;
        chunk = t % p.nchunk; t /= p.nchunk;
        blk = t % p.nblk;
        int64_t o = t / p.nblk;
        for (int d = p.nouter - 1; d >= 0; --d) {
            int64_t i = (d == 0) ? o : o % p.odim[d];
            if (d != 0) o /= p.odim[d];
            xo += i * p.xs[d]; yo += i * p.ys[d]; ro += i * p.rs[d];
        }
    }
    const int64_t i0 = (chunk * LI + li) * V;
    const bool col_ok = i0 < p.inner;
    const int64_t kb = blk * B + (int64_t)lk * RPT;

    const Tin *__restrict__ x = static_cast<const Tin *>(p.x) + xo + i0;
    Tout *__restrict__ y = static_cast<Tout *>(p.y) + yo + i0;

    float v[RPT][V];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        int64_t k = kb + r;
        if (col_ok && k < p.K) {
            VecIO<Tin>::load(x + k * p.xks, v[r]);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[r][j] = 0.0f;
        }
    }

#pragma unroll 1
    for (int s = 0; s < p.chain.n; ++s) {
        const StageDev &st = p.chain.st[s];
        if (st.kind == ST_BFP || st.kind == ST_SBFP || st.kind == ST_MXFP) {
            uint32_t m[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                m[j] = 0;
#pragma unroll
                for (int r = 0; r < RPT; ++r) m[j] = max(m[j], f2u(v[r][j]) & 0x7FFFFFFFu);
#pragma unroll
                for (int off = LI; off < 32; off <<= 1) m[j] = max(m[j], __shfl_xor_sync(0xFFFFFFFFu, m[j], off));
            }
            if (st.kind == ST_BFP && st.mode == R_NEAREST) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (st.fast && !st.asym && bfp_fast_ok(m[j])) {
                        BfpFast b = bfp_fast_block(m[j], st.wl);
                        if (st.fast16) {
#pragma unroll
                            for (int r = 0; r < RPT; ++r) v[r][j] = bfp_fast16_elem(v[r][j], b);
                        } else {
#pragma unroll
                            for (int r = 0; r < RPT; ++r) v[r][j] = bfp_fast_elem(v[r][j], b);
                        }
                        if (b.clamp) {
#pragma unroll
                            for (int r = 0; r < RPT; ++r) v[r][j] = bfp_clamp(v[r][j], b);
                        }
                    } else {
                        BfpBlock b = bfp_block(m[j], st.wl);
#pragma unroll
                        for (int r = 0; r < RPT; ++r) {
                            float q = bfp_elem<R_NEAREST>(v[r][j], b, st.sh, st.mask, 0u);
                            if (st.asym) q = bfp_asym_fix(q, v[r][j], b);
                            v[r][j] = q;
                        }
                    }
                }
            } else if (st.kind == ST_BFP) {
#pragma unroll
                for (int j = 0; j < V; ++j)
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        uint32_t rnd = 0;
                        if (st.mode == R_STOCHASTIC && p.rnd != nullptr && col_ok && kb + r < p.K)
                            rnd = __ldg(static_cast<const uint32_t *>(p.rnd) + ro + (kb + r) * p.rks + (i0 + j) * p.ris);
                        v[r][j] = bfp_elem_slow(v[r][j], m[j], st.wl, st.sh, st.mask, st.mode, st.asym, rnd);
                    }
            } else if (st.kind == ST_MXFP) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    MxBlock b = mx_block(m[j], st.mx_largest);
#pragma unroll
                    for (int r = 0; r < RPT; ++r) v[r][j] = mx_elem_ol(v[r][j], b.scale, &st.ff);
                }
            } else {
                const bool fast = sbfp_fast(st.sb);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    SbfpBlock b = sbfp_block_ol(m[j], st.sb);
                    if (fast) {
#pragma unroll
                        for (int r = 0; r < RPT; ++r) v[r][j] = sbfp_elem_fast(v[r][j], b, st.sb);
                    } else {
#pragma unroll
                        for (int r = 0; r < RPT; ++r) v[r][j] = sbfp_elem_slow(v[r][j], b.cmax, b.fs, &st.sb.xp);
                    }
                }
            }
        } else if (st.kind == ST_FLOAT) {
            if (st.ff.mode == R_NEAREST) {
#pragma unroll
                for (int r = 0; r < RPT; ++r)
#pragma unroll
                    for (int j = 0; j < V; ++j) v[r][j] = float_elem_nearest(v[r][j], st.ff);
            } else {
#pragma unroll
                for (int r = 0; r < RPT; ++r)
#pragma unroll
                    for (int j = 0; j < V; ++j) v[r][j] = float_elem_slow(v[r][j], &st.ff, 0u);
            }
        } else if (st.kind == ST_FIXED) {
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int j = 0; j < V; ++j) v[r][j] = fixed_elem_slow(v[r][j], &st.xf, st.affine, st.sc, st.zp, 0.5f);
        }
        if (st.requant) {
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int j = 0; j < V; ++j) v[r][j] = requant1<Tout>(v[r][j]);
        }
    }

#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        int64_t k = kb + r;
        if (col_ok && k < p.K) VecIO<Tout>::template store<V>(y + k * p.yks, v[r]);
    }
}

struct ColsCfg { int B, RPT, LK; };
// fp32 (V = 4) and 16-bit (V = 8) tilings; RPT * V <= 64 live values per thread
static const ColsCfg kCfg32[] = {{8, 2, 4}, {16, 4, 4}, {32, 8, 4}, {64, 8, 8}, {128, 16, 8}};
static const ColsCfg kCfg16[] = {{8, 1, 8}, {16, 2, 8}, {32, 4, 8}, {64, 8, 8}, {128, 8, 16}};

static const ColsCfg *find_cfg(int in_dt, int B)
{
    const ColsCfg *t = in_dt == 0 ? kCfg32 : kCfg16;
    for (int i = 0; i < 5; ++i) if (t[i].B == B) return &t[i];
    return nullptr;
}
bool cols_supported(int in_dt, int B) { return find_cfg(in_dt, B) != nullptr; }
int cols_tile_inner(int in_dt, int B)
{
    const ColsCfg *c = find_cfg(in_dt, B);
    return c ? (32 / c->LK) * (in_dt == 0 ? 4 : 8) : 0;
}

template <typename Tin, typename Tout> static cudaError_t launch_cols_t(int B, const ColsParams &p, cudaStream_t s)
{
    constexpr int wpc = kThreads / 32;
    int64_t grid = (p.n_tiles + wpc - 1) / wpc;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    dim3 g((unsigned)grid), b(kThreads);
    if constexpr (sizeof(Tin) == 4) {
        switch (B) {
        case 8: chain_cols_kernel<Tin, Tout, 2, 4><<<g, b, 0, s>>>(p); break;
        case 16: chain_cols_kernel<Tin, Tout, 4, 4><<<g, b, 0, s>>>(p); break;
        case 32: chain_cols_kernel<Tin, Tout, 8, 4><<<g, b, 0, s>>>(p); break;
        case 64: chain_cols_kernel<Tin, Tout, 8, 8><<<g, b, 0, s>>>(p); break;
        case 128: chain_cols_kernel<Tin, Tout, 16, 8><<<g, b, 0, s>>>(p); break;
        default: return cudaErrorInvalidValue;
        }
    } else {
        switch (B) {
        case 8: chain_cols_kernel<Tin, Tout, 1, 8><<<g, b, 0, s>>>(p); break;
        case 16: chain_cols_kernel<Tin, Tout, 2, 8><<<g, b, 0, s>>>(p); break;
        case 32: chain_cols_kernel<Tin, Tout, 4, 8><<<g, b, 0, s>>>(p); break;
        case 64: chain_cols_kernel<Tin, Tout, 8, 8><<<g, b, 0, s>>>(p); break;
        case 128: chain_cols_kernel<Tin, Tout, 8, 16><<<g, b, 0, s>>>(p); break;
        default: return cudaErrorInvalidValue;
        }
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_cols(int in_dt, int out_dt, int B, const ColsParams &p, cudaStream_t s)
{
    if (in_dt == 0 && out_dt == 0) return launch_cols_t<float, float>(B, p, s);
    if (in_dt == 1 && out_dt == 1) return launch_cols_t<__nv_bfloat16, __nv_bfloat16>(B, p, s);
    if (in_dt == 2 && out_dt == 2) return launch_cols_t<__half, __half>(B, p, s);
    if (in_dt == 1 && out_dt == 0) return launch_cols_t<__nv_bfloat16, float>(B, p, s);
    if (in_dt == 2 && out_dt == 0) return launch_cols_t<__half, float>(B, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace dmxq

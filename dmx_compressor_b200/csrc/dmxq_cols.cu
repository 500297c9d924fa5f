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

// tile -> (inner chunk, block index along K, element offsets of the outer coordinate in x / y / rand)
struct ColsTile {
    int64_t chunk, blk, xo, yo, ro;
};
__device__ __forceinline__ ColsTile cols_tile(const ColsParams &p, int64_t tile)
{
    ColsTile c;
    c.xo = c.yo = c.ro = 0;
    if (p.n_tiles <= 0xFFFFFFFFll) {  // 32-bit index arithmetic (64-bit div/mod costs ~100 instructions each)
        uint32_t t = (uint32_t)tile, nc = (uint32_t)p.nchunk, nb = (uint32_t)p.nblk;
        uint32_t q = t / nc; c.chunk = t - q * nc; t = q;
        q = t / nb; c.blk = t - q * nb;
        uint32_t o = q;
        for (int d = p.nouter - 1; d >= 0; --d) {
            uint32_t od = (uint32_t)p.odim[d];
            uint32_t i = (d == 0) ? o : o % od;
            if (d != 0) o /= od;
            c.xo += (int64_t)i * p.xs[d]; c.yo += (int64_t)i * p.ys[d]; c.ro += (int64_t)i * p.rs[d];
        }
    } else {
        int64_t t = tile;
        c.chunk = t % p.nchunk; t /= p.nchunk;
        c.blk = t % p.nblk;
        int64_t o = t / p.nblk;
        for (int d = p.nouter - 1; d >= 0; --d) {
            int64_t i = (d == 0) ? o : o % p.odim[d];
            if (d != 0) o /= p.odim[d];
            c.xo += i * p.xs[d]; c.yo += i * p.ys[d]; c.ro += i * p.rs[d];
        }
    }
    return c;
}

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

    const ColsTile tc = cols_tile(p, tile);
    const int64_t chunk = tc.chunk, blk = tc.blk, xo = tc.xo, yo = tc.yo, ro = tc.ro;
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

// ------------------------------------------------------------------------------------------------
// 16-bit specialisation: ONE symmetric nearest BFP stage, Tin == Tout (bf16 or fp16), wl small enough for the two-add
// element form (bfp_fast16_elem).  Same tiling as above, but the tile stays PACKED in registers: RPT raw 16-byte rows
// (4 registers each) instead of RPT x 8 floats, the per-column maxima are taken two columns at a time on the packed
// words (also through the shuffles), and every word is widened, rounded and narrowed in place.  Half the registers of
// the generic kernel (twice the resident warps) and half its instructions.
template <typename T> struct Pack16;
template <> struct Pack16<__nv_bfloat16> {
    static __device__ __forceinline__ void widen(uint32_t w, float &lo, float &hi) { lo = u2f(w << 16); hi = u2f(w & 0xFFFF0000u); }
    static __device__ __forceinline__ uint32_t narrow(float lo, float hi)
    {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    static __device__ __forceinline__ uint32_t max_bits(uint32_t m16) { return m16 << 16; }
};
template <> struct Pack16<__half> {
    static __device__ __forceinline__ void widen(uint32_t w, float &lo, float &hi)
    {
        float2 f = __half22float2(*reinterpret_cast<__half2 *>(&w));
        lo = f.x; hi = f.y;
    }
    static __device__ __forceinline__ uint32_t narrow(float lo, float hi)
    {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    static __device__ __forceinline__ uint32_t max_bits(uint32_t m16) { return f2u(__half2float(__ushort_as_half((unsigned short)m16))); }
};

template <typename T> static __device__ __noinline__ uint32_t bfp_word16_slow(uint32_t w, uint32_t ma, uint32_t mb, int wl, int sh, uint32_t mask)
{
    float lo, hi;
    Pack16<T>::widen(w, lo, hi);
    return Pack16<T>::narrow(bfp_elem_slow(lo, ma, wl, sh, mask, R_NEAREST, 0, 0u), bfp_elem_slow(hi, mb, wl, sh, mask, R_NEAREST, 0, 0u));
}

template <typename T, int RPT, int LK>
__global__ void __launch_bounds__(kThreads, 4) bfp_cols16_kernel(const __grid_constant__ ColsParams p)
{
    constexpr int V = 8;
    constexpr int LI = 32 / LK;
    constexpr int B = RPT * LK;
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (tile >= p.n_tiles) return;  // whole warp leaves together
    const int li = lane % LI, lk = lane / LI;

    const ColsTile tc = cols_tile(p, tile);
    const int64_t chunk = tc.chunk, blk = tc.blk, xo = tc.xo, yo = tc.yo;
    const int64_t i0 = (chunk * LI + li) * V;
    const bool col_ok = i0 < p.inner;
    const int64_t kb = blk * B + (int64_t)lk * RPT;
    const T *__restrict__ x = static_cast<const T *>(p.x) + xo + i0;
    T *__restrict__ y = static_cast<T *>(p.y) + yo + i0;
    const StageDev &st = p.chain.st[0];

    uint32_t w[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const uint4 t = (col_ok && kb + r < p.K) ? ldg_stream(x + (kb + r) * p.xks) : make_uint4(0u, 0u, 0u, 0u);
        w[r][0] = t.x; w[r][1] = t.y; w[r][2] = t.z; w[r][3] = t.w;
    }
    // per-column max|x| patterns, two columns per register
    uint32_t m2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        m2[c] = 0u;
#pragma unroll
        for (int r = 0; r < RPT; ++r) m2[c] = __vmaxu2(m2[c], w[r][c] & 0x7FFF7FFFu);
#pragma unroll
        for (int off = LI; off < 32; off <<= 1) m2[c] = __vmaxu2(m2[c], __shfl_xor_sync(0xFFFFFFFFu, m2[c], off));
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t ma = Pack16<T>::max_bits(m2[c] & 0xFFFFu), mb = Pack16<T>::max_bits(m2[c] >> 16);
        if (bfp_fast_ok(ma) && bfp_fast_ok(mb)) {
            const BfpFast ba = bfp_fast_block(ma, st.wl), bb = bfp_fast_block(mb, st.wl);
            if (ba.clamp || bb.clamp) {
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    float lo, hi;
                    Pack16<T>::widen(w[r][c], lo, hi);
                    w[r][c] = Pack16<T>::narrow(bfp_clamp(bfp_fast16_elem(lo, ba), ba), bfp_clamp(bfp_fast16_elem(hi, bb), bb));
                }
            } else {
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    float lo, hi;
                    Pack16<T>::widen(w[r][c], lo, hi);
                    w[r][c] = Pack16<T>::narrow(bfp_fast16_elem(lo, ba), bfp_fast16_elem(hi, bb));
                }
            }
        } else {  // a denormal / huge / non-finite block in this column pair: the literal integer path
#pragma unroll
            for (int r = 0; r < RPT; ++r) w[r][c] = bfp_word16_slow<T>(w[r][c], ma, mb, st.wl, st.sh, st.mask);
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (col_ok && kb + r < p.K) stg_stream(y + (kb + r) * p.yks, make_uint4(w[r][0], w[r][1], w[r][2], w[r][3]));
}

template <typename T> static bool launch_cols16(int B, const ColsParams &p, dim3 g, dim3 b, cudaStream_t s)
{
    const StageDev &st = p.chain.st[0];
    if (!(p.chain.n == 1 && st.kind == ST_BFP && st.mode == R_NEAREST && !st.asym && st.fast && st.fast16)) return false;
    switch (B) {
    case 8: bfp_cols16_kernel<T, 1, 8><<<g, b, 0, s>>>(p); return true;
    case 16: bfp_cols16_kernel<T, 2, 8><<<g, b, 0, s>>>(p); return true;
    case 32: bfp_cols16_kernel<T, 4, 8><<<g, b, 0, s>>>(p); return true;
    case 64: bfp_cols16_kernel<T, 8, 8><<<g, b, 0, s>>>(p); return true;
    case 128: bfp_cols16_kernel<T, 8, 16><<<g, b, 0, s>>>(p); return true;
    default: return false;
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
    if constexpr (sizeof(Tin) == 2 && std::is_same<Tin, Tout>::value) {
        if (launch_cols16<Tin>(B, p, g, b, s)) {
            count_launch();
            return cudaGetLastError();
        }
    }
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

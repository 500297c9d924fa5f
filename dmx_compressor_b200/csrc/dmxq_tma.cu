// dmxq_tma.cu -- EXPERIMENT (round 2, VERDICT item "one TMA experiment, measured"): symmetric nearest BFP along a STRIDED
// dim of a contiguous [outer, K, inner] tensor (the `v` cast of an attention block: blocks run down the sequence, `inner` =
// head dim is the contiguous one) with the tile moved by TMA instead of per-lane 16-byte loads.
//
//   * a 3-D tensor map (inner, K, outer) with box = 128 bytes x B rows x 1: ONE cp.async.bulk.tensor instruction per tile
//     brings a whole block-column tile (B = block size rows x 128 B) into shared memory, no per-lane address arithmetic,
//     out-of-range rows / columns zero-filled by the hardware (ragged K and inner need no code);
//   * a warp owns a ring of STAGES tiles with one mbarrier each; no CTA-wide barrier exists.  A lane owns one 4-byte word of
//     every row (fp32: one column, 16-bit: two columns as a packed pair), holds the B words in registers (one conflict-free
//     LDS.32 per row), reduces its column maxima in registers -- no shuffles at all --, rounds, writes the words back in
//     place, and one elected lane sends the tile back with a TMA store (bulk group), refilling the stage once the store has
//     read it.
// Entry point (not part of include/dmxq.h: measured by scripts/probe_tma.py against the production kernels
// bfp_cols16_kernel / chain_cols_kernel, results in DESIGN.md section 4):
//     int dmxq_x_bfp_cols_tma(const void *x, void *y, int dtype, int64_t outer, int64_t K, int64_t inner, int block, int precision, int config, void *stream)
#include <cuda.h>

#include "dmxq_stages.cuh"

namespace dmxq {

template <int B> struct TmaColsParams {
    int64_t n_tiles;
    uint32_t kblk, iblk;  // tiles along K and along inner
    int wl;
    StageDev st;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <typename T, int B, int kTmaWarps, int kTmaStages>
__global__ void __launch_bounds__(kTmaWarps * 32) bfp_cols_tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap my,
                                                                      const __grid_constant__ TmaColsParams<B> p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWB = 128;                 // bytes per tile row
    constexpr int TILE = B * ROWB;            // bytes per tile
    constexpr int W = ROWB / (int)sizeof(T);  // tile width in elements
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = smem_raw + (size_t)warp * kTmaStages * TILE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kTmaWarps * kTmaStages * TILE) + warp * kTmaStages;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int64_t first = (int64_t)blockIdx.x * kTmaWarps + warp, stride = (int64_t)gridDim.x * kTmaWarps;
    auto coords = [&](int64_t t, int &c0, int &c1, int &c2) {
        const uint32_t ib = (uint32_t)(t % p.iblk);
        const int64_t r = t / p.iblk;
        c0 = (int)(ib * W);
        c1 = (int)((r % p.kblk) * B);
        c2 = (int)(r / p.kblk);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) {
            const int64_t t = first + (int64_t)s * stride;
            if (t < p.n_tiles) {
                int c0, c1, c2;
                coords(t, c0, c1, c2);
                mbar_expect_tx(&bars[s], TILE);
                tma_load_3d(&mx, &bars[s], ring + s * TILE, c0, c1, c2);
            }
        }
    }
    __syncwarp();

    int64_t i = 0;
    for (int64_t t = first; t < p.n_tiles; t += stride, ++i) {
        const int s = (int)(i % kTmaStages);
        const uint32_t parity = (uint32_t)((i / kTmaStages) & 1);
        mbar_wait(&bars[s], parity);
        uint32_t *tile = reinterpret_cast<uint32_t *>(ring + s * TILE) + lane;
        uint32_t w[B];
#pragma unroll
        for (int r = 0; r < B; ++r) w[r] = tile[r * 32];
        if constexpr (sizeof(T) == 4) {
            uint32_t mq[4] = {0u, 0u, 0u, 0u};  // (four independent chains: the tile's critical path, not issue, bounds a warp)
#pragma unroll
            for (int r = 0; r < B; ++r) mq[r & 3] = max(mq[r & 3], w[r] & 0x7FFFFFFFu);
            const uint32_t m = max(max(mq[0], mq[1]), max(mq[2], mq[3]));
            if (p.st.fast && bfp_fast_ok(m)) {
                const BfpFast b = bfp_fast_block(m, p.wl);
#pragma unroll
                for (int r = 0; r < B; ++r) {
                    float q = bfp_fast_elem(u2f(w[r]), b);
                    if (b.clamp) q = bfp_clamp(q, b);
                    w[r] = f2u(q);
                }
            } else {
#pragma unroll  // (fully unrolled: a run-time index would put w[] into local memory)
                for (int r = 0; r < B; ++r) w[r] = f2u(bfp_elem_slow(u2f(w[r]), m, p.wl, p.st.sh, p.st.mask, R_NEAREST, 0, 0u));
            }
        } else {
            uint32_t mq[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int r = 0; r < B; ++r) mq[r & 3] = __vmaxu2(mq[r & 3], w[r] & 0x7FFF7FFFu);
            const uint32_t m2 = __vmaxu2(__vmaxu2(mq[0], mq[1]), __vmaxu2(mq[2], mq[3]));
            uint32_t ma, mb;
            if constexpr (std::is_same<T, __nv_bfloat16>::value) { ma = (m2 & 0xFFFFu) << 16; mb = m2 & 0xFFFF0000u; }
            else { ma = f2u(__half2float(__ushort_as_half((unsigned short)(m2 & 0xFFFFu)))); mb = f2u(__half2float(__ushort_as_half((unsigned short)(m2 >> 16)))); }
            const bool fa = p.st.fast && p.st.fast16 && bfp_fast_ok(ma), fb = p.st.fast && p.st.fast16 && bfp_fast_ok(mb);
            const BfpFast ba = bfp_fast_block(ma, p.wl), bb = bfp_fast_block(mb, p.wl);
#pragma unroll
            for (int r = 0; r < B; ++r) {
                float lo, hi;
                if constexpr (std::is_same<T, __nv_bfloat16>::value) { lo = u2f(w[r] << 16); hi = u2f(w[r] & 0xFFFF0000u); }
                else { float2 f = __half22float2(*reinterpret_cast<__half2 *>(&w[r])); lo = f.x; hi = f.y; }
                if (fa) { lo = bfp_fast16_elem(lo, ba); if (ba.clamp) lo = bfp_clamp(lo, ba); }
                else lo = bfp_elem_slow(lo, ma, p.wl, p.st.sh, p.st.mask, R_NEAREST, 0, 0u);
                if (fb) { hi = bfp_fast16_elem(hi, bb); if (bb.clamp) hi = bfp_clamp(hi, bb); }
                else hi = bfp_elem_slow(hi, mb, p.wl, p.st.sh, p.st.mask, R_NEAREST, 0, 0u);
                if constexpr (std::is_same<T, __nv_bfloat16>::value) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); w[r] = *reinterpret_cast<uint32_t *>(&h); }
                else { __half2 h = __floats2half2_rn(lo, hi); w[r] = *reinterpret_cast<uint32_t *>(&h); }
            }
        }
#pragma unroll
        for (int r = 0; r < B; ++r) tile[r * 32] = w[r];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            int c0, c1, c2;
            coords(t, c0, c1, c2);
            tma_store_3d(&my, ring + s * TILE, c0, c1, c2);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // refill the PREVIOUS stage: its store (every group but the newest) has finished reading shared memory
            if (i >= 1) {
                const int64_t tn = t - stride + (int64_t)kTmaStages * stride;
                if (tn < p.n_tiles) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int sp = (int)((i - 1) % kTmaStages);
                    coords(tn, c0, c1, c2);
                    mbar_expect_tx(&bars[sp], TILE);
                    tma_load_3d(&mx, &bars[sp], ring + sp * TILE, c0, c1, c2);
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA retires
    __syncwarp();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <typename T, int B, int kTmaWarps, int kTmaStages>
static int run_tma(const void *x, void *y, int64_t outer, int64_t K, int64_t inner, int precision, cudaStream_t s)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -2;
    constexpr int W = 128 / (int)sizeof(T);
    CUtensorMap mx, my;
    const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)K, (cuuint64_t)outer};
    const cuuint64_t strides[2] = {(cuuint64_t)inner * sizeof(T), (cuuint64_t)K * inner * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)B, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                  : (std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (enc(&mx, dt, 3, const_cast<void *>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -3;
    if (enc(&my, dt, 3, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -3;
    TmaColsParams<B> p;
    memset(&p, 0, sizeof(p));
    p.kblk = (uint32_t)((K + B - 1) / B);
    p.iblk = (uint32_t)((inner + W - 1) / W);
    p.n_tiles = outer * (int64_t)p.kblk * p.iblk;
    p.wl = precision;
    p.st.kind = ST_BFP; p.st.block = B; p.st.wl = precision; p.st.mode = R_NEAREST;
    p.st.sh = 23 - precision;  // (dmxq_api.cu decode_stage)
    p.st.mask = (1u << p.st.sh) - 1u;
    p.st.fast = precision <= 20;
    p.st.fast16 = (std::is_same<T, __nv_bfloat16>::value && precision <= 14) || (std::is_same<T, __half>::value && precision <= 11);
    const size_t smem = (size_t)kTmaWarps * kTmaStages * B * 128 + kTmaWarps * kTmaStages * sizeof(uint64_t);
    auto kern = bfp_cols_tma_kernel<T, B, kTmaWarps, kTmaStages>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -4;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTmaWarps * 32, smem);
    const int64_t want = (p.n_tiles + kTmaWarps - 1) / kTmaWarps;
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)148 * std::max(per_sm, 1));
    kern<<<grid, kTmaWarps * 32, smem, s>>>(mx, my, p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace dmxq

template <int WARPS, int STAGES>
static int run_cfg(const void *x, void *y, int dtype, int64_t outer, int64_t K, int64_t inner, int precision, cudaStream_t s)
{
    using namespace dmxq;
    if (dtype == 0) return run_tma<float, 64, WARPS, STAGES>(x, y, outer, K, inner, precision, s);
    if (dtype == 1) return run_tma<__nv_bfloat16, 64, WARPS, STAGES>(x, y, outer, K, inner, precision, s);
    if (dtype == 2) return run_tma<__half, 64, WARPS, STAGES>(x, y, outer, K, inner, precision, s);
    return -1;
}

// config: 0 = 2 warps x 3 stages (48 KB per CTA, 4 CTAs / 8 warps per SM), 1 = 4 warps x 2 stages (64 KB, 3 CTAs / 12 warps),
//         2 = 1 warp x 2 stages (16 KB, 13 CTAs / 13 warps), 3 = 2 warps x 2 stages (32 KB, 7 CTAs / 14 warps)
extern "C" int dmxq_x_bfp_cols_tma(const void *x, void *y, int dtype, int64_t outer, int64_t K, int64_t inner, int block, int precision, int config, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (block != 64 || precision < 2 || precision > 16 || (inner * (dtype == 0 ? 4 : 2)) % 16 != 0) return -1;
    switch (config) {
    case 0: return run_cfg<2, 3>(x, y, dtype, outer, K, inner, precision, s);
    case 1: return run_cfg<4, 2>(x, y, dtype, outer, K, inner, precision, s);
    case 2: return run_cfg<1, 2>(x, y, dtype, outer, K, inner, precision, s);
    case 3: return run_cfg<2, 2>(x, y, dtype, outer, K, inner, precision, s);
    default: return -1;
    }
}

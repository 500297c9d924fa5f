// dmxq_kernels.cuh -- device-side parameter blocks shared by the kernels (dmxq_kernels.cu)
// and the dispatcher (dmxq_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmxq_numerics.cuh"

namespace dmxq {

constexpr int kMaxStages = 4;
constexpr int kMaxOuter = 3;   // outer dims the tiled kernels decompose (after collapsing)
constexpr int kMaxDims = 8;

// One stage of a cast chain, fully decoded for the device.
struct StageDev {
    int kind;        // ST_*
    int block;       // block size (BFP/SBFP) or M (NM)
    int requant;     // 1: round the stage result to the output dtype before the next stage
                     //    (what consecutive CastTo.forward calls do, S/numerical/cast.py:306)
    // BFP
    int wl, sh, mode, asym;
    int fast;        // wl <= 20: the float-add fast path of dmxq_numerics.cuh is valid
    int fast16;      // stage 0 on a 16-bit source with wl <= 14 (bf16) / 11 (fp16): two-add variant
    uint32_t mask;
    // NM
    int n_prune;
    int nm_order;    // DMXQ_NM_ORDER_*: tie order of the group sort (stable, or torch's CUDA bitonic network)
    // FLOAT
    FloatFmt ff;
    // FIXED (+ per-tensor affine)
    FixedFmt xf;
    int affine;
    float sc, zp;
    // SBFP
    SbfpFmt sb;
    int sb_exp_bits;  // exponent bits of the scaler format (needed when the bias is derived on the device from an amax)
    // MXFP (element format in ff)
    float mx_largest;  // 2^(2^(exp_bits-1))
    // SCALE: x / vec[k] or x * vec[k] along the blocked dim (SmoothQuant's scale application)
    const float *vec;
    int vec_op;
};

struct ChainDev {
    int n;
    StageDev st[kMaxStages];
};

// Row-tiled kernel: the blocked dim is contiguous (stride 1).  A "row" is one index tuple of
// the remaining (outer) dims.
// n / d for n < 2^31 as one multiply-high and a shift (Granlund-Montgomery round-up method): the hardware has no
// integer divider, `n / d` with a run-time d is a ~20-instruction sequence per use.
struct FastDiv {
    uint32_t mul, shr, d;
#ifdef __CUDACC__
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1u ? n : __umulhi(n, mul) >> shr; }
#endif
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        uint32_t lg = 0;
        while ((1ull << lg) < d) ++lg;                 // ceil(log2 d)
        const unsigned p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
    }
    return f;
}

struct RowsParams {
    const void *x;
    void *y;
    const float *score;  // nullable, NM stage
    float *mask;         // nullable, NM stage
    const void *rnd;     // nullable, stochastic stage (int32 or fp32)
    int philox;          // 1: no random tensor; word i = philox_word(logical element index i, ph_stream, ph_seed)
    unsigned long long ph_seed, ph_stream;
    const float *qscale, *qzp;  // nullable: per-tensor FixedPoint affine parameters in device memory (K_FIXED)
    int64_t n_vec;       // total (padded) vectors = rows * vpr
    int64_t rows;
    int64_t K;
    uint32_t vpr;        // vectors per padded row (multiple of lanes per tile)
    uint32_t kvec;       // valid vectors per row = K / V
    int nouter;          // 1..kMaxOuter
    int64_t odim[kMaxOuter];
    int64_t xs[kMaxOuter], ys[kMaxOuter], ss[kMaxOuter], ms[kMaxOuter], rs[kMaxOuter];
    int64_t rks;         // rand stride along K
    FastDiv vpr_div, odim_div[kMaxOuter];  // valid while n_vec < 2^31
    ChainDev chain;
};

// Many flat tensors in one launch of the rows kernel (dmxq_cast_chain_multi).  The table rides in the kernel parameters:
// CTAs [cta0[i], cta0[i+1]) work on tensor i.
constexpr int kMultiMax = 64;
struct MultiTable {
    int n;
    uint32_t cta0[kMultiMax + 1];
    const void *x[kMultiMax];
    void *y[kMultiMax];
    int64_t n_vec[kMultiMax];
    const float *amax;    // nullable: device array of per-tensor amax (SBFP scaler bias derived in the kernel)
    int slot[kMultiMax];  // index of tensor i in `amax`
};

// Column-tiled kernel: the blocked dim is strided, another dim ("inner") is contiguous.
struct ColsParams {
    const void *x;
    void *y;
    const void *rnd;
    int64_t n_tiles;     // outer * nblk * nchunk
    int64_t K, inner;
    int64_t nblk;        // ceil(K / B)
    int64_t nchunk;      // ceil(inner / (LI * V))
    int64_t xks, yks, rks;  // strides along K
    int64_t ris;            // rand stride along inner
    int nouter;
    int64_t odim[kMaxOuter];
    int64_t xs[kMaxOuter], ys[kMaxOuter], rs[kMaxOuter];
    ChainDev chain;
    int blocked;  // index of the (single) blocked stage in chain, or -1
};

// Generic kernel: arbitrary strides, one thread per (block, everything-else) pair.
struct GenericParams {
    const void *x;
    void *y;
    const float *score;
    float *mask;
    const void *rnd;
    int64_t n_items;  // prod(other dims) * nblk
    int64_t K, nblk;
    int64_t xks, yks, sks, mks, rks;
    int nd;  // number of non-block dims
    int64_t dim[kMaxDims];
    int64_t xs[kMaxDims], ys[kMaxDims], ss[kMaxDims], ms[kMaxDims], rs[kMaxDims];
    StageDev st;
};

// Per-channel / group affine fixed point on a contiguous (outer, C, inner) tensor.
struct FixedChanParams {
    const void *x;
    void *y;
    const float *scale, *zp, *rnd;
    int64_t n, C, inner, group, nq;
    FixedFmt xf;
};

// L1 block_quantize apply: per-"channel" max bits given in `maxbits`.
struct BlockQParams {
    const float *x;
    float *y;
    const int32_t *rnd;
    const uint32_t *maxbits;
    int64_t n, C, inner;  // channel of element i = (i / inner) % C  (C == 1: whole tensor)
    int wl, sh, mode, symmetric;
    uint32_t mask;
};

// fused add: y = out(A(a) + B(b)); a, y flat contiguous; b = [o0, o1, inner] with arbitrary outer strides
struct AddParams {
    const void *a, *b;
    void *y;
    int64_t n_vec;        // vectors of a
    uint32_t inner_vec;   // vectors per contiguous inner run of b
    uint32_t d1;          // size of the inner outer-dim (o = o0 * d1 + o1)
    int64_t bs0, bs1;     // element strides of b for o0, o1
    int has_a, has_b, has_o;
    FloatFmt fa, fb, fo;
};

// dmxq_softmax_cast (dmxq_softmax.cu)
struct SoftmaxParams {
    const void *x, *b;
    void *y;
    int64_t rows;
    int n;                 // row length
    int64_t xs, ys;        // row strides of x / y (elements)
    // addend row of row r = (i0, i1, i2) with r = (i0 * d1 + i1) * d2 + i2: b + i0 * bs[0] + i1 * bs[1] + i2 * bs[2]
    uint32_t d1, d2;
    int64_t bs[3];
    int has_a, has_b, has_o;  // FLOAT casts of x, of the addend, of their sum (ResAdd's input / residual / output casts)
    FloatFmt fa, fb, fo;
    ChainDev chain;        // casts applied to the probabilities (chain.n may be 0)
};

struct MinMaxParams {
    const void *x;
    int dtype;
    int64_t outer, C, inner;
    int64_t xo, xc, xi;  // strides
    int *omin, *omax;    // ordered-int accumulators, aliased onto the float outputs
};

// launchers (dmxq_kernels.cu)
cudaError_t launch_rows(int in_dt, int out_dt, bool flat, int kind, const RowsParams &p, cudaStream_t s);  // kind: K_* of dmxq_rows.cuh
cudaError_t launch_rows_multi(int dt, int kind, const RowsParams &p, const MultiTable &t, cudaStream_t s);  // same-dtype flat tensors
bool rows_multi_supported(int kind);
cudaError_t launch_cols(int in_dt, int out_dt, int B, const ColsParams &p, cudaStream_t s);
bool cols_supported(int in_dt, int B);
int cols_tile_inner(int in_dt, int B);  // LI * V of the instantiation used for block size B
cudaError_t launch_generic(int in_dt, int out_dt, const GenericParams &p, cudaStream_t s);
cudaError_t launch_fixed_chan(int in_dt, int out_dt, const FixedChanParams &p, cudaStream_t s);
cudaError_t launch_blockq(const BlockQParams &p, cudaStream_t s);
cudaError_t launch_minmax(const MinMaxParams &p, cudaStream_t s);
cudaError_t launch_histc(int dt, const void *x, int64_t n, float lo, float hi, int bins, unsigned long long *counts, float *out_min,
                         float *out_max, cudaStream_t s);
cudaError_t launch_add(int dt, const AddParams &p, cudaStream_t s);
cudaError_t launch_softmax(int dt, const SoftmaxParams &p, cudaStream_t s);
cudaError_t launch_philox_fill(void *out, int64_t n, int as_float, unsigned long long seed, unsigned long long stream_id, cudaStream_t s);
cudaError_t launch_bfp_pack(int dt, const void *x, void *mant, uint8_t *exps, int64_t n, int B, int wl, cudaStream_t s);
cudaError_t launch_bfp_unpack(int dt, const void *mant, const uint8_t *exps, void *y, int64_t n, int B, int wl, cudaStream_t s);
cudaError_t launch_sbfp_pack(int dt, const void *x, void *mant, uint8_t *scalers, unsigned int *n_inexact, int64_t n, int B, const SbfpFmt &f, int sc_man,
                             int sc_exp, cudaStream_t s);
cudaError_t launch_sbfp_unpack(int dt, const void *mant, const uint8_t *scalers, void *y, int64_t n, int B, const SbfpFmt &f, int sc_man, cudaStream_t s);
cudaError_t launch_amax_multi(int dt, const MultiTable &t, float *out, cudaStream_t s);  // t.n_vec holds ELEMENT counts here
int64_t amax_multi_ctas(int dt, int64_t n_elems);
cudaError_t launch_fold_absmax(const float *mn, const float *mx, uint32_t *out, int64_t C, cudaStream_t s);
int64_t launch_count();

}  // namespace dmxq

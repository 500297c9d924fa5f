/*
 * dmxq.h -- C ABI of libdmxq, the B200-native (sm_100a) CastTo / Sparsify numerics path.
 *
 * This header is the drop-in boundary: every entry point replaces a native entry of the
 * reference (d-matrix-ai/dmx-compressor v0.1.11) or the python loop that sits directly on
 * top of it.  "S/" = src/dmx/compressor/, "Q/" = src/dmx/compressor/quant/ of the reference.
 *
 *   entry point            replaces (reference file:line)
 *   ---------------------  ------------------------------------------------------------------
 *   dmxq_bfp_qdq           BlockFloatingPoint.cast           S/numerical/format.py:304-372
 *                          -> block_quantize                 Q/quant_function.py:87-117
 *                          -> block_quantize_*_cuda          Q/quant_cuda/quant.cu:14-112
 *                          -> block_kernel_*                 Q/quant_cuda/block_kernel.cu:7-139
 *   dmxq_sbfp_qdq          ScaledBlockFloatingPoint.cast     S/numerical/format.py:453-479
 *   dmxq_float_qdq         FloatingPoint.cast                S/numerical/format.py:208-233
 *                          -> float_quantize_*_cuda          Q/quant_cuda/quant.cu:155-228
 *                          -> float_kernel_*                 Q/quant_cuda/float_kernel.cu:6-168
 *   dmxq_fixed_qdq         FixedPoint.cast + affine wrap     S/numerical/format.py:134-142,
 *                                                            S/numerical/cast.py:279-296
 *                          -> fixed_point_quantize_*_cuda    Q/quant_cuda/quant.cu:230-327
 *   dmxq_nm_prune          BlockTopK.forward + x*mask        S/sparse.py:163-180, 287-301
 *   dmxq_cast_chain        DmxModule.weight_hypernet chain   S/modeling/nn/core.py:178-198
 *                          (sparsify -> storage cast -> weight cast in ONE pass), and any
 *                          back-to-back CastTo pair (output cast -> next input cast)
 *   dmxq_add_cast          ResAdd.forward (casts + add fused)  S/modeling/nn/torch_modules.py:15-37
 *   dmxq_bfp_pack/unpack   packed BFP storage (QuantizeBFP/DequantizeBFP of the ONNX export, S/numerical/cast.py:34-55)
 *   dmxq_sbfp_pack/unpack  packed SBFP storage (bytes_per_elem S/numerical/format.py:481-486, ids S/numerical/onnx.py:53-67)
 *   dmxq_block_quantize    L1 block_quantize(x, wl, dim,...) Q/quant_cuda/quant.cu:14-112
 *   dmxq_minmax            MinMaxObserver.forward statistics S/numerical/observer.py:173-193
 *   dmxq_histc             HistogramObserver.forward: torch.histc (+ aminmax) S/numerical/observer.py:454-499
 *   dmxq_cast_chain_host   same as dmxq_cast_chain on HOST buffers (pipelined H2D/compute/D2H)
 *
 * Conventions
 *   - plain C: pointers, sizes, enums; no torch types.  All tensors are described by a
 *     dmxq_tensor view (device pointer + dtype + shape + element strides), because real call
 *     sites pass views (e.g. key.transpose(-2,-1), per-head slices) and the reference's
 *     .contiguous()/transpose/reshape copies are exactly the HBM traffic this path removes.
 *   - the input is never modified; the output is caller-allocated (no memset needed) and may
 *     have its own strides; x and y must have the same shape.  In-place (y == x, same
 *     strides) is allowed.
 *   - kernels are enqueued on the caller's stream (a cudaStream_t passed as void*); the
 *     calls never synchronise and never allocate, except the *_host entry points.
 *   - return value: DMXQ_OK (0) or a negative dmxq_status; dmxq_last_error() holds a
 *     thread-local message.  Language bindings turn non-zero into their exception type
 *     (the reference raises RuntimeError from TORCH_CHECK, Q/quant_cuda/quant_cuda.cpp:7-11).
 *   - numerics are bit-exact with the reference's CUDA kernels for deterministic rounding
 *     and "same random tensor in => same bits out" for stochastic rounding.
 */
#ifndef DMXQ_H_
#define DMXQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMXQ_ABI_VERSION 3
#define DMXQ_MAX_DIMS 8
#define DMXQ_MAX_STAGES 4

typedef enum dmxq_status {
    DMXQ_OK = 0,
    DMXQ_ERR_BAD_ARG = -1,      /* malformed argument (the reference would assert / TORCH_CHECK) */
    DMXQ_ERR_UNSUPPORTED = -2,  /* well-formed but outside what the kernels implement */
    DMXQ_ERR_CUDA = -3,         /* a CUDA runtime call failed; message holds cudaGetErrorString */
    DMXQ_ERR_NO_DEVICE = -4
} dmxq_status;

typedef enum dmxq_dtype { DMXQ_F32 = 0, DMXQ_BF16 = 1, DMXQ_F16 = 2 } dmxq_dtype;

/* ROUNDING_MODE of S/numerical/format.py:23-30 ("N","S","U","D") */
typedef enum dmxq_rounding {
    DMXQ_ROUND_NEAREST = 0,
    DMXQ_ROUND_STOCHASTIC = 1,
    DMXQ_ROUND_UP = 2,
    DMXQ_ROUND_DOWN = 3
} dmxq_rounding;

/* FixedPoint "nearest" differs between the reference's own two back ends:
 *   CUDA  (Q/quant_cuda/sim_helper.cu:18-27)  roundf            -> half away from zero
 *   CPU   (Q/quant_cpu/sim_helper.cpp:14-21)  nearbyint(a+.5f-.5) -> ties to even (+ float add)
 * DMXQ_TIE_AWAY is what a user of the reference gets on CUDA tensors (the default of the
 * python binding); DMXQ_TIE_EVEN reproduces the CPU extension bit for bit. */
typedef enum dmxq_tie { DMXQ_TIE_AWAY = 0, DMXQ_TIE_EVEN = 1 } dmxq_tie;

/* The reference's CPU and CUDA back ends also disagree in two places that are not kernels of its own but torch
 * operators it calls; both are selectable per stage, the python binding defaults to what a user of the reference
 * gets on CUDA tensors:
 *
 * SBFP block scale (S/numerical/format.py:461-463, `get_chunk_max(chunk) / self.man_scaling`): dividing a CUDA tensor
 *   by a python scalar, torch multiplies by the reciprocal instead (ATen div_true_kernel_cuda: inv_b = 1.0 / b in
 *   double, rounded to fp32, then a * inv_b); on CPU tensors it divides.  The two differ in the last bit for about one
 *   max in ten, which moves rounding ties of x / scale and of the scaler cast.
 *     DMXQ_SCALE_DIV    max / man_scaling          (CPU tensors)
 *     DMXQ_SCALE_RECIP  max * fp32(1 / man_scaling) (CUDA tensors)
 *
 * N:M tie order (S/sparse.py:172, `torch.argsort(score, dim=1)`, default = unstable): on CPU the sort is stable (ties
 *   keep index order, so the lower index is pruned first); on CUDA rows of <= 32 keys go through ATen's 32-slot
 *   bitonic network (bitonicSortKVInPlace, ATen/native/cuda/SortUtils.cuh: LT comparator, tied keys ARE exchanged by
 *   the ascending comparators), whose result for tied keys is a fixed but pattern-dependent permutation.  Rows of more
 *   than 32 keys use a stable sort on CUDA too.
 *     DMXQ_NM_ORDER_STABLE      torch.argsort(stable=True) == the reference on CPU tensors
 *     DMXQ_NM_ORDER_TORCH_CUDA  the bitonic network's order, bit for bit == the reference on CUDA tensors */
typedef enum dmxq_scale_mode { DMXQ_SCALE_DIV = 0, DMXQ_SCALE_RECIP = 1 } dmxq_scale_mode;
typedef enum dmxq_nm_order { DMXQ_NM_ORDER_STABLE = 0, DMXQ_NM_ORDER_TORCH_CUDA = 1 } dmxq_nm_order;

typedef struct dmxq_tensor {
    void *data;                    /* device pointer (host pointer for *_host entry points) */
    int32_t dtype;                 /* dmxq_dtype */
    int32_t ndim;                  /* 0..DMXQ_MAX_DIMS */
    int64_t shape[DMXQ_MAX_DIMS];
    int64_t stride[DMXQ_MAX_DIMS]; /* in elements, like torch.Tensor.stride() */
} dmxq_tensor;

/* One stage of a fused cast chain.  Unused fields are ignored. */
typedef enum dmxq_stage_kind {
    DMXQ_STAGE_NONE = 0,
    DMXQ_STAGE_NM = 1,    /* N:M prune        : block = M, n_keep                                    */
    DMXQ_STAGE_BFP = 2,   /* block floating pt: block, precision, symmetric, rounding                */
    DMXQ_STAGE_SBFP = 3,  /* scaled BFP       : block, precision(XP wl), clamp, rounding(XP), tie,   */
                          /*                    sc_* = scaler FloatingPoint format                   */
    DMXQ_STAGE_FLOAT = 4, /* low-bit float    : man, exp, bias, flush, is_unsigned, fp16_flush, rounding */
    DMXQ_STAGE_FIXED = 5, /* fixed point      : precision(wl), fraction(fl), clamp, symmetric,       */
                          /*                    rounding, tie, scale, zero_point (per-tensor affine) */
    DMXQ_STAGE_SCALE = 7, /* per-channel scale: x / vec[k] (vec_op 0) or x * vec[k] (vec_op 1), k = index along      */
                          /*                    block_dim, vec = fp32 device vector of vec_len = shape[block_dim]: the  */
                          /*                    SmoothQuant scale application (S/numerical/smoothquant.py:253-283:      */
                          /*                    `a / scale.view(..)`, `b * scale.view(..)`) as a pre-stage of the casts */
                          /*                    that follow it.  The quotient / product is an fp32 value (torch         */
                          /*                    promotes `tensor op fp32_vector`); with an fp32 output tensor it stays    */
                          /*                    one (scale_input), with a 16-bit output it is rounded to that dtype      */
                          /*                    before the next stage (scale_weight's `.to(wgt.dtype)`).  Rows layouts.  */
    DMXQ_STAGE_MXFP = 6   /* MX floating pt   : block, man, exp (element format E<exp>M<man>, bias    */
                          /*                    2^(exp-1)-1, subnormals kept, nearest); power-of-two  */
                          /*                    block scale 2^floor(log2 max) / 2^(2^(exp-1))         */
                          /*                    (MXFP.cast, S/numerical/format.py:545-564)            */
} dmxq_stage_kind;

typedef struct dmxq_stage {
    int32_t kind;       /* dmxq_stage_kind */
    int32_t block;      /* block size along block_dim (BFP/SBFP) or M (NM) */
    int32_t precision;  /* BFP precision / XP word length */
    int32_t fraction;   /* XP fraction bits */
    int32_t man, exp, bias; /* FLOAT (and unused otherwise) */
    int32_t flush;      /* FLOAT flush_subnormal */
    int32_t is_unsigned;/* FLOAT: abs() after the cast (format.py:233) */
    int32_t fp16_flush; /* FLOAT: extra |x| < 2^-14 -> +0 pass of "FP[1|5|10,15](FN)" (format.py:223-232) */
    int32_t symmetric;  /* BFP: 0 => make_mantissa_asymmetric post-pass (format.py:349-372); XP: symmetric range */
    int32_t clamp;      /* XP clamp */
    int32_t rounding;   /* dmxq_rounding */
    int32_t tie;        /* dmxq_tie (XP nearest) */
    int32_t n_keep;     /* NM: K of K:M */
    int32_t sc_man, sc_exp, sc_bias, sc_flush, sc_unsigned, sc_fp16_flush, sc_rounding; /* SBFP scaler format */
    float scale, zero_point; /* FIXED per-tensor affine (cast.py:293,296); scale = 1, zp = 0 for none */
    int32_t scale_mode; /* SBFP: dmxq_scale_mode */
    int32_t nm_order;   /* NM: dmxq_nm_order */
    const float *vec;   /* SCALE: fp32 device vector (must stay alive until the launch has run) */
    int32_t vec_len;    /* SCALE: its length = the extent of block_dim */
    int32_t vec_op;     /* SCALE: 0 divide, 1 multiply */
} dmxq_stage;

int dmxq_abi_version(void);
const char *dmxq_last_error(void);
const char *dmxq_status_string(int status);
/* number of kernels this library has launched in the calling process (bench "gpu_launches") */
int64_t dmxq_launch_count(void);

/* ---- fused chain: y = stage[n-1](... stage[0](x)) along block_dim, one pass over HBM ------
 * `score`  (nullable) drives an NM stage (S/sparse.py:289-293); NULL => score = |x|.
 * `mask`   (nullable) receives the fp32 0/1 mask of the NM stage (Sparsify.mask).
 * `rand`   (nullable) contiguous random tensor in logical element order for the (single)
 *          stochastic stage: int32 for BFP/FLOAT (quant.cu:40,160), fp32 in [0,1) for FIXED
 *          (quant.cu:244).                                                                   */
int dmxq_cast_chain(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim,
                    const dmxq_stage *stages, int n_stages,
                    const dmxq_tensor *score, const dmxq_tensor *mask, const void *rand,
                    void *stream);

/* ---- stochastic rounding without a random tensor.  The reference's launchers draw a full-size
 * random tensor per cast (randint_like / rand_like, Q/quant_cuda/quant.cu:40,118,160,244): 4 bytes written + 4 read per element
 * next to the 8 the cast moves.  dmxq_cast_chain_philox computes the words in registers instead:
 *     word(i) = Philox4x32-10(counter = (i / 4, stream_id), key = seed)[i % 4],   i = logical element index of x
 * (FixedPoint stages consume (word >> 8) * 2^-24, a uniform fp32 in [0, 1)).  dmxq_philox_fill writes exactly that stream into a
 * tensor (as_float: the FixedPoint form), so cast_chain_philox(x, seed, stream_id) == cast_chain(x, rand = philox_fill(...)) bit
 * for bit -- the property the tests pin, next to a numpy restatement of the generator.  Rows layouts only (blocked dim
 * contiguous); DMXQ_ERR_UNSUPPORTED otherwise.  `rand` of dmxq_cast_chain stays the way to reproduce the reference seed for seed. */
int dmxq_cast_chain_philox(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, const dmxq_stage *stages, int n_stages,
                           uint64_t seed, uint64_t stream_id, void *stream);
int dmxq_philox_fill(void *out, int64_t n, int as_float, uint64_t seed, uint64_t stream_id, void *stream);

/* ---- many tensors, one launch: the shards a rank owns in a sharded whole-model weight cast (SURVEY.md section 8e;
 * replaces the per-module python loop of DmxModel.fold_weights_and_biases, S/modeling/model.py:145-150, over
 * DmxModule.weight_hypernet, S/modeling/nn/core.py:178-198).  ys[i] = chain(xs[i]) for i < n_tensors, the same stages and
 * block_dim for all.  Tensors whose layout the row-tiled kernel takes flat (contiguous, blocked along the last dim, same
 * dtype in and out) share launches -- up to 64 tensors each, CTAs dealt to tensors by a prefix table inside the kernel
 * parameters, so nothing is allocated or copied -- every other tensor is launched exactly as dmxq_cast_chain would.
 * `amax` (nullable): device array of n_tensors floats, amax[i] = max|x| over the WHOLE tensor xs[i] is a shard of (the
 * result of a calibration all-reduce).  When given, an SBFP stage takes its scaler exponent bias from it on the device,
 *     bias = (2^sc_exp - 1) - floor(log2(amax[i] / man_scaling)),  clamped to what FloatingPoint accepts,
 * instead of stage.sc_bias: the statistics never visit the host between the all-reduce and the cast.  (The reference
 * delegates this choice to the private `numerics` module, S/numerical/format.py:13-20, 438-446: parity unpinned.) */
int dmxq_cast_chain_multi(const dmxq_tensor *xs, const dmxq_tensor *ys, int n_tensors, int block_dim,
                          const dmxq_stage *stages, int n_stages, const float *amax, void *stream);

/* ---- single-format conveniences (thin wrappers over dmxq_cast_chain) ---------------------- */
int dmxq_bfp_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, int block_size,
                 int precision, int symmetric, int rounding, const int32_t *rand, void *stream);
int dmxq_sbfp_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, int block_size,
                  int xp_precision, int xp_clamp, int xp_rounding, int xp_tie,
                  int sc_man, int sc_exp, int sc_bias, int sc_flush, int sc_unsigned,
                  int sc_fp16_flush, int sc_rounding, int scale_mode, void *stream);
int dmxq_float_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int man, int exp, int bias,
                   int flush_subnormal, int is_unsigned, int fp16_flush, int rounding,
                   const int32_t *rand, void *stream);
/* scale / zero_point: device arrays of n_qparams floats (NULL => no affine wrap).
 * n_qparams == 1: per tensor.  Otherwise the qparam of index c along ch_axis is
 * scale[c / group_size] (group_size == 1: per channel, cast.py:228-237; > 1: group
 * quantisation, cast.py:281-292). */
int dmxq_fixed_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int wl, int fl, int clamp,
                   int symmetric, int rounding, int tie, const float *scale,
                   const float *zero_point, int64_t n_qparams, int ch_axis, int64_t group_size,
                   const float *rand, void *stream);
int dmxq_nm_prune(const dmxq_tensor *x, const dmxq_tensor *score, const dmxq_tensor *y,
                  const dmxq_tensor *mask, int block_dim, int n_keep, int m, int nm_order, void *stream);

/* ---- fused residual add: y = out( A(a) + B(b) ) in one pass (ResAdd.forward of the reference,
 * S/modeling/nn/torch_modules.py:15-37, with its two input casts and its output cast folded into the
 * add).  Stages are elementwise FLOAT stages or NULL (no cast).  a and y: same shape, contiguous,
 * same dtype; b: same dtype, broadcastable to a (stride-0 dims allowed, e.g. an attention mask).
 * Intermediates are rounded to the tensor dtype exactly where the unfused module sequence rounds. */
int dmxq_add_cast(const dmxq_tensor *a, const dmxq_tensor *b, const dmxq_tensor *y, const dmxq_stage *stage_a,
                  const dmxq_stage *stage_b, const dmxq_stage *stage_out, void *stream);

/* ---- fused attention softmax: y = post( softmax( sum( A(x) + B(addend) ) ) ) along the last (contiguous) dim, ONE pass.
 * Replaces, for one attention block of a BASIC-mode model, the reference's module sequence
 *   ResAdd.forward (mask add with its input / residual / output casts, S/modeling/nn/torch_modules.py:15-37)
 *   -> Softmax.forward (input cast, torch softmax, output cast; S/modeling/nn/core.py:215-264 around torch.nn.Softmax)
 *   -> the consumer's input cast (ActActMatMul's BFP16 cast of the probabilities, S/numerical/cast.py:261-306)
 * The softmax itself reproduces torch's CUDA kernel for rows of 33..2048 elements (ATen PersistentSoftmax.cuh
 * softmax_warp_forward: per-lane sequential max / expf / sum over elements lane, lane + 32, ..., xor-butterfly reductions,
 * IEEE division, result rounded to the tensor dtype) BIT FOR BIT; other row lengths return DMXQ_ERR_UNSUPPORTED (the caller
 * keeps torch.softmax).  addend (nullable): same dtype, broadcastable to x, contiguous along the last dim (an attention
 * mask); stage_x / stage_addend / stage_sum: nearest + flush FLOAT stages or NULL, as in dmxq_add_cast.  post[0..n_post):
 * the casts applied to the probabilities (n_post may be 0), blocked formats along the row with whole blocks; each stage is
 * followed by the rounding to the tensor dtype, as consecutive CastTo.forward calls do.  x, y: same shape / dtype, contiguous. */
int dmxq_softmax_cast(const dmxq_tensor *x, const dmxq_tensor *addend, const dmxq_tensor *y, const dmxq_stage *stage_x,
                      const dmxq_stage *stage_addend, const dmxq_stage *stage_sum, const dmxq_stage *post, int n_post, void *stream);

/* ---- packed BFP storage: the real format behind the simulation ------------
 * The reference only ever materialises dequantised fp32 tensors, but it reports the packed size
 * (BlockFloatingPoint.bytes_per_elem, S/numerical/format.py:345-347) and names the packed ops in its ONNX
 * export (com.microsoft::QuantizeBFP / DequantizeBFP, S/numerical/cast.py:34-55).  These two entry points
 * are that pair: per block of `block_size` consecutive elements one uint8 shared exponent (the biased
 * fp32 exponent of the block max) and `block_size` signed mantissas of `precision` bits -- int8 each for
 * precision 5..8, two per byte (low nibble first) for precision <= 4.
 *   x / y: contiguous [rows, K], K % block_size == 0, block_size % 16 == 0; nearest rounding, symmetric.
 *   dmxq_bfp_unpack(dmxq_bfp_pack(x)) == dmxq_bfp_qdq(x) bit for bit for every block whose max is a normal
 *   number below 2^101; blocks outside that range (all-denormal, huge, non-finite) are stored as zeros. */
int dmxq_bfp_pack(const dmxq_tensor *x, void *mantissas, uint8_t *exponents, int block_size, int precision, void *stream);
int dmxq_bfp_unpack(const void *mantissas, const uint8_t *exponents, const dmxq_tensor *y, int block_size, int precision,
                    void *stream);

/* ---- packed SBFP storage (ScaledBlockFloatingPoint.bytes_per_elem, S/numerical/format.py:481-486;
 * ids DMX_SBFP_12_16_<bias>, S/numerical/onnx.py:53-67).  `fmt`: a DMXQ_STAGE_SBFP stage (the same description
 * dmxq_cast_chain takes).  Per block of fmt->block consecutive elements:
 *   one scaler byte    0 = zero scaler; otherwise (E << sc_man) | M with the scaler FloatingPoint value
 *                      (1 + M / 2^sc_man) * 2^(E - sc_bias), E >= 1 -- the byte of a real E<sc_exp>M<sc_man> number;
 *   fmt->block mantissas, sign-magnitude (top bit = sign of x, kept on zero results exactly as the simulated cast keeps
 *                      it), 4 bits each for precision <= 4 (two per byte, low nibble first), else 8 bits.
 * SBFP12_16: 4 bits per element + one byte per 16 = 0.5625 B/elem (the reference's bytes_per_elem reports 0.5703: it counts
 * a sign bit for the unsigned scaler, S/numerical/format.py:240-241).
 *   x / y: contiguous [rows, K], K % block == 0, block a power of two in 8..128; block format XP[p,0] clamped, nearest
 *   (half away: the reference's CUDA rule), p in 2..8; scaler format flushing subnormals, sc_exp + sc_man <= 8.
 *   dmxq_sbfp_unpack(dmxq_sbfp_pack(x)) == dmxq_sbfp_qdq(x) bit for bit for every block the byte can hold.  It cannot hold:
 *   non-finite blocks and blocks whose max / man_scaling underflows to zero (stored as zeros), and scalers above the
 *   exponent field's range -- the reference's simulated scaler saturates at 2^(2^(sc_exp-1)) whatever the bias, which for
 *   sc_bias > 2^(sc_exp-1) - 1 exceeds what the real byte reaches (stored saturated).  `n_inexact` (nullable, device
 *   counter, ACCUMULATED into) receives the number of such blocks: 0 means the round trip is exact. */
int dmxq_sbfp_pack(const dmxq_tensor *x, void *mantissas, uint8_t *scalers, const dmxq_stage *fmt, unsigned int *n_inexact,
                   void *stream);
int dmxq_sbfp_unpack(const void *mantissas, const uint8_t *scalers, const dmxq_tensor *y, const dmxq_stage *fmt, void *stream);

/* ---- L1 mirror: block_quantize(x, wl, dim, symmetric, rounding) of quant_cuda --------------
 * dim == -1: one exponent for the whole tensor; dim == 0: per row of view(size0,-1);
 * dim == d: per index of dimension d (quant.cu:14-34).  x, y contiguous fp32 (as the
 * reference requires).  `workspace`: device scratch of 3 * C 4-byte words, C = shape[dim]
 * (C = 1 for dim == -1). */
int dmxq_block_quantize(const dmxq_tensor *x, const dmxq_tensor *y, int wl, int dim, int symmetric,
                        int rounding, const int32_t *rand, void *workspace, void *stream);

/* ---- calibration statistics: amin / amax per tensor (ch_axis < 0) or per channel -----------
 * out_min/out_max: device arrays of 1 or shape[ch_axis] floats.  NaN propagates.  Exact and
 * order independent, so a sharded reduction followed by an all-reduce(MIN/MAX) is
 * bit-identical to the single-device result. */
int dmxq_minmax(const dmxq_tensor *x, int ch_axis, float *out_min, float *out_max, void *stream);

/* ---- max|x| of many tensors in one launch: out_amax[i] = max over xs[i] of |x| (fp32; NaN propagates), the per-shard
 * statistic of a sharded calibration pass (a tensor-wide amax for the SBFP scaler range; the symmetric
 * half of MinMaxObserver's statistic, S/numerical/observer.py:181-186).  Exact and order independent, so the per-shard
 * results followed by an all-reduce(MAX) equal the single-device value bit for bit.  xs: contiguous, 16-byte aligned,
 * any shapes, one dtype per launch group (mixed dtypes just cost extra launches).  out_amax: device array of n_tensors
 * floats, overwritten.  Feeds dmxq_cast_chain_multi(amax=...) without leaving the device. */
int dmxq_amax_multi(const dmxq_tensor *xs, int n_tensors, float *out_amax, void *stream);

/* ---- calibration histogram: the torch.histc call of HistogramObserver.forward (S/numerical/observer.py:470-491)
 * counts[b] += #{ v in x : lo <= v <= hi, b == min((int)((v - lo) * bins / (hi - lo)), bins - 1) }, the bin expression
 * evaluated in fp32 exactly as torch does; NaN and out-of-range values are dropped.  `counts`: device array of `bins`
 * 64-bit counters, ACCUMULATED into (zero it first for a fresh histogram; torch's float histogram is float(counts),
 * exact wherever torch's own float accumulation is).  lo < hi, both finite (torch's "min == max -> use the data's
 * range" rule needs the data's range first: the python wrapper resolves it with dmxq_minmax).
 * out_min / out_max (both or neither, device float[1]): also return amin / amax of x from the same pass -- the
 * torch.aminmax(x) every observer step needs -- NaN propagates.  x contiguous, 16-byte aligned, bins <= 12288. */
int dmxq_histc(const dmxq_tensor *x, int bins, float lo, float hi, unsigned long long *counts, float *out_min, float *out_max,
               void *stream);

/* ---- host-buffer entry (the e2e path): x_host / y_host are HOST pointers to contiguous
 * [rows, K] matrices blocked along K (pinned memory gives full PCIe speed).  The call
 * pipelines H2D copy, the chain kernel and D2H copy over internal streams in row chunks and
 * returns when y_host is complete. */
int dmxq_cast_chain_host(const void *x_host, void *y_host, int in_dtype, int out_dtype,
                         int64_t rows, int64_t K, const dmxq_stage *stages, int n_stages,
                         int device);
void *dmxq_host_alloc(int64_t bytes); /* pinned host memory; NULL on failure */
void dmxq_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* DMXQ_H_ */

// dmxq_api.cu -- the C ABI (include/dmxq.h): argument validation, stage decoding, view
// canonicalisation and kernel selection.  Host-only logic; kernels live in dmxq_kernels.cu.
//
// View canonicalisation: the non-blocked dims of a cast are pure batch dims, so they are
// sorted by input stride and merged whenever every tensor of the call (x, y, score, mask,
// rand) is jointly contiguous across the pair.  What remains is classified:
//   rows    blocked dim has stride 1            -> chain_rows_kernel (flat when fully merged)
//   cols    blocked dim strided, another dim 1  -> chain_cols_kernel
//   generic anything else                       -> chain_generic_kernel, one launch per stage
// so e.g. key.transpose(-2,-1) blocked along -2 is recognised as the contiguous case and the
// reference's .contiguous()/transpose copies (Q/quant_function.py:120,
// S/numerical/format.py:322-326) never happen.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/dmxq.h"
#include "dmxq_kernels.cuh"

using namespace dmxq;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(DMXQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

inline int dtype_size(int dt) { return dt == DMXQ_F32 ? 4 : 2; }
inline bool pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

bool dtype_pair_ok(int in_dt, int out_dt)
{
    if (in_dt < 0 || in_dt > 2 || out_dt < 0 || out_dt > 2) return false;
    return in_dt == out_dt || out_dt == DMXQ_F32;
}

// fixed_min_max, Q/quant_cuda/quant.cu:230-237
void fixed_min_max(int wl, int fl, bool symmetric, float *t_min, float *t_max)
{
    int sigma = -fl;
    *t_min = (float)-std::ldexp(1.0, wl - fl - 1);
    *t_max = (float)(-(double)*t_min - std::ldexp(1.0, sigma));
    if (symmetric) *t_min = (float)((double)*t_min + std::ldexp(1.0, sigma));
}

int decode_float(int man, int exp, int bias, int flush, int is_unsigned, int fp16_flush, int rounding, FloatFmt &f)
{
    // FloatingPoint.__init__ asserts, S/numerical/format.py:189-203
    if (man < 0 || man > 23) return fail(DMXQ_ERR_BAD_ARG, "number of mantissa bits simulatable by FP32 is between 0 and 23, got %d", man);
    if (exp < 1 || exp > 8) return fail(DMXQ_ERR_BAD_ARG, "number of exponent bits simulatable by FP32 is between 1 and 8, got %d", exp);
    if (bias > 127 || bias < -127) return fail(DMXQ_ERR_BAD_ARG, "exponent bias %d out of range", bias);
    if (rounding < 0 || rounding > 3) return fail(DMXQ_ERR_BAD_ARG, "invalid rounding mode %d", rounding);
    f.sh = 23 - man;
    f.mode = rounding;
    f.exact = 0;
    if (man >= 23) { f.sh = 0; f.mode = rounding == R_NEAREST ? R_NEAREST : R_DOWN; f.exact = 1; }  // reference UB; identity on its CUDA build
    f.mask = (1u << f.sh) - 1u;
    f.min_exp = -(bias - 1);
    f.shift_exp = (uint32_t)(127 + f.min_exp) << 23;
    f.max_store = (uint32_t)((1 << (exp - 1)) + 127);
    f.max_num = (f.max_store << 23) | ((0x007FFFFFu >> f.sh) << f.sh);
    f.flush = flush != 0;
    f.is_unsigned = is_unsigned != 0;
    f.fp16_flush = fp16_flush != 0;
    // fast path: nearest + flush + signed.  The extra fp16 pass (|q| < 2^-14 -> +0,
    // format.py:223-232) is implied by flush_subnormal whenever the format's own smallest normal
    // is >= 2^-14 (NaN inputs, the one exception, take the exact out-of-line path anyway).
    f.fastpath = f.mode == R_NEAREST && f.flush && !f.is_unsigned && (!f.fp16_flush || f.min_exp >= -14);
    f.nsub = f.mode == R_NEAREST && !f.flush && !f.is_unsigned && !f.fp16_flush && f.sh >= 2;
    f.magic = ((uint32_t)f.sh << 23) | 0x00400000u;
    return DMXQ_OK;
}

int decode_fixed(int wl, int fl, int clamp, int symmetric, int rounding, int tie, FixedFmt &f)
{
    if (wl < 1 || wl > 24) return fail(DMXQ_ERR_BAD_ARG, "highest integer precision simulated by FP32 is 25, got %d", wl);
    if (fl < -126 || fl > 126) return fail(DMXQ_ERR_UNSUPPORTED, "fixed point fraction %d outside [-126, 126]", fl);
    if (rounding < 0 || rounding > 3) return fail(DMXQ_ERR_BAD_ARG, "invalid rounding mode %d", rounding);
    f.up = std::ldexp(1.0f, fl);
    f.down = std::ldexp(1.0f, -fl);
    fixed_min_max(wl, fl, symmetric != 0, &f.t_min, &f.t_max);
    f.clamp = clamp != 0;
    f.mode = rounding;
    f.tie = tie;
    return DMXQ_OK;
}

int decode_stage(const dmxq_stage &s, StageDev &d)
{
    memset(&d, 0, sizeof(d));
    d.kind = s.kind;
    d.block = s.block;
    switch (s.kind) {
    case DMXQ_STAGE_NM:
        if (s.block < 1 || s.n_keep < 1 || s.n_keep > s.block)
            return fail(DMXQ_ERR_BAD_ARG, "N and M must be positive and N no greater than M (got %d:%d)", s.n_keep, s.block);
        if (s.block > 64) return fail(DMXQ_ERR_UNSUPPORTED, "N:M group size %d > 64", s.block);
        d.n_prune = s.block - s.n_keep;
        if (s.nm_order != DMXQ_NM_ORDER_STABLE && s.nm_order != DMXQ_NM_ORDER_TORCH_CUDA) return fail(DMXQ_ERR_BAD_ARG, "invalid nm_order %d", s.nm_order);
        // torch sorts rows of more than 32 keys with a stable sort on CUDA as well (Sort.cu: bitonic only for <= 32)
        d.nm_order = (s.nm_order == DMXQ_NM_ORDER_TORCH_CUDA && s.block <= 32 && s.block >= 2) ? 1 : 0;
        return DMXQ_OK;
    case DMXQ_STAGE_BFP:
        // BlockFloatingPoint.__init__ asserts, S/numerical/format.py:289-292
        if (s.precision < 2 || s.precision > 25) return fail(DMXQ_ERR_BAD_ARG, "highest integer precision simulated by FP32 is 25, got %d", s.precision);
        if (s.block < 1) return fail(DMXQ_ERR_BAD_ARG, "block size has to be positive, got %d", s.block);
        if (s.rounding < 0 || s.rounding > 3) return fail(DMXQ_ERR_BAD_ARG, "invalid rounding mode %d", s.rounding);
        if (s.block == 1) {  // format.py:312-320: borrow float_quantize, man = precision - 2
            d.kind = ST_FLOAT;
            return decode_float(std::min(s.precision - 2, 23), 8, 127, 0, 0, 0, s.rounding, d.ff);
        }
        if (s.precision > 22) return fail(DMXQ_ERR_UNSUPPORTED, "BFP precision %d > 22 with block size > 1 is undefined in the reference (negative shift)", s.precision);
        d.wl = s.precision;
        d.sh = 23 - s.precision;
        d.mask = (1u << d.sh) - 1u;
        d.mode = s.rounding;
        d.asym = s.symmetric ? 0 : 1;
        d.fast = s.precision <= 20;
        return DMXQ_OK;
    case DMXQ_STAGE_SBFP: {
        if (s.block < 1) return fail(DMXQ_ERR_BAD_ARG, "block size has to be positive, got %d", s.block);
        if (s.rounding == DMXQ_ROUND_STOCHASTIC) return fail(DMXQ_ERR_UNSUPPORTED, "stochastic SBFP block format is not supported");
        int rc = decode_fixed(s.precision, 0, s.clamp, 1, s.rounding, s.tie, d.sb.xp);
        if (rc) return rc;
        rc = decode_float(s.sc_man, s.sc_exp, s.sc_bias, s.sc_flush, s.sc_unsigned, s.sc_fp16_flush, s.sc_rounding, d.sb.sc);
        if (rc) return rc;
        if (d.sb.sc.mode == R_STOCHASTIC) return fail(DMXQ_ERR_UNSUPPORTED, "stochastic SBFP scaler format is not supported");
        d.sb.man_scaling = (float)((1 << (s.precision - 1)) - 1);
        d.sb.inv_man = 1.0f / d.sb.man_scaling;
        if (s.scale_mode != DMXQ_SCALE_DIV && s.scale_mode != DMXQ_SCALE_RECIP) return fail(DMXQ_ERR_BAD_ARG, "invalid scale_mode %d", s.scale_mode);
        d.sb_exp_bits = s.sc_exp;
        d.sb.recip = s.scale_mode == DMXQ_SCALE_RECIP;
        d.sb.inv_man_t = (float)(1.0 / (double)d.sb.man_scaling);  // ATen div_true_kernel_cuda: inv_b computed in double, used in fp32
        d.sb.no_clamp = !d.sb.xp.clamp || (d.sb.xp.t_max >= d.sb.man_scaling && d.sb.xp.t_min <= -d.sb.man_scaling);
        d.sb.sc_fast = d.sb.sc.mode == R_NEAREST && d.sb.sc.flush && (!d.sb.sc.fp16_flush || d.sb.sc.min_exp >= -14);
        return DMXQ_OK;
    }
    case DMXQ_STAGE_FLOAT:
        return decode_float(s.man, s.exp, s.bias, s.flush, s.is_unsigned, s.fp16_flush, s.rounding, d.ff);
    case DMXQ_STAGE_FIXED: {
        int rc = decode_fixed(s.precision, s.fraction, s.clamp, s.symmetric, s.rounding, s.tie, d.xf);
        if (rc) return rc;
        d.sc = s.scale;
        d.zp = s.zero_point;
        d.affine = !(s.scale == 1.0f && s.zero_point == 0.0f);
        return DMXQ_OK;
    }
    case DMXQ_STAGE_MXFP: {
        if (s.block < 1) return fail(DMXQ_ERR_BAD_ARG, "block size has to be positive, got %d", s.block);
        if (s.exp < 1 || s.exp > 7) return fail(DMXQ_ERR_UNSUPPORTED, "MXFP element exponent bits must be 1..7, got %d", s.exp);
        int rc = decode_float(s.man, s.exp, (1 << (s.exp - 1)) - 1, 0, 0, 0, DMXQ_ROUND_NEAREST, d.ff);  // format.py:585-592
        if (rc) return rc;
        d.mx_largest = std::ldexp(1.0f, 1 << (s.exp - 1));  // FloatingPoint.largest_representable_power_of_two, format.py:235-237
        return DMXQ_OK;
    }
    case DMXQ_STAGE_SCALE:
        if (!s.vec || s.vec_len < 1) return fail(DMXQ_ERR_BAD_ARG, "scale stage needs a device vector");
        if (s.vec_op != 0 && s.vec_op != 1) return fail(DMXQ_ERR_BAD_ARG, "scale stage: vec_op 0 (divide) or 1 (multiply)");
        if ((reinterpret_cast<uintptr_t>(s.vec) & 15) != 0) return fail(DMXQ_ERR_UNSUPPORTED, "scale stage: 16-byte aligned vector");
        d.vec = s.vec;
        d.vec_op = s.vec_op;
        d.block = s.vec_len;  // (checked against the extent of block_dim by the caller)
        return DMXQ_OK;
    default:
        return fail(DMXQ_ERR_BAD_ARG, "unknown stage kind %d", s.kind);
    }
}

inline bool stage_blocked(const StageDev &d) { return d.kind == ST_NM || d.kind == ST_BFP || d.kind == ST_SBFP || d.kind == ST_MXFP || d.kind == ST_SCALE; }
inline int stage_mode(const StageDev &d) { return d.kind == ST_FLOAT ? d.ff.mode : d.kind == ST_FIXED ? d.xf.mode : d.kind == ST_BFP ? d.mode : 0; }

struct Dim {
    int64_t n, xs, ys, ss, ms, rs;
};

struct Canon {
    std::vector<Dim> outer;  // non-blocked dims, sorted by |xs| descending and merged
    Dim k;                   // the blocked dim (n == 1, strides 1 when the chain is elementwise)
    bool has_k;
};

bool same_shape(const dmxq_tensor *a, const dmxq_tensor *b)
{
    if (a->ndim != b->ndim) return false;
    for (int i = 0; i < a->ndim; ++i) if (a->shape[i] != b->shape[i]) return false;
    return true;
}

void canonicalise(const dmxq_tensor *x, const dmxq_tensor *y, const dmxq_tensor *sc, const dmxq_tensor *mk,
                  bool has_rand, int kd, Canon &c)
{
    int nd = x->ndim;
    std::vector<int64_t> rstride(nd, 1);
    for (int i = nd - 2; i >= 0; --i) rstride[i] = rstride[i + 1] * x->shape[i + 1];
    c.outer.clear();
    c.has_k = kd >= 0;
    c.k = Dim{1, 1, 1, 1, 1, 1};
    for (int i = 0; i < nd; ++i) {
        Dim d{x->shape[i], x->stride[i], y->stride[i], sc ? sc->stride[i] : 0, mk ? mk->stride[i] : 0, has_rand ? rstride[i] : 0};
        if (i == kd) { c.k = d; continue; }
        if (d.n == 1) continue;
        c.outer.push_back(d);
    }
    std::stable_sort(c.outer.begin(), c.outer.end(), [](const Dim &a, const Dim &b) { return a.xs > b.xs; });
    // merge (i, i+1) when jointly contiguous in every tensor
    std::vector<Dim> m;
    for (const Dim &d : c.outer) {
        if (!m.empty()) {
            Dim &p = m.back();
            bool ok = p.xs == d.xs * d.n && p.ys == d.ys * d.n && (!sc || p.ss == d.ss * d.n) && (!mk || p.ms == d.ms * d.n) &&
                      (!has_rand || p.rs == d.rs * d.n);
            if (ok) { p.n *= d.n; p.xs = d.xs; p.ys = d.ys; p.ss = d.ss; p.ms = d.ms; p.rs = d.rs; continue; }
        }
        m.push_back(d);
    }
    c.outer.swap(m);
}

inline bool aligned(const void *p, int bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; }

int run_generic(const dmxq_tensor *x, const dmxq_tensor *y, const Canon &c, const ChainDev &chain, const float *score, float *mask,
                const void *rnd, cudaStream_t st)
{
    // one launch per stage; stage 0 reads x, later stages run in place on y
    for (int s = 0; s < chain.n; ++s) {
        GenericParams p;
        memset(&p, 0, sizeof(p));
        const StageDev &sd = chain.st[s];
        bool first = s == 0;
        p.x = first ? x->data : y->data;
        p.y = y->data;
        p.score = sd.kind == ST_NM ? score : nullptr;
        p.mask = sd.kind == ST_NM ? mask : nullptr;
        p.rnd = stage_mode(sd) == R_STOCHASTIC ? rnd : nullptr;
        p.st = sd;
        std::vector<Dim> dims = c.outer;
        Dim k = c.k;
        if (!stage_blocked(sd)) {
            // elementwise stage: the blocked dim is one more batch dim; keep the smallest stride fastest
            if (c.has_k || k.n > 1) dims.push_back(k);
            std::stable_sort(dims.begin(), dims.end(), [&](const Dim &a, const Dim &b) { return (first ? a.xs : a.ys) > (first ? b.xs : b.ys); });
            k = Dim{1, 1, 1, 1, 1, 1};
        }
        p.K = k.n;
        int64_t B = stage_blocked(sd) ? sd.block : 1;
        p.nblk = (p.K + B - 1) / B;
        p.xks = first ? k.xs : k.ys; p.yks = k.ys; p.sks = k.ss; p.mks = k.ms; p.rks = k.rs;
        p.nd = (int)dims.size();
        if (p.nd > kMaxDims) return fail(DMXQ_ERR_UNSUPPORTED, "too many dims");
        int64_t items = p.nblk;
        for (int i = 0; i < p.nd; ++i) {
            const Dim &d = dims[i];
            p.dim[i] = d.n; p.xs[i] = first ? d.xs : d.ys; p.ys[i] = d.ys; p.ss[i] = d.ss; p.ms[i] = d.ms; p.rs[i] = d.rs;
            items *= d.n;
        }
        p.n_items = items;
        int in_dt = first ? x->dtype : y->dtype;
        cudaError_t e = launch_generic(in_dt, y->dtype, p, st);
        if (e != cudaSuccess) return cuda_fail(e, "chain_generic_kernel");
    }
    return DMXQ_OK;
}

// the stages of one fused chain as the kernels consume them: decoded formats plus what depends on the position in the chain
// (significant bits of the values a stage sees, rounding to the tensor dtype between consecutive CastTo.forward calls)
int decode_chain(int x_dtype, int y_dtype, const dmxq_stage *stages, int n_stages, ChainDev &chain, bool *blocked_out, int *n_stoch_out)
{
    memset(&chain, 0, sizeof(chain));
    chain.n = n_stages;
    bool blocked = false;
    int n_stoch = 0;
    for (int s = 0; s < n_stages; ++s) {
        int rc = decode_stage(stages[s], chain.st[s]);
        if (rc) return rc;
        if (chain.st[s].kind == ST_BFP) {
            // significant bits of the values this stage sees: the source dtype's for stage 0 (and
            // after an N:M stage, which only zeroes elements), the output dtype's after a requant
            int src = -1;
            if (s == 0 || (s == 1 && chain.st[0].kind == ST_NM)) src = x_dtype;
            else if (chain.st[s - 1].requant) src = y_dtype;
            chain.st[s].fast16 = (src == DMXQ_BF16 && chain.st[s].wl <= 14) || (src == DMXQ_F16 && chain.st[s].wl <= 11);
        }
        if (s == 0 && chain.st[s].kind == ST_FLOAT && chain.st[s].ff.mode == R_NEAREST) {
            // a bf16 (7 mantissa bits) / fp16 (10) source is already representable: rounding is the identity
            int src_man = x_dtype == DMXQ_BF16 ? 7 : x_dtype == DMXQ_F16 ? 10 : 23;
            if (23 - chain.st[s].ff.sh >= src_man) chain.st[s].ff.exact = 1;
        }
        blocked |= stage_blocked(chain.st[s]);
        n_stoch += stage_mode(chain.st[s]) == R_STOCHASTIC;
        // consecutive CastTo.forward calls round to the tensor dtype in between (cast.py:306)
        chain.st[s].requant = (s + 1 < n_stages && y_dtype != DMXQ_F32 && chain.st[s].kind != ST_NM) ? 1 : 0;
    }
    if (blocked_out) *blocked_out = blocked;
    if (n_stoch_out) *n_stoch_out = n_stoch;
    return DMXQ_OK;
}

// qscale / qzp (nullable): device-resident per-tensor FixedPoint affine parameters; only the vectorised rows
// kernel consumes them -- when the layout needs another path the call returns kNeedFallback (no message) and
// dmxq_fixed_qdq falls back to fixed_chan_kernel.
constexpr int kNeedFallback = -100;

// what chain_impl decided for a tensor that the many-tensor launch can take over (flat rows layout, same dtype in and out)
struct RowsPlan {
    bool planned = false;
    int kind = -1, dt = -1;
    RowsParams p;
};

// In-kernel random words (dmxq_cast_chain_philox): the call travels through chain_impl with this tag in place of a random
// tensor, so every layout decision is the one an external tensor in logical element order would get; the rows kernels then
// compute the words instead of loading them.  Layouts that need the cols / generic kernels return DMXQ_ERR_UNSUPPORTED (the
// binding fills a tensor with dmxq_philox_fill -- the same stream -- and passes it explicitly).
alignas(16) static const char kPhiloxTag[16] = {0};
static thread_local unsigned long long g_ph_seed = 0, g_ph_stream = 0;

int chain_impl(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, const dmxq_stage *stages, int n_stages,
               const dmxq_tensor *score, const dmxq_tensor *mask, const void *rand, cudaStream_t st,
               const float *qscale = nullptr, const float *qzp = nullptr, const float *amax = nullptr, RowsPlan *plan = nullptr)
{
    if (!x || !y || !stages) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (n_stages < 1 || n_stages > DMXQ_MAX_STAGES) return fail(DMXQ_ERR_BAD_ARG, "n_stages must be 1..%d, got %d", DMXQ_MAX_STAGES, n_stages);
    if (x->ndim < 0 || x->ndim > DMXQ_MAX_DIMS) return fail(DMXQ_ERR_BAD_ARG, "ndim %d out of range", x->ndim);
    if (!same_shape(x, y)) return fail(DMXQ_ERR_BAD_ARG, "x and y must have the same shape");
    if (!dtype_pair_ok(x->dtype, y->dtype)) return fail(DMXQ_ERR_UNSUPPORTED, "unsupported dtype pair in=%d out=%d", x->dtype, y->dtype);
    if (score && (!same_shape(x, score) || score->dtype != DMXQ_F32)) return fail(DMXQ_ERR_BAD_ARG, "score must be fp32 with the shape of x");
    if (mask && (!same_shape(x, mask) || mask->dtype != DMXQ_F32)) return fail(DMXQ_ERR_BAD_ARG, "mask must be fp32 with the shape of x");

    ChainDev chain;
    bool blocked = false;
    int n_stoch = 0;
    {
        int rc = decode_chain(x->dtype, y->dtype, stages, n_stages, chain, &blocked, &n_stoch);
        if (rc) return rc;
    }
    if (n_stoch > 1) return fail(DMXQ_ERR_UNSUPPORTED, "at most one stochastic stage per chain");
    if (n_stoch == 1 && !rand) return fail(DMXQ_ERR_BAD_ARG, "stochastic rounding needs a random tensor");
    if (n_stoch == 0) rand = nullptr;

    int64_t numel = 1;
    for (int i = 0; i < x->ndim; ++i) {
        if (x->shape[i] < 0) return fail(DMXQ_ERR_BAD_ARG, "negative extent");
        numel *= x->shape[i];
    }
    int kd = -1;
    if (blocked) {
        if (x->ndim == 0) return fail(DMXQ_ERR_BAD_ARG, "blocked format on a 0-d tensor");
        kd = block_dim < 0 ? block_dim + x->ndim : block_dim;
        if (kd < 0 || kd >= x->ndim) return fail(DMXQ_ERR_BAD_ARG, "block_dim %d out of range for %d dims", block_dim, x->ndim);
        for (int s = 0; s < n_stages; ++s)
            if (chain.st[s].kind == ST_SCALE) {
                if (chain.st[s].block != x->shape[kd])
                    return fail(DMXQ_ERR_BAD_ARG, "scale vector of %d entries along a dim of extent %lld", chain.st[s].block, (long long)x->shape[kd]);
            }
        for (int s = 0; s < n_stages; ++s)
            if (chain.st[s].kind == ST_NM && x->shape[kd] % chain.st[s].block != 0)
                return fail(DMXQ_ERR_BAD_ARG, "score has size %lld at dimension %d, not a multiple of block size %d",
                            (long long)x->shape[kd], block_dim, chain.st[s].block);
    }
    if (numel == 0) return DMXQ_OK;
    if (!x->data || !y->data) return fail(DMXQ_ERR_BAD_ARG, "null data pointer");

    Canon c;
    canonicalise(x, y, score, mask, rand != nullptr, kd, c);
    const int in_sz = dtype_size(x->dtype), out_sz = dtype_size(y->dtype);
    const int V = 16 / in_sz;
    const float *score_p = score ? static_cast<const float *>(score->data) : nullptr;
    float *mask_p = mask ? static_cast<float *>(mask->data) : nullptr;

    if (!blocked) {
        // elementwise chain: use the stride-1 dim (if any) as the vector dim
        int vd = -1;
        for (int i = (int)c.outer.size() - 1; i >= 0; --i)
            if (c.outer[i].xs == 1 && c.outer[i].ys == 1) { vd = i; break; }
        if (vd >= 0) {
            c.k = c.outer[vd];
            c.outer.erase(c.outer.begin() + vd);
            c.has_k = true;
        }
    }

    // ---------------------------------------------------------------- rows path
    bool rows_ok = c.k.n >= 1 && (c.k.n == 1 || (c.k.xs == 1 && c.k.ys == 1)) && (int)c.outer.size() <= kMaxOuter &&
                   c.k.n % V == 0 && aligned(x->data, 16) && aligned(y->data, 16);
    if (rows_ok && score_p) rows_ok = c.k.ss == 1 && aligned(score_p, 16);
    if (rows_ok && mask_p) rows_ok = c.k.ms == 1 && aligned(mask_p, 16);
    int64_t tile = V;
    if (rows_ok) {
        for (const Dim &d : c.outer) {
            rows_ok &= (d.xs * in_sz) % 16 == 0 && (d.ys * out_sz) % 16 == 0;
            if (score_p) rows_ok &= d.ss % 4 == 0;
            if (mask_p) rows_ok &= d.ms % 4 == 0;
        }
        for (int s = 0; s < chain.n && rows_ok; ++s) {
            const StageDev &sd = chain.st[s];
            if (sd.kind == ST_BFP || sd.kind == ST_SBFP || sd.kind == ST_MXFP) {
                rows_ok &= sd.block % V == 0 && pow2(sd.block / V) && sd.block / V <= 32;
                tile = std::max<int64_t>(tile, sd.block);
            } else if (sd.kind == ST_NM) {
                int M = sd.block;
                rows_ok &= pow2(M) && M >= 2 && (M <= V ? true : (M / V <= 4));
                if (sd.nm_order) rows_ok &= M <= V && M <= 8;  // the network emulation is in-thread; other group sizes: generic kernel
                tile = std::max<int64_t>(tile, M);
            }
        }
    }
    if (rows_ok) {
        RowsParams p;
        memset(&p, 0, sizeof(p));
        p.x = x->data; p.y = y->data; p.score = score_p; p.mask = mask_p; p.rnd = rand;
        if (rand == kPhiloxTag) { p.rnd = nullptr; p.philox = 1; p.ph_seed = g_ph_seed; p.ph_stream = g_ph_stream; }
        p.K = c.k.n;
        p.kvec = (uint32_t)(p.K / V);
        int64_t tiles_per_row = (p.K + tile - 1) / tile;
        int64_t vpr = tiles_per_row * (tile / V);
        if (vpr > 0x7FFFFFFFll) rows_ok = false;
        p.vpr = (uint32_t)vpr;
        int64_t rows = 1;
        p.nouter = std::max<int>(1, (int)c.outer.size());
        p.odim[0] = 1;
        for (size_t i = 0; i < c.outer.size(); ++i) {
            const Dim &d = c.outer[i];
            p.odim[i] = d.n; p.xs[i] = d.xs; p.ys[i] = d.ys; p.ss[i] = d.ss; p.ms[i] = d.ms; p.rs[i] = d.rs;
            rows *= d.n;
        }
        p.rows = rows;
        p.rks = c.k.rs;
        p.chain = chain;
        p.vpr_div = make_fastdiv(p.vpr);
        for (int i = 0; i < p.nouter; ++i) p.odim_div[i] = make_fastdiv((uint32_t)std::min<int64_t>(std::max<int64_t>(p.odim[i], 1), 0x7FFFFFFF));
        bool flat = c.outer.size() <= 1 && p.K % tile == 0;
        if (flat && c.outer.size() == 1) {
            const Dim &d = c.outer[0];
            flat = d.xs == p.K && d.ys == p.K && (!score_p || d.ss == p.K) && (!mask_p || d.ms == p.K) && (!rand || (d.rs == p.K && c.k.rs == 1));
        } else if (flat && rand) {
            flat = c.k.rs == 1;
        }
        p.n_vec = flat ? rows * p.K / V : rows * vpr;
        // kernel specialisation (K_* of dmxq_rows.cuh)
        auto bfp_ns = [](const StageDev &d) { return d.kind == ST_BFP && d.mode == R_NEAREST && !d.asym; };
        auto float_fast = [](const StageDev &d) { return d.kind == ST_FLOAT && d.ff.fastpath; };
        int kind = 1;  // K_CHAIN
        const StageDev &s0 = chain.st[0];
        const bool stoch1 = (s0.kind == ST_BFP && s0.mode == R_STOCHASTIC && !s0.asym) || (s0.kind == ST_FLOAT && s0.ff.mode == R_STOCHASTIC) ||
                            (s0.kind == ST_FIXED && s0.xf.mode == R_STOCHASTIC && !s0.affine && !qscale);
        if (rand && !score_p && !mask_p && flat && chain.n == 1 && stoch1 && aligned(rand, 16)) kind = 11;  // K_BFP_STOCH
        else if (score_p || mask_p || rand) kind = 0;  // K_AUX
        else if (chain.n == 1 && bfp_ns(chain.st[0])) kind = 2;  // K_BFP
        else if (chain.n == 1 && chain.st[0].kind == ST_FLOAT && chain.st[0].ff.mode == R_NEAREST) kind = 3;  // K_FLOAT
        else if (chain.n == 2 && float_fast(chain.st[0]) && bfp_ns(chain.st[1])) kind = 4;  // K_FLOAT_BFP
        else if (chain.n == 2 && chain.st[0].kind == ST_NM && bfp_ns(chain.st[1])) kind = 5;  // K_NM_BFP
        else if (chain.n == 1 && chain.st[0].kind == ST_SBFP && chain.st[0].sb.xp.mode == R_NEAREST && chain.st[0].sb.xp.tie == TIE_AWAY) kind = 6;  // K_SBFP
        else if (chain.n == 1 && chain.st[0].kind == ST_FIXED && chain.st[0].xf.mode == R_NEAREST && chain.st[0].xf.tie == TIE_AWAY) kind = 7;  // K_FIXED
        else if (chain.n == 1 && chain.st[0].kind == ST_NM) kind = 8;  // K_NM
        else if (chain.n == 1 && chain.st[0].kind == ST_MXFP) kind = 9;  // K_MXFP
        else if (chain.n == 1 && chain.st[0].kind == ST_BFP && chain.st[0].mode == R_NEAREST && chain.st[0].asym) kind = 10;  // K_BFP_ASYM
        // 2:4 -> BFP on a bf16 / fp16 tensor (whole-model weight casts): the straight-line specialisation
        if (kind == 5 && x->dtype == y->dtype && x->dtype != DMXQ_F32 && chain.st[0].block == 4 && chain.st[0].n_prune == 2 &&
            chain.st[1].fast && chain.st[1].fast16) kind = 12;  // K_NM24_BFP
        for (int s = 0; s < chain.n; ++s)
            if (chain.st[s].kind == ST_NM && chain.st[s].nm_order && kind != 0) kind = 1;  // torch tie order: runtime chain (nm_stage)
        if (qscale) {
            if (kind != 7) return kNeedFallback;
            p.qscale = qscale; p.qzp = qzp;
        }
        const int special = kind;
        bool amax_sbfp = false;
        if (amax) {
            for (int s = 0; s < chain.n; ++s) amax_sbfp |= chain.st[s].kind == ST_SBFP;
            if (amax_sbfp && !(kind == 6 && chain.st[0].sb.sc_fast && plan && flat && x->dtype == y->dtype))
                return fail(DMXQ_ERR_UNSUPPORTED, "a device-resident amax drives the SBFP scaler bias only for a single SBFP stage (nearest, half away, "
                                                  "flushing scaler format) on a contiguous tensor with whole blocks and the same dtype in and out");
        }
        if (rows_ok && plan && flat && x->dtype == y->dtype && rows_multi_supported(special)) {
            plan->planned = true; plan->kind = special; plan->dt = x->dtype; plan->p = p;
            return DMXQ_OK;
        }
        if (rows_ok) {
            cudaError_t e = launch_rows(x->dtype, y->dtype, flat, special, p, st);
            if (e != cudaSuccess) return cuda_fail(e, "chain_rows_kernel");
            return DMXQ_OK;
        }
    }

    if (qscale) return kNeedFallback;
    for (int s = 0; s < chain.n; ++s)
        if (chain.st[s].kind == ST_SCALE)
            return fail(DMXQ_ERR_UNSUPPORTED, "a scale stage needs the rows layout (channel dim contiguous, 16-byte aligned, whole vectors)");
    if (rand == kPhiloxTag)
        return fail(DMXQ_ERR_UNSUPPORTED, "in-kernel random words need the rows layout (blocked dim contiguous, 16-byte aligned, whole vectors); "
                                          "pass a tensor filled by dmxq_philox_fill instead");
    if (amax)
        for (int s = 0; s < chain.n; ++s)
            if (chain.st[s].kind == ST_SBFP)
                return fail(DMXQ_ERR_UNSUPPORTED, "a device-resident amax needs the rows layout (blocked dim contiguous, 16-byte aligned, whole blocks)");

    // ---------------------------------------------------------------- cols path
    if (blocked && !score_p && !mask_p && c.k.n > 1) {
        int B = 0;
        bool ok = true;
        for (int s = 0; s < chain.n; ++s) {
            const StageDev &sd = chain.st[s];
            if (sd.kind == ST_NM) ok = false;
            if (sd.kind == ST_BFP || sd.kind == ST_SBFP || sd.kind == ST_MXFP) {
                if (B == 0) B = sd.block; else if (B != sd.block) ok = false;
            }
        }
        int vd = -1;
        for (int i = (int)c.outer.size() - 1; i >= 0 && ok; --i)
            if (c.outer[i].xs == 1 && c.outer[i].ys == 1 && (!rand || true)) { vd = i; break; }
        ok = ok && vd >= 0 && B > 0 && cols_supported(x->dtype, B) && (int)c.outer.size() - 1 <= kMaxOuter;
        if (ok) {
            const Dim in = c.outer[vd];
            // 16-byte vectors of V input elements per lane; outputs are V * out_sz >= 16 bytes
            ok = in.n % V == 0 && aligned(x->data, 16) && aligned(y->data, 16) && (c.k.xs * in_sz) % 16 == 0 &&
                 (c.k.ys * out_sz) % 16 == 0;
            for (size_t i = 0; i < c.outer.size() && ok; ++i)
                if ((int)i != vd) ok = (c.outer[i].xs * in_sz) % 16 == 0 && (c.outer[i].ys * out_sz) % 16 == 0;
            if (ok) {
                ColsParams p;
                memset(&p, 0, sizeof(p));
                p.x = x->data; p.y = y->data; p.rnd = rand;
                p.K = c.k.n; p.inner = in.n;
                p.nblk = (p.K + B - 1) / B;
                int ti = cols_tile_inner(x->dtype, B);
                p.nchunk = (in.n + ti - 1) / ti;
                p.xks = c.k.xs; p.yks = c.k.ys; p.rks = c.k.rs; p.ris = in.rs;
                int64_t outer = 1;
                int no = 0;
                for (size_t i = 0; i < c.outer.size(); ++i) {
                    if ((int)i == vd) continue;
                    const Dim &d = c.outer[i];
                    p.odim[no] = d.n; p.xs[no] = d.xs; p.ys[no] = d.ys; p.rs[no] = d.rs;
                    outer *= d.n;
                    ++no;
                }
                p.nouter = no;
                p.n_tiles = outer * p.nblk * p.nchunk;
                p.chain = chain;
                cudaError_t e = launch_cols(x->dtype, y->dtype, B, p, st);
                if (e != cudaSuccess) return cuda_fail(e, "chain_cols_kernel");
                return DMXQ_OK;
            }
        }
    }

    // ---------------------------------------------------------------- generic path
    if (!c.has_k) c.k = Dim{1, 1, 1, 1, 1, 1};
    return run_generic(x, y, c, chain, score_p, mask_p, rand, st);
}

dmxq_tensor flat_view(void *data, int dtype, int64_t rows, int64_t K)
{
    dmxq_tensor t;
    memset(&t, 0, sizeof(t));
    t.data = data; t.dtype = dtype; t.ndim = 2;
    t.shape[0] = rows; t.shape[1] = K; t.stride[0] = K; t.stride[1] = 1;
    return t;
}

// pipelined host path ---------------------------------------------------------------------------
struct HostCtx {
    int device = -1;
    static constexpr int NB = 3;
    cudaStream_t stream[NB] = {nullptr, nullptr, nullptr};
    void *din[NB] = {nullptr, nullptr, nullptr};
    void *dout[NB] = {nullptr, nullptr, nullptr};
    size_t cap_in = 0, cap_out = 0;
};
std::mutex g_host_mu;
HostCtx g_host;

}  // namespace

extern "C" {

int dmxq_abi_version(void) { return DMXQ_ABI_VERSION; }
const char *dmxq_last_error(void) { return g_err; }
int64_t dmxq_launch_count(void) { return dmxq::launch_count(); }

const char *dmxq_status_string(int status)
{
    switch (status) {
    case DMXQ_OK: return "ok";
    case DMXQ_ERR_BAD_ARG: return "bad argument";
    case DMXQ_ERR_UNSUPPORTED: return "unsupported";
    case DMXQ_ERR_CUDA: return "CUDA error";
    case DMXQ_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
    }
}

int dmxq_cast_chain(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, const dmxq_stage *stages, int n_stages,
                    const dmxq_tensor *score, const dmxq_tensor *mask, const void *rand, void *stream)
{
    return chain_impl(x, y, block_dim, stages, n_stages, score, mask, rand, static_cast<cudaStream_t>(stream));
}

int dmxq_cast_chain_philox(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, const dmxq_stage *stages, int n_stages,
                           uint64_t seed, uint64_t stream_id, void *stream)
{
    g_ph_seed = seed;
    g_ph_stream = stream_id;
    return chain_impl(x, y, block_dim, stages, n_stages, nullptr, nullptr, kPhiloxTag, static_cast<cudaStream_t>(stream));
}

int dmxq_philox_fill(void *out, int64_t n, int as_float, uint64_t seed, uint64_t stream_id, void *stream)
{
    if (n < 0 || (n > 0 && !out)) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (n == 0) return DMXQ_OK;
    if (!aligned(out, 16)) return fail(DMXQ_ERR_UNSUPPORTED, "philox_fill: 16-byte aligned output");
    cudaError_t e = launch_philox_fill(out, n, as_float, seed, stream_id, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "philox_fill_kernel");
    return DMXQ_OK;
}

int dmxq_cast_chain_multi(const dmxq_tensor *xs, const dmxq_tensor *ys, int n_tensors, int block_dim, const dmxq_stage *stages,
                          int n_stages, const float *amax, void *stream)
{
    if (n_tensors < 0 || (n_tensors > 0 && (!xs || !ys))) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MultiTable t;
    RowsParams p0;
    int kind0 = -1, dt0 = -1;
    t.n = 0; t.cta0[0] = 0; t.amax = amax;
    auto flush = [&]() -> int {
        if (t.n == 0) return DMXQ_OK;
        cudaError_t e = launch_rows_multi(dt0, kind0, p0, t, st);
        t.n = 0; t.cta0[0] = 0;
        return e == cudaSuccess ? DMXQ_OK : cuda_fail(e, "chain_rows_multi_kernel");
    };
    for (int i = 0; i < n_tensors; ++i) {
        RowsPlan pl;
        int rc = chain_impl(&xs[i], &ys[i], block_dim, stages, n_stages, nullptr, nullptr, nullptr, st, nullptr, nullptr,
                            amax ? amax + i : nullptr, &pl);
        if (rc) { flush(); return rc; }
        if (!pl.planned) continue;  // empty tensor, or a layout the rows kernel does not take: already launched on its own
        const int64_t per_cta = (int64_t)256 * 4;  // kThreads * kUnroll of dmxq_stages.cuh
        const int64_t ctas = (pl.p.n_vec + per_cta - 1) / per_cta;
        if (ctas > 0x3FFFFFFFll) {  // a tensor this large gets its own launch
            cudaError_t e = launch_rows(pl.dt, pl.dt, true, pl.kind, pl.p, st);
            if (e != cudaSuccess) { flush(); return cuda_fail(e, "chain_rows_kernel"); }
            continue;
        }
        if (t.n > 0 && (pl.kind != kind0 || pl.dt != dt0 || t.n == kMultiMax || (int64_t)t.cta0[t.n] + ctas > 0x7FFFFFFFll)) {
            rc = flush();
            if (rc) return rc;
        }
        if (t.n == 0) { p0 = pl.p; kind0 = pl.kind; dt0 = pl.dt; }
        t.x[t.n] = pl.p.x; t.y[t.n] = pl.p.y; t.n_vec[t.n] = pl.p.n_vec; t.slot[t.n] = i;
        t.cta0[t.n + 1] = t.cta0[t.n] + (uint32_t)ctas;
        ++t.n;
    }
    return flush();
}

int dmxq_amax_multi(const dmxq_tensor *xs, int n_tensors, float *out_amax, void *stream)
{
    if (n_tensors < 0 || (n_tensors > 0 && (!xs || !out_amax))) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (n_tensors == 0) return DMXQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out_amax, 0, sizeof(float) * (size_t)n_tensors, st);  // max over |x| bit patterns starts at +0
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    MultiTable t;
    t.n = 0; t.cta0[0] = 0; t.amax = nullptr;
    int dt0 = -1;
    auto flush = [&]() -> int {
        if (t.n == 0) return DMXQ_OK;
        cudaError_t e2 = launch_amax_multi(dt0, t, out_amax, st);
        t.n = 0; t.cta0[0] = 0;
        return e2 == cudaSuccess ? DMXQ_OK : cuda_fail(e2, "amax_multi_kernel");
    };
    for (int i = 0; i < n_tensors; ++i) {
        const dmxq_tensor &x = xs[i];
        if (x.dtype < 0 || x.dtype > 2) { flush(); return fail(DMXQ_ERR_BAD_ARG, "bad dtype"); }
        int64_t n = 1, expect = 1;
        for (int d = x.ndim - 1; d >= 0; --d) {
            if (x.shape[d] != 1 && x.stride[d] != expect) { flush(); return fail(DMXQ_ERR_UNSUPPORTED, "dmxq_amax_multi needs contiguous tensors"); }
            expect *= x.shape[d];
            n *= x.shape[d];
        }
        if (n == 0) continue;
        if (!x.data || !aligned(x.data, 16)) { flush(); return fail(DMXQ_ERR_UNSUPPORTED, "dmxq_amax_multi needs 16-byte aligned data"); }
        const int64_t ctas = amax_multi_ctas(x.dtype, n);
        if (t.n > 0 && (x.dtype != dt0 || t.n == kMultiMax || (int64_t)t.cta0[t.n] + ctas > 0x7FFFFFFFll)) {
            int rc = flush();
            if (rc) return rc;
        }
        if (ctas > 0x7FFFFFFFll) return fail(DMXQ_ERR_UNSUPPORTED, "tensor too large");
        dt0 = x.dtype;
        t.x[t.n] = x.data; t.y[t.n] = nullptr; t.n_vec[t.n] = n /* elements */; t.slot[t.n] = i;
        t.cta0[t.n + 1] = t.cta0[t.n] + (uint32_t)ctas;
        ++t.n;
    }
    return flush();
}

int dmxq_bfp_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, int block_size, int precision, int symmetric,
                 int rounding, const int32_t *rand, void *stream)
{
    dmxq_stage s;
    memset(&s, 0, sizeof(s));
    s.kind = DMXQ_STAGE_BFP; s.block = block_size; s.precision = precision; s.symmetric = symmetric; s.rounding = rounding;
    return chain_impl(x, y, block_dim, &s, 1, nullptr, nullptr, rand, static_cast<cudaStream_t>(stream));
}

int dmxq_sbfp_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int block_dim, int block_size, int xp_precision, int xp_clamp,
                  int xp_rounding, int xp_tie, int sc_man, int sc_exp, int sc_bias, int sc_flush, int sc_unsigned,
                  int sc_fp16_flush, int sc_rounding, int scale_mode, void *stream)
{
    dmxq_stage s;
    memset(&s, 0, sizeof(s));
    s.kind = DMXQ_STAGE_SBFP; s.block = block_size; s.precision = xp_precision; s.clamp = xp_clamp; s.rounding = xp_rounding;
    s.tie = xp_tie; s.sc_man = sc_man; s.sc_exp = sc_exp; s.sc_bias = sc_bias; s.sc_flush = sc_flush; s.sc_unsigned = sc_unsigned;
    s.sc_fp16_flush = sc_fp16_flush; s.sc_rounding = sc_rounding; s.scale_mode = scale_mode;
    return chain_impl(x, y, block_dim, &s, 1, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int dmxq_float_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int man, int exp, int bias, int flush_subnormal, int is_unsigned,
                   int fp16_flush, int rounding, const int32_t *rand, void *stream)
{
    dmxq_stage s;
    memset(&s, 0, sizeof(s));
    s.kind = DMXQ_STAGE_FLOAT; s.man = man; s.exp = exp; s.bias = bias; s.flush = flush_subnormal; s.is_unsigned = is_unsigned;
    s.fp16_flush = fp16_flush; s.rounding = rounding;
    return chain_impl(x, y, -1, &s, 1, nullptr, nullptr, rand, static_cast<cudaStream_t>(stream));
}

int dmxq_fixed_qdq(const dmxq_tensor *x, const dmxq_tensor *y, int wl, int fl, int clamp, int symmetric, int rounding, int tie,
                   const float *scale, const float *zero_point, int64_t n_qparams, int ch_axis, int64_t group_size,
                   const float *rand, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!x || !y) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (scale == nullptr || zero_point == nullptr) {
        dmxq_stage s;
        memset(&s, 0, sizeof(s));
        s.kind = DMXQ_STAGE_FIXED; s.precision = wl; s.fraction = fl; s.clamp = clamp; s.symmetric = symmetric;
        s.rounding = rounding; s.tie = tie; s.scale = 1.0f; s.zero_point = 0.0f;
        return chain_impl(x, y, -1, &s, 1, nullptr, nullptr, rand, st);
    }
    if (n_qparams == 1 && rounding == DMXQ_ROUND_NEAREST && tie == DMXQ_TIE_AWAY) {
        // per-tensor parameters: the vectorised rows kernel reads them from device memory
        dmxq_stage s;
        memset(&s, 0, sizeof(s));
        s.kind = DMXQ_STAGE_FIXED; s.precision = wl; s.fraction = fl; s.clamp = clamp; s.symmetric = symmetric;
        s.rounding = rounding; s.tie = tie; s.scale = 1.0f; s.zero_point = 0.0f;
        int rc = chain_impl(x, y, -1, &s, 1, nullptr, nullptr, nullptr, st, scale, zero_point);
        if (rc != kNeedFallback) return rc;
    }
    // device-resident qparams: contiguous tensors only (what observers produce them for)
    if (!same_shape(x, y)) return fail(DMXQ_ERR_BAD_ARG, "x and y must have the same shape");
    if (!dtype_pair_ok(x->dtype, y->dtype)) return fail(DMXQ_ERR_UNSUPPORTED, "unsupported dtype pair");
    if (rounding == DMXQ_ROUND_STOCHASTIC && !rand) return fail(DMXQ_ERR_BAD_ARG, "stochastic rounding needs a random tensor");
    int64_t n = 1, expect = 1;
    for (int i = x->ndim - 1; i >= 0; --i) {
        if (x->shape[i] != 1 && (x->stride[i] != expect || y->stride[i] != expect))
            return fail(DMXQ_ERR_UNSUPPORTED, "affine fixed-point cast with device qparams needs contiguous x and y");
        expect *= x->shape[i];
        n *= x->shape[i];
    }
    FixedChanParams p;
    memset(&p, 0, sizeof(p));
    int rc = decode_fixed(wl, fl, clamp, symmetric, rounding, tie, p.xf);
    if (rc) return rc;
    if (n_qparams < 1) return fail(DMXQ_ERR_BAD_ARG, "n_qparams must be >= 1");
    p.x = x->data; p.y = y->data; p.scale = scale; p.zp = zero_point; p.rnd = rounding == DMXQ_ROUND_STOCHASTIC ? rand : nullptr;
    p.n = n; p.nq = n_qparams;
    if (n_qparams == 1) {
        p.C = 1; p.inner = std::max<int64_t>(n, 1); p.group = 1;
    } else {
        int ax = ch_axis < 0 ? ch_axis + x->ndim : ch_axis;
        if (ax < 0 || ax >= x->ndim) return fail(DMXQ_ERR_BAD_ARG, "ch_axis %d out of range", ch_axis);
        p.C = x->shape[ax];
        p.inner = 1;
        for (int i = ax + 1; i < x->ndim; ++i) p.inner *= x->shape[i];
        p.group = group_size > 0 ? group_size : 1;
        if ((p.C + p.group - 1) / p.group > n_qparams) return fail(DMXQ_ERR_BAD_ARG, "%lld qparams do not cover %lld channels in groups of %lld", (long long)n_qparams, (long long)p.C, (long long)p.group);
    }
    if (n == 0) return DMXQ_OK;
    cudaError_t e = launch_fixed_chan(x->dtype, y->dtype, p, st);
    if (e != cudaSuccess) return cuda_fail(e, "fixed_chan_kernel");
    return DMXQ_OK;
}

int dmxq_nm_prune(const dmxq_tensor *x, const dmxq_tensor *score, const dmxq_tensor *y, const dmxq_tensor *mask, int block_dim,
                  int n_keep, int m, int nm_order, void *stream)
{
    dmxq_stage s;
    memset(&s, 0, sizeof(s));
    s.kind = DMXQ_STAGE_NM; s.block = m; s.n_keep = n_keep; s.nm_order = nm_order;
    return chain_impl(x, y, block_dim, &s, 1, score, mask, nullptr, static_cast<cudaStream_t>(stream));
}

int dmxq_add_cast(const dmxq_tensor *a, const dmxq_tensor *b, const dmxq_tensor *y, const dmxq_stage *stage_a,
                  const dmxq_stage *stage_b, const dmxq_stage *stage_out, void *stream)
{
    if (!a || !b || !y) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (!same_shape(a, y) || a->dtype != y->dtype || a->dtype != b->dtype) return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: a, b, y must share a dtype and a, y a shape");
    if (a->dtype < 0 || a->dtype > 2 || b->ndim > a->ndim) return fail(DMXQ_ERR_BAD_ARG, "add_cast: bad dtype / rank");
    AddParams p;
    memset(&p, 0, sizeof(p));
    const dmxq_stage *sts[3] = {stage_a, stage_b, stage_out};
    FloatFmt *fmts[3] = {&p.fa, &p.fb, &p.fo};
    int *has[3] = {&p.has_a, &p.has_b, &p.has_o};
    for (int i = 0; i < 3; ++i) {
        if (!sts[i]) continue;
        StageDev d;
        int rc = decode_stage(*sts[i], d);
        if (rc) return rc;
        if (d.kind != ST_FLOAT || !d.ff.fastpath) return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: only nearest+flush FLOAT stages fuse");
        // every value a stage sees here is representable in the tensor dtype (inputs, and the sum after its
        // rounding to the dtype): with >= that many mantissa bits in the format, rounding is the identity
        const int src_man = a->dtype == DMXQ_BF16 ? 7 : a->dtype == DMXQ_F16 ? 10 : 23;
        if (23 - d.ff.sh >= src_man) d.ff.exact = 1;
        *fmts[i] = d.ff;
        *has[i] = 1;
    }
    const int nd = a->ndim, V = 16 / dtype_size(a->dtype);
    int64_t n = 1, expect = 1;
    for (int i = nd - 1; i >= 0; --i) {
        if (a->shape[i] != 1 && (a->stride[i] != expect || y->stride[i] != expect)) return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: a and y must be contiguous");
        expect *= a->shape[i];
        n *= a->shape[i];
    }
    if (n == 0) return DMXQ_OK;
    // b right-aligned against a; broadcast dims get stride 0
    std::vector<int64_t> bst(nd, 0);
    for (int i = 0; i < b->ndim; ++i) {
        int ai = nd - b->ndim + i;
        if (b->shape[i] == a->shape[ai]) bst[ai] = b->shape[i] == 1 ? 0 : b->stride[i];
        else if (b->shape[i] == 1) bst[ai] = 0;
        else return fail(DMXQ_ERR_BAD_ARG, "add_cast: b is not broadcastable to a");
    }
    // inner: longest suffix over which b is contiguous like a
    int64_t inner = 1;
    int d = nd - 1;
    for (; d >= 0; --d) {
        if (a->shape[d] == 1) continue;
        if (bst[d] != inner) break;
        inner *= a->shape[d];
    }
    // remaining dims [0..d]: collapse into at most two (size, stride) pairs
    std::vector<std::pair<int64_t, int64_t>> outer;  // outermost first
    for (int i = 0; i <= d; ++i) {
        if (a->shape[i] == 1) continue;
        if (!outer.empty() && outer.back().second == bst[i] * a->shape[i]) { outer.back().first *= a->shape[i]; outer.back().second = bst[i]; }
        else if (!outer.empty() && outer.back().second == 0 && bst[i] == 0) outer.back().first *= a->shape[i];
        else outer.emplace_back(a->shape[i], bst[i]);
    }
    if (outer.size() > 2 || inner % V != 0 || !aligned(a->data, 16) || !aligned(b->data, 16) || !aligned(y->data, 16))
        return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: layout not supported by the fused kernel");
    for (auto &o : outer) if ((o.second * dtype_size(a->dtype)) % 16 != 0) return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: misaligned broadcast stride");
    if (inner / V > 0xFFFFFFFFll) return fail(DMXQ_ERR_UNSUPPORTED, "add_cast: inner run too long");
    p.a = a->data; p.b = b->data; p.y = y->data;
    p.n_vec = n / V;
    p.inner_vec = (uint32_t)(inner / V);
    p.d1 = 1; p.bs0 = 0; p.bs1 = 0;
    if (outer.size() == 1) { p.d1 = (uint32_t)outer[0].first; p.bs1 = outer[0].second; }
    if (outer.size() == 2) { p.d1 = (uint32_t)outer[1].first; p.bs1 = outer[1].second; p.bs0 = outer[0].second; }
    cudaError_t e = launch_add(a->dtype, p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "add_cast_kernel");
    return DMXQ_OK;
}

int dmxq_softmax_cast(const dmxq_tensor *x, const dmxq_tensor *addend, const dmxq_tensor *y, const dmxq_stage *stage_x,
                      const dmxq_stage *stage_addend, const dmxq_stage *stage_sum, const dmxq_stage *post, int n_post, void *stream)
{
    if (!x || !y) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (!same_shape(x, y) || x->dtype != y->dtype) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: x and y must share shape and dtype");
    if (x->dtype < 0 || x->dtype > 2 || x->ndim < 1 || x->ndim > DMXQ_MAX_DIMS) return fail(DMXQ_ERR_BAD_ARG, "softmax_cast: bad dtype / rank");
    if (n_post < 0 || n_post > DMXQ_MAX_STAGES || (n_post > 0 && !post)) return fail(DMXQ_ERR_BAD_ARG, "softmax_cast: 0..%d output stages", DMXQ_MAX_STAGES);
    if (!addend && (stage_x || stage_addend || stage_sum)) return fail(DMXQ_ERR_BAD_ARG, "softmax_cast: add stages without an addend");
    const int nd = x->ndim, V = 16 / dtype_size(x->dtype);
    const int64_t n = x->shape[nd - 1];
    // torch's persistent-warp softmax with a full warp per row (ATen host_softmax: dim_size <= 2048 and <= 8 KiB per row)
    if (n <= 32 || n > 2048) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: rows of 33..2048 elements (got %lld)", (long long)n);
    if (n % V != 0) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: row length must be a multiple of %d", V);
    SoftmaxParams p;
    memset(&p, 0, sizeof(p));
    int64_t rows = 1, expect = n;
    for (int i = nd - 2; i >= 0; --i) {
        if (x->shape[i] != 1 && (x->stride[i] != expect || y->stride[i] != expect)) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: x and y must be contiguous");
        expect *= x->shape[i];
        rows *= x->shape[i];
    }
    if (x->stride[nd - 1] != 1 && n > 1) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: the softmax dim must be contiguous");
    if (rows == 0) return DMXQ_OK;
    if (!x->data || !y->data || !aligned(x->data, 16) || !aligned(y->data, 16)) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: 16-byte aligned data");
    if (rows > 0xFFFFFFFFll) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: too many rows");
    p.x = x->data; p.y = y->data; p.rows = rows; p.n = (int)n; p.xs = n; p.ys = n;
    p.d1 = p.d2 = 1;
    if (addend) {
        const dmxq_tensor *b = addend;
        if (b->dtype != x->dtype || b->ndim > nd || b->ndim < 1) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: addend dtype / rank");
        if (!b->data || !aligned(b->data, 16)) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: 16-byte aligned addend");
        if (b->shape[b->ndim - 1] != n || b->stride[b->ndim - 1] != 1) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: the addend must span the softmax dim contiguously");
        // outer dims of x (right-aligned against the addend; broadcast dims get stride 0), size-1 dims dropped, neighbours that
        // the addend walks like one dim merged, at most three left
        std::vector<std::pair<int64_t, int64_t>> outer;  // (size, addend stride), outermost first
        for (int i = 0; i < nd - 1; ++i) {
            if (x->shape[i] == 1) continue;
            int64_t bst = 0;
            const int bi = i - (nd - b->ndim);
            if (bi >= 0) {
                if (b->shape[bi] == x->shape[i]) bst = b->stride[bi];
                else if (b->shape[bi] != 1) return fail(DMXQ_ERR_BAD_ARG, "softmax_cast: the addend is not broadcastable to x");
            }
            if (!outer.empty() && outer.back().second == bst * x->shape[i]) { outer.back().first *= x->shape[i]; outer.back().second = bst; }
            else outer.emplace_back(x->shape[i], bst);
        }
        if (outer.size() > 3) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: addend broadcast pattern needs more than three outer dims");
        for (auto &o : outer)
            if ((o.second * dtype_size(x->dtype)) % 16 != 0) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: misaligned addend stride");
        while (outer.size() < 3) outer.insert(outer.begin(), std::make_pair<int64_t, int64_t>(1, 0));
        if (outer[1].first > 0xFFFFFFFFll || outer[2].first > 0xFFFFFFFFll) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: extent too large");
        p.b = b->data;
        p.d1 = (uint32_t)outer[1].first; p.d2 = (uint32_t)outer[2].first;
        p.bs[0] = outer[0].second; p.bs[1] = outer[1].second; p.bs[2] = outer[2].second;
        const dmxq_stage *sts[3] = {stage_x, stage_addend, stage_sum};
        FloatFmt *fmts[3] = {&p.fa, &p.fb, &p.fo};
        int *has[3] = {&p.has_a, &p.has_b, &p.has_o};
        for (int i = 0; i < 3; ++i) {
            if (!sts[i]) continue;
            StageDev d;
            int rc = decode_stage(*sts[i], d);
            if (rc) return rc;
            if (d.kind != ST_FLOAT || !d.ff.fastpath) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: only nearest+flush FLOAT stages fuse into the add");
            const int src_man = x->dtype == DMXQ_BF16 ? 7 : x->dtype == DMXQ_F16 ? 10 : 23;
            if (23 - d.ff.sh >= src_man) d.ff.exact = 1;
            *fmts[i] = d.ff;
            *has[i] = 1;
        }
    }
    if (n_post > 0) {
        bool blocked = false;
        int n_stoch = 0;
        int rc = decode_chain(x->dtype, y->dtype, post, n_post, p.chain, &blocked, &n_stoch);
        if (rc) return rc;
        if (n_stoch) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: stochastic stages are not fused");
        for (int s = 0; s < n_post; ++s) {
            const StageDev &d = p.chain.st[s];
            if (d.kind == ST_NM) return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: N:M stages are not fused");
            if (stage_blocked(d) && (d.block % V != 0 || !pow2(d.block / V) || d.block / V > 32 || n % d.block != 0))
                return fail(DMXQ_ERR_UNSUPPORTED, "softmax_cast: block size %d needs whole blocks of whole 16-byte vectors along the row", d.block);
        }
    }
    cudaError_t e = launch_softmax(x->dtype, p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "softmax_cast_kernel");
    return DMXQ_OK;
}

static int packed_args(const dmxq_tensor *t, const void *mant, const void *exps, int block_size, int precision, int64_t *n)
{
    if (!t || !mant || !exps) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (t->dtype < 0 || t->dtype > 2) return fail(DMXQ_ERR_BAD_ARG, "bad dtype");
    if (precision < 2 || precision > 8) return fail(DMXQ_ERR_UNSUPPORTED, "packed BFP supports precision 2..8, got %d", precision);
    const int V = 16 / dtype_size(t->dtype);
    if (block_size < 16 || block_size % 16 != 0 || !pow2(block_size / V) || block_size / V > 32)
        return fail(DMXQ_ERR_UNSUPPORTED, "packed BFP needs a power-of-two block size in 16..%d, got %d", 32 * V, block_size);
    int64_t total = 1, expect = 1;
    for (int i = t->ndim - 1; i >= 0; --i) {
        if (t->shape[i] != 1 && t->stride[i] != expect) return fail(DMXQ_ERR_UNSUPPORTED, "packed BFP needs a contiguous tensor");
        expect *= t->shape[i];
        total *= t->shape[i];
    }
    if (t->ndim < 1 || t->shape[t->ndim - 1] % block_size != 0) return fail(DMXQ_ERR_BAD_ARG, "last dim must be a multiple of the block size");
    if (total && !aligned(t->data, 16)) return fail(DMXQ_ERR_UNSUPPORTED, "packed BFP needs 16-byte aligned data");
    *n = total;
    return DMXQ_OK;
}

int dmxq_bfp_pack(const dmxq_tensor *x, void *mantissas, uint8_t *exponents, int block_size, int precision, void *stream)
{
    int64_t n = 0;
    int rc = packed_args(x, mantissas, exponents, block_size, precision, &n);
    if (rc || n == 0) return rc;
    cudaError_t e = launch_bfp_pack(x->dtype, x->data, mantissas, exponents, n, block_size, precision, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "bfp_pack_kernel");
    return DMXQ_OK;
}

int dmxq_bfp_unpack(const void *mantissas, const uint8_t *exponents, const dmxq_tensor *y, int block_size, int precision, void *stream)
{
    int64_t n = 0;
    int rc = packed_args(y, mantissas, exponents, block_size, precision, &n);
    if (rc || n == 0) return rc;
    cudaError_t e = launch_bfp_unpack(y->dtype, mantissas, exponents, y->data, n, block_size, precision, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "bfp_unpack_kernel");
    return DMXQ_OK;
}

static int sbfp_packed_args(const dmxq_tensor *t, const void *mant, const void *scalers, const dmxq_stage *fmt, StageDev &d, int64_t *n)
{
    if (!t || !mant || !scalers || !fmt) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (t->dtype < 0 || t->dtype > 2) return fail(DMXQ_ERR_BAD_ARG, "bad dtype");
    if (fmt->kind != DMXQ_STAGE_SBFP) return fail(DMXQ_ERR_BAD_ARG, "packed SBFP needs a DMXQ_STAGE_SBFP format description");
    int rc = decode_stage(*fmt, d);
    if (rc) return rc;
    if (fmt->precision < 2 || fmt->precision > 8) return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP supports block precision 2..8, got %d", fmt->precision);
    if (d.sb.xp.mode != R_NEAREST || d.sb.xp.tie != TIE_AWAY || !d.sb.xp.clamp)
        return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP needs a clamped nearest (half away) block format");
    if (!d.sb.sc.flush || fmt->sc_man + fmt->sc_exp > 8 || fmt->sc_man < 0 || fmt->sc_exp < 1)
        return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP needs a subnormal-flushing scaler format of at most 8 bits (got E%dM%d)", fmt->sc_exp, fmt->sc_man);
    const int B = fmt->block;
    if (B < 8 || B > 128 || !pow2(B)) return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP needs a power-of-two block size in 8..128, got %d", B);
    int64_t total = 1, expect = 1;
    for (int i = t->ndim - 1; i >= 0; --i) {
        if (t->shape[i] != 1 && t->stride[i] != expect) return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP needs a contiguous tensor");
        expect *= t->shape[i];
        total *= t->shape[i];
    }
    if (t->ndim < 1 || t->shape[t->ndim - 1] % B != 0) return fail(DMXQ_ERR_BAD_ARG, "last dim must be a multiple of the block size");
    if (total && !aligned(t->data, 16)) return fail(DMXQ_ERR_UNSUPPORTED, "packed SBFP needs 16-byte aligned data");
    *n = total;
    return DMXQ_OK;
}

int dmxq_sbfp_pack(const dmxq_tensor *x, void *mantissas, uint8_t *scalers, const dmxq_stage *fmt, unsigned int *n_inexact, void *stream)
{
    int64_t n = 0;
    StageDev d;
    int rc = sbfp_packed_args(x, mantissas, scalers, fmt, d, &n);
    if (rc || n == 0) return rc;
    cudaError_t e = launch_sbfp_pack(x->dtype, x->data, mantissas, scalers, n_inexact, n, fmt->block, d.sb, fmt->sc_man, fmt->sc_exp,
                                     static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "sbfp_pack_kernel");
    return DMXQ_OK;
}

int dmxq_sbfp_unpack(const void *mantissas, const uint8_t *scalers, const dmxq_tensor *y, const dmxq_stage *fmt, void *stream)
{
    int64_t n = 0;
    StageDev d;
    int rc = sbfp_packed_args(y, mantissas, scalers, fmt, d, &n);
    if (rc || n == 0) return rc;
    cudaError_t e = launch_sbfp_unpack(y->dtype, mantissas, scalers, y->data, n, fmt->block, d.sb, fmt->sc_man, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "sbfp_unpack_kernel");
    return DMXQ_OK;
}

int dmxq_minmax(const dmxq_tensor *x, int ch_axis, float *out_min, float *out_max, void *stream)
{
    if (!x || !out_min || !out_max) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (x->dtype < 0 || x->dtype > 2) return fail(DMXQ_ERR_BAD_ARG, "bad dtype");
    // contiguous (outer, C, inner) addressing; strided inputs: make them contiguous upstream
    int64_t n = 1, expect = 1;
    for (int i = x->ndim - 1; i >= 0; --i) {
        if (x->shape[i] != 1 && x->stride[i] != expect) return fail(DMXQ_ERR_UNSUPPORTED, "dmxq_minmax needs a contiguous tensor");
        expect *= x->shape[i];
        n *= x->shape[i];
    }
    MinMaxParams p;
    memset(&p, 0, sizeof(p));
    p.x = x->data; p.dtype = x->dtype;
    if (ch_axis < 0) {
        p.outer = 1; p.C = 1; p.inner = n; p.xo = 0; p.xc = 0; p.xi = 1;
    } else {
        if (ch_axis >= x->ndim) return fail(DMXQ_ERR_BAD_ARG, "ch_axis %d out of range", ch_axis);
        p.C = x->shape[ch_axis];
        p.inner = 1;
        for (int i = ch_axis + 1; i < x->ndim; ++i) p.inner *= x->shape[i];
        p.outer = p.C * p.inner ? n / (p.C * p.inner) : 0;
        p.xi = 1; p.xc = p.inner; p.xo = p.C * p.inner;
    }
    p.omin = reinterpret_cast<int *>(out_min);
    p.omax = reinterpret_cast<int *>(out_max);
    cudaError_t e = launch_minmax(p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "minmax_kernel");
    return DMXQ_OK;
}

int dmxq_histc(const dmxq_tensor *x, int bins, float lo, float hi, unsigned long long *counts, float *out_min, float *out_max,
               void *stream)
{
    if (!x || !counts) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if ((out_min == nullptr) != (out_max == nullptr)) return fail(DMXQ_ERR_BAD_ARG, "out_min and out_max go together");
    if (x->dtype < 0 || x->dtype > 2) return fail(DMXQ_ERR_BAD_ARG, "bad dtype");
    if (bins < 1 || bins > 12288) return fail(DMXQ_ERR_UNSUPPORTED, "bins must be in [1, 12288] (got %d)", bins);
    if (!(lo < hi) || std::isinf(lo) || std::isinf(hi))
        return fail(DMXQ_ERR_BAD_ARG, "histogram range [%g, %g] must be finite and non-empty", (double)lo, (double)hi);
    int64_t n = 1, expect = 1;
    for (int i = x->ndim - 1; i >= 0; --i) {
        if (x->shape[i] != 1 && x->stride[i] != expect) return fail(DMXQ_ERR_UNSUPPORTED, "dmxq_histc needs a contiguous tensor");
        expect *= x->shape[i];
        n *= x->shape[i];
    }
    if (n > 0 && reinterpret_cast<uintptr_t>(x->data) % 16) return fail(DMXQ_ERR_UNSUPPORTED, "dmxq_histc needs 16-byte aligned data");
    cudaError_t e = launch_histc(x->dtype, x->data, n, lo, hi, bins, counts, out_min, out_max, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "histc_kernel");
    return DMXQ_OK;
}

int dmxq_block_quantize(const dmxq_tensor *x, const dmxq_tensor *y, int wl, int dim, int symmetric, int rounding,
                        const int32_t *rand, void *workspace, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!x || !y || !workspace) return fail(DMXQ_ERR_BAD_ARG, "null argument");
    if (x->dtype != DMXQ_F32 || y->dtype != DMXQ_F32) return fail(DMXQ_ERR_BAD_ARG, "x is not a single precision Floating Point Tensor");
    if (!same_shape(x, y)) return fail(DMXQ_ERR_BAD_ARG, "x and y must have the same shape");
    if (wl < 2 || wl > 22) return fail(DMXQ_ERR_UNSUPPORTED, "block_quantize wl %d outside [2, 22]", wl);
    if (rounding < 0 || rounding > 3) return fail(DMXQ_ERR_BAD_ARG, "invalid rounding mode, %d", rounding);
    if (rounding == DMXQ_ROUND_STOCHASTIC && !rand) return fail(DMXQ_ERR_BAD_ARG, "stochastic rounding needs a random tensor");
    int64_t n = 1, expect = 1;
    for (int i = x->ndim - 1; i >= 0; --i) {
        if (x->shape[i] != 1 && (x->stride[i] != expect || y->stride[i] != expect)) return fail(DMXQ_ERR_BAD_ARG, "a must be contiguous");
        expect *= x->shape[i];
        n *= x->shape[i];
    }
    if (n == 0) return DMXQ_OK;
    if (dim < -1 || dim >= x->ndim) return fail(DMXQ_ERR_BAD_ARG, "dim %d out of range", dim);
    // per-slice max|x| via the exact min/max reduction: max|x| = max(-min, max) on bit patterns
    BlockQParams p;
    memset(&p, 0, sizeof(p));
    p.x = static_cast<const float *>(x->data); p.y = static_cast<float *>(y->data); p.rnd = rounding == DMXQ_ROUND_STOCHASTIC ? rand : nullptr;
    p.n = n; p.wl = wl; p.sh = 23 - wl; p.mask = (1u << p.sh) - 1u; p.mode = rounding; p.symmetric = symmetric != 0;
    if (dim == -1) { p.C = 1; p.inner = n; }
    else {
        p.C = x->shape[dim]; p.inner = 1;
        for (int i = dim + 1; i < x->ndim; ++i) p.inner *= x->shape[i];
        // dim == 0 of the reference (view(size0,-1)) is the same addressing with inner = n / size0
    }
    // |x| maxima: run minmax on the tensor, then fold into bit patterns with a tiny kernel-free trick:
    // max|x| bits = max(bits(max) if max>=0, bits(-min)) -- done on device by blockq via two arrays.
    // workspace layout: [C] uint32 maxbits | [C] float min | [C] float max
    uint32_t *maxbits = static_cast<uint32_t *>(workspace);
    float *mn = reinterpret_cast<float *>(maxbits + p.C);
    float *mx = mn + p.C;
    int rc = dmxq_minmax(x, dim == -1 ? -1 : dim, mn, mx, stream);
    if (rc) return rc;
    cudaError_t e = launch_fold_absmax(mn, mx, maxbits, p.C, st);
    if (e != cudaSuccess) return cuda_fail(e, "fold_absmax");
    p.maxbits = maxbits;
    e = launch_blockq(p, st);
    if (e != cudaSuccess) return cuda_fail(e, "blockq_kernel");
    return DMXQ_OK;
}

void *dmxq_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void dmxq_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int dmxq_cast_chain_host(const void *x_host, void *y_host, int in_dtype, int out_dtype, int64_t rows, int64_t K,
                         const dmxq_stage *stages, int n_stages, int device)
{
    if (!x_host || !y_host) return fail(DMXQ_ERR_BAD_ARG, "null host pointer");
    if (!dtype_pair_ok(in_dtype, out_dtype)) return fail(DMXQ_ERR_UNSUPPORTED, "unsupported dtype pair");
    if (rows < 0 || K <= 0) return fail(DMXQ_ERR_BAD_ARG, "bad shape");
    if (rows == 0) return DMXQ_OK;
    std::lock_guard<std::mutex> lock(g_host_mu);
    struct DeviceGuard {  // the caller's current device is restored on every exit path
        int prev = -1;
        DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
        ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    } guard;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    HostCtx &h = g_host;
    const size_t in_row = (size_t)K * dtype_size(in_dtype), out_row = (size_t)K * dtype_size(out_dtype);
    // chunk: ~32 MiB of input, whole rows
    int64_t rpc = std::max<int64_t>(1, (int64_t)((32u << 20) / in_row));
    rpc = std::min(rpc, rows);
    size_t need_in = (size_t)rpc * in_row, need_out = (size_t)rpc * out_row;
    if (h.device != device || h.cap_in < need_in || h.cap_out < need_out) {
        for (int i = 0; i < HostCtx::NB; ++i) {
            if (h.din[i]) cudaFree(h.din[i]);
            if (h.dout[i]) cudaFree(h.dout[i]);
            h.din[i] = h.dout[i] = nullptr;
            if (!h.stream[i] || h.device != device) {
                if (h.stream[i]) cudaStreamDestroy(h.stream[i]);
                e = cudaStreamCreateWithFlags(&h.stream[i], cudaStreamNonBlocking);
                if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
            }
            e = cudaMalloc(&h.din[i], need_in);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
            e = cudaMalloc(&h.dout[i], need_out);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        }
        h.device = device; h.cap_in = need_in; h.cap_out = need_out;
    }
    int rc = DMXQ_OK;
    int64_t chunk = 0;
    for (int64_t r0 = 0; r0 < rows && rc == DMXQ_OK; r0 += rpc, ++chunk) {
        int b = (int)(chunk % HostCtx::NB);
        int64_t nr = std::min(rpc, rows - r0);
        cudaStream_t s = h.stream[b];
        e = cudaMemcpyAsync(h.din[b], static_cast<const char *>(x_host) + (size_t)r0 * in_row, (size_t)nr * in_row, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D"); break; }
        dmxq_tensor tx = flat_view(h.din[b], in_dtype, nr, K), ty = flat_view(h.dout[b], out_dtype, nr, K);
        rc = chain_impl(&tx, &ty, 1, stages, n_stages, nullptr, nullptr, nullptr, s);
        if (rc) break;
        e = cudaMemcpyAsync(static_cast<char *>(y_host) + (size_t)r0 * out_row, h.dout[b], (size_t)nr * out_row, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { rc = cuda_fail(e, "D2H"); break; }
    }
    for (int i = 0; i < HostCtx::NB; ++i) {
        e = cudaStreamSynchronize(h.stream[i]);
        if (e != cudaSuccess && rc == DMXQ_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

}  // extern "C"

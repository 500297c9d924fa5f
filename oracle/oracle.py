"""numpy front-end of the CPU oracle (oracle/dmxq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs as the *checker*; never by the product package
(dmx_compressor_b200 has no CPU fallback and raises if libdmxq.so is missing).

Parity pin (tests/test_oracle.py): bit-equal to the reference's compiled quant_cpu
(oracle/_ref/ref_quant_cpu.so), to the golden vectors generated from the reference's own
python (tests/golden/*.npz, made by tests/golden/make_golden.py) and to the reference's
inline KATs.  The asymmetric-BFP post-pass and the composition helpers below are written
in numpy and follow S/numerical/format.py line by line.

Shorthands accepted by :func:`cast` are the reference's own (S/numerical/format.py):
``SAME``, ``XP[p,f](CSR)``, ``FP[s|e|m,b](FR)``, ``BFP[p|8]{B}(SR)``, ``SBFP<XP..><FP..>{B}``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "dmxq_oracle.c")
_LIB = os.path.join(HERE, "_build", "libdmxq_oracle.so")

ROUNDING = {"nearest": 0, "stochastic": 1, "up": 2, "down": 3, "N": 0, "S": 1, "U": 2, "D": 3}
TIE_AWAY, TIE_EVEN = 0, 1  # reference CUDA roundf  /  reference CPU nearbyint(a + .5f - .5)


def build(force: bool = False) -> str:
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           "-frounding-math", _SRC, "-o", _LIB, "-lm"])
    return _LIB


_lib = None
_i64, _int, _fp, _ip = C.c_int64, C.c_int, C.c_void_p, C.c_void_p


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_block_quantize_rows.argtypes = [_fp, _fp, _i64, _i64, _int, _int, _int, _ip]
        L.orc_bfp_cast.argtypes = [_fp, _fp, _i64, _i64, _i64, _i64, _int, _int, _ip]
        L.orc_float_cast.argtypes = [_fp, _fp, _i64] + [_int] * 7 + [_ip]
        L.orc_fixed_quantize.argtypes = [_fp, _fp, _i64] + [_int] * 6 + [_fp]
        L.orc_fixed_cast_affine.argtypes = [_fp, _fp, _i64, _i64, _i64] + [_int] * 6 + [_fp, _fp, _i64, _i64, _fp]
        L.orc_sbfp_cast.argtypes = [_fp, _fp, _i64, _i64, _i64, _i64] + [_int] * 12
        L.orc_nm_prune.argtypes = [_fp, _fp, _fp, _fp, _i64, _i64, _i64, _int, _int, _int]
        L.orc_argsort_cuda_order.argtypes = [_fp, _i64, _int, _fp]
        L.orc_argsort_cuda_order.restype = None
        L.orc_mxfp_cast.argtypes = [_fp, _fp, _i64, _i64, _i64, _i64, _int, _int]
        L.orc_mxfp_cast.restype = None
        L.orc_minmax.argtypes = [_fp, _i64, _i64, _i64, _fp, _fp]
        L.orc_histc.argtypes = [_fp, _i64, _int, C.c_float, C.c_float, _fp]
        L.orc_histc.restype = None
        L.orc_bf16_to_f32.argtypes = [_fp, _fp, _i64]
        L.orc_f32_to_bf16.argtypes = [_fp, _fp, _i64]
        L.orc_fixed_min_max.argtypes = [_int, _int, _int, _fp, _fp]
        for f in ("orc_block_quantize_rows", "orc_bfp_cast", "orc_float_cast", "orc_fixed_quantize",
                  "orc_fixed_cast_affine", "orc_sbfp_cast", "orc_nm_prune", "orc_minmax", "orc_bf16_to_f32",
                  "orc_f32_to_bf16", "orc_fixed_min_max"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _okI(shape, dim):
    """(outer, K, inner) of a contiguous array of `shape` blocked along `dim`."""
    dim = dim % len(shape)
    outer = int(np.prod(shape[:dim], dtype=np.int64))
    inner = int(np.prod(shape[dim + 1:], dtype=np.int64))
    return outer, int(shape[dim]), inner


# --------------------------------------------------------------------------- L1 (quant_function.py)
def block_quantize_rows(x, wl, symmetric=True, rounding="nearest", rand=None):
    """reference block_quantize(x2d, wl, dim=0): one shared exponent per row of x2d."""
    x = _f32(x)
    assert x.ndim == 2
    y = np.empty_like(x)
    r = None if rand is None else np.ascontiguousarray(rand, dtype=np.int32)
    lib().orc_block_quantize_rows(_p(x), _p(y), x.shape[0], x.shape[1], wl, int(symmetric), ROUNDING[rounding], _p(r))
    return y


def float_quantize(x, exp, man, bias=None, flush_subnormal=True, rounding="nearest", rand=None,
                   unsigned=False, fp16_flush=False):
    x = _f32(x)
    y = np.empty_like(x)
    if bias is None:
        bias = 2 ** (exp - 1) - 1
    r = None if rand is None else np.ascontiguousarray(rand, dtype=np.int32)
    lib().orc_float_cast(_p(x), _p(y), x.size, man, exp, bias, int(flush_subnormal), int(unsigned), int(fp16_flush),
                         ROUNDING[rounding], _p(r))
    return y


def fixed_point_quantize(x, wl, fl, clamp=True, symmetric=False, rounding="nearest", tie=TIE_EVEN, rand=None):
    x = _f32(x)
    y = np.empty_like(x)
    r = None if rand is None else _f32(rand)
    lib().orc_fixed_quantize(_p(x), _p(y), x.size, wl, fl, int(clamp), int(symmetric), ROUNDING[rounding], tie, _p(r))
    return y


# --------------------------------------------------------------------------- L2 (format.py / cast.py / sparse.py)
def bfp_cast(x, block_dim=-1, block_size=64, precision=8, symmetric=True, rounding="nearest", rand=None):
    """BlockFloatingPoint.cast (S/numerical/format.py:304-343) on an fp32 array."""
    x = _f32(x)
    if block_size == 1:  # format.py:312-320
        return float_quantize(x, exp=8, man=precision - 2, bias=127, flush_subnormal=False, rounding=rounding, rand=rand)
    o, K, i = _okI(x.shape, block_dim)
    y = np.empty_like(x)
    r = None if rand is None else np.ascontiguousarray(rand, dtype=np.int32)
    lib().orc_bfp_cast(_p(x), _p(y), o, K, i, block_size, precision, ROUNDING[rounding], _p(r))
    if not symmetric:
        y = _make_mantissa_asymmetric(y, x, block_dim, block_size, precision)
    return y


def _make_mantissa_asymmetric(q, x, block_dim, bs, n):
    """BlockFloatingPoint.make_mantissa_asymmetric, S/numerical/format.py:349-372, applied per
    chunk exactly as format.py:337-339 does (the early return at :363-364 is per chunk)."""
    qm = np.moveaxis(q, block_dim, -1)
    xm = np.moveaxis(x, block_dim, -1)
    shp = qm.shape
    q2 = np.ascontiguousarray(qm).reshape(-1, shp[-1])
    x2 = np.ascontiguousarray(xm).reshape(-1, shp[-1])
    out = q2.copy()
    for k0 in range(0, shp[-1], bs):
        qc, xc = q2[:, k0:k0 + bs], x2[:, k0:k0 + bs]
        man, exp = np.frexp(qc)
        exp = exp.astype(np.int32)
        exp[(exp == 0) & (man == 0)] = -200
        max_exp = exp.max(-1, keepdims=True) - n + 1
        with np.errstate(all="ignore"):
            int_man = (man * np.power(np.float32(2.0), (exp - max_exp).astype(np.float32))).astype(np.float32)
            int_man = np.trunc(int_man).astype(np.int32)
            edge = int_man == -(2 ** (n - 1) - 1)
            if not edge.any():
                continue
            quantum = np.power(np.float32(2.0), max_exp.astype(np.float32)).astype(np.float32)
            old_err = (qc - xc).astype(np.float32)
            cand_err = (old_err - np.broadcast_to(quantum, old_err.shape)).astype(np.float32)
            sub = edge & (np.abs(cand_err) <= np.abs(old_err))
            int_man = int_man - sub.astype(np.int32)
            out[:, k0:k0 + bs] = (int_man.astype(np.float32) * quantum).astype(np.float32)
    return np.moveaxis(out.reshape(shp), -1, block_dim % q.ndim)


def float_cast(x, mantissa, exponent, bias, flush_subnormal=True, unsigned=False, rounding="nearest", rand=None):
    """FloatingPoint.cast (S/numerical/format.py:208-233) on an fp32 array."""
    x = _f32(x)
    rep = _fp_repr(mantissa, exponent, bias, flush_subnormal, unsigned, rounding)
    if rep == "FP[1|8|23,127](_N)":  # format.py:209-212 identity shortcut for native fp32
        return x.copy()
    return float_quantize(x, exponent, mantissa, bias, flush_subnormal, rounding, rand, unsigned=unsigned,
                          fp16_flush=(rep == "FP[1|5|10,15](FN)"))


def _fp_repr(m, e, b, flush, unsigned, rounding):
    r = {"nearest": "N", "stochastic": "S", "up": "U", "down": "D"}[rounding]
    return f"FP[{'0' if unsigned else '1'}|{e}|{m},{b}]({'F' if flush else '_'}{r})"


def fixed_cast(x, precision, fraction, clamp=True, symmetric=True, rounding="nearest", tie=TIE_EVEN,
               scale=None, zero_point=None, ch_axis=None, group_size=None, rand=None):
    """CastTo.forward for FixedPoint incl. the affine wrap (S/numerical/cast.py:279-296)."""
    x = _f32(x)
    y = np.empty_like(x)
    if scale is None:
        scale, zero_point = np.ones(1, np.float32), np.zeros(1, np.float32)
    scale, zero_point = _f32(np.atleast_1d(scale)), _f32(np.atleast_1d(zero_point))
    nq = scale.size
    if nq == 1 or ch_axis is None:
        o, Cc, i, group = 1, 1, x.size, 1
    else:
        o, Cc, i = _okI(x.shape, ch_axis)
        group = group_size if group_size else 1
    r = None if rand is None else _f32(rand)
    lib().orc_fixed_cast_affine(_p(x), _p(y), o, Cc, i, precision, fraction, int(clamp), int(symmetric),
                                ROUNDING[rounding], tie, _p(scale), _p(zero_point), nq, group, _p(r))
    return y


def sbfp_cast(x, block_dim=-1, block_size=16, xp_precision=4, xp_clamp=True, xp_rounding="nearest", tie=TIE_EVEN,
              sc_mantissa=4, sc_exponent=4, sc_bias=7, sc_flush=True, sc_unsigned=True, sc_rounding="nearest", scale_recip=None):
    """ScaledBlockFloatingPoint.cast (S/numerical/format.py:453-479) on an fp32 array.  ``scale_recip``: the block scale
    as torch computes it for CUDA tensors (max * fp32(1/man_scaling)) instead of max / man_scaling; default = it goes
    with the tie rule (TIE_AWAY == "the reference on CUDA tensors", TIE_EVEN == "on CPU tensors")."""
    if scale_recip is None:
        scale_recip = tie == TIE_AWAY
    x = _f32(x)
    o, K, i = _okI(x.shape, block_dim)
    y = np.empty_like(x)
    fp16_flush = _fp_repr(sc_mantissa, sc_exponent, sc_bias, sc_flush, sc_unsigned, sc_rounding) == "FP[1|5|10,15](FN)"
    lib().orc_sbfp_cast(_p(x), _p(y), o, K, i, block_size, xp_precision, int(xp_clamp), ROUNDING[xp_rounding], tie,
                        sc_mantissa, sc_exponent, sc_bias, int(sc_flush), int(sc_unsigned), int(fp16_flush),
                        ROUNDING[sc_rounding], int(scale_recip))
    return y


def sbfp_pack(x, block_size=16, xp_precision=4, tie=TIE_AWAY, sc_mantissa=4, sc_exponent=4, sc_bias=7, scale_recip=None):
    """Packed SBFP storage of a [..., K] fp32 array, blocks along the last dim (the format of include/dmxq.h
    ``dmxq_sbfp_pack``), restated from the pieces of ScaledBlockFloatingPoint.cast (S/numerical/format.py:453-479):
    per block ``cmax = max|x| / man_scaling`` (:462-464), mantissa = ``block_format.cast(chunk / cmax)`` (:468), scaler =
    ``scaler_format.cast(cmax)`` (:469).  Stored: the mantissa as sign-magnitude (sign of x), the scaler as the exponent |
    mantissa fields of its E<sc_exponent>M<sc_mantissa> value (0 = zero scaler).
    -> (mantissas uint8, scalers uint8, n_inexact): n_inexact counts blocks the bytes cannot hold."""
    x = _f32(x)
    K = x.shape[-1]
    assert K % block_size == 0
    man_scaling = np.float32(2 ** (xp_precision - 1) - 1)
    blk = x.reshape(x.shape[:-1] + (K // block_size, block_size))
    with np.errstate(all="ignore"):
        m = np.abs(blk).max(-1, keepdims=True)
        finite = np.isfinite(blk).all(-1, keepdims=True)
        if scale_recip is None:
            scale_recip = tie == TIE_AWAY
        cmax = (m * np.float32(1.0 / float(man_scaling))).astype(np.float32) if scale_recip else (m / man_scaling).astype(np.float32)
        on = finite & (cmax > 0)
        fs = float_cast(np.where(on, cmax, np.float32(0)), sc_mantissa, sc_exponent, sc_bias, True, True, "nearest")
        q = fixed_cast((blk / np.where(on, cmax, np.float32(1))).astype(np.float32), xp_precision, 0, True, True, "nearest", tie=tie)
    mag = np.where(on, np.abs(q), 0).astype(np.uint8)
    sign = (np.signbit(blk) & finite).astype(np.uint8)
    sh = 23 - sc_mantissa
    base = (((127 - (sc_bias - 1)) << 23) >> sh) - (1 << sc_mantissa)
    code = np.where(fs == 0, 0, (fs.view(np.uint32).astype(np.int64) >> sh) - base)
    code_max = (1 << (sc_exponent + sc_mantissa)) - 1
    inexact = ~finite | ((m > 0) & ~(cmax > 0)) | (code > code_max)
    code = np.minimum(code, code_max).astype(np.uint8)
    if xp_precision <= 4:
        t = (mag | (sign << 3)).reshape(x.shape[:-1] + (K // 2, 2))
        mant = (t[..., 0] | (t[..., 1] << 4)).astype(np.uint8)
    else:
        mant = (mag | (sign << 7)).reshape(x.shape).astype(np.uint8)
    return mant, code.reshape(x.shape[:-1] + (K // block_size,)), int(inexact.sum())


def sbfp_unpack(mant, scal, block_size=16, xp_precision=4, sc_mantissa=4, sc_bias=7):
    """dequantise :func:`sbfp_pack` bytes: sign * (magnitude * scaler), the last step of format.py:468-469"""
    mant, scal = np.asarray(mant, np.uint8), np.asarray(scal, np.uint8).astype(np.int64)
    if xp_precision <= 4:
        t = np.stack([mant & 0xF, mant >> 4], -1).reshape(mant.shape[:-1] + (-1,))
        mag, sg = (t & 7).astype(np.float32), (t >> 3).astype(bool)
    else:
        mag, sg = (mant & 0x7F).astype(np.float32), (mant >> 7).astype(bool)
    E, M = scal >> sc_mantissa, scal & ((1 << sc_mantissa) - 1)
    fs = np.where(scal == 0, 0.0, np.ldexp(1.0 + M / float(1 << sc_mantissa), E - sc_bias)).astype(np.float32)
    y = (mag.reshape(mag.shape[:-1] + (-1, block_size)) * fs[..., None]).astype(np.float32).reshape(mag.shape)
    return np.where(sg, -y, y).astype(np.float32)


def mxfp_cast(x, block_dim=-1, block_size=32, mantissa=3, exponent=4):
    """MXFP.cast (S/numerical/format.py:545-564) on an fp32 array."""
    x = _f32(x)
    o, K, i = _okI(x.shape, block_dim)
    y = np.empty_like(x)
    lib().orc_mxfp_cast(_p(x), _p(y), o, K, i, block_size, mantissa, exponent)
    return y


def argsort_cuda_order(keys):
    """torch.argsort(keys, dim=1) as torch computes it on CUDA for rows of <= 32 keys (unstable bitonic network)."""
    keys = _f32(keys)
    rows, m = keys.shape
    assert m <= 32
    out = np.empty((rows, m), np.int32)
    lib().orc_argsort_cuda_order(_p(keys), rows, m, _p(out))
    return out


NM_STABLE, NM_TORCH_CUDA = 0, 1


def nm_prune(x, n_keep, m, block_dim=-1, score=None, return_mask=False, nm_order=NM_STABLE):
    """Sparsify.forward with a BlockTopK sparseness (S/sparse.py:163-180, 287-301).  ``nm_order``: tie order of the
    group sort -- stable (the reference on CPU tensors) or torch's CUDA order (groups of <= 32)."""
    x = _f32(x)
    o, K, i = _okI(x.shape, block_dim)
    assert K % m == 0, f"score has size {K} at dimension {block_dim}, not a multiple of block size {m}"
    y = np.empty_like(x)
    mask = np.empty_like(x) if return_mask else None
    s = None if score is None else _f32(score)
    lib().orc_nm_prune(_p(x), _p(s), _p(y), _p(mask), o, K, i, n_keep, m, int(nm_order))
    return (y, mask) if return_mask else y


def minmax(x, ch_axis=None):
    x = _f32(x)
    if ch_axis is None:
        o, Cc, i = 1, 1, x.size
    else:
        o, Cc, i = _okI(x.shape, ch_axis)
    mn, mx = np.empty(Cc, np.float32), np.empty(Cc, np.float32)
    lib().orc_minmax(_p(x), o, Cc, i, _p(mn), _p(mx))
    return mn, mx


def histc(x, bins=100, min=0, max=0):
    """torch.histc(x, bins, min, max) on fp32 data -> float32 histogram (min == max: the data's own range; still
    equal: widened by one either side -- ATen's rule)."""
    x = _f32(x).reshape(-1)
    lo, hi = np.float32(min), np.float32(max)
    if lo == hi and x.size:
        lo, hi = x.min(), x.max()
    if lo == hi:
        lo, hi = np.float32(lo - 1), np.float32(hi + 1)
    if not (np.isfinite(lo) and np.isfinite(hi)):
        raise RuntimeError(f"range of [{lo}, {hi}] is not finite")
    counts = np.zeros(bins, np.int64)
    lib().orc_histc(_p(x), x.size, bins, float(lo), float(hi), _p(counts))
    return counts.astype(np.float32)


def bf16_to_f32(u16):
    u16 = np.ascontiguousarray(u16, dtype=np.uint16)
    y = np.empty(u16.shape, np.float32)
    lib().orc_bf16_to_f32(_p(u16), _p(y), u16.size)
    return y


def f32_to_bf16(x):
    x = _f32(x)
    y = np.empty(x.shape, np.uint16)
    lib().orc_f32_to_bf16(_p(x), _p(y), x.size)
    return y


# --------------------------------------------------------------------------- shorthand front door
_RX_XP = re.compile(r"^XP\[(\d+),([-+]?\d+)\]\((\w)(\w)(\w)\)$")
_RX_FP = re.compile(r"^FP\[(\d)\|(\d+)\|(\d+),([-+]?\d+)\]\((\w)([A-Za-z])\)$")
_RX_BFP = re.compile(r"^BFP\[(\d+)\|8\]\{(\d+)\}\((\w)([A-Za-z])\)$")
_RX_SBFP = re.compile(r"^SBFP<(.+?)><(.+?)>\{(\d+)\}$")
_RX_MXFP = re.compile(r"^MXFP(\d+)\[E(\d+)M(\d+)\]\{(\d+)\}$")
_RX_MXINT = re.compile(r"^MXINT(\d+)\{(\d+)\}$")
_RMODE = {"N": "nearest", "S": "stochastic", "U": "up", "D": "down"}


def cast(x, shorthand, block_dim=-1, tie=TIE_EVEN, rand=None, **affine):
    """`Format.from_shorthand(shorthand).cast(x, block_dim)` on an fp32 numpy array."""
    sh = shorthand
    if sh.startswith("SAME"):
        return _f32(x).copy()
    m = _RX_XP.match(sh)
    if m:
        return fixed_cast(x, int(m[1]), int(m[2]), m[3] == "C", m[4] == "S", _RMODE[m[5]], tie=tie, rand=rand, **affine)
    m = _RX_FP.match(sh)
    if m:
        return float_cast(x, int(m[3]), int(m[2]), int(m[4]), m[5] == "F", m[1] == "0", _RMODE[m[6]], rand=rand)
    m = _RX_BFP.match(sh)
    if m:
        return bfp_cast(x, block_dim, int(m[2]), int(m[1]), m[3] == "S", _RMODE[m[4]], rand=rand)
    m = _RX_SBFP.match(sh)
    if m:
        xp, fp = _RX_XP.match(m[1]), _RX_FP.match(m[2])
        return sbfp_cast(x, block_dim, int(m[3]), int(xp[1]), xp[3] == "C", _RMODE[xp[5]], tie,
                         int(fp[3]), int(fp[2]), int(fp[4]), fp[5] == "F", fp[1] == "0", _RMODE[fp[6]])
    m = _RX_MXFP.match(sh)
    if m:
        return mxfp_cast(x, block_dim, int(m[4]), int(m[3]), int(m[2]))
    m = _RX_MXINT.match(sh)
    if m:  # MXINT(BlockFloatingPoint), S/numerical/format.py:605-627
        return bfp_cast(x, block_dim, int(m[2]), int(m[1]), True, "nearest")
    raise ValueError(f"unrecognized format shorthand: {sh}")


# ---------------------------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), restated in numpy to pin the
# stream dmxq_philox_fill / dmxq_cast_chain_philox produce: word i = Philox(counter = (i // 4, stream_id), key = seed)[i % 4].
def philox_words(n, seed, stream_id=0):
    import numpy as np

    q = np.arange((n + 3) // 4, dtype=np.uint64)
    c0 = (q & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (q >> np.uint64(32)).astype(np.uint64)
    c2 = np.full_like(c0, stream_id & 0xFFFFFFFF)
    c3 = np.full_like(c0, (stream_id >> 32) & 0xFFFFFFFF)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2  # 32 x 32 -> 64 bit products (no overflow in uint64)
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).reshape(-1)[:n].astype(np.uint32)


def philox_unit_floats(n, seed, stream_id=0):
    import numpy as np

    return ((philox_words(n, seed, stream_id) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)

"""Functional front-end of the CUDA cast path: torch tensors in, torch tensors out, every
call one trip through the C ABI (dmx_compressor_b200/_lib.py -> libdmxq.so).

Output allocation policy (the reference returns a fresh tensor, Q/quant_cuda/quant.cu:60):
``torch.empty_like(x)`` with preserved strides when x is dense, so strided views (e.g.
``key.transpose(-2, -1)``) are cast in place of their layout with no ``.contiguous()`` copy.
"""
from __future__ import annotations

import ctypes as C
import math
import struct
from typing import Optional, Sequence

import torch

from . import _lib as L


class _guard:
    """device guard that costs nothing when x already lives on the current device"""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _out_like(x: torch.Tensor, dtype: Optional[torch.dtype]) -> torch.Tensor:
    dtype = dtype or x.dtype
    if x.is_contiguous() or x.numel() == 0:
        return torch.empty(x.shape, dtype=dtype, device=x.device)
    return torch.empty_like(x, dtype=dtype)  # preserve_format: keeps dense permuted layouts


def make_stage(**kw) -> L.Stage:
    s = L.Stage()
    s.scale = 1.0
    s.zero_point = 0.0
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def bfp_stage(block_size: int, precision: int, symmetric: bool = True, rounding: str = "nearest") -> L.Stage:
    return make_stage(kind=L.ST_BFP, block=block_size, precision=precision, symmetric=int(symmetric),
                      rounding=L.ROUND[rounding])


def float_stage(mantissa: int, exponent: int, bias: int, flush_subnormal: bool = True, unsigned: bool = False,
                fp16_flush: bool = False, rounding: str = "nearest") -> L.Stage:
    return make_stage(kind=L.ST_FLOAT, man=mantissa, exp=exponent, bias=bias, flush=int(flush_subnormal),
                      is_unsigned=int(unsigned), fp16_flush=int(fp16_flush), rounding=L.ROUND[rounding])


def fixed_stage(precision: int, fraction: int, clamp: bool = True, symmetric: bool = True, rounding: str = "nearest",
                tie: int = L.TIE_AWAY, scale: float = 1.0, zero_point: float = 0.0) -> L.Stage:
    return make_stage(kind=L.ST_FIXED, precision=precision, fraction=fraction, clamp=int(clamp), symmetric=int(symmetric),
                      rounding=L.ROUND[rounding], tie=tie, scale=scale, zero_point=zero_point)


def _scale_mode(tie: int, scale_mode: Optional[int]) -> int:
    """SBFP block-scale rule: by default it goes with the tie rule -- both are "which of the reference's two back ends":
    TIE_AWAY + SCALE_RECIP = the reference on CUDA tensors, TIE_EVEN + SCALE_DIV = on CPU tensors (include/dmxq.h)"""
    return (L.SCALE_RECIP if tie == L.TIE_AWAY else L.SCALE_DIV) if scale_mode is None else scale_mode


def sbfp_stage(block_size: int, xp_precision: int, xp_clamp: bool, xp_rounding: str, tie: int, sc_mantissa: int,
               sc_exponent: int, sc_bias: int, sc_flush: bool, sc_unsigned: bool, sc_fp16_flush: bool = False,
               sc_rounding: str = "nearest", scale_mode: Optional[int] = None) -> L.Stage:
    return make_stage(scale_mode=_scale_mode(tie, scale_mode), kind=L.ST_SBFP, block=block_size, precision=xp_precision, clamp=int(xp_clamp),
                      rounding=L.ROUND[xp_rounding], tie=tie, sc_man=sc_mantissa, sc_exp=sc_exponent, sc_bias=sc_bias,
                      sc_flush=int(sc_flush), sc_unsigned=int(sc_unsigned), sc_fp16_flush=int(sc_fp16_flush),
                      sc_rounding=L.ROUND[sc_rounding])


def mxfp_stage(block_size: int, mantissa: int, exponent: int) -> L.Stage:
    return make_stage(kind=L.ST_MXFP, block=block_size, man=mantissa, exp=exponent)


def scale_stage(vec: torch.Tensor, multiply: bool = False) -> L.Stage:
    """SmoothQuant's scale application as the first stage of a chain: x / vec[k] (or x * vec[k]) along block_dim, in fp32 --
    `a / scale.view(..)`, `b * scale.view(..)` of S/numerical/smoothquant.py:253-283.  The chain's output is fp32 (torch promotes
    `tensor / fp32_vector`): call cast_chain(..., out_dtype=torch.float32).  The stage keeps ``vec`` alive."""
    L.require_cuda(vec, "vec")
    if vec.dtype != torch.float32 or vec.dim() != 1 or not vec.is_contiguous():
        raise RuntimeError("dmxq: the scale vector must be a contiguous 1-d fp32 CUDA tensor")
    s = make_stage(kind=L.ST_SCALE, vec=vec.data_ptr(), vec_len=vec.numel(), vec_op=int(multiply))
    s._keep = vec
    return s


def nm_stage(n_keep: int, m: int, nm_order: int = L.NM_STABLE) -> L.Stage:
    """N:M prune stage.  ``nm_order``: tie order inside a group -- NM_STABLE (torch.argsort(stable=True), what the reference
    gets on CPU tensors) or NM_TORCH_CUDA (what it gets on CUDA tensors; see include/dmxq.h)."""
    return make_stage(kind=L.ST_NM, block=m, n_keep=n_keep, nm_order=nm_order)


def philox_fill(shape, seed: int, stream_id: int = 0, as_float: bool = False, device="cuda") -> torch.Tensor:
    """the random stream the kernels compute in registers under ``philox=(seed, stream_id)``, as a tensor (dmxq_philox_fill):
    int32 words, or -- ``as_float``, what FixedPoint's stochastic rounding consumes -- uniform fp32 values in [0, 1)"""
    out = torch.empty(shape, dtype=torch.float32 if as_float else torch.int32, device=device)
    with _guard(out.device):
        rc = L.lib.dmxq_philox_fill(out.data_ptr(), out.numel(), int(as_float), seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFFFFFFFFFF, L.stream_ptr(out.device))
    L.check(rc, "dmxq_philox_fill")
    return out


def cast_chain(x: torch.Tensor, stages: Sequence[L.Stage], block_dim: int = -1, out: Optional[torch.Tensor] = None,
               out_dtype: Optional[torch.dtype] = None, score: Optional[torch.Tensor] = None,
               mask: Optional[torch.Tensor] = None, rand: Optional[torch.Tensor] = None, philox=None) -> torch.Tensor:
    """y = stages[-1](... stages[0](x)) in one pass over HBM (dmxq_cast_chain).

    Stochastic stages take their random words from ``rand`` (a tensor, as the reference draws one) or -- ``philox=(seed,
    stream_id)`` -- compute them in the kernel (Philox4x32-10 over the logical element index, dmxq_cast_chain_philox): no
    random tensor is written or read.  Both give the same result for ``rand = philox_fill(x.shape, seed, stream_id)``."""
    L.require_cuda(x)
    n = len(stages)
    if not 1 <= n <= L.MAX_STAGES:
        raise RuntimeError(f"dmxq: a chain holds 1..{L.MAX_STAGES} stages, got {n}")
    y = out if out is not None else _out_like(x, out_dtype)
    arr = C.byref(stages[0]) if n == 1 else (L.Stage * n)(*stages)  # one stage: its own struct is the array
    vx, vy = L.view(x), L.view(y)
    if philox is not None and rand is None and score is None and mask is None:
        seed, sid = philox
        with _guard(x.device):
            rc = L.lib.dmxq_cast_chain_philox(C.byref(vx), C.byref(vy), block_dim, arr, n, seed & 0xFFFFFFFFFFFFFFFF, sid & 0xFFFFFFFFFFFFFFFF, L.stream_ptr(x.device))
        if rc == -2:  # a layout the rows kernels do not take: the same stream as an explicit tensor
            fixed = any(s.kind == L.ST_FIXED and s.rounding == L.ROUND["stochastic"] for s in stages)
            rand = philox_fill(x.shape, seed, sid, as_float=fixed, device=x.device)
        else:
            L.check(rc, "dmxq_cast_chain_philox")
            return y
    elif philox is not None and rand is None:
        seed, sid = philox
        fixed = any(s.kind == L.ST_FIXED and s.rounding == L.ROUND["stochastic"] for s in stages)
        rand = philox_fill(x.shape, seed, sid, as_float=fixed, device=x.device)
    vs = vm = None
    if score is not None:
        L.require_cuda(score, "score")
        score = score if score.dtype == torch.float32 else score.float()
        vs = C.byref(L.view(score))
    if mask is not None:
        vm = C.byref(L.view(mask))
    rp = None
    if rand is not None:
        L.require_cuda(rand, "rand")
        if not rand.is_contiguous() or rand.shape != x.shape or rand.element_size() != 4:
            raise RuntimeError("dmxq: rand must be a contiguous 4-byte tensor with the shape of x")
        rp = rand.data_ptr()
    with _guard(x.device):
        rc = L.lib.dmxq_cast_chain(C.byref(vx), C.byref(vy), block_dim, arr, n, vs, vm, rp, L.stream_ptr(x.device))
    L.check(rc, "dmxq_cast_chain")
    return y


def _groups(n: int):
    """(start, size) of the per-call tensor groups of the many-tensor entry points: 8, 24, then 64 each"""
    i = 0
    for g in (8, 24):
        if i < n:
            g = min(g, n - i)
            yield i, g
            i += g
    while i < n:
        g = min(64, n - i)
        yield i, g
        i += g


def cast_chain_multi(xs: Sequence[torch.Tensor], stages: Sequence[L.Stage], block_dim: int = -1,
                     outs: Optional[Sequence[torch.Tensor]] = None, amax: Optional[torch.Tensor] = None):
    """ys[i] = chain(xs[i]) for many tensors in as few launches as possible (dmxq_cast_chain_multi): the shards one rank
    owns in a sharded whole-model weight cast.  ``amax``: optional fp32 CUDA vector, one tensor-wide amax per entry of
    ``xs`` (e.g. straight out of an all-reduce); an SBFP stage then takes its scaler exponent bias from it on the device."""
    n = len(xs)
    if n == 0:
        return []
    for x in xs:
        L.require_cuda(x)
    ns = len(stages)
    if not 1 <= ns <= L.MAX_STAGES:
        raise RuntimeError(f"dmxq: a chain holds 1..{L.MAX_STAGES} stages, got {ns}")
    ys = list(outs) if outs is not None else [_out_like(x, None) for x in xs]
    if len(ys) != n:
        raise RuntimeError("dmxq: outs must have one tensor per input")
    dev = xs[0].device
    if any(t.device != dev for t in list(xs) + ys):
        raise RuntimeError("dmxq: all tensors of one call must live on one device")
    ap = None
    if amax is not None:
        L.require_cuda(amax, "amax")
        if amax.dtype != torch.float32 or amax.numel() != n or not amax.is_contiguous() or amax.device != dev:
            raise RuntimeError("dmxq: amax must be a contiguous fp32 CUDA vector with one entry per tensor")
        ap = amax.data_ptr()
    arr = (L.Stage * ns)(*stages)
    stream = L.stream_ptr(dev)
    # one C call per group of up to 64 tensors (= one launch for flat tensors), a small first group: the GPU starts almost at
    # once and works on group k while the host is still describing group k + 1 (~3 us per tensor in python)
    rc = 0
    with _guard(dev):
        for i, g in _groups(n):
            vx, kx = L.views(xs[i:i + g])
            vy, ky = L.views(ys[i:i + g])
            rc = L.lib.dmxq_cast_chain_multi(vx, vy, g, block_dim, arr, ns, None if ap is None else ap + 4 * i, stream)
            if rc:
                break
    L.check(rc, "dmxq_cast_chain_multi")
    return ys


# Where stochastic rounding takes its random words from when the caller passes no tensor:
#   "torch"  (default) draw a full-size tensor with torch's generator exactly as the reference's CUDA launchers do
#            (randint_like / rand_like): the same torch seed reproduces the reference's stream;
#   "philox" compute them in the cast kernel (dmxq_cast_chain_philox): no random tensor exists; seed = torch's initial seed
#            unless given, one stream id per call (a counter), so results are reproducible per (seed, call order).
_STOCH = {"source": "torch", "seed": None, "calls": 0}


def stochastic_source(source: str = "torch", seed: Optional[int] = None) -> None:
    if source not in ("torch", "philox"):
        raise ValueError("source must be 'torch' or 'philox'")
    _STOCH.update(source=source, seed=seed, calls=0)


def _next_philox():
    seed = _STOCH["seed"] if _STOCH["seed"] is not None else torch.initial_seed()
    _STOCH["calls"] += 1
    return seed, _STOCH["calls"]


def bfp_qdq(x, block_dim=-1, block_size=64, precision=8, symmetric=True, rounding="nearest", rand=None, out=None,
            out_dtype=None):
    """BlockFloatingPoint.cast (reference S/numerical/format.py:304-372) via dmxq_bfp_qdq."""
    L.require_cuda(x)
    if rounding == "stochastic" and rand is None and _STOCH["source"] == "philox":
        return cast_chain(x, [bfp_stage(block_size, precision, symmetric, rounding)], block_dim, out=out, out_dtype=out_dtype, philox=_next_philox())
    if rounding == "stochastic" and rand is None:
        rand = torch.randint_like(x, 2**31 - 1, dtype=torch.int32)  # Q/quant_cuda/quant.cu:40
    y = out if out is not None else _out_like(x, out_dtype)
    vx, vy = L.view(x), L.view(y)
    with _guard(x.device):
        rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vy), block_dim, block_size, precision, int(symmetric),
                                L.ROUND[rounding], rand.data_ptr() if rand is not None else None, L.stream_ptr(x.device))
    L.check(rc, "dmxq_bfp_qdq")
    return y


def sbfp_qdq(x, block_dim=-1, block_size=16, xp_precision=4, xp_clamp=True, xp_rounding="nearest", tie=L.TIE_AWAY,
             sc_mantissa=4, sc_exponent=4, sc_bias=7, sc_flush=True, sc_unsigned=True, sc_fp16_flush=False,
             sc_rounding="nearest", out=None, out_dtype=None, scale_mode=None):
    """ScaledBlockFloatingPoint.cast (reference S/numerical/format.py:453-479) via dmxq_sbfp_qdq."""
    L.require_cuda(x)
    y = out if out is not None else _out_like(x, out_dtype)
    vx, vy = L.view(x), L.view(y)
    with _guard(x.device):
        rc = L.lib.dmxq_sbfp_qdq(C.byref(vx), C.byref(vy), block_dim, block_size, xp_precision, int(xp_clamp),
                                 L.ROUND[xp_rounding], tie, sc_mantissa, sc_exponent, sc_bias, int(sc_flush),
                                 int(sc_unsigned), int(sc_fp16_flush), L.ROUND[sc_rounding], _scale_mode(tie, scale_mode),
                                 L.stream_ptr(x.device))
    L.check(rc, "dmxq_sbfp_qdq")
    return y


def float_qdq(x, mantissa, exponent, bias, flush_subnormal=True, unsigned=False, fp16_flush=False, rounding="nearest",
              rand=None, out=None, out_dtype=None):
    """FloatingPoint.cast (reference S/numerical/format.py:208-233) via dmxq_float_qdq."""
    L.require_cuda(x)
    if rounding == "stochastic" and rand is None and _STOCH["source"] == "philox":
        return cast_chain(x, [float_stage(mantissa, exponent, bias, flush_subnormal, unsigned, fp16_flush, rounding)], -1, out=out,
                          out_dtype=out_dtype, philox=_next_philox())
    if rounding == "stochastic" and rand is None:
        rand = torch.randint_like(x, 2**31 - 1, dtype=torch.int32)  # Q/quant_cuda/quant.cu:160
    y = out if out is not None else _out_like(x, out_dtype)
    vx, vy = L.view(x), L.view(y)
    with _guard(x.device):
        rc = L.lib.dmxq_float_qdq(C.byref(vx), C.byref(vy), mantissa, exponent, bias, int(flush_subnormal), int(unsigned),
                                  int(fp16_flush), L.ROUND[rounding], rand.data_ptr() if rand is not None else None,
                                  L.stream_ptr(x.device))
    L.check(rc, "dmxq_float_qdq")
    return y


def fixed_qdq(x, precision, fraction, clamp=True, symmetric=True, rounding="nearest", tie=L.TIE_AWAY, scale=None,
              zero_point=None, ch_axis=-1, group_size=None, rand=None, out=None, out_dtype=None):
    """FixedPoint.cast incl. CastTo's affine wrap (reference S/numerical/format.py:134-142,
    S/numerical/cast.py:279-296) via dmxq_fixed_qdq.  scale / zero_point: fp32 CUDA tensors."""
    L.require_cuda(x)
    if rounding == "stochastic" and rand is None:
        rand = torch.rand_like(x, dtype=torch.float32)  # Q/quant_cuda/quant.cu:244
    sp = zp = None
    nq = 0
    if scale is not None:
        scale = scale.to(device=x.device, dtype=torch.float32).contiguous().view(-1)
        zero_point = zero_point.to(device=x.device, dtype=torch.float32).contiguous().view(-1)
        nq = scale.numel()
        sp, zp = scale.data_ptr(), zero_point.data_ptr()
        if not x.is_contiguous():
            x = x.contiguous()
    y = out if out is not None else _out_like(x, out_dtype)
    vx, vy = L.view(x), L.view(y)
    with _guard(x.device):
        rc = L.lib.dmxq_fixed_qdq(C.byref(vx), C.byref(vy), precision, fraction, int(clamp), int(symmetric), L.ROUND[rounding],
                                  tie, sp, zp, nq, ch_axis, group_size or 1, rand.data_ptr() if rand is not None else None,
                                  L.stream_ptr(x.device))
    L.check(rc, "dmxq_fixed_qdq")
    return y


def nm_prune(x, n_keep, m, block_dim=-1, score=None, return_mask=False, out=None, out_dtype=None, nm_order=L.NM_STABLE):
    """Sparsify.forward with BlockTopK (reference S/sparse.py:163-180, 287-301) via dmxq_nm_prune."""
    L.require_cuda(x)
    y = out if out is not None else _out_like(x, out_dtype)
    mask = torch.empty(x.shape, dtype=torch.float32, device=x.device) if return_mask else None
    vx, vy = L.view(x), L.view(y)
    vs = vm = None
    if score is not None:
        L.require_cuda(score, "score")
        score = score if score.dtype == torch.float32 else score.float()
        vs = C.byref(L.view(score))
    if mask is not None:
        vm = C.byref(L.view(mask))
    with _guard(x.device):
        rc = L.lib.dmxq_nm_prune(C.byref(vx), vs, C.byref(vy), vm, block_dim, n_keep, m, nm_order, L.stream_ptr(x.device))
    L.check(rc, "dmxq_nm_prune")
    return (y, mask) if return_mask else y


def add_cast(a, b, stage_a=None, stage_b=None, stage_out=None, out=None):
    """y = out(A(a) + B(b)): ResAdd with its boundary casts in one pass (dmxq_add_cast).  Raises
    RuntimeError('unsupported') for layouts / stages the fused kernel does not take."""
    L.require_cuda(a, "a")
    L.require_cuda(b, "b")
    y = out if out is not None else torch.empty(a.shape, dtype=a.dtype, device=a.device)
    va, vb, vy = L.view(a), L.view(b), L.view(y)
    ptr = lambda s: None if s is None else C.byref(s)
    with _guard(a.device):
        rc = L.lib.dmxq_add_cast(C.byref(va), C.byref(vb), C.byref(vy), ptr(stage_a), ptr(stage_b), ptr(stage_out), L.stream_ptr(a.device))
    L.check(rc, "dmxq_add_cast")
    return y


def softmax_cast(x, post=(), addend=None, stage_x=None, stage_addend=None, stage_sum=None, out=None):
    """y = post(softmax(x [+ addend], dim=-1)) in one pass (dmxq_softmax_cast): torch's CUDA softmax bit for bit, with the
    attention-mask add and its casts in front (``addend`` + the three FLOAT stages, as ``add_cast``) and the casts of the
    probabilities behind (``post``: a sequence of stages applied along the last dim, each followed by the rounding to
    x.dtype).  Rows of 33..2048 elements; raises RuntimeError('... unsupported ...') otherwise."""
    L.require_cuda(x)
    if addend is not None:
        L.require_cuda(addend, "addend")
    y = out if out is not None else torch.empty(x.shape, dtype=x.dtype, device=x.device)
    vx, vy = L.view(x), L.view(y)
    vb = L.view(addend) if addend is not None else None
    n = len(post)
    arr = (L.Stage * n)(*post) if n else None
    ptr = lambda s: None if s is None else C.byref(s)
    with _guard(x.device):
        rc = L.lib.dmxq_softmax_cast(C.byref(vx), ptr(vb), C.byref(vy), ptr(stage_x), ptr(stage_addend), ptr(stage_sum), arr, n, L.stream_ptr(x.device))
    L.check(rc, "dmxq_softmax_cast")
    return y


def softmax_supported(x, dim=-1) -> bool:
    """whether dmxq_softmax_cast takes this tensor (else: torch.softmax)"""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16, torch.float16) and x.dim() >= 1):
        return False
    if dim not in (-1, x.dim() - 1):
        return False
    n = x.shape[-1]
    return 32 < n <= 2048 and n % (16 // x.element_size()) == 0 and x.is_contiguous() and x.data_ptr() % 16 == 0 and x.numel() > 0


def bfp_pack(x, block_size=64, precision=8):
    """-> (mantissas, exponents): packed BFP storage of a contiguous [..., K] tensor (dmxq_bfp_pack).
    mantissas: int8 [..., K] (precision 5..8) or uint8 [..., K/2] (precision <= 4, two nibbles per byte);
    exponents: uint8 [..., K / block_size]."""
    L.require_cuda(x)
    x = x.contiguous()
    K = x.shape[-1]
    mant = torch.empty(x.shape[:-1] + ((K // 2,) if precision <= 4 else (K,)), dtype=torch.uint8 if precision <= 4 else torch.int8, device=x.device)
    exps = torch.empty(x.shape[:-1] + (K // block_size,), dtype=torch.uint8, device=x.device)
    vx = L.view(x)
    with _guard(x.device):
        rc = L.lib.dmxq_bfp_pack(C.byref(vx), mant.data_ptr(), exps.data_ptr(), block_size, precision, L.stream_ptr(x.device))
    L.check(rc, "dmxq_bfp_pack")
    return mant, exps


def _check_packed(mant, side, block_size, precision, side_name):
    """shape / layout contract of a packed (mantissas, per-block bytes) pair -> K; the C entry points take raw pointers, so
    a mismatched pair would read out of bounds on the device"""
    K = mant.shape[-1] * (2 if precision <= 4 else 1)
    if block_size <= 0 or K % block_size != 0:
        raise RuntimeError(f"packed storage: last dim {K} is not a multiple of the block size {block_size}")
    want = tuple(mant.shape[:-1]) + (K // block_size,)
    if tuple(side.shape) != want or side.dtype != torch.uint8 or mant.dtype not in (torch.uint8, torch.int8):
        raise RuntimeError(f"packed storage: {side_name} must be uint8 of shape {want} for mantissas of shape {tuple(mant.shape)}, "
                           f"got {side.dtype} {tuple(side.shape)}")
    if not (mant.is_contiguous() and side.is_contiguous()) or side.device != mant.device:
        raise RuntimeError(f"packed storage: mantissas and {side_name} must be contiguous tensors on one device")
    return K


def bfp_unpack(mant, exps, block_size=64, precision=8, dtype=torch.float32):
    """dequantise packed BFP storage (dmxq_bfp_unpack): bit-identical to bfp_qdq of the original tensor."""
    L.require_cuda(mant, "mantissas")
    L.require_cuda(exps, "exponents")
    K = _check_packed(mant, exps, block_size, precision, "exponents")
    y = torch.empty(mant.shape[:-1] + (K,), dtype=dtype, device=mant.device)
    vy = L.view(y)
    with _guard(mant.device):
        rc = L.lib.dmxq_bfp_unpack(mant.data_ptr(), exps.data_ptr(), C.byref(vy), block_size, precision, L.stream_ptr(mant.device))
    L.check(rc, "dmxq_bfp_unpack")
    return y


def sbfp_pack(x, stage, return_inexact=False, out=None, inexact=None):
    """-> (mantissas, scalers[, n_inexact]): packed SBFP storage of a contiguous [..., K] tensor (dmxq_sbfp_pack).
    `stage`: the SBFP stage struct (``ScaledBlockFloatingPoint.stage()`` / ``sbfp_stage``).
    mantissas: uint8, sign-magnitude, [..., K/2] (block precision <= 4, two nibbles per byte) or [..., K];
    scalers: uint8 [..., K / block], the scaler's exponent | mantissa fields (0 = zero scaler).
    n_inexact (device int32 scalar, on request): blocks the bytes cannot hold (see include/dmxq.h); 0 = exact.
    `out`: optional preallocated (mantissas, scalers) pair of those shapes.  `inexact`: optional existing int32 device
    counter to ACCUMULATE into (implies return_inexact)."""
    L.require_cuda(x)
    x = x.contiguous()
    K, nib = x.shape[-1], stage.precision <= 4
    mshape, sshape = x.shape[:-1] + ((K // 2,) if nib else (K,)), x.shape[:-1] + (K // max(stage.block, 1),)
    if out is not None:
        mant, scal = out
        for t, shp, nm in ((mant, mshape, "mantissas"), (scal, sshape, "scalers")):
            L.require_cuda(t, nm)
            if t.dtype != torch.uint8 or tuple(t.shape) != tuple(shp) or not t.is_contiguous():
                raise RuntimeError(f"sbfp_pack: out {nm} must be a contiguous uint8 tensor of shape {tuple(shp)}")
    else:
        mant = torch.empty(mshape, dtype=torch.uint8, device=x.device)
        scal = torch.empty(sshape, dtype=torch.uint8, device=x.device)
    if inexact is not None:
        L.require_cuda(inexact, "inexact")
        if inexact.dtype != torch.int32 or inexact.numel() != 1:
            raise RuntimeError("sbfp_pack: inexact must be a one-element int32 tensor")
        return_inexact = True
    bad = inexact if inexact is not None else (torch.zeros((), dtype=torch.int32, device=x.device) if return_inexact else None)
    vx = L.view(x)
    with _guard(x.device):
        rc = L.lib.dmxq_sbfp_pack(C.byref(vx), mant.data_ptr(), scal.data_ptr(), C.byref(stage), bad.data_ptr() if return_inexact else None,
                                  L.stream_ptr(x.device))
    L.check(rc, "dmxq_sbfp_pack")
    return (mant, scal, bad) if return_inexact else (mant, scal)


def sbfp_unpack(mant, scal, stage, dtype=torch.float32):
    """dequantise packed SBFP storage (dmxq_sbfp_unpack): bit-identical to the SBFP cast of the original tensor."""
    L.require_cuda(mant, "mantissas")
    L.require_cuda(scal, "scalers")
    K = _check_packed(mant, scal, stage.block, stage.precision, "scalers")
    y = torch.empty(mant.shape[:-1] + (K,), dtype=dtype, device=mant.device)
    vy = L.view(y)
    with _guard(mant.device):
        rc = L.lib.dmxq_sbfp_unpack(mant.data_ptr(), scal.data_ptr(), C.byref(vy), C.byref(stage), L.stream_ptr(mant.device))
    L.check(rc, "dmxq_sbfp_unpack")
    return y


def minmax(x, ch_axis: Optional[int] = None, out=None):
    """amin / amax per tensor or per channel (MinMaxObserver statistics) via dmxq_minmax.  ``out``: optional (min, max) pair of
    contiguous fp32 CUDA vectors (or slices of larger ones) to write into."""
    L.require_cuda(x)
    x = x.contiguous()
    c = 1 if ch_axis is None else x.shape[ch_axis]
    if out is not None:
        mn, mx = out
        for t in (mn, mx):
            if t.dtype != torch.float32 or t.numel() != c or not t.is_contiguous() or t.device != x.device:
                raise RuntimeError(f"dmxq: minmax out tensors must be contiguous fp32 vectors of {c} elements on the input's device")
    else:
        mn = torch.empty(c, dtype=torch.float32, device=x.device)
        mx = torch.empty(c, dtype=torch.float32, device=x.device)
    vx = L.view(x)
    with _guard(x.device):
        rc = L.lib.dmxq_minmax(C.byref(vx), -1 if ch_axis is None else ch_axis % x.dim(), mn.data_ptr(), mx.data_ptr(),
                               L.stream_ptr(x.device))
    L.check(rc, "dmxq_minmax")
    return mn, mx


def amax_multi(xs: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """max|x| of every tensor of ``xs`` as one fp32 CUDA vector, in one launch per 64 tensors (dmxq_amax_multi)."""
    n = len(xs)
    dev = xs[0].device if n else (out.device if out is not None else torch.device("cuda"))
    res = out if out is not None else torch.empty(n, dtype=torch.float32, device=dev)
    if n == 0:
        return res
    for x in xs:
        L.require_cuda(x)
        if not x.is_contiguous() or x.device != dev:
            raise RuntimeError("dmxq: amax_multi needs contiguous tensors on one device")
    if res.dtype != torch.float32 or res.numel() != n or not res.is_contiguous() or res.device != dev:
        raise RuntimeError("dmxq: amax_multi out must be a contiguous fp32 vector with one entry per tensor")
    stream = L.stream_ptr(dev)
    rc = 0
    with _guard(dev):
        for i, g in _groups(n):
            vx, keep = L.views(xs[i:i + g])
            rc = L.lib.dmxq_amax_multi(vx, g, res.data_ptr() + 4 * i, stream)
            if rc:
                break
    L.check(rc, "dmxq_amax_multi")
    return res


def _f32(v) -> float:
    """a python number rounded to fp32 (what a torch Scalar becomes inside an fp32 kernel)"""
    return struct.unpack("f", struct.pack("f", float(v)))[0]


def histc(x, bins: int = 100, min=0, max=0, return_minmax: bool = False):
    """``torch.histc(x.float(), bins, min, max)`` via dmxq_histc (the call HistogramObserver.forward makes,
    reference S/numerical/observer.py:470-491): fp32 histogram of ``bins`` equal-width bins over [min, max], values
    outside the range and NaN dropped, ``min == max`` -> the data's own range (and +-1 if that is empty too).
    ``return_minmax``: also return amin / amax of ``x`` (0-d fp32 tensors) computed in the same pass."""
    L.require_cuda(x)
    x = x.contiguous()
    lo, hi = _f32(min), _f32(max)
    mn = mx = None
    if lo == hi and x.numel():
        mn, mx = minmax(x)
        lo, hi = torch.cat([mn, mx]).tolist()  # one 8-byte read-back
        have_minmax = True
    else:
        have_minmax = False
    if lo == hi:
        lo, hi = _f32(lo - 1), _f32(hi + 1)
    if math.isinf(lo) or math.isinf(hi) or lo != lo or hi != hi:
        raise RuntimeError(f"range of [{lo}, {hi}] is not finite")
    if lo > hi:
        raise RuntimeError("max must be larger than min")
    counts = torch.zeros(bins, dtype=torch.int64, device=x.device)
    want = return_minmax and not have_minmax
    if want:
        mn = torch.empty(1, dtype=torch.float32, device=x.device)
        mx = torch.empty(1, dtype=torch.float32, device=x.device)
    vx = L.view(x)
    with _guard(x.device):
        rc = L.lib.dmxq_histc(C.byref(vx), bins, lo, hi, counts.data_ptr(), mn.data_ptr() if want else None,
                              mx.data_ptr() if want else None, L.stream_ptr(x.device))
    L.check(rc, "dmxq_histc")
    hist = counts.to(torch.float32)
    if return_minmax:
        return hist, mn.reshape(()), mx.reshape(())
    return hist


def block_quantize_l1(x, wl, dim=-1, symmetric=True, rounding="stochastic", rand=None):
    """quant_cuda.block_quantize_<rounding>(x, wl, dim, symmetric) (reference Q/quant_cuda/quant.cu:36-112)."""
    L.require_cuda(x)
    if x.dtype != torch.float32:
        raise RuntimeError("x is not a single precision Floating Point Tensor")
    if not x.is_contiguous():
        raise RuntimeError("a must be contiguous")
    if rounding == "stochastic" and rand is None:
        rand = torch.randint_like(x, 2**31 - 1, dtype=torch.int32)
    y = torch.empty_like(x)
    c = 1 if dim == -1 else x.shape[dim]
    ws = torch.empty(3 * max(c, 1), dtype=torch.int32, device=x.device)
    vx, vy = L.view(x), L.view(y)
    with _guard(x.device):
        rc = L.lib.dmxq_block_quantize(C.byref(vx), C.byref(vy), wl, dim, int(symmetric), L.ROUND[rounding],
                                       rand.data_ptr() if rand is not None else None, ws.data_ptr(), L.stream_ptr(x.device))
    L.check(rc, "dmxq_block_quantize")
    return y


def cast_chain_host(x_host: torch.Tensor, y_host: torch.Tensor, stages: Sequence[L.Stage], device: int = 0) -> torch.Tensor:
    """The e2e entry: host [rows, K] in, host out (dmxq_cast_chain_host); pinned tensors run at PCIe speed."""
    if x_host.is_cuda or y_host.is_cuda or not x_host.is_contiguous() or not y_host.is_contiguous():
        raise RuntimeError("cast_chain_host needs contiguous host tensors")
    rows = x_host.numel() // x_host.shape[-1] if x_host.numel() else 0
    n = len(stages)
    arr = (L.Stage * n)(*stages)
    rc = L.lib.dmxq_cast_chain_host(x_host.data_ptr(), y_host.data_ptr(), L.dtype_code(x_host.dtype),
                                    L.dtype_code(y_host.dtype), rows, x_host.shape[-1], arr, n, device)
    L.check(rc, "dmxq_cast_chain_host")
    return y_host

"""Drop-in installation into an *installed* reference package (``dmx.compressor``).

    import dmx.compressor            # the reference, unmodified
    import dmx_compressor_b200.plugin as dmxq
    dmxq.install()                   # CUDA tensors now take the libdmxq path

What gets patched -- exactly the plugin points of SURVEY.md section 8b, nothing above them, so
``CastTo`` / ``CastToFormat`` (STE backward) / ``DmxModule`` / ``DmxModel.from_torch`` /
``config_rules`` keep running the reference's own code:

  dmx.compressor.numerical.format.BlockFloatingPoint.cast         -> dmxq_bfp_qdq
  dmx.compressor.numerical.format.ScaledBlockFloatingPoint.cast   -> dmxq_sbfp_qdq
  dmx.compressor.numerical.format.FloatingPoint.cast              -> dmxq_float_qdq
  dmx.compressor.numerical.format.FixedPoint.cast                 -> dmxq_fixed_qdq (half-away, the reference's CUDA rule)
  dmx.compressor.numerical.format.MXINT.cast                      -> (inherits BlockFloatingPoint)
  dmx.compressor.sparse.Sparsify.forward (BlockTopK patterns)     -> dmxq_nm_prune
  dmx.compressor.quant.{fixed_point,block,float}_quantize         -> L1 mirrors (dmx_compressor_b200.quant)
  {BlockFloatingPoint,ScaledBlockFloatingPoint}.pack / .unpack     -> NEW methods: packed storage (dmxq_bfp_pack / dmxq_sbfp_pack)

Only CUDA tensors are redirected; a CPU tensor still runs the reference's own CPU extension
(this package contains no CPU implementation).  ``uninstall()`` restores the originals.
"""
from __future__ import annotations

import importlib
from typing import Dict, Tuple

import torch

from . import _lib as L
from . import ops, quant

_saved: Dict[Tuple[object, str], object] = {}


_MISSING = object()  # the attribute did not exist before install(): uninstall() deletes it again


def _patch(obj, name, new):
    key = (obj, name)
    if key not in _saved:
        _saved[key] = obj.__dict__.get(name, _MISSING) if isinstance(obj, type) else getattr(obj, name, _MISSING)
    setattr(obj, name, new)


def install(package: str = "dmx.compressor", fuse_castto: bool = True, tie_order: str = "torch_cuda") -> None:
    """``tie_order``: N:M tie order of the patched ``Sparsify.forward`` -- "torch_cuda" (default) reproduces what the
    unpatched reference computes for CUDA tensors (torch's unstable bitonic argsort) bit for bit, "stable" is the
    reference's CPU order (and the faster kernel).  FixedPoint ties (half away) and the SBFP block scale (multiplied by
    the rounded reciprocal) always follow the reference's CUDA behaviour: only CUDA tensors are redirected."""
    assert tie_order in ("torch_cuda", "stable")
    nm_order = L.NM_TORCH_CUDA if tie_order == "torch_cuda" else L.NM_STABLE
    fmt = importlib.import_module(package + ".numerical.format")
    sparse = importlib.import_module(package + ".sparse")
    refquant = importlib.import_module(package + ".quant")
    qf = importlib.import_module(package + ".quant.quant_function")

    o_bfp, o_sbfp, o_fp, o_xp = (fmt.BlockFloatingPoint.cast, fmt.ScaledBlockFloatingPoint.cast, fmt.FloatingPoint.cast,
                                 fmt.FixedPoint.cast)

    def bfp_cast(self, x, block_dim):
        if not x.is_cuda:
            return o_bfp(self, x, block_dim)
        return ops.bfp_qdq(x, block_dim, self.block_size, self.precision, self.symmetric, self.rounding, out_dtype=torch.float32)

    def sbfp_cast(self, x, block_dim):
        if not x.is_cuda:
            return o_sbfp(self, x, block_dim)
        if not self.scaler_format_exponent_bias_determined:
            self.determine_scaler_exponent_bias_from(x)
            self.scaler_format_exponent_bias_determined = True
        b, s = self.block_format, self.scaler_format
        return ops.sbfp_qdq(x, block_dim, self.block_size, b.precision, b.clamp, b.rounding, L.TIE_AWAY, s.mantissa, s.exponent,
                            s.bias, s.flush_subnormal, s.unsigned, repr(s) == "FP[1|5|10,15](FN)", s.rounding,
                            out_dtype=torch.float32)

    def fp_cast(self, x, *args):
        if not x.is_cuda:
            return o_fp(self, x, *args)
        r = repr(self)
        if (x.dtype == torch.float32 and r == "FP[1|8|23,127](_N)") or (x.dtype == torch.float16 and r == "FP[1|5|10,15](_N)"):
            return x.abs() if self.unsigned else x
        return ops.float_qdq(x, self.mantissa, self.exponent, self.bias, self.flush_subnormal, self.unsigned,
                             r == "FP[1|5|10,15](FN)", self.rounding, out_dtype=torch.float32)

    def xp_cast(self, x, *args):
        if not x.is_cuda:
            return o_xp(self, x, *args)
        return ops.fixed_qdq(x, self.precision, self.fraction, self.clamp, self.symmetric, self.rounding, L.TIE_AWAY,
                             out_dtype=torch.float32)

    _patch(fmt.BlockFloatingPoint, "cast", bfp_cast)
    _patch(fmt.ScaledBlockFloatingPoint, "cast", sbfp_cast)
    _patch(fmt.FloatingPoint, "cast", fp_cast)
    _patch(fmt.FixedPoint, "cast", xp_cast)

    # packed storage -- the real formats whose size the reference's `bytes_per_elem` reports (format.py:345-347, 481-486):
    # NEW methods on the reference's classes (nothing is overridden), CUDA tensors only
    def sbfp_stage_of(f):
        b, s = f.block_format, f.scaler_format
        return ops.sbfp_stage(f.block_size, b.precision, b.clamp, b.rounding, L.TIE_AWAY, s.mantissa, s.exponent, s.bias,
                              s.flush_subnormal, s.unsigned, repr(s) == "FP[1|5|10,15](FN)", s.rounding)

    def bfp_pack(self, x):
        """-> (mantissas, exponents): dmxq_bfp_pack; blocks along the last dim"""
        assert self.symmetric and self.rounding == "nearest" and self.precision <= 8, "packed storage: symmetric nearest, <= 8 bits"
        return ops.bfp_pack(x, self.block_size, self.precision)

    def bfp_unpack(self, mantissas, exponents, dtype=torch.float32):
        """dequantise packed storage (dmxq_bfp_unpack): equals ``cast`` of the original tensor bit for bit"""
        return ops.bfp_unpack(mantissas, exponents, self.block_size, self.precision, dtype)

    def sbfp_pack(self, x, return_inexact=False):
        """-> (mantissas, scalers[, n_inexact]): dmxq_sbfp_pack; blocks along the last dim"""
        if not self.scaler_format_exponent_bias_determined:
            self.determine_scaler_exponent_bias_from(x)
            self.scaler_format_exponent_bias_determined = True
        return ops.sbfp_pack(x, sbfp_stage_of(self), return_inexact)

    def sbfp_unpack(self, mantissas, scalers, dtype=torch.float32):
        """dequantise packed storage (dmxq_sbfp_unpack): equals ``cast`` of the original tensor bit for bit"""
        return ops.sbfp_unpack(mantissas, scalers, sbfp_stage_of(self), dtype)

    _patch(fmt.BlockFloatingPoint, "pack", bfp_pack)
    _patch(fmt.BlockFloatingPoint, "unpack", bfp_unpack)
    _patch(fmt.ScaledBlockFloatingPoint, "pack", sbfp_pack)
    _patch(fmt.ScaledBlockFloatingPoint, "unpack", sbfp_unpack)

    # Sparsify.forward: BlockTopK mask + apply in one kernel (other patterns: reference code)
    o_fwd = sparse.Sparsify.forward

    def sparsify_forward(self, x):
        sp = self.sparseness
        if not (x.is_cuda and isinstance(sp, sparse.BlockTopK)) or (self.training and torch.is_grad_enabled()):
            return o_fwd(self, x)
        if self.plastic:
            score = self.score_func(self.score, x)
            self.plastic = False
        else:
            score = self.score
        out_dtype = torch.promote_types(x.dtype, score.dtype)
        y, mask = ops.nm_prune(x, sp.K, sp.block_size, sp.block_dim, score=score.detach(), return_mask=True,
                               out_dtype=torch.float32 if out_dtype == torch.float32 else x.dtype, nm_order=nm_order)
        self.mask = mask.to(score.dtype)
        return y

    _patch(sparse.Sparsify, "forward", sparsify_forward)

    # L1 functions, wherever the reference bound them by name
    def _route(ours, theirs):
        def f(x, *a, **k):
            return ours(x, *a, **k) if x.is_cuda else theirs(x, *a, **k)
        f.__name__ = theirs.__name__
        f.__doc__ = theirs.__doc__
        return f

    for name in ("fixed_point_quantize", "block_quantize", "float_quantize"):
        routed = _route(getattr(quant, name), getattr(qf, name))
        for mod in (qf, refquant, fmt):
            if hasattr(mod, name):
                _patch(mod, name, routed)

    # HistogramObserver.forward: aminmax + histc over the observed tensor on dmxq_histc / dmxq_minmax
    # (one fused pass per steady-state step); the state it leaves in histogram / min_val / max_val and
    # everything downstream (calculate_qparams, the clipping search) stay the reference's own code.
    obs = importlib.import_module(package + ".numerical.observer")
    o_hfwd = obs.HistogramObserver.forward
    from .numerical.observer import histogram_step

    def histogram_forward(self, x_orig):
        if not x_orig.is_cuda:
            return o_hfwd(self, x_orig)
        return histogram_step(self, x_orig)

    _patch(obs.HistogramObserver, "forward", histogram_forward)

    if fuse_castto:
        # CastTo.forward's x.float() ... .to(dtype) round trip fused into the kernel for the plain
        # (no pre-transform, no observer, non-FixedPoint) case; everything else: reference code.
        cast = importlib.import_module(package + ".numerical.cast")
        o_cfwd = cast.CastTo.forward

        class _Fused(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, stage, block_dim):
                ctx.set_materialize_grads(False)
                return ops.cast_chain(x, [stage], block_dim)

            @staticmethod
            def backward(ctx, g):
                return g, None, None

        def stage_of(f):
            if isinstance(f, fmt.ScaledBlockFloatingPoint):
                return sbfp_stage_of(f)
            if isinstance(f, fmt.BlockFloatingPoint):
                return ops.bfp_stage(f.block_size, f.precision, f.symmetric, f.rounding)
            if isinstance(f, fmt.FloatingPoint):
                return ops.float_stage(f.mantissa, f.exponent, f.bias, f.flush_subnormal, f.unsigned,
                                       repr(f) == "FP[1|5|10,15](FN)", f.rounding)
            return None

        def castto_forward(self, x):
            f = self.format
            plain = (isinstance(x, torch.Tensor) and x.is_cuda and x.is_floating_point() and not self.pre_transform
                     and isinstance(f, (fmt.BlockFloatingPoint, fmt.ScaledBlockFloatingPoint, fmt.FloatingPoint))
                     and getattr(f, "rounding", "nearest") != "stochastic")
            if not plain or self.observer_enabled[0] == 1 or self.fake_quant_enabled[0] != 1:
                return o_cfwd(self, x)
            if isinstance(f, fmt.ScaledBlockFloatingPoint) and not f.scaler_format_exponent_bias_determined:
                return o_cfwd(self, x)
            if isinstance(f, fmt.FloatingPoint) and ((x.dtype == torch.float32 and repr(f) == "FP[1|8|23,127](_N)")
                                                     or (x.dtype == torch.float16 and repr(f) == "FP[1|5|10,15](_N)")):
                return o_cfwd(self, x)
            self.physical_dtype = x.dtype
            return _Fused.apply(x, stage_of(f), self.block_dim)

        _patch(cast.CastTo, "forward", castto_forward)


def uninstall() -> None:
    for (obj, name), orig in list(_saved.items()):
        if orig is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, orig)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)

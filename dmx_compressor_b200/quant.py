"""L1 quantize functions with the reference's exact signatures
(reference src/dmx/compressor/quant/quant_function.py:47-152): ``fixed_point_quantize``,
``block_quantize``, ``float_quantize``.  fp32 CUDA tensors in, fresh fp32 tensors out; the
work is done by libdmxq instead of the vendored QPyTorch kernels.
"""
import torch

from . import _lib as L
from . import ops

__all__ = ["fixed_point_quantize", "block_quantize", "float_quantize"]


def assert_wl_fl(wl, fl, stage=""):
    if wl == -1 and fl != -1:
        raise ValueError("fixed point {} wl {}, fl {}".format(stage, wl, fl))


def fixed_point_quantize(x, wl, fl, clamp=True, symmetric=False, rounding="stochastic"):
    assert isinstance(x, torch.Tensor)
    assert rounding in ["stochastic", "nearest", "up", "down"]
    assert_wl_fl(wl, fl)
    return ops.fixed_qdq(x.contiguous(), wl, fl, clamp, symmetric, rounding, tie=L.TIE_AWAY, out_dtype=torch.float32)


def block_quantize(x, wl, dim=-1, symmetric=True, rounding="stochastic"):
    assert isinstance(x, torch.Tensor), "x is not a single precision Floating Point Tensor"
    assert rounding in ["stochastic", "nearest", "down", "up"], "invalid rounding mode, {}".format(rounding)
    return ops.block_quantize_l1(x.contiguous(), wl, dim, symmetric, rounding)


def float_quantize(x, exp, man, bias=None, flush_subnormal=True, rounding="stochastic"):
    assert isinstance(x, torch.Tensor), "x is not a single precision Floating Point Tensor"
    assert rounding in ["stochastic", "nearest"], "invalid rounding mode, {}".format(rounding)
    if bias is None:
        bias = 2 ** (exp - 1) - 1
    return ops.float_qdq(x.contiguous(), man, exp, bias, flush_subnormal, False, False, rounding, out_dtype=torch.float32)

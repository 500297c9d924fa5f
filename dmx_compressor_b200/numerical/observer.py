"""Calibration observers feeding FixedPoint casts their scale / zero-point -- host-side mirror
of the reference's observer surface (reference src/dmx/compressor/numerical/observer.py:59-210).

Only the min/max family is on the CUDA path: the running amin/amax statistics come from
``dmxq_minmax`` (exact, order independent => a sharded reduction + all-reduce(MIN/MAX) equals
the single-device result bit for bit, see dmx_compressor_b200/parallel.py).  The histogram
search of the reference (observer.py:213-582) is a host-side calibration-time loop and is out
of scope.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.ao.quantization.observer import ObserverBase
from torch.ao.quantization.utils import check_min_max_valid, is_per_channel

from .. import ops
from .format import FixedPoint, Format

_SCHEMES = (torch.per_tensor_affine, torch.per_tensor_symmetric, torch.per_channel_affine, torch.per_channel_symmetric,
            torch.per_channel_affine_float_qparams)


def _get_qmin_qmax(fmt: Format) -> Tuple[Optional[int], Optional[int]]:
    """integer range of a clamped zero-fraction FixedPoint (reference observer.py:13-21)."""
    if isinstance(fmt, FixedPoint) and fmt.fraction == 0 and fmt.clamp:
        hi = 2 ** (fmt.precision - 1) - 1
        lo = -hi if fmt.symmetric else -hi - 1
        return lo, hi
    return None, None


class DMXObserverBase(ObserverBase):
    eps: torch.Tensor

    def __init__(self, dtype: Format, qscheme: torch.qscheme = torch.per_tensor_affine, factory_kwargs=None,
                 eps: float = torch.finfo(torch.float32).eps, **kwargs) -> None:
        assert isinstance(dtype, Format), f"illegal format {dtype}"
        super().__init__(dtype=dtype, **kwargs)
        assert qscheme in _SCHEMES, f"unsupported quantization scheme {qscheme}"
        self.qscheme = qscheme
        self.register_buffer("eps", torch.tensor([eps], **torch.nn.factory_kwargs(factory_kwargs)))
        self.quant_min, self.quant_max = _get_qmin_qmax(self.dtype)

    def _calculate_qparams(self, min_val: torch.Tensor, max_val: torch.Tensor):
        """scale / zero-point from running min / max (reference observer.py:59-115): symmetric =>
        scale = max(|min|, |max|) / ((qmax - qmin) / 2), zp = 0; affine => scale = (max - min) /
        (qmax - qmin), zp = clamp(qmin - round(min / scale)); both floored at eps."""
        if not check_min_max_valid(min_val, max_val):
            return torch.tensor([1.0], device=min_val.device.type), torch.tensor([0], device=min_val.device.type)
        qmin, qmax = self.quant_min, self.quant_max
        lo = torch.clamp(min_val, max=0.0)
        hi = torch.clamp(max_val, min=0.0)
        dev = lo.device
        eps = self.eps.to(dev)
        zero_point = torch.zeros(lo.size(), dtype=torch.int64, device=dev)
        if self.qscheme in (torch.per_tensor_symmetric, torch.per_channel_symmetric):
            scale = torch.max(torch.max(-lo, hi) / (float(qmax - qmin) / 2), eps)
        elif self.qscheme == torch.per_channel_affine_float_qparams:
            scale = (max_val - min_val) / float(qmax - qmin)
            scale = torch.where(scale > eps, scale, torch.ones_like(scale))
            zero_point = -1 * min_val / scale
        else:
            scale = torch.max((hi - lo) / float(qmax - qmin), eps)
            zero_point = torch.clamp(qmin - torch.round(lo / scale).to(torch.int), qmin, qmax)
        if scale.dim() == 0:
            scale = scale.reshape(1)
        if zero_point.dim() == 0:
            zero_point = zero_point.reshape(1)
        return scale, zero_point

    def extra_repr(self):
        return f"quant_min = {self.quant_min}, quant_max = {self.quant_max}"


class DummyObserver(DMXObserverBase):
    r"""Observer that observes nothing (reference observer.py:121-136)."""

    def __init__(self, dtype: Format, ch_axis: int = -1, **kwargs) -> None:
        super().__init__(dtype=dtype, **kwargs)
        self.dtype = dtype
        self.ch_axis = ch_axis

    def forward(self, x):
        return x

    def calculate_qparams(self):
        return self._calculate_qparams(torch.empty(0), torch.empty(0))


class MinMaxObserver(DMXObserverBase):
    r"""Running min / max, per tensor or per channel (reference observer.py:139-210); the
    reduction itself is the ``dmxq_minmax`` kernel."""

    min_val: torch.Tensor
    max_val: torch.Tensor

    def __init__(self, dtype: Format = None, qscheme: Optional[torch.qscheme] = torch.per_tensor_affine, ch_axis: int = -1,
                 factory_kwargs=None, eps: Optional[float] = torch.finfo(torch.float32).eps, **kwargs) -> None:
        if qscheme == torch.per_channel_affine_float_qparams:
            raise NotImplementedError("MinMaxObserver does not support qscheme: torch.per_channel_affine_float_qparams")
        if dtype is None:
            dtype = Format.from_shorthand("XP[8,0](CSN)")
        super().__init__(dtype=dtype, qscheme=qscheme, factory_kwargs=factory_kwargs, eps=eps, **kwargs)
        self.ch_axis = ch_axis
        fk = torch.nn.factory_kwargs(factory_kwargs)
        self.register_buffer("min_val", torch.tensor(float("inf"), **fk))
        self.register_buffer("max_val", torch.tensor(float("-inf"), **fk))

    def forward(self, x_orig):
        if x_orig.numel() == 0:
            return x_orig
        x = x_orig.detach()
        per_ch = is_per_channel(self.qscheme)
        mn, mx = ops.minmax(x, self.ch_axis % x.dim() if per_ch else None)
        if not per_ch:
            mn, mx = mn.reshape(()), mx.reshape(())
        self.min_val = self.min_val.to(mn.device)
        self.max_val = self.max_val.to(mx.device)
        mn, mx = torch.min(mn, self.min_val), torch.max(mx, self.max_val)
        if self.min_val.shape:
            self.min_val.copy_(mn)
            self.max_val.copy_(mx)
        else:
            self.min_val.data = mn
            self.max_val.data = mx
        return x_orig

    def calculate_qparams(self):
        return self._calculate_qparams(self.min_val, self.max_val)

    def extra_repr(self):
        return super().extra_repr() + f", min_val = {self.min_val}, max_val = {self.max_val}"

    def reset_min_max_vals(self):
        self.min_val.copy_(torch.tensor(float("inf")))
        self.max_val.copy_(torch.tensor(float("-inf")))

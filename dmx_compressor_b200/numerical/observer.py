"""Calibration observers feeding FixedPoint casts their scale / zero-point -- host-side mirror
of the reference's observer surface (reference src/dmx/compressor/numerical/observer.py:59-210).

The passes over the observed tensor are CUDA kernels: running amin/amax come from ``dmxq_minmax``
(exact, order independent => a sharded reduction + all-reduce(MIN/MAX) equals the single-device
result bit for bit, see dmx_compressor_b200/parallel.py) and the HistogramObserver's
``torch.aminmax`` + ``torch.histc`` pair from ``dmxq_histc`` (one fused pass in steady state).
What happens to the 2048 bins afterwards (re-binning onto a wider range, the clipping-range
search) is O(bins) host logic on small torch tensors, as in the reference (observer.py:213-582).
"""
from __future__ import annotations

import bisect
import struct
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


def _f32(v: float) -> float:
    return struct.unpack("f", struct.pack("f", v))[0]


# ---- one HistogramObserver step (module level: dmx_compressor_b200.plugin binds it onto the reference's own class) ----
def _widened(obs, lo: torch.Tensor, hi: torch.Tensor):
    """Grow [lo, hi] (only upwards) so that its width is a whole multiple of the current bin grid refined
    ``upsample_rate`` times; returns the new bounds, that multiple and the fine-grid offset of the old
    minimum (reference observer.py:390-413).  All arithmetic on 0-d fp32 tensors, as there."""
    fine = (obs.max_val - obs.min_val) / (obs.bins * obs.upsample_rate)
    span = obs.bins * fine
    factor = int(torch.ceil((hi - lo) / span).item())
    hi = hi + (factor * span - (hi - lo))
    first = int(torch.round((obs.min_val - lo) / fine).item())
    return lo, hi, factor, first


def _rebinned(obs, fresh: torch.Tensor, factor: int, first: int) -> torch.Tensor:
    """Spread the stored histogram uniformly onto the fine grid, drop it at its offset inside the widened range
    and re-integrate it ``factor`` fine cells per new bin (double-precision prefix sums), then add the fresh
    batch's histogram (reference observer.py:415-452)."""
    n, up = obs.bins, obs.upsample_rate
    dense = torch.zeros(n * factor, device=fresh.device)
    dense[first:n * up + first] = obs.histogram.repeat_interleave(up)
    prefix = torch.cumsum(dense, 0, dtype=torch.double)[factor - 1::factor]
    before = torch.zeros(n, device=fresh.device)
    before[1:n] = prefix[0:-1]
    return fresh + ((prefix - before) / up).to(torch.float)


def _store(obs, hist, lo, hi):
    obs.histogram.detach_().resize_(hist.shape)
    obs.histogram.copy_(hist)
    obs.min_val.detach_().resize_(lo.shape)
    obs.min_val.copy_(lo)
    obs.max_val.detach_().resize_(hi.shape)
    obs.max_val.copy_(hi)


def _count(obs, what):
    stats = getattr(obs, "stats", None)
    if stats is not None:
        stats[what] += 1


def histogram_step(obs, x_orig: torch.Tensor, histc=ops.histc, minmax=ops.minmax) -> torch.Tensor:
    """HistogramObserver.forward (reference observer.py:454-499) on the dmxq kernels.  ``int(tensor)`` truncates toward
    zero and raises OverflowError / ValueError on inf / NaN exactly as it does in the reference."""
    if x_orig.numel() == 0:
        return x_orig
    x = x_orig.detach()
    lo0, hi0 = obs.min_val.item(), obs.max_val.item()
    if (lo0 == float("inf") and hi0 == float("-inf")) or lo0 == hi0:
        mn, mx = minmax(x)
        mn, mx = mn.reshape(()), mx.reshape(())
        _store(obs, histc(x, obs.bins, min=int(mn), max=int(mx)), mn, mx)
        return x_orig
    # steady state: bin into the range the running min/max predict and learn the batch's amin/amax in the same pass
    lo, hi, factor, first = _widened(obs, obs.min_val, obs.max_val)
    a, b = int(lo), int(hi)
    hist = None
    if a != b:
        hist, mn, mx = histc(x, obs.bins, min=a, max=b, return_minmax=True)
    else:  # empty integer range: histc falls back on this batch's own range, which needs its amin/amax first
        mn, mx = minmax(x)
        mn, mx = mn.reshape(()), mx.reshape(())
    mn, mx = mn.to(obs.min_val.device), mx.to(obs.max_val.device)
    new_lo, new_hi = torch.min(mn, obs.min_val), torch.max(mx, obs.max_val)
    if hist is not None and new_lo.item() == lo0 and new_hi.item() == hi0:
        _count(obs, "fused_steps")
    else:
        lo, hi, factor, first = _widened(obs, new_lo, new_hi)
        hist = histc(x, obs.bins, min=int(lo), max=int(hi))
        _count(obs, "rebinned_steps")
    hist = hist.to(obs.histogram.device)
    if lo.item() == lo0 and hi.item() == hi0:
        hist = hist + obs.histogram
    else:
        hist = _rebinned(obs, hist, factor, first)
    _store(obs, hist, lo, hi)
    return x_orig


class HistogramObserver(DMXObserverBase):
    r"""Running histogram of the observed values + an L2-optimal clipping range for the
    quantizer (reference observer.py:213-582, itself after torch.ao's HistogramObserver;
    per-tensor schemes only).  The default observer of ``CastTo.enable_calibration``.

    Per step the reference makes three passes over ``x`` (``x.float()``, ``aminmax``,
    ``histc``).  Here the first step is ``dmxq_minmax`` + ``dmxq_histc`` on ``x`` in its own
    dtype, and every later step ONE ``dmxq_histc`` launch that bins into the range the
    running min/max predict while folding the new amin/amax in the same pass; only when the new
    batch widens the range is the histogram pass repeated over the widened range.  Range
    quirks are the reference's: the bin range is ``[int(min), int(max)]`` (truncated toward
    zero, observer.py:470,488) with torch.histc's "min == max => use the data's range" rule.
    """

    histogram: torch.Tensor
    min_val: torch.Tensor
    max_val: torch.Tensor

    # kernel entry points; the CPU unit tests of the host logic swap these for the oracle's
    _histc = staticmethod(ops.histc)
    _minmax = staticmethod(ops.minmax)

    def __init__(self, bins: int = 2048, upsample_rate: int = 128, dtype: Format = None,
                 qscheme: Optional[torch.qscheme] = torch.per_tensor_affine, ch_axis: int = -1, factory_kwargs=None,
                 eps=torch.finfo(torch.float32).eps, **kwargs) -> None:
        if qscheme not in (torch.per_tensor_affine, torch.per_tensor_symmetric):
            raise NotImplementedError(
                "HistogramObserver's qscheme only support torch.per_tensor_symmetric or torch.per_tensor_affine.")
        if dtype is None:
            dtype = Format.from_shorthand("XP[8,0](CSN)")
        super().__init__(dtype=dtype, qscheme=qscheme, factory_kwargs=factory_kwargs, eps=eps, **kwargs)
        fk = torch.nn.factory_kwargs(factory_kwargs)
        self.bins = bins
        self.upsample_rate = upsample_rate
        self.ch_axis = ch_axis
        self.register_buffer("histogram", torch.zeros(bins, **fk))
        self.register_buffer("min_val", torch.tensor(float("inf"), **fk))
        self.register_buffer("max_val", torch.tensor(float("-inf"), **fk))
        self.stats = {"fused_steps": 0, "rebinned_steps": 0}

    # ---- accumulate -------------------------------------------------------------------------
    def forward(self, x_orig: torch.Tensor) -> torch.Tensor:
        return histogram_step(self, x_orig, self._histc, self._minmax)

    # ---- clipping-range search ----------------------------------------------------------------
    @staticmethod
    def _cube_span(a, b, density):
        """density * integral_a^b t^2 dt"""
        return density * ((b * b * b - a * a * a) / 3)

    def _clip_error(self, first: int, last: int) -> float:
        """Expected squared quantization error when source bins first..last span the quantizer's 2^precision
        levels, every source bin taken as a uniform density (reference observer.py:269-329)."""
        levels = 2 ** self.dtype.precision
        w = (self.max_val.item() - self.min_val.item()) / self.bins
        q = w * (last - first + 1) / levels
        if q == 0.0:
            return 0.0
        dev = self.histogram.device
        idx = torch.arange(self.bins, device=dev)
        left = (idx - first) * w
        right = left + w
        lv_left = torch.clamp(torch.div(left, q, rounding_mode="floor"), 0, levels - 1)
        lv_right = torch.clamp(torch.div(right, q, rounding_mode="floor"), 0, levels - 1)
        density = self.histogram / w
        half = q / 2
        # the part of the bin inside its first level, the whole levels in between, the part inside its last level
        err = torch.zeros(self.bins, device=dev)
        err += self._cube_span(left - (lv_left + 0.5) * q, torch.ones(self.bins, device=dev) * half, density)
        err += (lv_right - lv_left - 1) * self._cube_span(torch.tensor(-half), torch.tensor(half), density)
        err += self._cube_span(torch.tensor(-half), right - (lv_right * q + half), density)
        return err.sum().item()

    def _non_linear_param_search(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Greedy outlier clipping: shave 1e-5 of the probability mass off whichever tail gives up more bins, as
        long as the L2 error estimate keeps falling (reference observer.py:331-388).  The quantile scans run on a
        host copy of the prefix sums (bisect instead of one device read per bin)."""
        assert self.histogram.size()[0] == self.bins, "bins mistmatch"
        w = (self.max_val - self.min_val) / self.bins
        total = torch.sum(self.histogram).item()
        cdf = torch.cumsum(self.histogram, dim=0).tolist()
        monotone = all(a <= b for a, b in zip(cdf, cdf[1:]))  # parallel prefix sums may wobble by an ulp

        def scan_up(thr, i):  # first i' >= i with i' == last or cdf[i'] >= thr
            if monotone:
                return min(bisect.bisect_left(cdf, thr, i), last)
            while i < last and cdf[i] < thr:
                i += 1
            return i

        def scan_down(thr, i):  # last i' <= i with i' == first or cdf[i'] <= thr
            if monotone:
                return max(bisect.bisect_right(cdf, thr, first, i + 1) - 1, first)
            while i > first and cdf[i] > thr:
                i -= 1
            return i

        step, lo_q, hi_q = 1e-5, 0.0, 1.0
        first, last = 0, self.bins - 1
        best = float("inf")
        while lo_q < hi_q:
            nlo, nhi = lo_q + step, hi_q - step
            # tensor-vs-python comparisons round the scalar to fp32 first
            l, r = scan_up(_f32(nlo * total), first), scan_down(_f32(nhi * total), last)
            cand = (first, last)
            if (l - first) > (last - r):
                cand, lo_q = (l, last), nlo
            else:
                cand, hi_q = (first, r), nhi
            if cand == (first, last):
                continue
            e = self._clip_error(*cand)
            if e > best:
                break
            best, (first, last) = e, cand
        return self.min_val + w * first, self.min_val + w * (last + 1)

    def calculate_qparams(self):
        if self.min_val == float("inf") and self.max_val == float("-inf"):
            return torch.tensor([1.0], device=self.min_val.device.type), torch.tensor([0], device=self.min_val.device.type)
        assert self.bins == len(self.histogram), "histogram length does not match the observer's bins"
        return self._calculate_qparams(*self._non_linear_param_search())

    def extra_repr(self):
        return super().extra_repr() + f", min_val = {self.min_val}, max_val = {self.max_val}"

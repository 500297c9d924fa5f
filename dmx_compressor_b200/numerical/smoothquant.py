"""SmoothQuant statistics and scale (host-side mirror of reference src/dmx/compressor/numerical/smoothquant.py).

SURVEY.md section 8 (f3): the per-channel ``maxabs`` of activation and weight is the O(n) part and runs
on the ``dmxq_minmax`` kernel (amin / amax per channel in one pass; maxabs = max(-amin, amax), exact and
order independent, so row-sharded tensors reduce with ``parallel.allreduce_minmax``).  ``compute_scale``
works on C-element vectors.  Applying the scale (``a / s``, ``b * s``) is a plain broadcast torch op here,
exactly as in the reference; SmoothQuant is disabled in every BASELINE configuration.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn

from .. import ops
from .format import Format


def maxabs(x: torch.Tensor, dim: int, minmax=None) -> torch.Tensor:
    """``torch.amax(x.abs(), dim=<all but dim>)`` (reference smoothquant.py:285-299) from one dmxq_minmax pass."""
    mn, mx = (minmax or ops.minmax)(x, dim % x.dim())
    return torch.maximum(-mn, mx).to(x.dtype)


class SmoothQuant(nn.Module):
    r"""Migrates quantization difficulty from input A of a matmul to input B: A / s, B * s with
    s = maxabs(A)^m / maxabs(B)^(1-m) per channel (reference smoothquant.py:7-338)."""

    calibrating: bool = False
    _minmax = staticmethod(ops.minmax)  # kernel entry point (the CPU tests of the host logic swap it for the oracle)

    def __init__(self, a_ch_axis: int, b_ch_axis: int, a_dynamic: bool = False, b_dynamic: bool = False,
                 migration_strength: float = 0.5, scale_format: Union[str, Format] = "SAME", scale_min: float = 1e-5, **kwargs) -> None:
        super().__init__()
        self.a_ch_axis = a_ch_axis
        self.b_ch_axis = b_ch_axis
        self.register_buffer("a_dynamic", torch.tensor([int(a_dynamic)], dtype=torch.long))
        self.register_buffer("b_dynamic", torch.tensor([int(b_dynamic)], dtype=torch.long))
        self.register_buffer("enabled", torch.tensor([0], dtype=torch.long))
        self.register_buffer("migration_strength", torch.tensor([migration_strength]))
        self.register_buffer("scale_min", torch.tensor([scale_min]))
        self.register_buffer("scale", torch.empty(0))
        self.register_buffer("a_maxabs", torch.empty(0), persistent=False)
        self.register_buffer("b_maxabs", torch.empty(0))
        from .cast import CastTo

        self.scale_cast = CastTo()
        self.set_scale_format(scale_format)
        # host mirrors of the flag buffers, so that a forward does not read the device to learn that nothing is enabled
        self._on, self._dyn, self._fused = False, bool(a_dynamic), False

    # ---- switches ---------------------------------------------------------------------------
    def enable(self, enabled: bool = True) -> None:
        self.enabled[0] = 1 if enabled else 0
        self._on = bool(enabled)

    def disable(self) -> None:
        self.enable(False)

    def set_dynamic(self, a_dynamic: bool = True, b_dynamic: bool = True) -> None:
        self.a_dynamic[0] = 1 if a_dynamic else 0
        self.b_dynamic[0] = 1 if b_dynamic else 0
        self._dyn = bool(a_dynamic)

    def set_scale_format(self, format: Union[str, Format] = "SAME") -> None:
        self.scale_cast.set_format(format)

    def set_migration_strength(self, migration_strength: float) -> None:
        if not 0.0 <= migration_strength <= 1.0:
            raise ValueError(f"migration_strength should be between 0 and 1, got {migration_strength}")
        self.migration_strength[0] = migration_strength

    def reset_scale(self) -> None:
        self.scale.data = torch.empty(0)

    def reset_a_maxabs(self) -> None:
        self.a_maxabs.data = torch.empty(0)

    def reset_b_maxabs(self) -> None:
        self.b_maxabs.data = torch.empty(0)

    @property
    def a_maxabs_exists(self) -> bool:
        return self.a_maxabs.numel() > 0

    @property
    def b_maxabs_exists(self) -> bool:
        return self.b_maxabs.numel() > 0

    # ---- shape helpers (running statistics of inputs whose channel count varies) -------------
    @staticmethod
    def _slicing(x: torch.Tensor, dims) -> torch.Tensor:
        if x.dim() != len(dims):
            raise RuntimeError("Input tensor should have the same number of dimensions as slicing dimensions")
        return x[tuple(slice(0, d) for d in dims)]

    @staticmethod
    def _padding(x: torch.Tensor, dims) -> torch.Tensor:
        if x.dim() != len(dims):
            raise RuntimeError("Input tensor should have the same number of dimensions as padding dimensions")
        pad = []
        for have, want in zip(reversed(x.shape), reversed(dims)):
            pad += [0, want - have]
        return nn.functional.pad(x, tuple(pad), "constant", 0)

    def _proper_shape(self, x: torch.Tensor, dim: int) -> torch.Size:
        sz = [1] * x.dim()
        sz[dim] = self.scale.numel()
        return torch.Size(sz)

    # ---- the transform ------------------------------------------------------------------------
    def scale_a(self, a: torch.Tensor) -> torch.Tensor:
        if self._on:
            a = a.to(self.scale.device) / self.scale.view(self._proper_shape(a, self.a_ch_axis))
        return a

    def scale_b(self, b: torch.Tensor) -> torch.Tensor:
        if self._on:
            b = b.to(self.scale.device) * self.scale.view(self._proper_shape(b, self.b_ch_axis))
        return b

    def _maxabs(self, x: torch.Tensor, dim: int) -> torch.Tensor:
        return maxabs(x, dim, self._minmax)

    def compute_scale(self, a_maxabs: torch.Tensor, b_maxabs: torch.Tensor) -> None:
        """s = clamp(a^m / clamp(b, eps)^(1-m), eps), then through the scale format (reference smoothquant.py:301-321)"""
        eps_dev, m_dev = self.scale_min.device, self.migration_strength.device
        b_maxabs = b_maxabs.to(eps_dev).clamp(min=self.scale_min)
        s = (a_maxabs.to(m_dev) ** self.migration_strength) / (b_maxabs.to(m_dev) ** (1.0 - self.migration_strength))
        self.scale = self.scale_cast(s.to(eps_dev).clamp(min=self.scale_min))

    def forward(self, a: torch.Tensor, b: torch.Tensor):
        with torch.no_grad():
            cur_a, cur_b = self._maxabs(a, self.a_ch_axis), self._maxabs(b, self.b_ch_axis)
            if not self.a_maxabs_exists or self.a_dynamic[0] == 1:
                self.a_maxabs = cur_a
            else:
                self.a_maxabs = torch.maximum(self._padding(cur_a, self.a_maxabs.size()), self.a_maxabs)
            if not self.b_maxabs_exists or self.b_dynamic[0] == 1:
                self.b_maxabs = cur_b
            else:
                self.b_maxabs = torch.maximum(self._padding(cur_b, self.b_maxabs.size()), self.b_maxabs)
            self.compute_scale(self._slicing(self.a_maxabs, cur_a.size()), self._slicing(self.b_maxabs, cur_b.size()))
            return self.scale_a(a), self.scale_b(b)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if prefix + "scale" in state_dict:
            self.scale = state_dict[prefix + "scale"]  # its shape is only known after calibration
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._on, self._dyn = bool(self.enabled[0].item()), bool(self.a_dynamic[0].item())
        if hasattr(self, "fused_to_weight"):
            self._fused = bool(self.fused_to_weight[0].item())

    def extra_repr(self) -> str:
        return (f"migration_strength = {self.migration_strength.item()}, a_ch_axis = {self.a_ch_axis}, b_ch_axis = {self.b_ch_axis}, "
                f"scale_format = {self.scale_cast.format}, dynamic = ({self.a_dynamic.bool().item()}, {self.b_dynamic.bool().item()})")


class ActivationWeightSmoothQuant(SmoothQuant):
    r"""Activation x weight flavour used by modules with a weight (reference smoothquant.py:372-545)."""

    def __init__(self, ch_axis: int, win_ch_axis: int, migration_strength: float = 0.5, scale_format: Union[str, Format] = "SAME",
                 dynamic: bool = False, scale_min: float = 1e-5, **kwargs) -> None:
        super().__init__(a_ch_axis=ch_axis, b_ch_axis=win_ch_axis, migration_strength=migration_strength, scale_format=scale_format,
                         a_dynamic=dynamic, b_dynamic=False, scale_min=scale_min, **kwargs)
        self.ch_axis = ch_axis
        self.win_ch_axis = win_ch_axis
        self.register_buffer("fused_to_weight", torch.tensor([0], dtype=torch.long))

    def set_dynamic(self, dynamic: bool = True) -> None:
        if dynamic and self.fused_to_weight[0] == 1:
            raise RuntimeError("SmoothQuant cannot be dynamic as scale has been fused to weight already")
        super().set_dynamic(a_dynamic=dynamic, b_dynamic=False)

    def reset_weight_maxabs(self) -> None:
        self.reset_b_maxabs()

    @property
    def dynamic(self) -> torch.Tensor:
        return self.a_dynamic

    @property
    def weight_maxabs_computed(self) -> bool:
        return self.b_maxabs_exists

    @property
    def input_maxabs_exists(self) -> bool:
        return self.a_maxabs_exists

    def scale_weight(self, wgt):
        return self.scale_b(wgt).to(wgt.device).to(wgt.dtype)

    def scale_input(self, inp):
        return self.scale_a(inp).to(inp.device)

    def fuse_to_weight(self, wgt: torch.Tensor) -> None:
        wgt.data = self.scale_weight(wgt.data)
        self.fused_to_weight[0] = 1
        self._fused = True

    def compute_scale(self, inp_maxabs: torch.Tensor) -> None:
        super().compute_scale(inp_maxabs, self.weight_maxabs)

    def forward(self, inp: torch.Tensor, wgt: torch.Tensor) -> None:
        """refresh the statistics and the scale from one (activation, weight) pair (reference smoothquant.py:518-538:
        the statistics live in plain attributes there, so each call starts from this pair's own maxima unless a
        previous ``input_maxabs`` was registered as ``a_maxabs``)"""
        with torch.no_grad():
            if not self.weight_maxabs_computed:
                self.weight_maxabs = self._maxabs(wgt, self.win_ch_axis)
            cur = self._maxabs(inp, self.ch_axis)
            if not self.input_maxabs_exists or self.dynamic[0] == 1:
                self.input_maxabs = cur
            else:
                self.input_maxabs = torch.maximum(cur, self.input_maxabs)
            self.compute_scale(self.input_maxabs)

    def extra_repr(self) -> str:
        return (f"migration_strength = {self.migration_strength.item()}, ch_axis = {self.ch_axis}, win_ch_axis = {self.win_ch_axis}, "
                f"scale_format = {self.scale_cast.format}, dynamic = {self.dynamic.bool().item()}")

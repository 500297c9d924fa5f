"""CastTo -- the simulated numerical cast module, host-side mirror of the reference's plugin
surface (reference src/dmx/compressor/numerical/cast.py:20-358): same class names, constructor
arguments, buffers (``scale``, ``zero_point``, ``fake_quant_enabled``, ``observer_enabled`` --
inherited from torch's FakeQuantize exactly like the reference, so state_dicts interchange),
``set_format`` / ``set_pre_transform`` / ``enable_calibration`` API and forward semantics.

What differs is the work per forward.  The reference runs ``x.float()`` -> python loop of
native calls -> ``.to(dtype)`` (cast.py:262-306).  Here a forward is ONE kernel launch: dtype
widening, the block reduction, the rounding, the FixedPoint affine wrap and the narrowing back
to the tensor dtype all happen in registers (``dmxq_cast_chain`` / ``dmxq_fixed_qdq``).
"""
from __future__ import annotations

import math
import warnings
from collections import OrderedDict
from typing import Dict, Optional, Union

import torch
from torch.autograd import Function
from torch.quantization.fake_quantize import FakeQuantize

from .. import elide, ops
from .format import FixedPoint, Format, Same
from .observer import DummyObserver, HistogramObserver, MinMaxObserver, ObserverBase  # noqa: F401


class CastToFormat(Function):
    r"""Straight-through-estimator cast (reference cast.py:20-32): forward = ``fmt.cast``,
    backward = identity."""

    @staticmethod
    def forward(ctx, x, fmt, block_dim):
        ctx.set_materialize_grads(False)
        return fmt.cast(x, block_dim)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _FusedCast(Function):
    """STE cast that also fuses CastTo.forward's dtype round trip (x.float() ... .to(dtype),
    reference cast.py:262,306) into the kernel: the result has x's dtype."""

    @staticmethod
    def forward(ctx, x, fmt, block_dim):
        ctx.set_materialize_grads(False)
        return ops.cast_chain(x, [fmt.stage()], block_dim)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _FusedFixedAffine(Function):
    """x / sc + zp -> FixedPoint -> (q - zp) * sc (reference cast.py:279-296) in one kernel, with
    scale / zero-point read from device memory (no host sync)."""

    @staticmethod
    def forward(ctx, x, fmt, sc, zp, ch_axis, group_size):
        ctx.set_materialize_grads(False)
        return ops.fixed_qdq(x, fmt.precision, fmt.fraction, fmt.clamp, fmt.symmetric, fmt.rounding, fmt.tie, sc, zp,
                             ch_axis=ch_axis, group_size=group_size)

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None, None


class CastToDict(torch.nn.ModuleDict):
    r"""Ordered set of CastTo modules applied to a module's tensor arguments (reference cast.py:58-134)."""

    def forward(self, x, *args, output=False, **kwargs):
        keys = list(self.keys())
        if output:
            if isinstance(x, (tuple, list)):
                return type(x)(self[keys[i]](a) for i, a in enumerate(x))
            return self[keys[0]](x, lazy_ok=True)  # an output cast may be deferred to its consumer under elision
        i = 1
        new_args, new_kwargs = [], {}
        for a in args:
            if isinstance(a, torch.Tensor):
                new_args.append(self[keys[i]](a))
                i += 1
            else:
                new_args.append(a)
        for k, v in kwargs.items():
            new_kwargs[k] = self[k + "_cast"](v) if isinstance(v, torch.Tensor) else v
        return self[keys[0]](x), new_args, new_kwargs

    def pack_to_dict(self, param):
        keys = list(self.keys())
        if isinstance(param, (tuple, list)):
            param = {keys[i]: p if p is not None else "SAME" for i, p in enumerate(param)}
        elif not isinstance(param, dict):
            raise ValueError("format needs to be a dict, tuple or list!")
        if len(param) != len(self):
            warnings.warn("length of format to set is not equal to length of input_casts, some CastTos might not be set "
                          f"properly!\nlen({param}!={len(self)})")
        return param

    def set_pre_transform(self, pre_transforms: Union[Dict, tuple, list]):
        for k, t in self.pack_to_dict(pre_transforms).items():
            self[k].set_pre_transform(t)

    def set_format(self, format: Union[Dict, tuple, list]):
        for k, f in self.pack_to_dict(format).items():
            if k not in self.keys():
                raise RuntimeError(f"No CastTo with key {k}!")
            self[k].set_format(f)

    def disable_fake_quant(self):
        for k in self.keys():
            self[k].disable_fake_quant()

    def enable_fake_quant(self):
        for k in self.keys():
            self[k].enable_fake_quant()

    def enable_observer(self):
        for k in self.keys():
            self[k].enable_observer()

    def disable_observer(self):
        for k in self.keys():
            self[k].disable_observer()


class CastTo(FakeQuantize):
    r"""Simulated numerical cast to a target format (reference cast.py:136-358)."""

    def __init__(self, format="SAME", observer=DummyObserver, group_size=None, block_dim=-1, **fake_quantize_kwargs):
        self.set_format(format)
        super().__init__(observer=observer, dtype=self.format, **fake_quantize_kwargs)
        if group_size:
            assert torch.ao.quantization.utils.is_per_tensor(self.qscheme), "group_size must be used with per tensor quantization scheme"
        self.group_size = group_size if group_size else None
        self.physical_dtype = None
        self.block_dim = block_dim
        self.enable_fake_quant()
        self.disable_observer()
        self.pre_transform = {}

    # The enable flags live in device buffers (FakeQuantize state, kept for state_dict
    # compatibility); reading them in forward() would force a device->host sync per cast, as
    # the reference does (cast.py:271,277).  Host-side mirrors are kept in step instead.
    def enable_fake_quant(self, enabled: bool = True) -> None:
        super().enable_fake_quant(enabled)
        self._fq_on = bool(enabled)

    def enable_observer(self, enabled: bool = True) -> None:
        super().enable_observer(enabled)
        self._obs_on = bool(enabled)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._fq_on = bool(self.fake_quant_enabled[0] == 1)
        self._obs_on = bool(self.observer_enabled[0] == 1)

    def set_format(self, format: Union[str, torch.dtype, Format]):
        if isinstance(format, str):
            format = Format.from_shorthand(format)
        self.format = format
        if hasattr(self, "dtype"):
            self.dtype = format
            self.activation_post_process.dtype = format

    def set_pre_transform(self, pre_transform: Dict):
        self.pre_transform = pre_transform.copy()
        if isinstance(self.pre_transform.get("format"), str):
            self.pre_transform["format"] = Format.from_shorthand(self.pre_transform["format"])

    # ------------------------------------------------------------------ calibration (reference cast.py:179-226)
    def _observer_step(self, x):
        self.activation_post_process.to(x.device)
        if self.group_size:
            if not hasattr(self, "activation_post_processes"):
                n_groups = math.ceil(x.shape[self.ch_axis] / self.group_size)
                self.activation_post_processes = [
                    self.activation_post_process.__class__(dtype=self.format, qscheme=self.qscheme, ch_axis=self.ch_axis).to(x.device)
                    for _ in range(n_groups)]
            scale, zero_point, mins, maxs = [], [], [], []
            for obs, chunk in zip(self.activation_post_processes, torch.split(x, self.group_size, dim=self.ch_axis)):
                obs(chunk)
                s, zp = obs.calculate_qparams()
                scale.append(s)
                zero_point.append(zp)
                mins.append(obs.min_val)
                maxs.append(obs.max_val)
            _scale, _zero_point = torch.tensor(scale), torch.tensor(zero_point)
            self.activation_post_process.min_val = torch.tensor(mins)
            self.activation_post_process.max_val = torch.tensor(maxs)
        else:
            self.activation_post_process(x.detach())
            _scale, _zero_point = self.calculate_qparams()
        _scale, _zero_point = _scale.to(self.scale.device), _zero_point.to(self.zero_point.device)
        if self.scale.shape != _scale.shape:
            self.scale = torch.zeros_like(_scale)
            self.zero_point = torch.zeros_like(_zero_point)
        self.scale.copy_(_scale)
        self.zero_point.copy_(_zero_point)

    def apply_shaping_seq(self, x, shaping_list):
        undo = []
        for op, args in shaping_list:
            shp = x.size()
            if op == "view":
                x = x.reshape(*args)
                undo.append(("view", shp))
            elif op == "permute":
                x = x.permute(*args)
                undo.append(("permute", torch.LongTensor(args).argsort().tolist()))
            elif op == "flatten":
                x = x.flatten(*args)
                undo.append(("view", shp))
            else:
                raise Exception(f"unknown shape op {op}")
        return x, undo[::-1]

    # ------------------------------------------------------------------ the hot path (reference cast.py:261-306)
    def _cast(self, x):
        fmt = self.format
        if isinstance(fmt, Same):
            return CastToFormat.apply(x, fmt, self.block_dim)  # x.clone(), reference format.py:89-90
        if isinstance(fmt, FixedPoint):
            ch_axis, group = -1, None
            if self.is_per_channel:
                ch_axis = self.ch_axis % x.dim()
                n = x.shape[ch_axis]
                sc, zp = self.scale[:n], self.zero_point[:n]
            elif self.group_size:
                ch_axis, group = self.ch_axis % x.dim(), self.group_size
                sc, zp = self.scale, self.zero_point
            else:
                sc, zp = self.scale, self.zero_point
            if not (torch.is_grad_enabled() and x.requires_grad):  # no tape needed: skip the autograd.Function hop
                return ops.fixed_qdq(x, fmt.precision, fmt.fraction, fmt.clamp, fmt.symmetric, fmt.rounding, fmt.tie, sc, zp,
                                     ch_axis=ch_axis, group_size=group)
            return _FusedFixedAffine.apply(x, fmt, sc, zp, ch_axis, group)
        if isinstance(fmt, Format):
            if fmt.stage() is None:
                return CastToFormat.apply(x, fmt, self.block_dim)
            if hasattr(fmt, "_identity_for") and fmt._identity_for(x.dtype) and not fmt.unsigned:
                return x
            if not (torch.is_grad_enabled() and x.requires_grad):
                return ops.cast_chain(x, [fmt.stage()], self.block_dim)
            return _FusedCast.apply(x, fmt, self.block_dim)
        return super().forward(x)  # plain torch.dtype fake-quant

    def _forward_elided(self, x, lazy_ok=False):
        """value-identical fast path, only under elide.enabled() + no_grad (see elide.py)"""
        fmt = self.format
        if not self._fq_on or isinstance(fmt, Same):
            elide.stats["elided"] += 1
            return x  # (a deferred cast passes through untouched)
        pend = x if isinstance(x, elide.Lazy) else None
        key = elide.format_key(fmt, self.block_dim if fmt.blocked else None)
        if key is None:
            return None
        if pend is not None:
            if pend._key == key:  # producer's output format == this input format: cast once
                elide.stats["elided"] += 1
                return pend.materialise()
            if pend._kind != "cast":  # an operation is pending in front of the cast (softmax / add): its own fused kernel, or run it first
                y = pend._fuse(fmt.stage(), self.block_dim) if (pend._fuse is not None and pend._real is None) else None
                if y is None:
                    y = ops.cast_chain(pend.materialise(), [fmt.stage()], self.block_dim)
                else:
                    elide.stats["elided"] += 1
                elide.stats["casts"] += 1
                elide.tag(y, key)
                return y
            ckey = ("chain", pend._key, key)
            y = elide.memo_get(pend._raw, ckey)
            if y is None:  # output cast fused with this input cast: ONE pass over the tensor
                y = ops.cast_chain(pend._raw, [pend._fmt.stage(), fmt.stage()], self.block_dim)
                elide.stats["casts"] += 1
                elide.stats["elided"] += 1
                elide.memo_put(pend._raw, ckey, y)
                elide.tag(y, key)
            return y
        if elide.is_tagged(x, key) or (hasattr(fmt, "_identity_for") and fmt._identity_for(x.dtype) and not fmt.unsigned):
            elide.stats["elided"] += 1
            return x
        if lazy_ok and key[0] == "FP" and fmt.flush_subnormal and not fmt.unsigned and type(x) is torch.Tensor:
            return elide.Lazy(x, fmt, self.block_dim, key)  # defer: the consumer decides how to run it
        y = elide.memo_get(x, key)
        if y is None:
            y = ops.cast_chain(x, [fmt.stage()], self.block_dim)
            elide.stats["casts"] += 1
            elide.memo_put(x, key, y)
            elide.tag(y, key)
        return y

    def forward(self, x, lazy_ok=False):
        self.__dict__["physical_dtype"] = x.dtype  # (plain attribute; nn.Module.__setattr__ costs microseconds per call)
        if (elide.active() and not torch.is_grad_enabled() and not self.pre_transform and not self._obs_on
                and isinstance(x, torch.Tensor) and x.is_cuda and x.is_floating_point()):
            y = self._forward_elided(x, lazy_ok and elide.defer_output_casts)
            if y is not None:
                return y
        if isinstance(x, elide.Lazy):
            x = x.materialise()
        undo = shortcut = None
        if "shaping" in self.pre_transform:
            x, undo = self.apply_shaping_seq(x, self.pre_transform["shaping"])
        if "noquant_shortcut" in self.pre_transform:
            shortcut = x[self.pre_transform["noquant_shortcut"]].clone()
        if "format" in self.pre_transform:
            x = CastToFormat.apply(x, self.pre_transform["format"], self.block_dim)
        if self._obs_on and x is not None and not isinstance(self.format, Same):
            self._observer_step(x)
        if self._fq_on:
            x = self._cast(x)
        if shortcut is not None:
            x = x.clone()
            x[self.pre_transform["noquant_shortcut"]] = shortcut
        if undo is not None:
            x, _ = self.apply_shaping_seq(x, undo)
        return x.to(self.physical_dtype)

    def enable_calibration(self, state: bool = True, observer_cls: ObserverBase = HistogramObserver,
                           qscheme_to_overload: Optional[torch.qscheme] = None, group_size: int = None, ch_axis: int = None) -> None:
        """reference cast.py:308-340; the default observer is the HistogramObserver, as there."""
        if state:
            if ch_axis is not None:
                self.ch_axis = self.activation_post_process.ch_axis = ch_axis
            if qscheme_to_overload is not None:
                self.qscheme = qscheme_to_overload
                self.is_per_channel = torch.ao.quantization.utils.is_per_channel(qscheme_to_overload)
            self.group_size = group_size if group_size else None
            if self.group_size:
                assert torch.ao.quantization.utils.is_per_tensor(qscheme_to_overload), "group quantization is to be used with per tensor quantization"
            self.activation_post_process = observer_cls(dtype=self.format, qscheme=self.qscheme, ch_axis=self.ch_axis)
            self.disable_fake_quant()
            self.enable_observer()
        else:
            self.enable_fake_quant()
            self.disable_observer()

    def get_precision(self) -> Optional[int]:
        if isinstance(self.format, (Same, torch.dtype)):
            if self.physical_dtype is not None:
                return torch.finfo(self.physical_dtype).bits
            raise RuntimeError("physical_dtype has not been inferred, pass some data through first")
        return self.format.bit_precision

    def extra_repr(self):
        if self.format.blocked:
            return (f"format = dtype = {repr(self.format)}, block_dim = {self.block_dim} \nfake_quant_enabled = "
                    f"{bool(self.fake_quant_enabled)},pre_transform = {self.pre_transform}")
        return (f"format = dtype = {repr(self.format)}, qscheme = {self.qscheme}, ch_axis = {self.ch_axis} \n"
                f"fake_quant_enabled = {bool(self.fake_quant_enabled)}, observer_enabled = {bool(self.observer_enabled)}, "
                f"scale = {self.scale.cpu().numpy()}, zero_point = {self.zero_point.cpu().numpy()}, group_size = "
                f"{self.group_size}, pre_transform = {self.pre_transform}")

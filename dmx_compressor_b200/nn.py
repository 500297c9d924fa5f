"""The callers of the cast path -- a compact mirror of the reference's ``DmxModule`` runtime for
the module types the BASELINE configs exercise (reference
src/dmx/compressor/modeling/nn/core.py:34-264, torch_modules.py).

Same contract as the reference: every module owns ``input_casts`` / ``output_casts``
(``CastToDict``), and parameterised modules also ``weight_storage_cast``, ``weight_cast``,
``bias_cast``, ``accum_cast`` and a ``weight_sparsifier``; ``configure(dict)`` takes the same
keys as ``DmxModule.configure`` (core.py:65-108); ``forward`` is
``input_casts -> _forward (weight hypernet inside) -> output_casts -> back to the input dtype``
(core.py:215-264); ``fold_weight_and_bias`` applies the parameter casts once (core.py:146-176).
SmoothQuant (statistics on the dmxq_minmax kernel) hooks in where the reference puts it (core.py:188-191,
227-230).  Graph tracing, approximation functions, plugins and perf proxies are out of scope.

Two opt-in accelerations that the reference does not have, both
value-preserving, see dmx_compressor_b200/elide.py:
  * weight-cast caching while parameters are unchanged,
  * elision of casts that are provably no-ops (SAME clones, idempotent format repeats) and
    sharing of one cast result between consumers of the same tensor.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn.functional as F

from . import elide, fused, ops
from .numerical import CastTo, CastToDict, Format, Same
from .numerical.smoothquant import ActivationWeightSmoothQuant
from .sparse import BlockTopK, Dense, LazySparsify, Sparsify


def _param_state(t):
    if t is None or isinstance(t, torch.nn.parameter.UninitializedParameter):
        return None
    return (t.data_ptr(), t._version)


def _cast_state(c):
    """what a CastTo's result depends on besides its input (key material of the weight cache)"""
    if c is None:
        return None
    pt = c.pre_transform
    return (repr(c.format), bool(c._fq_on), int(c.block_dim), tuple(sorted((k, repr(v)) for k, v in pt.items())) if pt else None,
            _param_state(getattr(c, "scale", None)), _param_state(getattr(c, "zero_point", None)), getattr(c, "group_size", None),
            getattr(getattr(c, "format", None), "tie", None))


class DmxModule:
    r"""Mixin that adds the boundary casts and the weight hypernet to a torch.nn.Module
    (reference NumericalCastMixin cast.py:401-467 + WeightSparseMixin sparse.py:366-421 +
    DmxModule core.py)."""

    ch_axis = win_ch_axis = wout_ch_axis = None

    def _init_dmx(self) -> None:
        self.align_boundary_dtype = True
        self.input_casts = CastToDict(OrderedDict({"input_cast": CastTo(ch_axis=self.ch_axis)}))
        self.output_casts = CastToDict(OrderedDict({"output_cast": CastTo()}))
        pnames = [n for n, _ in self.named_parameters(recurse=False)]
        has_w = "weight" in pnames
        self.accum_cast = CastTo() if isinstance(self, (torch.nn.Linear, torch.nn.modules.conv._ConvNd)) else None
        self.weight_storage_cast = CastTo(ch_axis=self.wout_ch_axis) if has_w else None
        self.weight_cast = CastTo(ch_axis=self.wout_ch_axis) if has_w else None
        self.bias_cast = CastTo() if "bias" in pnames else None
        self.weight_sparsifier = LazySparsify() if has_w else None
        self.smoothquant = (ActivationWeightSmoothQuant(self.ch_axis, self.win_ch_axis)
                            if has_w and self.ch_axis is not None and self.win_ch_axis is not None else None)
        self._wcache = None
        self.__dict__.pop("_bcache", None)

    # ------------------------------------------------------------------ configuration (core.py:65-108)
    def configure(self, config) -> None:
        if "input_formats" in config:
            self.input_casts.set_format(format=config["input_formats"])
        if "pre_input_transform" in config:
            self.input_casts.set_pre_transform(config["pre_input_transform"])
        if "output_formats" in config:
            self.output_casts.set_format(format=config["output_formats"])
        if "pre_output_transform" in config:
            self.output_casts.set_pre_transform(config["pre_output_transform"])
        if self.accum_cast is not None and "accum_format" in config:
            self.accum_cast.set_format(format=config["accum_format"])
        if self.weight_storage_cast is not None and "weight_storage_format" in config:
            self.weight_storage_cast.set_format(format=config["weight_storage_format"])
        if self.weight_cast is not None and "weight_format" in config:
            self.weight_cast.set_format(format=config["weight_format"])
        if self.weight_cast is not None and "pre_weight_transform" in config:
            self.weight_cast.set_pre_transform(config["pre_weight_transform"])
        if self.bias_cast is not None and "bias_format" in config:
            self.bias_cast.set_format(format=config["bias_format"])
        if self.weight_sparsifier is not None and "weight_sparseness" in config:
            self.weight_sparsifier.configure(sparseness=config["weight_sparseness"])
        if self.weight_sparsifier is not None and "weight_score_func" in config:
            self.weight_sparsifier.configure(score_func=config["weight_score_func"])
        if self.smoothquant is not None and "smoothquant_scale_format" in config:
            self.smoothquant.set_scale_format(format=config["smoothquant_scale_format"])
        self._wcache = None
        self.__dict__.pop("_bcache", None)

    transform = configure

    @property
    def input_formats(self):
        return [c.format for c in self.input_casts.values()]

    @property
    def output_formats(self):
        return [c.format for c in self.output_casts.values()]

    @property
    def accum_format(self):
        return self.accum_cast.format if self.accum_cast is not None else None

    @property
    def weight_format(self):
        return self.weight_cast.format if self.weight_cast is not None else None

    @property
    def bias_format(self):
        return self.bias_cast.format if self.bias_cast is not None else None

    @property
    def weight_sparseness(self):
        return self.weight_sparsifier.sparseness if self.weight_sparsifier is not None else None

    # ------------------------------------------------------------------ weight hypernet (core.py:178-213)
    def _sq_stage(self, t, ch_axis, multiply):
        """SmoothQuant's scale as a chain pre-stage (ops.scale_stage) when its channel axis is t's last dim, else None"""
        sq = self.smoothquant
        sc = getattr(sq, "scale", None)
        if not (isinstance(sc, torch.Tensor) and sc.is_cuda and sc.dtype == torch.float32 and t.dim() >= 1 and ch_axis in (-1, t.dim() - 1)
                and sc.numel() == t.shape[-1] and sc.device == t.device):
            return None
        return ops.scale_stage(sc.reshape(-1).contiguous(), multiply=multiply)

    def _fusable_hypernet_stages(self, scale=None):
        """sparsify -> [SmoothQuant scale] -> storage cast -> weight cast as dmxq stages, or None when some piece needs
        the module-by-module path (score parameter, pre-transforms, observers, FixedPoint affine)."""
        stages = []
        sp = self.weight_sparsifier
        if sp is not None and not isinstance(sp.sparseness, Dense):
            from .sparse import abs_score

            if not isinstance(sp.sparseness, BlockTopK) or sp.sparseness.block_dim not in (-1, self.weight.dim() - 1):
                return None
            if not (sp.plastic and sp.score_func is abs_score):
                return None
            stages.append(ops.nm_stage(sp.sparseness.K, sp.sparseness.block_size, sp.sparseness.nm_order))
        if scale is not None:
            stages.append(scale)
        for c in (self.weight_storage_cast, self.weight_cast):
            if c is not None and c._obs_on and not isinstance(c.format, Same):
                return None  # calibrating: the observer must see the weight (even with fake-quant off)
            if c is None or isinstance(c.format, Same) or not c._fq_on:
                continue
            if c.pre_transform or not isinstance(c.format, Format):
                return None
            st = c.format.stage() if not hasattr(c.format, "tie") else None  # FixedPoint carries affine state
            if st is None or (c.format.blocked and c.block_dim not in (-1, self.weight.dim() - 1)):
                return None
            stages.append(st)
        return stages

    def _fused_scaled_input(self, x):
        """scale_input followed by the (single, plain) input cast as one cast chain -> fp32 like the reference's promoted quotient;
        None when some piece needs the module-by-module path"""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.is_floating_point() and len(self.input_casts) == 1):
            return None
        c = next(iter(self.input_casts.values()))
        f = c.format
        if c.pre_transform or c._obs_on or not c._fq_on or isinstance(f, Same) or not isinstance(f, Format) or hasattr(f, "tie"):
            return None
        if f.blocked and c.block_dim not in (-1, x.dim() - 1):
            return None
        sc = self._sq_stage(x, self.smoothquant.a_ch_axis, False)
        if sc is None:
            return None
        try:
            c.__dict__["physical_dtype"] = torch.float32
            return ops.cast_chain(x, [sc, f.stage()], -1, out_dtype=torch.float32)
        except RuntimeError:
            return None

    @property
    def effective_weight(self):  # reference sparse.py:389-395
        return self.weight_sparsifier(self.weight) if self.weight_sparsifier is not None else self.weight

    def _smoothing_weight(self) -> bool:
        sq = self.smoothquant
        return sq is not None and sq._on and not sq._fused

    def weight_hypernet(self, _w):
        if self._smoothing_weight():  # sparsify -> smoothquant scale -> storage cast -> weight cast (core.py:184-196)
            if elide.active() and not torch.is_grad_enabled() and _w.is_cuda:
                # ONE kernel: [prune ->] w * scale (rounded to w.dtype, scale_weight's `.to(wgt.dtype)`) -> storage cast -> weight cast
                sc = self._sq_stage(_w, self.smoothquant.b_ch_axis, True)
                sp = self.weight_sparsifier
                dense = sp is None or isinstance(sp.sparseness, Dense)
                stages = self._fusable_hypernet_stages(sc) if (sc is not None and dense) else None
                if stages:
                    try:
                        return ops.cast_chain(_w, stages, -1)
                    except RuntimeError:
                        pass  # a layout the rows kernels do not take: module by module below
            if self.weight_sparsifier is not None:
                _w = self.weight_sparsifier(_w)
            _w = self.smoothquant.scale_weight(_w)
            if self.weight_storage_cast is not None:
                _w = self.weight_storage_cast(_w)
            return self.weight_cast(_w) if self.weight_cast is not None else _w
        if elide.active() and not torch.is_grad_enabled():
            stages = self._fusable_hypernet_stages()
            if stages is not None:
                if not stages:
                    return _w
                sp = self.weight_sparsifier
                if sp is not None and not isinstance(sp.sparseness, Dense):
                    # everything the sparsifier's own forward would leave behind: the lazily created score parameter (same
                    # RNG draw), the mask (Sparsify.mask / density / mask_str consumers) and the one-shot `plastic` flag
                    if isinstance(sp, LazySparsify) and sp.has_uninitialized_params():
                        sp._infer_parameters(sp, (_w,), {})
                    mask = torch.empty(_w.shape, dtype=torch.float32, device=_w.device)
                    out = ops.cast_chain(_w, stages, -1, mask=mask)  # ONE kernel: prune (+ mask) -> storage cast -> weight cast
                    sp.mask = mask
                    sp.plastic = False  # the reference rewires once, on this forward
                    return out
                return ops.cast_chain(_w, stages, -1)  # ONE kernel: storage cast -> weight cast
        if self.weight_sparsifier is not None:
            _w = self.weight_sparsifier(_w)
        if self.weight_storage_cast is not None:
            _w = self.weight_storage_cast(_w)
        if self.weight_cast is not None:
            _w = self.weight_cast(_w)
        return _w

    @property
    def _weight(self):
        if elide.active() and not torch.is_grad_enabled() and not self._smoothing_weight():
            w = self.weight
            casts = (self.weight_storage_cast, self.weight_cast)
            if any(c is not None and c._obs_on for c in casts):
                return self.weight_hypernet(w)  # calibrating: the observers must see every forward
            sp = self.weight_sparsifier
            # everything the result depends on: the weight (storage + in-place version), each weight-path cast (format,
            # fake-quant switch, block dim, pre-transform, FixedPoint qparams) and the sparsifier (pattern, tie order, the
            # one-shot `plastic` flag and its score function, the score parameter's version)
            key = (w.data_ptr(), w._version, tuple(w.shape)) + tuple(_cast_state(c) for c in casts) + (
                None if sp is None else (repr(sp.sparseness), getattr(sp.sparseness, "tie_order", None), sp.plastic,
                                         id(getattr(sp, "score_func", None)) if sp.plastic else None, _param_state(getattr(sp, "score", None))),)
            if self._wcache is not None and self._wcache[0] == key:
                return self._wcache[1]
            out = self.weight_hypernet(w)
            self._wcache = (key, out)
            return out
        return self.weight_hypernet(self.weight)

    @property
    def _bias(self):
        b = getattr(self, "bias", None)
        if b is None:
            return None
        c = self.bias_cast
        if c is None:
            return b
        if elide.active() and not torch.is_grad_enabled() and not c._obs_on:
            # like the weight: cast once while the bias, its cast and the cast's switches are unchanged (72 tiny launches per
            # OPT-125m forward otherwise)
            key = (b.data_ptr(), b._version, tuple(b.shape), _cast_state(c))
            ent = self.__dict__.get("_bcache")
            if ent is not None and ent[0] == key:
                return ent[1]
            out = elide.materialise(c(b))
            self.__dict__["_bcache"] = (key, out)
            return out
        return c(b)

    def fold_weight_and_bias(self) -> None:
        with torch.no_grad():
            if self.bias_cast is not None and not isinstance(self.bias_format, Same):
                self.bias.data = self.bias_cast(self.bias.data)
                self.bias_cast = CastTo(format=Same())
            if self.weight_cast is not None:
                self.weight.data = self.weight_hypernet(self.weight.data)
                if self.smoothquant is not None:  # its scale is now part of the stored weight (core.py:161-166)
                    self.smoothquant.fused_to_weight[0] = 1
                    self.smoothquant._fused = True
                self.weight_sparsifier = LazySparsify(sparseness=Dense())
                self.weight_storage_cast = CastTo(format=Same())
                self.weight_cast = CastTo(format=Same())
            self._wcache = None
            self.__dict__.pop("_bcache", None)
        self.__dict__.pop("_bcache", None)

    # ------------------------------------------------------------------ forward (core.py:215-264)
    def forward(self, input, *args, **kwargs):
        _dtype = input.dtype
        sq = self.smoothquant
        _input = None
        if sq is not None and (sq._on or sq._dyn or sq.calibrating):  # core.py:227-230
            input = elide.materialise(input)
            if sq._dyn or sq.calibrating:
                sq(input, self.effective_weight)
            if sq._on and elide.active() and not torch.is_grad_enabled() and not args and not kwargs:
                _input = self._fused_scaled_input(input)  # input / scale and the input cast in ONE kernel
            if _input is None:
                input = sq.scale_input(input)
        if _input is None:
            _input, args, kwargs = self.input_casts(input, *args, **kwargs)
        _output = self._forward(_input, *args, **kwargs)
        output = self.output_casts(_output, output=True)
        if self.align_boundary_dtype:
            if isinstance(output, (tuple, list)):
                output = type(output)(a if a.dtype == _dtype else a.to(_dtype) for a in output)
            elif output.dtype != _dtype:  # (dtype is metadata: a deferred cast is not forced here)
                output = output.to(_dtype)
        return output


class Linear(DmxModule, torch.nn.Linear):
    ch_axis, win_ch_axis, wout_ch_axis = -1, -1, 0

    def __init__(self, in_features, out_features, bias=True, **kw):
        super().__init__(in_features, out_features, bias=bias, **kw)
        self._init_dmx()
        self.input_casts.input_cast.block_dim = -1
        self.weight_cast.block_dim = -1
        if self.bias_cast is not None:
            self.bias_cast.block_dim = -1

    def _forward(self, _input):  # reference torch_modules.py:346-360
        if isinstance(self.accum_format, Same):
            _weight = self._weight.to(_input.dtype)
            _bias = None if self._bias is None else self._bias.to(_input.dtype)
            return F.linear(_input, _weight, _bias)
        _weight = self._weight
        _product = self.accum_cast(torch.matmul(_input.to(_weight.dtype), _weight.t()))
        return torch.add(_product, self._bias) if self.bias is not None else _product


class Conv2d(DmxModule, torch.nn.Conv2d):
    ch_axis, win_ch_axis, wout_ch_axis = 1, 1, 0

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()
        self.input_casts.input_cast.block_dim = 1
        self.weight_cast.block_dim = 1
        if self.bias_cast is not None:
            self.bias_cast.block_dim = -1

    def _forward(self, _input):  # reference torch_modules.py:679-688
        _weight = self._weight
        _conv = self.accum_cast(self._conv_forward(_input.to(_weight.dtype), _weight, None))
        return torch.add(_conv, self._bias.unsqueeze(-1).unsqueeze(-1)) if self.bias is not None else _conv


class Embedding(DmxModule, torch.nn.Embedding):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()

    def forward(self, input):
        out = F.embedding(input, self._weight, self.padding_idx, self.max_norm, self.norm_type, self.scale_grad_by_freq, self.sparse)
        return self.output_casts(out, output=True)


class ResAdd(DmxModule, torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._init_dmx()
        self.input_casts = CastToDict(OrderedDict({"input_cast": CastTo(), "residual_cast": CastTo()}))

    def _forward(self, _input, _residual):
        return _input + _residual

    def _fusable(self, a, b):
        """(stage_a, stage_b, stage_out, out_key) when input casts + add + output cast can run as ONE kernel
        (dmxq_add_cast): plain nearest+flush FLOAT formats (or SAME), no observers / pre-transforms."""
        from .numerical.format import FloatingPoint

        if not (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.is_cuda and b.is_cuda and a.dtype == b.dtype
                and a.is_floating_point() and a.is_contiguous() and a.dim() >= b.dim()):
            return None
        casts = (self.input_casts.input_cast, self.input_casts.residual_cast, self.output_casts.output_cast)
        stages, raws = [], []
        for c, t in zip(casts, (a, b, None)):
            f = c.format
            if c.pre_transform or c._obs_on:
                return None
            fast = isinstance(f, FloatingPoint) and f.rounding == "nearest" and f.flush_subnormal and not f.unsigned
            key = elide.format_key(f, None) if fast else None
            if isinstance(t, elide.Lazy):
                # a deferred producer cast folds into the add when it is this very input format
                # (F(F(x)) == F(x)); otherwise it has to run on its own first
                if fast and t._key == key and c._fq_on and t._kind == "cast":
                    stages.append(f.stage())
                    raws.append(t._raw)
                    continue
                t = t.materialise()
            if t is not None:
                raws.append(t)
            if isinstance(f, Same) or not c._fq_on:
                stages.append(None)
            elif fast:
                if t is not None and not elide.is_tagged(t, key) and 0 in t.stride() and t.numel() > 0:
                    # broadcast operand (e.g. the attention mask, the same object in every layer): cast its
                    # un-expanded base once, remember it, and feed the kernel an already-cast operand
                    tc = elide.memo_get(t, key)
                    if tc is None:
                        base = t[tuple(slice(0, 1) if (st == 0 and n > 1) else slice(None) for n, st in zip(t.shape, t.stride()))]
                        tc = ops.cast_chain(base, [f.stage()], -1).expand(t.shape)
                        elide.stats["casts"] += 1
                        elide.memo_put(t, key, tc, pinned=True)
                        elide.tag(tc, key)
                    raws[-1] = t = tc
                stages.append(None if (t is not None and elide.is_tagged(t, key)) else f.stage())  # already in format: skip
            else:
                return None
        f_out = casts[2].format
        return stages[0], stages[1], stages[2], (None if stages[2] is None else elide.format_key(f_out, None)), raws[0], raws[1]

    def forward(self, input, residual):
        if elide.active() and not torch.is_grad_enabled():
            plan = self._fusable(input, residual)
            if plan is not None:
                if plan[2] is not None and elide.defer_output_casts and fused.add_supported(plan[4], plan[5]):
                    # hand the add itself to the consumer: a Softmax folds it into its kernel, anything else runs dmxq_add_cast
                    elide.stats["elided"] += 2
                    return fused.lazy_add(plan, self.output_casts.output_cast.format)
                try:
                    y = ops.add_cast(plan[4], plan[5], plan[0], plan[1], plan[2])
                except RuntimeError:
                    y = None  # layout the fused kernel does not take: module-by-module path below
                if y is not None:
                    elide.stats["elided"] += 2
                    elide.tag(y, plan[3])
                    return y
        return DmxModule.forward(self, elide.materialise(input), elide.materialise(residual))


class Mul(DmxModule, torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._init_dmx()
        self.input_casts = CastToDict(OrderedDict({"input_cast": CastTo(), "multiplier_cast": CastTo()}))

    def _forward(self, _input, multiplier):
        return _input * multiplier


class ActActMatMul(DmxModule, torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._init_dmx()
        self.input_casts = CastToDict(OrderedDict({"input_cast": CastTo(block_dim=-1), "multiplier_cast": CastTo(block_dim=-2)}))

    def _forward(self, _input, _multiplier):
        return torch.matmul(_input, _multiplier)


class Softmax(DmxModule, torch.nn.Softmax):
    def __init__(self, dim: int = -1):
        super().__init__(dim=dim)
        self._init_dmx()

    def _forward(self, _input):
        # torch's CUDA softmax bit for bit (rows of 33..2048 elements), 1.2 - 3x faster; anything else: torch
        if ops.softmax_supported(_input, self.dim) and not (torch.is_grad_enabled() and _input.requires_grad):
            return ops.softmax_cast(_input)
        return F.softmax(_input, dim=self.dim)

    def forward(self, input):
        if elide.active() and not torch.is_grad_enabled() and isinstance(input, torch.Tensor) and input.is_cuda:
            y = self._forward_fused(input)
            if y is not None:
                return y
        return DmxModule.forward(self, input)

    def _forward_fused(self, input):
        """input cast -> softmax -> output cast with the softmax (and a pending mask add in front of it) handed to the consumer"""
        from .numerical.format import FloatingPoint

        ic, oc = self.input_casts.input_cast, self.output_casts.output_cast
        if ic.pre_transform or oc.pre_transform or ic._obs_on or oc._obs_on or self.dim not in (-1, input.dim() - 1):
            return None
        fo = oc.format
        o_on = oc._fq_on and not isinstance(fo, Same)
        o_fast = isinstance(fo, FloatingPoint) and fo.rounding == "nearest" and fo.flush_subnormal and not fo.unsigned
        if o_on and not o_fast:
            return None
        i_on = ic._fq_on and not isinstance(ic.format, Same)
        took = fused.take_add(input, i_on, elide.format_key(ic.format, None) if i_on else None)
        if took is not None:
            x, add = took
        else:
            x, add = elide.materialise(ic(input)), None
        y = fused.softmax(x, add, fo, fo.stage() if o_on else None, elide.format_key(fo, None) if o_on else None)
        if y is None and took is None:  # the input cast has run (and is tagged / memoised): finish module by module
            y = oc(F.softmax(x, dim=self.dim), lazy_ok=True)
        return y


class LayerNorm(DmxModule, torch.nn.LayerNorm):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()

    def _forward(self, _input):
        return F.layer_norm(_input, self.normalized_shape, self._weight, self._bias, self.eps)


class ReLU(DmxModule, torch.nn.ReLU):
    def __init__(self, inplace: bool = False):
        super().__init__(inplace=False)
        self._init_dmx()

    def _forward(self, _input):
        return F.relu(_input)


class GELU(DmxModule, torch.nn.GELU):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()

    def _forward(self, _input):
        return F.gelu(_input, approximate=self.approximate)


class Dropout(DmxModule, torch.nn.Dropout):
    def __init__(self, p: float = 0.5, inplace: bool = False):
        super().__init__(p=p, inplace=False)
        self._init_dmx()

    def _forward(self, _input):
        return F.dropout(_input, self.p, self.training, False)

    def forward(self, input):
        # inference-mode dropout with SAME casts is the identity: under elision do not even touch the tensor
        # (keeps a deferred producer cast alive for the real consumer)
        if (elide.active() and not self.training and not torch.is_grad_enabled()
                and all(isinstance(c.format, Same) or not c._fq_on for c in (self.input_casts.input_cast, self.output_casts.output_cast))):
            return input
        return DmxModule.forward(self, input)


class MaxPool2d(DmxModule, torch.nn.MaxPool2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()

    def _forward(self, _input):
        return F.max_pool2d(_input, self.kernel_size, self.stride, self.padding, self.dilation, self.ceil_mode, self.return_indices)


class AvgPool2d(DmxModule, torch.nn.AvgPool2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._init_dmx()

    def _forward(self, _input):
        return F.avg_pool2d(_input, self.kernel_size, self.stride, self.padding, self.ceil_mode, self.count_include_pad, self.divisor_override)


class Tanh(DmxModule, torch.nn.Tanh):
    def __init__(self):
        super().__init__()
        self._init_dmx()

    def _forward(self, _input):
        return torch.tanh(_input)


# --------------------------------------------------------------------------------------------
# format aliases and rule sets (reference src/dmx/compressor/__init__.py:20-105, 142-483)
_F = Format.from_shorthand
format = SimpleNamespace(
    SAME=_F("SAME"), FLOAT32=_F("FP[1|8|23,127](_N)"), FLOAT16=_F("FP[1|5|10,15](FN)"), BFLOAT16=_F("FP[1|8|7,127](FN)"),
    AFLOAT8=_F("FP[1|4|3,7](_N)"), BFLOAT8=_F("FP[1|5|2,15](_N)"), INT8=_F("XP[8,0](CSN)"), INT4=_F("XP[4,0](CSN)"),
    BFP32_1=_F("BFP[24|8]{1}(SN)"),
    **{f"BFP{p + 8}_{b}": _F(f"BFP[{p}|8]{{{b}}}(SN)") for p in (16, 8, 6, 4) for b in (128, 64, 32, 16) if not (p == 16 and b == 128)},
    **{f"BFP{p + 8}A_{b}": _F(f"BFP[{p}|8]{{{b}}}(_N)") for p in (8, 6, 4) for b in (128, 64, 32, 16)},
    SBFP12_16=_F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"),
    **{f"SBFP12_16_{b}": _F(f"SBFP<XP[4,0](CSN)><FP[0|4|4,{b}](FN)>{{16}}") for b in range(4, 19)},
    **{f"MXINT{p}_K{b}": _F(f"MXINT{p}{{{b}}}") for p in (8, 6, 4) for b in (128, 64, 32)},
)
sparseness = SimpleNamespace(BTK8_4_LD="BTOPK{4:8,-1}(U)", BTK8_4_FD="BTOPK{4:8,1}(U)", BTK8_2_LD="BTOPK{2:8,-1}(U)", BTK8_2_FD="BTOPK{2:8,1}(U)")


class DmxConfigRule(SimpleNamespace):
    r"""(module_types, name_re, module_config) -> applied to every matching submodule
    (reference modeling/model.py:721-792)."""

    def __init__(self, module_types=(), name_re: str = "", module_config: Optional[dict] = None):
        super().__init__(module_types=tuple(module_types), name_re=name_re, module_config=module_config or {})

    def apply_to(self, model: torch.nn.Module) -> None:
        import re

        for n, m in model.named_modules():
            if isinstance(m, DmxModule) and isinstance(m, self.module_types) and re.match(self.name_re, n):
                m.configure(self.module_config)


_ACT = (ReLU, GELU, Tanh, Softmax, LayerNorm)
config_rules = SimpleNamespace(
    BASELINE=[
        DmxConfigRule((Linear, Conv2d), module_config=dict(input_formats=[format.SAME], weight_format=format.SAME, bias_format=format.SAME, output_formats=[format.SAME])),
        DmxConfigRule((ResAdd, ActActMatMul, Mul), module_config=dict(input_formats=[format.SAME, format.SAME], output_formats=[format.SAME])),
        DmxConfigRule((Embedding,), module_config=dict(output_formats=[format.SAME])),
        DmxConfigRule(_ACT + (MaxPool2d, AvgPool2d, Dropout), module_config=dict(input_formats=[format.SAME], output_formats=[format.SAME])),
    ],
    BASIC=[
        DmxConfigRule((Linear, Conv2d), module_config=dict(input_formats=[format.BFP16_64], weight_format=format.BFP16_64, bias_format=format.BFP32_1, output_formats=[format.FLOAT16])),
        DmxConfigRule((ResAdd,), module_config=dict(input_formats=[format.FLOAT16, format.FLOAT16], output_formats=[format.FLOAT16])),
        DmxConfigRule((ActActMatMul,), module_config=dict(input_formats=[format.BFP16_64, format.BFP16_64], output_formats=[format.FLOAT16])),
        DmxConfigRule((Embedding,), module_config=dict(output_formats=[format.FLOAT16])),
        DmxConfigRule(_ACT + (MaxPool2d, AvgPool2d), module_config=dict(input_formats=[format.FLOAT16], output_formats=[format.FLOAT16])),
    ],
    SBFP_WEIGHT_STORAGE=[DmxConfigRule((Linear, Conv2d), module_config=dict(weight_storage_format=format.SBFP12_16))],
)


def configure(model: torch.nn.Module, *rules) -> torch.nn.Module:
    """model.transform(None, *rules) of the reference (modeling/model.py:61-80)."""
    for r in rules:
        for rr in (r if isinstance(r, (list, tuple)) else [r]):
            rr.apply_to(model)
    return model


def to_basic_mode(model: torch.nn.Module) -> torch.nn.Module:
    return configure(model, config_rules.BASELINE, config_rules.BASIC)


def fold_weights_and_biases(model: torch.nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, DmxModule):
            m.fold_weight_and_bias()

"""Opt-in elision of provably redundant casts.

The reference re-executes every cast of every module on every forward: ``SAME`` casts clone
(format.py:89-90), a FLOAT16 output cast is followed by the next module's FLOAT16 input cast
on the very same tensor, q/k/v projections each re-cast the same LayerNorm output, and weights
are re-cast per forward (core.py:200-205).  None of that changes a value:

  * ``SAME``                      -> identity (no clone),
  * idempotent formats (FLOAT*, symmetric nearest BFP) applied to a tensor
    that *is* the unmodified result of the same cast -> identity,
  * the same (tensor, format, block_dim) requested again while the tensor is unmodified -> the
    earlier result (a small strong-ref memo, cleared by ``reset()`` / on leaving the context),
  * weights: cached per (storage, version, formats) in ``DmxModule._weight`` (nn.py).

Safety: tags are keyed by the Python tensor object (identity checked through a weakref, so an
address reused by a new tensor can never inherit a tag) *and* by ``tensor._version`` (in-place
writes invalidate).  Elision is only consulted under ``torch.no_grad()``; results may alias
their inputs, which is the one observable difference from the reference (it always returns a
fresh tensor), hence opt-in:

    with dmx_compressor_b200.elide.enabled():
        y = model(x)
"""
from __future__ import annotations

import contextlib
import weakref
from collections import OrderedDict

import torch

_ON = False
defer_output_casts = True   # under elision, FLOAT output casts are handed to their consumer (see Lazy)
_TAGS = {}   # id(tensor) -> (weakref, version, set(keys))
_MEMO = OrderedDict()   # (id(x), version, key) -> (weakref(x), y)
_MEMO_MAX = 8
_PINNED = OrderedDict()  # same, for small broadcast operands (masks) that must survive a whole forward
_PINNED_MAX = 4
stats = {"elided": 0, "memo_hits": 0, "casts": 0}


def active() -> bool:
    return _ON


def enable(on: bool = True) -> None:
    global _ON
    _ON = bool(on)
    if not on:
        reset()


def reset() -> None:
    _TAGS.clear()
    _MEMO.clear()
    _PINNED.clear()


@contextlib.contextmanager
def enabled():
    prev = _ON
    enable(True)
    try:
        yield
    finally:
        enable(prev)
        reset()


def format_key(fmt, block_dim):
    """hashable identity of a cast, or None when the format is not safely idempotent."""
    from .numerical.format import BlockFloatingPoint, FloatingPoint

    if isinstance(fmt, FloatingPoint) and fmt.rounding == "nearest":
        return ("FP", repr(fmt))
    if isinstance(fmt, BlockFloatingPoint) and fmt.symmetric and fmt.rounding == "nearest":
        return ("BFP", repr(fmt), block_dim)
    return None


def is_tagged(x: torch.Tensor, key, ndim_block_dim=None) -> bool:
    t = _TAGS.get(id(x))
    if t is None:
        return False
    ref, ver, keys = t
    if ref() is not x or ver != x._version:
        _TAGS.pop(id(x), None)
        return False
    return key in keys


def tag(y: torch.Tensor, key) -> None:
    if key is None:
        return
    i = id(y)
    t = _TAGS.get(i)
    if t is not None and t[0]() is y and t[1] == y._version:
        t[2].add(key)
        return
    _TAGS[i] = (weakref.ref(y, lambda _r, i=i: _TAGS.pop(i, None)), y._version, {key})


# ------------------------------------------------------------------------------------------------
# deferred output casts
_META_PROPS = ("shape", "dtype", "device", "is_cuda", "is_cpu", "ndim", "requires_grad", "_version", "layout", "is_leaf",
               "grad_fn", "names", "is_sparse", "is_quantized", "is_meta", "is_nested", "is_mkldnn", "is_xpu")
_META_METHODS = ("dim", "size", "stride", "numel", "nelement", "element_size", "is_contiguous", "is_floating_point", "is_complex",
                 "storage_offset", "get_device", "ndimension", "is_pinned", "is_shared", "is_inference", "__len__", "__format__")
_PASS = None


def _passthrough():
    global _PASS
    if _PASS is None:
        _PASS = {getattr(torch.Tensor, n).__get__ for n in _META_PROPS if hasattr(torch.Tensor, n)}
        _PASS |= {getattr(torch.Tensor, n) for n in _META_METHODS if hasattr(torch.Tensor, n)}
    return _PASS


class Lazy(torch.Tensor):
    """An output cast that has not run yet: carries the uncast tensor and the pending format.

    A module's output cast is the producer half of a pair whose consumer half is the next module's
    input cast.  Deferring it lets the consumer run ONE kernel: the same idempotent format ->
    a single cast; a different format -> the fused chain [pending, own]; a ResAdd -> the pending
    cast folds into the fused add.  Any *other* use -- any torch function or method that needs
    values -- materialises the cast first through ``__torch_function__``, so code between modules
    (views, scaling, user ops) always sees exactly the values the eager cast would have produced.
    Only shape / dtype / device style metadata is answered without materialising."""

    @staticmethod
    def __new__(cls, raw, fmt, block_dim, key, materialise=None, kind="cast", fuse=None, add=None):
        r = torch.Tensor._make_subclass(cls, raw, False)
        r._raw, r._fmt, r._bd, r._key, r._real, r._mat = raw, fmt, block_dim, key, None, materialise
        # kind "cast": `raw` is the uncast value, a consumer may chain [fmt, its own format] over it.  Other kinds carry an
        # OPERATION in front of the pending cast -- "softmax" (raw = the softmax input) / "add" (raw = the first addend, `add` =
        # (a, b, stage_a, stage_b, stage_out)): only `materialise` -- or `fuse(next_stage, block_dim)`, which returns the
        # operation + pending cast + the consumer's cast in one kernel, or None -- may produce values from them.
        r._kind, r._fuse, r._add = kind, fuse, add
        return r

    def materialise(self) -> torch.Tensor:
        if self._real is None:
            if self._mat is not None:
                self._real = self._mat(self._raw)
            else:
                from . import ops

                self._real = ops.cast_chain(self._raw, [self._fmt.stage()], self._bd)
            stats["casts"] += 1
            tag(self._real, self._key)
        return self._real

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _passthrough():
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func is torch.Tensor.to and len(args) == 2 and not kwargs and isinstance(args[0], Lazy) and args[1] == args[0].dtype:
            return args[0]  # `.to(own dtype)` (DmxModule.forward's boundary alignment, core.py:258-263) returns the tensor itself
        stats["lazy_materialised_by_torch_op"] = stats.get("lazy_materialised_by_torch_op", 0) + 1
        return func(*_unlazy(args), **_unlazy(kwargs))


def _unlazy(o):
    if isinstance(o, Lazy):
        return o.materialise()
    if isinstance(o, (list, tuple)):
        return type(o)(_unlazy(e) for e in o)
    if isinstance(o, dict):
        return {k: _unlazy(v) for k, v in o.items()}
    return o


def materialise(x):
    """the plain tensor behind a (possibly) deferred cast; tuples / lists / dicts are walked"""
    return _unlazy(x)


def memo_get(x: torch.Tensor, key):
    k = (id(x), x._version, key)
    for table in (_MEMO, _PINNED):
        ent = table.get(k)
        if ent is not None and ent[0]() is x:
            stats["memo_hits"] += 1
            table.move_to_end(k)
            return ent[1]
    return None


def memo_put(x: torch.Tensor, key, y: torch.Tensor, pinned: bool = False) -> None:
    table, cap = (_PINNED, _PINNED_MAX) if pinned else (_MEMO, _MEMO_MAX)
    table[(id(x), x._version, key)] = (weakref.ref(x), y)
    while len(table) > cap:
        table.popitem(last=False)

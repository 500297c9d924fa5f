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
_TAGS = {}   # id(tensor) -> (weakref, version, set(keys))
_MEMO = OrderedDict()   # (id(x), version, key) -> (weakref(x), y)
_MEMO_MAX = 8
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


def memo_get(x: torch.Tensor, key):
    ent = _MEMO.get((id(x), x._version, key))
    if ent is not None and ent[0]() is x:
        stats["memo_hits"] += 1
        _MEMO.move_to_end((id(x), x._version, key))
        return ent[1]
    return None


def memo_put(x: torch.Tensor, key, y: torch.Tensor) -> None:
    _MEMO[(id(x), x._version, key)] = (weakref.ref(x), y)
    while len(_MEMO) > _MEMO_MAX:
        _MEMO.popitem(last=False)

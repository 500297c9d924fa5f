"""Deferred operations of the attention block under cast elision: the attention-mask add and the
softmax are handed to their consumer like deferred output casts are (elide.Lazy), so that

    ResAdd (scores + mask, three FLOAT16 casts) -> Softmax (FLOAT16 in / out) -> the BFP16 input cast of P.V

runs as ONE launch of dmxq_softmax_cast instead of three full passes over the [B, H, S, S] scores.  Shared by the mirror
modules (nn.py) and by the plugin's wrappers around the reference's own modules (plugin.py); every piece is value-identical
to the module-by-module sequence (tests/test_softmax_gpu.py, tests/test_model_gpu.py, tests/test_plugin_gpu.py)."""
from __future__ import annotations

import torch

from . import elide as E
from . import ops


def lazy_add(plan, out_fmt):
    """ResAdd under elision: (stage_a, stage_b, stage_out, out_key, a, b) -> the add as a pending operation in front of its output
    cast.  The next CastTo / ResAdd / any torch op materialises it with dmxq_add_cast, exactly as the eager fused add; a Softmax
    module folds it into its own kernel."""
    sa, sb, so, key, a, b = plan

    def mat(raw):
        try:
            return ops.add_cast(a, b, sa, sb, so)
        except RuntimeError:  # a layout the fused add does not take after all: the same values, pass by pass
            xa = ops.cast_chain(a, [sa], -1) if sa is not None else a
            xb = ops.cast_chain(b.contiguous(), [sb], -1) if sb is not None else b
            y = xa + xb
            return ops.cast_chain(y, [so], -1) if so is not None else y

    return E.Lazy(a, out_fmt, None, key, materialise=mat, kind="add", add=(a, b, sa, sb, so))


def add_supported(a, b):
    """what dmxq_add_cast takes (so that deferring the add can never turn a fused add into the module-by-module path later)"""
    if not (a.is_contiguous() and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and a.numel() > 0):
        return False
    v = 16 // a.element_size()
    # b right-aligned against a: its innermost run must be whole vectors, broadcast strides 16-byte aligned
    if b.dim() == 0 or b.stride(-1) != 1 or b.shape[-1] != a.shape[-1] or a.shape[-1] % v != 0:
        return False
    return all((s * b.element_size()) % 16 == 0 for s in b.stride()[:-1])


def take_add(inp, in_on, in_key):
    """the operands of a pending add that a Softmax may fold in: its input cast must be off, or the very format the add's output
    cast already applies (idempotent).  -> (x, (b, stage_a, stage_b, stage_out)) or None"""
    if isinstance(inp, E.Lazy) and inp._kind == "add" and inp._real is None and (not in_on or (in_key is not None and in_key == inp._key)):
        a, b, sa, sb, so = inp._add
        return a, (b, sa, sb, so)
    return None


def _run(x, post, add):
    if add is None:
        return ops.softmax_cast(x, post)
    b, sa, sb, so = add
    return ops.softmax_cast(x, post, addend=b, stage_x=sa, stage_addend=sb, stage_sum=so)


def softmax(x, add, out_fmt, out_stage, out_key):
    """softmax(x [+ pending add]) followed by the module's output cast.  out_stage None: no output cast (SAME / disabled) -> the
    tensor; otherwise a pending "softmax" Lazy whose consumer may add its own cast to the same launch.  Returns None when the
    fused kernel does not take x (row length, layout): the caller keeps torch.softmax."""
    if not ops.softmax_supported(x, -1):
        return None
    if out_stage is None:
        try:
            return _run(x, [], add)
        except RuntimeError:
            return None

    def mat(raw):
        return _run(raw, [out_stage], add)

    def fuse(next_stage, block_dim):
        if block_dim not in (-1, x.dim() - 1):
            return None
        try:
            return _run(x, [out_stage, next_stage], add)
        except RuntimeError:  # e.g. the consumer's blocks do not tile the row: run the softmax + output cast, then its cast
            return None

    if add is not None:  # probe once that the kernel takes this addend layout before promising the fusion
        b = add[0]
        if not (isinstance(b, torch.Tensor) and b.is_cuda and b.dtype == x.dtype and b.dim() <= x.dim() and b.shape[-1] == x.shape[-1]
                and b.stride(-1) == 1 and b.data_ptr() % 16 == 0 and all((s * b.element_size()) % 16 == 0 for s in b.stride()[:-1])):
            return None
    return E.Lazy(x, out_fmt, None, out_key, materialise=mat, kind="softmax", fuse=fuse)

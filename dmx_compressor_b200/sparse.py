"""Weight sparsity -- host-side mirror of the reference's ``Sparseness`` / ``Sparsify`` surface
(reference src/dmx/compressor/sparse.py:23-344) with the N:M (``BlockTopK``) mask + apply
routed to libdmxq (``dmxq_nm_prune``): one kernel computes the per-group ranks, writes the
0/1 mask and the masked tensor, instead of the reference's argsort / ones_like / scatter_ /
transpose / multiply sequence with its int64 index tensor (sparse.py:169-178, :300).

Tie rule (the reference relies on ``torch.argsort``): lowest score pruned first, equal scores
-> lowest index first, NaN scores rank highest.  tests/test_parity_gpu.py pins it against
torch's CUDA argsort.
"""
from __future__ import annotations

import re
from typing import Any

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.nn.modules.lazy import LazyModuleMixin
from torch.nn.parameter import UninitializedParameter

from . import _lib as L
from . import ops

__ALL__ = ["Sparseness", "Dense", "TopK", "BlockTopK", "Bernoulli", "Sparsify", "LazySparsify", "abs_score"]


def abs_score(score, x):
    """The magnitude score function ``lambda s, x: x.abs()``.  Passing *this* function to
    ``Sparsify.configure(score_func=...)`` lets the kernel derive the score from x in registers
    (no score tensor is materialised or read)."""
    return x.abs()


class Sparseness:
    r"""Sparsity pattern descriptor: child classes implement ``get_mask`` and the shorthand
    (reference sparse.py:23-64)."""

    blocked: bool
    density = None

    def __init__(self, mask_gradient=False):
        self.mask_gradient = torch.as_tensor(mask_gradient)

    def get_mask(self, *input: Any):
        raise NotImplementedError

    @classmethod
    def from_shorthand(cls, sh: str):
        if sh.startswith("DENSE"):
            return Dense.from_shorthand(sh)
        elif sh.startswith("TOPK"):
            return TopK.from_shorthand(sh)
        elif sh.startswith("BTOPK"):
            return BlockTopK.from_shorthand(sh)
        elif sh.startswith("BERN"):
            return Bernoulli.from_shorthand(sh)
        raise ValueError(f"unrecognized sparseness shorthand: {sh}")


class Dense(Sparseness):
    r"""No sparsity (reference sparse.py:67-94)."""

    blocked = False
    density = 1.0

    def get_mask(self, score):
        return torch.ones_like(score)

    @classmethod
    def from_shorthand(cls, sh: str):
        return cls()

    def __str__(self) -> str:
        return "Dense: no sparsity"

    def __repr__(self) -> str:
        return "DENSE"


class TopK(Sparseness):
    r"""Global unstructured top-K (reference sparse.py:97-144).  Not an N:M pattern: a global
    sort, outside the kernel path; provided with torch ops."""

    blocked = False
    _RX = re.compile(r"^TOPK\{(?P<density>[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)\}\((?P<mask_grad>[A-Za-z])\)$")

    def __init__(self, density=0.5, mask_gradient=False):
        super().__init__(mask_gradient)
        assert 0 <= density <= 1.0, "density has to be between 0 and 1"
        self.density = density

    def get_mask(self, score):
        flat = score.detach().reshape(-1)
        n_prune = int(score.numel() * (1.0 - self.density))
        order = torch.argsort(flat, dim=0, stable=True)
        mask = torch.ones_like(flat)
        mask[order[:n_prune]] = 0
        return mask.view_as(score)

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise ValueError(f"unrecognized sparseness shorthand: {sh}")
        return cls(density=float(m["density"]), mask_gradient=m["mask_grad"] == "M")

    def __str__(self) -> str:
        return f"Global TopK sparseness: density = {self.density}"

    def __repr__(self) -> str:
        return f"TOPK{{{self.density}}}({'M' if self.mask_gradient else 'U'})"


class BlockTopK(Sparseness):
    r"""K non-zeros out of every ``block_size`` consecutive elements along ``block_dim``
    (reference sparse.py:147-204)."""

    blocked = True
    _RX = re.compile(r"^BTOPK\{(?P<K>\d+):(?P<block_size>\d+),(?P<block_dim>[-+]?\d+)\}\((?P<mask_grad>[A-Za-z])\)$")

    # Tie order inside a group.  The reference sorts with torch.argsort (sparse.py:172), whose order for tied scores
    # differs between its two back ends: stable on CPU tensors, an unstable bitonic network on CUDA tensors (groups of
    # <= 32).  "stable" (default: deterministic, lower index pruned first) or "torch_cuda" (bit-identical to what the
    # reference produces on a GPU; include/dmxq.h DMXQ_NM_ORDER_*).  A class-level default, overridable per instance.
    tie_order = "stable"

    def __init__(self, K=4, block_size=8, block_dim=-1, mask_gradient=False, tie_order=None):
        super().__init__(mask_gradient)
        assert 0 < K <= block_size, "N and M must be positive and N no greater than M"
        self.K = K
        self.block_size = block_size
        self.block_dim = block_dim
        self.density = self.K / self.block_size
        if tie_order is not None:
            assert tie_order in ("stable", "torch_cuda"), f"unknown tie order {tie_order!r}"
            self.tie_order = tie_order

    @property
    def nm_order(self) -> int:
        return L.NM_TORCH_CUDA if self.tie_order == "torch_cuda" else L.NM_STABLE

    def get_mask(self, score):
        """mask only (reference BlockTopK.forward, sparse.py:163-180)."""
        _, mask = ops.nm_prune(score.detach(), self.K, self.block_size, self.block_dim, score=score.detach(), return_mask=True,
                               nm_order=self.nm_order)
        return mask.to(score.dtype)

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise ValueError(f"unrecognized sparseness shorthand: {sh}")
        return cls(K=int(m["K"]), block_size=int(m["block_size"]), block_dim=int(m["block_dim"]), mask_gradient=m["mask_grad"] == "M")

    def __str__(self) -> str:
        return f"Block TopK sparseness: pattern = {self.K}:{self.block_size}, block dimension = {self.block_dim}"

    def __repr__(self) -> str:
        return f"BTOPK{{{self.K}:{self.block_size},{self.block_dim}}}({'M' if self.mask_gradient else 'U'})"


class Bernoulli(Sparseness):
    r"""Bernoulli sampler for supermasking (reference sparse.py:207-242); RNG sampler, outside
    the kernel path."""

    blocked = False

    def get_mask(self, score):
        return torch.bernoulli(score.detach().clamp(0, 1))

    @classmethod
    def from_shorthand(cls, sh: str):
        return cls()

    def __str__(self) -> str:
        return "Bernoulli sparseness"

    def __repr__(self) -> str:
        return "BERN"


class _NMPrune(Function):
    """y = x * mask(score) in one kernel; backward routes gradients like the reference's
    ``x * mask`` with a straight-through mask (sparse.py:182-184, 295-300)."""

    @staticmethod
    def forward(ctx, x, score, sp, weight_grad, mask_grad):
        # x * mask promotes to the mask's dtype, which is the score's (sparse.py:173-178, 300)
        out_dtype = x.dtype if score is None else torch.promote_types(x.dtype, score.dtype)
        y, mask = ops.nm_prune(x, sp.K, sp.block_size, sp.block_dim, score=score, return_mask=True,
                               out_dtype=torch.float32 if out_dtype == torch.float32 else x.dtype, nm_order=sp.nm_order)
        ctx.save_for_backward(x, mask)
        ctx.flags = (weight_grad, mask_grad, score is not None)
        ctx.mark_non_differentiable(mask)
        return y, mask

    @staticmethod
    def backward(ctx, g, _gm):
        x, mask = ctx.saved_tensors
        wg, mg, has_score = ctx.flags
        gx = (g * mask).to(x.dtype) if (wg and ctx.needs_input_grad[0]) else None
        gs = (g * x) if (mg and has_score and ctx.needs_input_grad[1]) else None
        return gx, gs, None, None, None


class Sparsify(nn.Module):
    r"""Sparsification module (reference sparse.py:245-315)."""

    def __init__(self, tensor_shape, sparseness="DENSE", backward_mode="STE", score_func=None):
        super().__init__()
        self.score = nn.Parameter(torch.rand(tensor_shape), requires_grad=True)
        self.mask = None
        self.configure(sparseness, backward_mode, score_func)
        self.plastic = False

    def configure(self, sparseness=None, backward_mode=None, score_func=None):
        if sparseness is not None:
            if not isinstance(sparseness, Sparseness):
                sparseness = Sparseness.from_shorthand(sparseness)
            if not hasattr(self, "sparseness") or repr(sparseness) != repr(self.sparseness):
                self.sparseness = sparseness
        if backward_mode is not None:
            self.backward_mode = backward_mode
            self.enable_weight_gradient = backward_mode.lower() in {"ste", "joint"}
            self.enable_mask_gradient = backward_mode.lower() in {"supermask", "joint"}
        if score_func is not None:
            self.score_func = score_func
            self.plastic = True  # rewire on the next forward() (sparse.py:281-285)

    def update_mask(self, score):
        self.mask = self.sparseness.get_mask(score)

    def forward(self, x):
        sp = self.sparseness
        if isinstance(sp, Dense):
            return x
        implicit_abs = False
        if self.plastic:
            implicit_abs = self.score_func is abs_score
            score = None if implicit_abs else self.score_func(self.score, x)
            self.plastic = False
        else:
            score = self.score
        if isinstance(sp, BlockTopK):
            wg = self.enable_weight_gradient if self.training else True
            mg = self.enable_mask_gradient if self.training else True
            y, mask = _NMPrune.apply(x, score, sp, wg, mg)
            self.mask = mask
            return y
        # non-N:M patterns: mask with torch ops, outside the kernel path
        if implicit_abs:
            score = x.abs()
        self.update_mask(score)
        if self.training:
            x = x if self.enable_weight_gradient else x.detach()
        return x * self.mask

    @property
    def density(self) -> float:
        if self.sparseness.density is not None:
            return self.sparseness.density
        self.update_mask(self.score)
        return self.mask.data.sum() / self.mask.numel()

    def extra_repr(self):
        return f"sparseness = {self.sparseness.__repr__()}, backward_mode = {self.backward_mode}"


class LazySparsify(LazyModuleMixin, Sparsify):
    r"""Shape-deferred Sparsify (reference sparse.py:318-344)."""

    cls_to_become = Sparsify
    score: UninitializedParameter

    def __init__(self, sparseness="DENSE", backward_mode="STE", score_func=None) -> None:
        super().__init__(torch.Size([0]), sparseness, backward_mode, score_func)
        self.score = UninitializedParameter()
        self.configure(sparseness, backward_mode, score_func)

    def reset_parameters(self) -> None:
        if not self.has_uninitialized_params():
            nn.init.uniform_(self.score)

    def initialize_parameters(self, x: torch.Tensor) -> None:
        self.tensor_shape = x.shape
        if self.has_uninitialized_params():
            with torch.no_grad():
                if isinstance(self.sparseness, Dense):
                    self.score.materialize(torch.Size([0]))
                else:
                    self.score.materialize(self.tensor_shape)
                    self.reset_parameters()

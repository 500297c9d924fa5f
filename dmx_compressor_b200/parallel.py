"""Multi-GPU side of the path: whole-model weight casting sharded by
parameter / row range, and calibration statistics reduced with ONE batched all-reduce.

The reference has no distributed code at all; this is the capability BASELINE configs #4/#5
name.  One process per GPU (``torchrun``), ``torch.distributed`` for the plumbing.

* Every block (BFP/SBFP), element (FP/XP) and M-group (N:M) is independent, so a weight
  ``[out, in]`` blocked along ``in`` can be cut at any row boundary and different parameters
  are independent: sharding needs NO data-path collective.  ``plan_shards`` packs whole
  tensors onto ranks (longest-processing-time first) and row-splits the few tensors that are
  too big to balance (lm_head, embeddings).
* The only exchange is calibration statistics (per-tensor / per-channel amin & amax feeding
  ``MinMaxObserver._calculate_qparams``, or a tensor-wide amax): each rank reduces its shard
  with ``dmxq_minmax``; all statistics of all tensors are packed into one fp32 buffer and
  reduced with a single ``all_reduce(MAX)`` over ``[max..., -min...]`` (NCCL over NVLink on the
  box, gloo in the CPU tests).  min/max are exact and order independent, so the sharded
  result is bit-identical to the single-GPU result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch


@dataclass(frozen=True)
class Shard:
    name: str
    row0: int   # first row (dim 0) owned by the rank
    row1: int   # one past the last row
    rows: int   # total rows of the tensor

    @property
    def whole(self) -> bool:
        return self.row0 == 0 and self.row1 == self.rows


def plan_shards(shapes: Dict[str, Sequence[int]], world_size: int, split_threshold: float = 0.5,
                row_align: int = 1) -> List[List[Shard]]:
    """-> per rank, the list of shards it owns.

    Tensors with more than ``split_threshold * total / world_size`` elements are cut into
    ``world_size`` contiguous row ranges (whole rows, multiples of ``row_align`` rows, so no
    block or N:M group straddles ranks); the rest are assigned whole, largest first, to the
    currently lightest rank."""
    numel = {n: int(torch.Size(s).numel()) for n, s in shapes.items()}
    total = sum(numel.values())
    plan: List[List[Shard]] = [[] for _ in range(world_size)]
    load = [0] * world_size
    limit = split_threshold * total / max(world_size, 1)
    whole = []
    for n, s in shapes.items():
        rows = int(s[0]) if len(s) > 0 else 1
        if world_size > 1 and numel[n] > limit and rows >= world_size * row_align:
            per_row = numel[n] // max(rows, 1)
            units = rows // row_align
            base, extra = divmod(units, world_size)
            r0 = 0
            for r in range(world_size):
                take = (base + (1 if r < extra else 0)) * row_align
                r1 = rows if r == world_size - 1 else r0 + take
                plan[r].append(Shard(n, r0, r1, rows))
                load[r] += (r1 - r0) * per_row
                r0 = r1
        else:
            whole.append(n)
    for n in sorted(whole, key=lambda k: (-numel[k], k)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        rows = int(shapes[n][0]) if len(shapes[n]) > 0 else 1
        plan[r].append(Shard(n, 0, rows, rows))
        load[r] += numel[n]
    return plan


def plan_imbalance(plan: List[List[Shard]], shapes: Dict[str, Sequence[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = []
    for shards in plan:
        t = 0
        for sh in shards:
            s = shapes[sh.name]
            t += (sh.row1 - sh.row0) * (int(torch.Size(s).numel()) // max(sh.rows, 1))
        loads.append(t)
    mean = sum(loads) / len(loads)
    return max(loads) / mean if mean else 1.0


def cast_shards(shards: Sequence[Shard], materialise: Callable[[Shard], torch.Tensor], stages, block_dim: int = -1,
                keep: bool = False):
    """Cast every shard this rank owns with one fused chain kernel each.  ``materialise`` returns
    the shard's rows as a CUDA tensor (generated or loaded shard-locally).  Returns
    (algorithmic bytes processed, {name: result} if keep)."""
    from . import ops

    out = {}
    nbytes = 0
    for sh in shards:
        w = materialise(sh)
        y = ops.cast_chain(w, stages, block_dim)
        nbytes += 2 * w.numel() * w.element_size()
        if keep:
            out[(sh.name, sh.row0)] = y
    return nbytes, out


def replicate_shards(plan: List[List[Shard]], local: Dict[Tuple[str, int], torch.Tensor], shapes: Dict[str, Sequence[int]], dtype, device,
                     group=None) -> Dict[str, torch.Tensor]:
    """Give every rank every cast tensor ("no gather unless the caller wants a replicated result").
    ``local`` = the ``{(name, row0): tensor}`` this rank got from ``cast_shards(..., keep=True)``.  Each shard is
    broadcast by its owner straight into its rows of the preallocated full tensor -- no staging copy, no padding for
    the uneven row splits; on the box this is NCCL over NVLink and is reported separately from the cast throughput."""
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    full = {name: torch.empty(tuple(shape), dtype=dtype, device=device) for name, shape in shapes.items()}
    for owner, shards in enumerate(plan):
        for sh in shards:
            rows = full[sh.name][sh.row0:sh.row1]
            if owner == rank:
                rows.copy_(local[(sh.name, sh.row0)])
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.broadcast(rows, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    return full


# ------------------------------------------------------------------------------------------------
# batched statistics all-reduce
def local_minmax(t: torch.Tensor, ch_axis: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """amin / amax of a local shard: the dmxq_minmax kernel (CUDA tensors only; there is no CPU path --
    the gloo unit tests of the reduction logic inject their own statistic through ``local=``)."""
    from . import ops

    return ops.minmax(t, ch_axis)


def allreduce_minmax(stats: Sequence[Tuple[torch.Tensor, torch.Tensor]], group=None):
    """One collective for all tensors: pack [max_0.., -min_0.., isnan_0..] and all_reduce(MAX).
    ``stats``: per tensor (min, max) fp32 vectors (length 1 or #channels), identical lengths on
    every rank.  Returns the reduced list in the same order.  NaN propagates (as in torch.aminmax): it travels as an
    explicit 0/1 flag segment of the same buffer -- +-inf stay what they are, so a channel that really holds both
    +inf and -inf reduces to (-inf, +inf) exactly as on a single GPU."""
    import torch.distributed as dist

    if not stats:
        return []
    dev = stats[0][0].device
    sizes = [mn.numel() for mn, _ in stats]
    mx0 = torch.cat([mx.reshape(-1).float() for _, mx in stats]).to(dev)
    mn0 = torch.cat([mn.reshape(-1).float() for mn, _ in stats]).to(dev)
    nan = torch.isnan(mx0) | torch.isnan(mn0)
    ninf = torch.full_like(mx0, float("-inf"))
    buf = torch.cat([torch.where(nan, ninf, mx0), torch.where(nan, ninf, -mn0), nan.float()])  # a NaN entry is neutral in the value slots
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    n = sum(sizes)
    mx_all, mn_all, was_nan = buf[:n], -buf[n:2 * n], buf[2 * n:] > 0
    mx_all = torch.where(was_nan, torch.full_like(mx_all, float("nan")), mx_all)
    mn_all = torch.where(was_nan, torch.full_like(mn_all, float("nan")), mn_all)
    out, o = [], 0
    for s in sizes:
        out.append((mn_all[o:o + s].clone(), mx_all[o:o + s].clone()))
        o += s
    return out


def sharded_minmax(tensors: Sequence[torch.Tensor], ch_axes: Optional[Sequence[Optional[int]]] = None, group=None,
                   local: Callable = local_minmax):
    """amin/amax of row-sharded tensors as if they were whole: local kernel + one all-reduce."""
    ch_axes = ch_axes or [None] * len(tensors)
    return allreduce_minmax([local(t, a) for t, a in zip(tensors, ch_axes)], group)


def shard_stats(plan: List[List[Shard]], rank: int, tensors: Sequence[torch.Tensor], group=None, local: Callable = local_minmax):
    """Per-tensor (amin, amax) for the shards `tensors` that `plan[rank]` lists, as if every
    tensor were whole.  Tensors owned entirely by one rank need no communication; the row-split
    ones (the same set, in the same order, on every rank) share ONE all-reduce."""
    mine = plan[rank]
    local = [local(t, None) for t in tensors]
    split_names = sorted({sh.name for shards in plan for sh in shards if not sh.whole})
    if split_names:
        pos = {n: i for i, n in enumerate(split_names)}
        dev = tensors[0].device if tensors else torch.device("cpu")
        mn = torch.full((len(split_names),), float("inf"), device=dev)
        mx = torch.full((len(split_names),), float("-inf"), device=dev)
        for sh, (a, b) in zip(mine, local):
            if not sh.whole:
                mn[pos[sh.name]] = a[0]
                mx[pos[sh.name]] = b[0]
        (rmn, rmx), = allreduce_minmax([(mn, mx)], group)
        local = [(rmn[pos[sh.name]].reshape(1), rmx[pos[sh.name]].reshape(1)) if not sh.whole else st
                 for sh, st in zip(mine, local)]
    return local


_AMAX_PLANS: dict = {}


def shard_amax(plan: List[List[Shard]], rank: int, tensors: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None, group=None,
               local: Callable = None) -> torch.Tensor:
    """Tensor-wide max|x| for every shard `plan[rank]` lists (as if each tensor were whole), as ONE fp32 device vector --
    the form ``ops.cast_chain_multi(..., amax=...)`` consumes, so calibration statistics go from the reduction into the cast
    kernel without ever visiting the host.  One ``dmxq_amax_multi`` launch per 64 shards; the row-split tensors
    (the same set, in the same order, on every rank) share ONE ``all_reduce(MAX)``.  A NaN anywhere in a tensor makes its
    amax +inf (the cast then keeps the format's default scaler bias)."""
    import torch.distributed as dist

    from . import ops

    mine = plan[rank]
    n = len(mine)
    dev = tensors[0].device if n else torch.device("cpu")
    key = (id(plan), rank, str(dev))
    ent = _AMAX_PLANS.get(key)
    if ent is None or ent[0] is not plan:
        split_names = sorted({sh.name for shards in plan for sh in shards if not sh.whole})
        pos = {nm: i for i, nm in enumerate(split_names)}
        src = [i for i, sh in enumerate(mine) if not sh.whole]
        ent = (plan, len(split_names), torch.tensor(src, dtype=torch.long, device=dev),
               torch.tensor([pos[mine[i].name] for i in src], dtype=torch.long, device=dev),
               torch.empty(2, max(n, 1), dtype=torch.float32, device=dev))
        if len(_AMAX_PLANS) > 64:
            _AMAX_PLANS.clear()
        _AMAX_PLANS[key] = ent
    _, n_split, src_idx, dst_idx, mnmx = ent
    amax = out if out is not None else torch.empty(n, dtype=torch.float32, device=dev)
    if n and local is None:
        ops.amax_multi(tensors, out=amax)  # one launch per 64 shards
        torch.nan_to_num_(amax, nan=float("inf"))
    elif n:  # the gloo unit tests inject a CPU statistic
        for i, t in enumerate(tensors):
            a, b = local(t, None)
            mnmx[0, i], mnmx[1, i] = a.reshape(()), b.reshape(())
        torch.maximum(-mnmx[0, :n], mnmx[1, :n], out=amax)
        torch.nan_to_num_(amax, nan=float("inf"))
    if n_split and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buf = torch.zeros(n_split, dtype=torch.float32, device=dev)
        if src_idx.numel():
            buf[dst_idx] = amax[src_idx]
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
        if src_idx.numel():
            amax[src_idx] = buf[dst_idx]
    return amax


def qparams_from_minmax(mn: torch.Tensor, mx: torch.Tensor, fmt, symmetric: bool = True, eps: float = torch.finfo(torch.float32).eps):
    """scale / zero-point exactly as MinMaxObserver._calculate_qparams (reference
    S/numerical/observer.py:59-115) computes them from (reduced) statistics."""
    from .numerical.observer import _get_qmin_qmax

    qmin, qmax = _get_qmin_qmax(fmt)
    lo, hi = torch.clamp(mn, max=0.0), torch.clamp(mx, min=0.0)
    e = torch.tensor([eps], device=mn.device)
    if symmetric:
        scale = torch.max(torch.max(-lo, hi) / (float(qmax - qmin) / 2), e)
        zp = torch.zeros_like(scale, dtype=torch.int64)
    else:
        scale = torch.max((hi - lo) / float(qmax - qmin), e)
        zp = torch.clamp(qmin - torch.round(lo / scale).to(torch.int), qmin, qmax)
    return scale, zp


def sbfp_scaler_bias_from_amax(amax: float, scaler_exponent_bits: int = 4, man_scaling: int = 7) -> int:
    """Exponent bias of the SBFP scaler format from a tensor-wide amax.

    The reference delegates this choice to d-Matrix's private ``numerics`` module
    (S/numerical/format.py:13-20, 438-446), absent from the public repo -- parity is unpinned.
    Rule used here (hardware-faithful reading of an E-bit exponent field): the largest block
    scaler of the tensor, amax / man_scaling, must fall in the top binade the field can encode,
    2^((2^E - 1) - bias); smaller block scalers then use the binades below it."""
    import math

    default = 2 ** (scaler_exponent_bits - 1) - 1
    if not (amax > 0) or math.isinf(amax) or math.isnan(amax):
        return default
    top = math.floor(math.log2(amax / man_scaling))
    bias = (2 ** scaler_exponent_bits - 1) - top
    lo = 127 if scaler_exponent_bits == 8 else -128 + 2 ** scaler_exponent_bits
    return int(max(lo, min(127, bias)))

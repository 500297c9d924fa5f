"""Host-side multi-GPU logic on CPU: shard planning and the batched statistics all-reduce over a
world_size-2 gloo group (the N > 1 path of SURVEY.md section 8e).  No cast kernels run here."""
import os
import sys

import pytest
import numpy as np
import torch
import torch.multiprocessing as mp

from dmx_compressor_b200 import parallel as P

def _cpu_minmax(t, ch_axis=None):
    """stand-in for the dmxq_minmax kernel so that the *reduction logic* can be exercised over gloo without a GPU"""
    if ch_axis is None:
        return t.float().amin().reshape(1), t.float().amax().reshape(1)
    dims = [d for d in range(t.dim()) if d != ch_axis % t.dim()]
    return t.float().amin(dims), t.float().amax(dims)


LLAMA8B = {f"layers.{i}.{n}": s for i in range(4) for n, s in
           (("q", (4096, 4096)), ("k", (1024, 4096)), ("v", (1024, 4096)), ("o", (4096, 4096)),
            ("gate", (14336, 4096)), ("up", (14336, 4096)), ("down", (4096, 14336)))}
LLAMA8B["lm_head"] = (128256, 4096)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_covers_every_row_exactly_once(world):
    plan = P.plan_shards(LLAMA8B, world, row_align=1)
    seen = {}
    for rank, shards in enumerate(plan):
        for sh in shards:
            assert 0 <= sh.row0 < sh.row1 <= sh.rows == LLAMA8B[sh.name][0]
            seen.setdefault(sh.name, []).append((sh.row0, sh.row1))
    assert set(seen) == set(LLAMA8B)
    for n, rs in seen.items():
        rs.sort()
        assert rs[0][0] == 0 and rs[-1][1] == LLAMA8B[n][0]
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:])), f"{n}: gaps or overlaps {rs}"
    assert P.plan_imbalance(plan, LLAMA8B) < 1.15


def test_plan_row_alignment_for_dim0_blocks():
    # conv-style block_dim = 0 tensors must be cut at multiples of the block / group size
    plan = P.plan_shards({"w": (1000, 64)}, 4, split_threshold=0.1, row_align=8)
    for shards in plan:
        for sh in shards:
            assert sh.row0 % 8 == 0 and (sh.row1 % 8 == 0 or sh.row1 == 1000)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    full = [torch.randn(64, 48, generator=g) * 3, torch.randn(32, 16, generator=g), torch.randn(10, 8, generator=g)]
    full[2][3, 4] = float("nan")
    shards = [t.chunk(world, 0)[rank] for t in full]
    got = P.sharded_minmax(shards, [None, 1, None], local=_cpu_minmax)
    q.put((rank, [(a.numpy().copy(), b.numpy().copy()) for a, b in got]))  # (plain arrays: tensors travel by fd and race with exit)
    dist.destroy_process_group()


def test_sharded_minmax_equals_single_device_bitwise():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    g = torch.Generator().manual_seed(7)
    full = [torch.randn(64, 48, generator=g) * 3, torch.randn(32, 16, generator=g), torch.randn(10, 8, generator=g)]
    want = [(full[0].amin().reshape(1), full[0].amax().reshape(1)), (full[1].amin(0), full[1].amax(0))]
    for rank in range(world):
        for (mn, mx), (wmn, wmx) in zip(res[rank][:2], want):
            assert np.array_equal(mn, wmn.numpy()) and np.array_equal(mx, wmx.numpy())
        assert np.isnan(res[rank][2][0]).all() and np.isnan(res[rank][2][1]).all()  # NaN propagates


def test_qparams_match_observer_formula():
    from dmx_compressor_b200.numerical import Format, MinMaxObserver

    fmt = Format.from_shorthand("XP[8,0](CSN)")
    x = torch.randn(1000) * 5
    for scheme, sym in ((torch.per_tensor_symmetric, True), (torch.per_tensor_affine, False)):
        obs = MinMaxObserver(dtype=fmt, qscheme=scheme)
        obs.min_val, obs.max_val = x.amin(), x.amax()
        s, z = obs.calculate_qparams()
        s2, z2 = P.qparams_from_minmax(x.amin().reshape(1), x.amax().reshape(1), fmt, symmetric=sym)
        assert torch.equal(s, s2) and torch.equal(z.to(torch.int64), z2.to(torch.int64))


def test_sbfp_bias_rule():
    assert P.sbfp_scaler_bias_from_amax(7 * 2.0**8) == 7       # top binade 2^8 <-> E4 bias 7
    assert P.sbfp_scaler_bias_from_amax(7 * 2.0**-3) == 18
    assert P.sbfp_scaler_bias_from_amax(0.0) == 7


def _worker_stats(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"big": (64, 32), "a": (8, 32), "b": (6, 32), "c": (4, 32)}
    plan = P.plan_shards(shapes, world, split_threshold=0.5)
    g = torch.Generator().manual_seed(11)
    full = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    mine = [full[sh.name][sh.row0:sh.row1] for sh in plan[rank]]
    stats = P.shard_stats(plan, rank, mine, local=_cpu_minmax)
    q.put((rank, [(sh.name, float(a), float(b)) for sh, (a, b) in zip(plan[rank], stats)]))
    dist.destroy_process_group()


def test_shard_stats_mixed_whole_and_split_tensors():
    """ranks own different whole tensors plus a slice of the split one: one all-reduce, exact result"""
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_stats, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    shapes = {"big": (64, 32), "a": (8, 32), "b": (6, 32), "c": (4, 32)}
    g = torch.Generator().manual_seed(11)
    full = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    seen = set()
    for rank in range(world):
        for name, mn, mx in res[rank]:
            assert mn == float(full[name].min()) and mx == float(full[name].max()), name
            seen.add(name)
    assert seen == set(shapes)


def _worker_replicate(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"big": (65, 16), "a": (8, 16), "b": (6, 16), "c": (4, 16)}  # 65 rows: uneven split
    plan = P.plan_shards(shapes, world, split_threshold=0.5)
    g = torch.Generator().manual_seed(13)
    full = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    # (the cast itself is a GPU kernel; what is exercised here is the exchange: every rank contributes "its" rows doubled)
    local = {(sh.name, sh.row0): 2 * full[sh.name][sh.row0:sh.row1] for sh in plan[rank]}
    out = P.replicate_shards(plan, local, shapes, torch.float32, "cpu")
    q.put((rank, {n: t.numpy().copy() for n, t in out.items()}))
    dist.destroy_process_group()


def test_replicate_shards_over_gloo():
    """every rank ends up with every tensor, rows in place, uneven row splits included"""
    world, port = 2, 33500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_replicate, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    shapes = {"big": (65, 16), "a": (8, 16), "b": (6, 16), "c": (4, 16)}
    g = torch.Generator().manual_seed(13)
    full = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    for rank in range(world):
        assert set(res[rank]) == set(shapes)
        for n in shapes:
            assert np.array_equal(res[rank][n], (2 * full[n]).numpy()), (rank, n)


def _worker_amax(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"big": (64, 32), "big2": (48, 32), "a": (8, 32), "b": (6, 32), "c": (4, 32), "nan": (5, 32)}
    plan = P.plan_shards(shapes, world, split_threshold=0.4)
    g = torch.Generator().manual_seed(17)
    full = {n: torch.randn(s, generator=g) * (i + 1) for i, (n, s) in enumerate(shapes.items())}
    full["nan"][2, 3] = float("nan")
    full["big"][60, 1] = -1e6  # the tensor-wide amax lives in the LAST rank's rows, and is a negative value
    mine = [full[sh.name][sh.row0:sh.row1] for sh in plan[rank]]
    for _ in range(2):  # second call: the cached index plan
        amax = P.shard_amax(plan, rank, mine, local=_cpu_minmax)
    q.put((rank, [(sh.name, float(a)) for sh, a in zip(plan[rank], amax)]))
    dist.destroy_process_group()


def test_shard_amax_one_vector_one_allreduce():
    """shard_amax: per-shard statistic -> one device vector in plan order, the row-split tensors reduced over the ranks"""
    world, port = 2, 35500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_amax, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    shapes = {"big": (64, 32), "big2": (48, 32), "a": (8, 32), "b": (6, 32), "c": (4, 32), "nan": (5, 32)}
    g = torch.Generator().manual_seed(17)
    full = {n: torch.randn(s, generator=g) * (i + 1) for i, (n, s) in enumerate(shapes.items())}
    full["big"][60, 1] = -1e6
    seen = {}
    for rank in range(world):
        for name, a in res[rank]:
            want = float("inf") if name == "nan" else float(full[name].abs().max())
            assert a == want, (name, a, want)
            seen[name] = seen.get(name, 0) + 1
    assert set(seen) == set(shapes) and seen["big"] == world and seen["a"] == 1


def test_allreduce_minmax_keeps_real_infinities():
    """a channel that really holds +inf and -inf is (-inf, +inf), not NaN; a NaN channel is NaN (single process: no group)"""
    mn = torch.tensor([-float("inf"), -1.0, float("nan"), 2.0])
    mx = torch.tensor([float("inf"), 3.0, float("nan"), float("inf")])
    (rmn, rmx), = P.allreduce_minmax([(mn, mx)])
    assert rmn[0] == -float("inf") and rmx[0] == float("inf")
    assert rmn[1] == -1.0 and rmx[1] == 3.0
    assert torch.isnan(rmn[2]) and torch.isnan(rmx[2])
    assert rmn[3] == 2.0 and rmx[3] == float("inf")

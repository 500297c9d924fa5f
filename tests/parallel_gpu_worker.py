"""torchrun worker of tests/test_parallel_gpu.py: the sharded whole-model weight cast on N GPUs over NCCL must reproduce the
single-GPU cast bit for bit -- 2:4 -> BFP12 (no collective) and SBFP with the scaler bias taken from the
batched amax all-reduce."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200 import parallel as P
    from dmx_compressor_b200.numerical import Format

    shapes = {"lm_head": (4099, 512), "big": (2048, 512)}
    shapes.update({f"layers.{i}.{n}": s for i in range(3) for n, s in (("q", (256, 512)), ("k", (64, 512)), ("gate", (896, 512)), ("down", (512, 896)))})
    plan = P.plan_shards(shapes, world, split_threshold=0.5)
    g = torch.Generator().manual_seed(5)
    full = {n: (torch.randn(s, generator=g) * 0.02 * (1 + i % 5)).to(torch.bfloat16) for i, (n, s) in enumerate(shapes.items())}
    full["big"][2047, 3] = 11.0  # the tensor-wide amax sits in the last rank's rows
    mine = plan[rank]
    ws = [full[sh.name][sh.row0:sh.row1].to(dev) for sh in mine]
    ok = True
    # config #4: 2:4 -> BFP12, no collective
    st = [ops.nm_stage(2, 4), Format.from_shorthand("BFP[4|8]{64}(SN)").stage()]
    ys = ops.cast_chain_multi(ws, st, -1)
    rep = P.replicate_shards(plan, {(sh.name, sh.row0): y for sh, y in zip(mine, ys)}, shapes, torch.bfloat16, dev)
    for n in shapes:
        want = ops.cast_chain(full[n].to(dev), st, -1)
        ok &= bool(torch.equal(rep[n].view(torch.int16), want.view(torch.int16)))
    # config #5: SBFP, scaler bias from the all-reduced amax, never leaving the device
    amax = P.shard_amax(plan, rank, ws)
    sb = [Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()]
    ys = ops.cast_chain_multi(ws, sb, -1, amax=amax)
    rep = P.replicate_shards(plan, {(sh.name, sh.row0): y for sh, y in zip(mine, ys)}, shapes, torch.bfloat16, dev)
    biases = set()
    for n in shapes:
        a = float(full[n].float().abs().max())
        b = P.sbfp_scaler_bias_from_amax(a)
        biases.add(b)
        want = ops.cast_chain(full[n].to(dev), [Format.from_shorthand(f"SBFP<XP[4,0](CSN)><FP[0|4|4,{b}](FN)>{{16}}").stage()], -1)
        ok &= bool(torch.equal(rep[n].view(torch.int16), want.view(torch.int16)))
    ok &= len(biases) > 1
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_EQUALS_SINGLE_GPU" if int(t) == 1 else "MISMATCH", world, flush=True)
    dist.destroy_process_group()
    return 0 if int(t) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())

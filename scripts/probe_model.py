"""whole-model (Llama-3-8B shapes, bf16) weight cast: kernel kinds side by side, and the footprint effect (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dmx_compressor_b200 import ops, parallel as P
from dmx_compressor_b200.numerical import Format
dev = torch.device("cuda", 0)
shapes = bench.llama_shapes("8b")
plan = P.plan_shards(shapes, 1)
ws = [bench.make_shard(sh.name, shapes[sh.name], sh.row0, sh.row1, dev, torch.bfloat16) for sh in plan[0]]
outs = [torch.empty_like(w) for w in ws]
nb = sum(4 * w.numel() for w in ws)
F = Format.from_shorthand
chains = {"BFP16": [F("BFP[8|8]{64}(SN)").stage()], "BFP12": [F("BFP[4|8]{64}(SN)").stage()], "2:4": [ops.nm_stage(2, 4)],
          "2:4->BFP12": [ops.nm_stage(2, 4), F("BFP[4|8]{64}(SN)").stage()], "2:4->BFP16": [ops.nm_stage(2, 4), F("BFP[8|8]{64}(SN)").stage()],
          "SBFP": [F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()]}
def t(fn, n=3):
    fn(); best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for name, st in chains.items():
    ms = t(lambda: ops.cast_chain_multi(ws, st, -1, outs=outs))
    print(f"whole model {name:12s} {nb / ms / 1e6:7.0f} GB/s  {ms:.3f} ms", flush=True)
# in place (y == x): half the footprint
st = chains["2:4->BFP12"]
ms = t(lambda: ops.cast_chain_multi(ws, st, -1, outs=ws))
print(f"whole model 2:4->BFP12 in place {nb / ms / 1e6:7.0f} GB/s  {ms:.3f} ms")
# first 1/8 of the tensors only (footprint 4 GB)
k = len(ws) // 8
nb8 = sum(4 * w.numel() for w in ws[:k])
ms = t(lambda: ops.cast_chain_multi(ws[:k], st, -1, outs=outs[:k]))
print(f"first {k} tensors 2:4->BFP12 {nb8 / ms / 1e6:7.0f} GB/s  {ms:.3f} ms")
ms = t(lambda: torch._foreach_copy_(outs, ws))
print(f"torch._foreach_copy_ whole model {nb / ms / 1e6:7.0f} GB/s  {ms:.3f} ms")

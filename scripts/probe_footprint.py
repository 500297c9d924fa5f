"""HBM-bound BFP16 cast vs working-set size and tensor count (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = torch.device("cuda", 0)
st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
def t(fn, n=3):
    fn(); best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for gib in (1, 4, 8, 16):
    n = gib * (1 << 30) // 2
    x = torch.empty(n // 4096, 4096, device=dev, dtype=torch.bfloat16).normal_()
    y = torch.empty_like(x)
    ms = t(lambda: ops.cast_chain(x, st, -1, out=y))
    print(f"one tensor {gib:2d} GiB in + {gib} GiB out: {4 * n / ms / 1e6:7.0f} GB/s", flush=True)
    ms = t(lambda: y.copy_(x))
    print(f"   torch copy_                      : {4 * n / ms / 1e6:7.0f} GB/s", flush=True)
    for parts in (16, 256):
        r = x.shape[0] // parts
        xs = [x[i * r:(i + 1) * r] for i in range(parts)]
        ys = [y[i * r:(i + 1) * r] for i in range(parts)]
        ms = t(lambda: ops.cast_chain_multi(xs, st, -1, outs=ys))
        print(f"   multi, {parts:3d} slices              : {4 * n / ms / 1e6:7.0f} GB/s", flush=True)
    del x, y
    torch.cuda.empty_cache()

"""why is the whole-model 2:4 -> BFP12 cast slower than the standalone kernel?  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = torch.device("cuda", 0)
st = [ops.nm_stage(2, 4), Format.from_shorthand("BFP[4|8]{64}(SN)").stage()]
sb = [Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()]

def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

g = torch.Generator(device=dev).manual_seed(0)
R = 65536
for name, chain in (("2:4->BFP12", st), ("SBFP", sb)):
    x_rows = (torch.randn(R, 4096, device=dev, generator=g) * torch.pow(2.0, torch.randint(-8, 9, (R, 1), device=dev, generator=g).float())).bfloat16()
    x_w = (torch.randn(R, 4096, device=dev, generator=g) * 0.02).bfloat16()
    y = torch.empty_like(x_w)
    nb = 4 * x_w.numel()
    for label, x in (("row-scaled", x_rows), ("weights*0.02", x_w)):
        ms = t(lambda: ops.cast_chain(x, chain, -1, out=y))
        print(name, label, "single", round(nb / ms / 1e6), "GB/s")
        ms = t(lambda: ops.cast_chain_multi([x], chain, -1, outs=[y]))
        print(name, label, "multi x1", round(nb / ms / 1e6), "GB/s")
    # many tensors: 8 x [8192,4096] slices of the same buffer vs separately allocated
    xs = [x_w[i * 8192:(i + 1) * 8192] for i in range(8)]
    ys = [y[i * 8192:(i + 1) * 8192] for i in range(8)]
    ms = t(lambda: ops.cast_chain_multi(xs, chain, -1, outs=ys))
    print(name, "multi x8 slices", round(nb / ms / 1e6), "GB/s")
    xs2 = [s.clone() for s in xs]; ys2 = [torch.empty_like(s) for s in xs]
    ms = t(lambda: ops.cast_chain_multi(xs2, chain, -1, outs=ys2))
    print(name, "multi x8 separate allocations", round(nb / ms / 1e6), "GB/s")
    # llama-like mix: 32 x (k [1024,4096], q [4096,4096], gate [14336,4096])
    mix = []
    for i in range(12):
        for r in (1024, 4096, 14336):
            mix.append((torch.randn(r, 4096, device=dev, generator=g) * 0.02).bfloat16())
    outs = [torch.empty_like(m) for m in mix]
    nbm = sum(4 * m.numel() for m in mix)
    ms = t(lambda: ops.cast_chain_multi(mix, chain, -1, outs=outs), 5)
    print(name, "multi llama-like mix", round(nbm / ms / 1e6), "GB/s", round(ms, 3), "ms")
    ms = t(lambda: [ops.cast_chain(m, chain, -1, out=o) for m, o in zip(mix, outs)], 5)
    print(name, "per-tensor loop, same mix", round(nbm / ms / 1e6), "GB/s", round(ms, 3), "ms")
    del mix, outs, xs2, ys2

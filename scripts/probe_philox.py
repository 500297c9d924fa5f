"""stochastic BFP16: external random tensor (as the reference draws it) vs in-kernel Philox, 2^28 elements (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = "cuda:0"
n = 1 << 28


def t(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(n // 4096, 4096, device=dev).to(dt)
    y = torch.empty_like(x)
    r = torch.empty(x.shape, dtype=torch.int32, device=dev)
    for sh in ("BFP[8|8]{64}(SS)", "FP[1|4|3,7](_S)", "XP[8,0](CSS)"):
        st = [Format.from_shorthand(sh).stage()]
        fixed = sh.startswith("XP")
        rr = torch.rand(x.shape, device=dev) if fixed else r
        draw = (lambda: torch.rand(x.shape, device=dev, out=rr)) if fixed else (lambda: r.random_(0, 2**31 - 1))
        t_draw = t(draw)
        t_ext = t(lambda: ops.cast_chain(x, st, -1, out=y, rand=rr))
        t_ph = t(lambda: ops.cast_chain(x, st, -1, out=y, philox=(1234, 5)))
        es = x.element_size()
        print(f"{str(dt):15s} {sh:20s} torch draw {t_draw:6.3f} ms + cast with tensor {t_ext:6.3f} ms = {t_draw + t_ext:6.3f} ms | in-kernel Philox {t_ph:6.3f} ms "
              f"= {2 * es * n / t_ph / 1e6:6.0f} GB/s of the cast's own bytes ({(t_draw + t_ext) / t_ph:4.2f}x)", flush=True)

"""Device-side throughput of the N:M -> BFP weight-cast kernels (development aid, not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
from probe_bw import report

dev = "cuda:0"
n = 2**28
for dt in (torch.bfloat16, torch.float16, torch.float32):
    x = (torch.randn(n // 4096, 4096, device=dev) * 0.02).to(dt)
    y = torch.empty_like(x)
    es = x.element_size()
    for sh in ("BFP[4|8]{64}(SN)", "BFP[8|8]{64}(SN)", "BFP[4|8]{128}(SN)"):
        st = [ops.nm_stage(2, 4), Format.from_shorthand(sh).stage()]
        report(f"2:4 -> {sh} {dt}", 2 * n * es, lambda: ops.cast_chain(x, st, -1, out=y))
    report(f"2:4 alone {dt}", 2 * n * es, lambda: ops.cast_chain(x, [ops.nm_stage(2, 4)], -1, out=y))
    del x, y

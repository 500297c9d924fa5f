"""Device-side throughput of the packed SBFP12_16 storage kernels (development aid, not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
from probe_bw import report  # noqa: F401  (runs its own sweep only under __main__)

dev = "cuda:0"
n = 2**28
f = Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")
st = f.stage()
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(n // 4096, 4096, device=dev).to(dt)
    es = x.element_size()
    packed = n // 2 + n // 16
    y = torch.empty_like(x)
    report(f"SBFP12_16 qdq {dt}", 2 * n * es, lambda: ops.cast_chain(x, [st], -1, out=y))
    report(f"SBFP12_16 pack -> nibbles + scaler byte {dt}", n * es + packed, lambda: ops.sbfp_pack(x, st))
    m, s = ops.sbfp_pack(x, st)
    report(f"SBFP12_16 unpack {dt}", n * es + packed, lambda: ops.sbfp_unpack(m, s, st, dtype=dt))
    del x, y, m, s

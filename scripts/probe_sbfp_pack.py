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
    sc, zp = torch.full((1,), 0.037, device=dev), torch.full((1,), 3.0, device=dev)
    report(f"INT8 calibrated (device scale, zero-point) {dt}", 2 * n * es, lambda: ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=sc, zero_point=zp, out=y))
    xc = x.view(-1, 4096)[:16384]
    yc = torch.empty_like(xc)
    scc, zpc = torch.rand(4096, device=dev) * 0.05 + 0.01, torch.zeros(4096, device=dev)
    scr, zpr = torch.rand(16384, device=dev) * 0.05 + 0.01, torch.zeros(16384, device=dev)
    report(f"INT8 per-column qparams [16384,4096] {dt}", 2 * xc.numel() * es, lambda: ops.fixed_qdq(xc, 8, 0, True, True, "nearest", scale=scc, zero_point=zpc, ch_axis=1, out=yc))
    report(f"INT8 per-row qparams [16384,4096] {dt}", 2 * xc.numel() * es, lambda: ops.fixed_qdq(xc, 8, 0, True, True, "nearest", scale=scr, zero_point=zpr, ch_axis=0, out=yc))
    del x, y, m, s

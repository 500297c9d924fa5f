"""Tiny driver for ncu captures: runs each profiled cast a few times (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 2**28
st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
if which in ("all", "rows_f32"):
    x = torch.randn(n // 4096, 4096, device=dev); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, st, -1, out=y)
if which in ("all", "rows_bf16"):
    x = torch.randn(n // 4096, 4096, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, st, -1, out=y)
if which in ("all", "cols_f32"):
    x = torch.randn(96, 2048, 1024, device=dev); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, st, -2, out=y)
if which in ("all", "float16_f32"):
    x = torch.randn(n // 4096, 4096, device=dev); y = torch.empty_like(x)
    st2 = [Format.from_shorthand("FP[1|5|10,15](FN)").stage()]
    for _ in range(3): ops.cast_chain(x, st2, -1, out=y)
torch.cuda.synchronize()
if which in ("nm_bf16",):
    x = torch.randn(n // 4096, 4096, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, [ops.nm_stage(2, 4)], -1, out=y)
if which in ("nm_bfp_bf16",):
    x = torch.randn(n // 4096, 4096, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, [ops.nm_stage(2, 4), Format.from_shorthand("BFP[4|8]{64}(SN)").stage()], -1, out=y)
if which in ("float_bfp_bf16",):
    x = torch.randn(n // 4096, 4096, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, [Format.from_shorthand("FP[1|5|10,15](FN)").stage(), Format.from_shorthand("BFP[8|8]{64}(SN)").stage()], -1, out=y)
torch.cuda.synchronize()
if which in ("cols_bf16",):
    x = torch.randn(96, 2048, 2048, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, st, -2, out=y)
if which in ("sbfp_bf16", "sbfp_f32"):
    x = torch.randn(n // 4096, 4096, device=dev)
    if which == "sbfp_bf16": x = x.bfloat16()
    y = torch.empty_like(x)
    for _ in range(3): ops.cast_chain(x, [Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()], -1, out=y)
torch.cuda.synchronize()
if which in ("int8cal_bf16", "int8row_bf16", "int8cal_f32", "int8row_f32"):
    x = torch.randn(n // 4096, 4096, device=dev)
    if which.endswith("bf16"): x = x.bfloat16()
    y = torch.empty_like(x)
    if "cal" in which:
        sc, zp = torch.full((1,), 0.037, device=dev), torch.full((1,), 3.0, device=dev)
        for _ in range(3): ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=sc, zero_point=zp, out=y)
    else:
        sc, zp = torch.rand(n // 4096, device=dev) * 0.05 + 0.01, torch.zeros(n // 4096, device=dev)
        for _ in range(3): ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=sc, zero_point=zp, ch_axis=0, out=y)
if which in ("mxfp_bf16", "fp8_bf16", "int8_bf16"):
    x = torch.randn(n // 4096, 4096, device=dev).bfloat16(); y = torch.empty_like(x)
    sh = {"mxfp_bf16": "MXFP8[E4M3]{32}", "fp8_bf16": "FP[1|4|3,7](_N)", "int8_bf16": "XP[8,0](CSN)"}[which]
    for _ in range(3): ops.cast_chain(x, [Format.from_shorthand(sh).stage()], -1, out=y)
torch.cuda.synchronize()

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

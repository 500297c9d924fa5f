"""ncu driver: one launch of each dmxq_softmax_cast variant on the OPT-125m attention shape (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
F = lambda sh: Format.from_shorthand(sh).stage()
dev = "cuda:0"
for dt in (torch.bfloat16, torch.float32):
    x = (torch.randn(96, 2048, 2048, device=dev) * 3).to(dt)
    mask = torch.full((2048, 2048), float("-inf"), device=dev).triu(1).to(dt)[None, None].expand(8, 1, 2048, 2048).contiguous()
    x4 = x.view(8, 12, 2048, 2048)
    y = torch.empty_like(x)
    f16 = F("FP[1|5|10,15](FN)")
    post = [f16, F("BFP[8|8]{64}(SN)")]
    ops.softmax_cast(x, out=y)
    ops.softmax_cast(x, post, out=y)
    ops.softmax_cast(x4, post, addend=mask.expand(8, 12, 2048, 2048), stage_x=f16, stage_addend=f16, stage_sum=f16, out=y.view_as(x4))
    torch.softmax(x, -1)
    torch.cuda.synchronize()

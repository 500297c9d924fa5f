"""dmxq_softmax_cast vs torch.softmax (+ the unfused add / cast passes around it) on the OPT-125m attention shape (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
F = lambda sh: Format.from_shorthand(sh).stage()
dev = "cuda:0"


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for dt in (torch.bfloat16, torch.float32):
    x = (torch.randn(96, 2048, 2048, device=dev) * 3).to(dt)
    mask = torch.full((2048, 2048), float("-inf"), device=dev).triu(1).to(dt)[None, None].expand(8, 1, 2048, 2048).contiguous()
    x4 = x.view(8, 12, 2048, 2048)
    y = torch.empty_like(x)
    f16 = F("FP[1|5|10,15](FN)")
    post = [f16, F("BFP[8|8]{64}(SN)")]
    nbytes = 2 * x.numel() * x.element_size()
    rows = [("torch.softmax", lambda: torch.softmax(x, -1)),
            ("dmxq softmax", lambda: ops.softmax_cast(x, out=y)),
            ("dmxq softmax -> FLOAT16 -> BFP16", lambda: ops.softmax_cast(x, post, out=y)),
            ("dmxq mask add + casts -> softmax -> FLOAT16 -> BFP16", lambda: ops.softmax_cast(x4, post, addend=mask.expand(8, 12, 2048, 2048), stage_x=f16, stage_addend=f16, stage_sum=f16, out=y.view_as(x4))),
            ("unfused: add_cast, torch.softmax, cast_chain", lambda: ops.cast_chain(torch.softmax(ops.add_cast(x4, mask.expand(8, 12, 2048, 2048), f16, f16, f16), -1), post, -1))]
    for name, fn in rows:
        ms = t(fn)
        print(f"{str(dt):16s} {name:60s} {ms:7.3f} ms   {nbytes / ms / 1e6:7.0f} GB/s (read + write once)", flush=True)

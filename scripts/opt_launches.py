"""one elided BASIC forward of the OPT stack for an ncu launch list (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import elide, opt
dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
q, p = opt.build_pair(device="cuda:0", dtype=dt)
ids = torch.randint(0, 50272, (8, 2048), device="cuda:0")
with torch.no_grad(), elide.enabled():
    q(ids); q(ids)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    q(ids)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()

"""ncu driver: the fused attention kernel as the OPT-125m bf16 forward launches it (development aid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dmx_compressor_b200 import elide, opt
q, p = opt.build_pair(device="cuda:0", dtype=torch.bfloat16)
ids = torch.randint(0, 50272, (8, 2048), device="cuda:0")
with torch.no_grad(), elide.enabled():
    for _ in range(2):
        elide.materialise(q(ids))
torch.cuda.synchronize()

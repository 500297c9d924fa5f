"""which cast chains one elided OPT-125m bf16 forward issues: (stage kinds, shape, block_dim) -> count (development aid)"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import _lib as L, elide, ops, opt
q, p = opt.build_pair(device="cuda:0", dtype=torch.bfloat16)
ids = torch.randint(0, 50272, (8, 2048), device="cuda:0")
census = collections.Counter()
orig = ops.cast_chain
NAMES = {1: "NM", 2: "BFP", 3: "SBFP", 4: "FLOAT", 5: "FIXED", 6: "MXFP", 7: "SCALE"}


def spy(x, stages, block_dim=-1, **kw):
    desc = tuple(f"{NAMES[s.kind]}(b{s.block},p{s.precision},m{s.man},e{s.exp},fl{s.flush})" for s in stages)
    census[(desc, tuple(x.shape), tuple(x.stride()), block_dim, str(kw.get("out_dtype")))] += 1
    return orig(x, stages, block_dim, **kw)


with torch.no_grad(), elide.enabled():
    for _ in range(2):
        elide.materialise(q(ids))
    ops.cast_chain = spy
    import dmx_compressor_b200.numerical.cast as C1, dmx_compressor_b200.nn as N1, dmx_compressor_b200.elide as E1
    elide.materialise(q(ids))
    ops.cast_chain = orig
for k, v in sorted(census.items(), key=lambda kv: -kv[1]):
    print(v, k)

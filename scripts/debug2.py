import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import BlockFloatingPoint
from test_parity_gpu import _rand
DEV="cuda:0"
dt=torch.float16; wl,bs=8,64
x = _rand((96, 1024), 100 + wl + bs, spread=12).to(dt)
x.view(-1)[::5] = torch.round(x.view(-1)[::5].float() * 16).to(dt) / 16
x[5] = 0
xd = x.to(DEV)
f = BlockFloatingPoint(wl, bs)
want = ops.cast_chain(xd, [f.stage()], -1)
mant, exps = ops.bfp_pack(xd, bs, wl)
got = ops.bfp_unpack(mant, exps, bs, wl, dtype=dt)
bad = (got.view(torch.int16) != want.view(torch.int16))
print("bad", int(bad.sum()), "of", bad.numel(), "nonfinite in x:", int((~torch.isfinite(x)).sum()))
idx = torch.nonzero(bad)[:10]
for r, c in idx.tolist():
    blk = xd[r, (c//64)*64:(c//64+1)*64].float()
    print(r, c, "x=%g want=%g got=%g blockmax=%g exp=%d mant=%d" % (float(xd[r,c]), float(want[r,c]), float(got[r,c]), float(blk.abs().max()), int(exps[r,c//64]), int(mant[r,c])))

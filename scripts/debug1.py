import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from util import case, f32
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
DEV="cuda:0"
for name in ("xp/7","xp/8"):
    m,d = case(name)
    x = f32(d["x"]).reshape(m["shape"])
    f = Format.from_shorthand(m["fmt"]); f.tie=1
    y = f.cast(torch.from_numpy(x).to(DEV)).cpu().numpy()
    w = f32(d["y"]).reshape(m["shape"])
    bad = np.argwhere(y.view(np.uint32)!=w.view(np.uint32))[:10]
    print(name, m["fmt"], len(bad))
    for i in bad: print("  x=%r got=%r want=%r"%(x[tuple(i)], y[tuple(i)], w[tuple(i)]))
torch.manual_seed(0)
for m,k in ((4,2),(8,4)):
    x = torch.round(torch.randn(4096, m, device=DEV)*2)/2
    score = x.abs()
    idx = torch.argsort(score, dim=1)[:, :m-k]
    mask = torch.ones_like(score).scatter_(dim=1, index=idx, value=0)
    want = x*mask
    got = ops.nm_prune(x, k, m, -1)
    bad = (got.view(torch.int32)!=want.view(torch.int32)).any(1)
    print(f"{k}:{m} rows differing: {int(bad.sum())}/4096")
    for r in torch.nonzero(bad)[:6].flatten().tolist():
        print("  score", score[r].tolist(), "torch idx", idx[r].tolist(), "ours keep", (got[r]!=0).int().tolist(), "torch keep", mask[r].int().tolist())
    idx2 = torch.argsort(score, dim=1, stable=True)[:, :m-k]
    mask2 = torch.ones_like(score).scatter_(dim=1, index=idx2, value=0)
    print("  vs stable argsort differing rows:", int(((x*mask2).view(torch.int32)!=got.view(torch.int32)).any(1).sum()))
    # bigger rows count
    xb = torch.round(torch.randn(1<<20, m, device=DEV)*2)/2
    idx = torch.argsort(xb.abs(), dim=1)[:, :m-k]
    idxs = torch.argsort(xb.abs(), dim=1, stable=True)[:, :m-k]
    print("  unstable==stable for 1M rows:", bool((idx==idxs).all()))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev="cuda:0"
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
F16 = Format.from_shorthand("FP[1|5|10,15](FN)").stage()
BFP = Format.from_shorthand("BFP[8|8]{64}(SN)").stage()
for dt in (torch.float32, torch.bfloat16):
    a = torch.randn(8,12,2048,2048, device=dev).to(dt); y = torch.empty_like(a)
    S=2048
    mask = torch.full((S,S), torch.finfo(dt).min, device=dev, dtype=dt).triu(1)[None,None].expand(8,1,S,S)
    b2 = torch.randn_like(a)
    es = a.element_size(); n=a.numel()
    t = timeit(lambda: torch.add(a, mask, out=y)); print(f"{dt} torch add mask        {t:.3f} ms  {2*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: ops.add_cast(a, mask, F16, F16, F16, out=y)); print(f"{dt} add_cast mask (3 casts) {t:.3f} ms  {2*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: ops.add_cast(a, mask, None, None, F16, out=y)); print(f"{dt} add_cast mask (out cast) {t:.3f} ms  {2*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: torch.add(a, b2, out=y)); print(f"{dt} torch add same        {t:.3f} ms  {3*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: ops.add_cast(a, b2, F16, F16, F16, out=y)); print(f"{dt} add_cast same (3 casts) {t:.3f} ms  {3*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: ops.cast_chain(a, [F16, BFP], -1, out=y)); print(f"{dt} FLOAT16->BFP16 chain    {t:.3f} ms  {2*n*es/t/1e6:.0f} GB/s")
    t = timeit(lambda: torch.softmax(a, -1, out=y) if False else y.copy_(torch.softmax(a,-1))); print(f"{dt} torch softmax+copy      {t:.3f} ms")
    t = timeit(lambda: torch.softmax(a, -1)); print(f"{dt} torch softmax      {t:.3f} ms")

"""host-side cost of one cast call through the module path (development aid)"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from dmx_compressor_b200 import elide, ops  # noqa: E402
from dmx_compressor_b200.numerical import CastTo, Format  # noqa: E402

x = torch.randn(64, 768, device="cuda")
for name, fn in (("CastTo(BFP16_64)", CastTo("BFP[8|8]{64}(SN)").cuda()), ("CastTo(FLOAT16)", CastTo("FP[1|5|10,15](FN)").cuda()),
                 ("ops.cast_chain prebuilt stage", None)):
    if fn is None:
        st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
        fn = lambda t: ops.cast_chain(t, st, -1)  # noqa: E731
    with torch.no_grad():
        for _ in range(200):
            fn(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5000):
            fn(x)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
    print(f"{name:32s} {1e6 * (t1 - t0) / 5000:6.2f} us per call (host)")
c = CastTo("BFP[8|8]{64}(SN)").cuda()
pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    for _ in range(3000):
        c(x)
    pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

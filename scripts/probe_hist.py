"""throughput of the calibration passes: dmxq_histc (with / without the fused amin/amax) and dmxq_minmax vs torch"""
import sys

import torch

sys.path.insert(0, ".")
from dmx_compressor_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


n = 2**28
for dt in (torch.float32, torch.bfloat16):
    x = (torch.randn(n, device="cuda") * 3).to(dt)
    nbytes = x.numel() * x.element_size()
    for name, fn in (("histc", lambda: ops.histc(x, 2048, min=-7, max=9)),
                     ("histc+minmax", lambda: ops.histc(x, 2048, min=-7, max=9, return_minmax=True)),
                     ("minmax", lambda: ops.minmax(x)),
                     ("torch.histc(x.float())", lambda: torch.histc(x.float(), 2048, min=-7, max=9)),
                     ("torch.aminmax", lambda: torch.aminmax(x))):
        ms = timeit(fn)
        print(f"{str(dt):16s} {name:24s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s (input bytes)")
    # narrow distribution: most values fall into a handful of bins (shared-memory atomic contention)
    xn = (torch.randn(n, device="cuda") * 0.01).to(dt)
    ms = timeit(lambda: ops.histc(xn, 2048, min=-7, max=9))
    print(f"{str(dt):16s} {'histc (3 hot bins)':24s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")

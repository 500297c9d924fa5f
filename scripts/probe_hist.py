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

# per-channel statistics (MinMaxObserver per_channel, SmoothQuant maxabs)
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(16384, 4096, device="cuda").to(dt)
    nbytes = x.numel() * x.element_size()
    for name, fn in (("minmax ch_axis=0 [16384,4096]", lambda: ops.minmax(x, 0)), ("minmax ch_axis=1 [16384,4096]", lambda: ops.minmax(x, 1)),
                     ("torch amin+amax dim=1", lambda: (x.amin(1), x.amax(1))), ("torch amin+amax dim=0", lambda: (x.amin(0), x.amax(0)))):
        ms = timeit(fn)
        print(f"{str(dt):16s} {name:32s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")
    a = torch.randn(8, 2048, 3072, device="cuda").to(dt)
    ms = timeit(lambda: ops.minmax(a, 2))
    print(f"{str(dt):16s} {'minmax ch_axis=2 [8,2048,3072]':32s} {ms:8.3f} ms  {a.numel() * a.element_size() / ms / 1e6:8.1f} GB/s")

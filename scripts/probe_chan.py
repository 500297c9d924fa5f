"""per-channel / group FixedPoint casts with device qparams (development aid)"""
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


for dt in (torch.float32, torch.bfloat16):
    x = (torch.randn(16384, 4096, device="cuda") * 3).to(dt)
    y = torch.empty_like(x)
    nbytes = 2 * x.numel() * x.element_size()
    for name, ax, nq, gs in (("per-channel axis 0", 0, 16384, None), ("per-channel axis 1", 1, 4096, None), ("groups of 128 along axis 1", 1, 32, 128),
                             ("groups of 64 rows (axis 0)", 0, 256, 64)):
        sc = torch.rand(nq, device="cuda") * 0.05 + 0.01
        zp = torch.randint(-3, 4, (nq,), device="cuda").float()
        ms = timeit(lambda: ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=sc, zero_point=zp, ch_axis=ax, group_size=gs, out=y))
        print(f"{str(dt):16s} INT8 {name:30s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")

# L1 block_quantize(x, wl, dim) (quant_cuda.block_quantize_nearest mirror)
x = torch.randn(16384, 4096, device="cuda")
for dim in (-1, 0, 1):
    ms = timeit(lambda: ops.block_quantize_l1(x, 8, dim, True, "nearest"))
    print(f"L1 block_quantize dim={dim:2d} fp32 [16384,4096]   {ms:8.3f} ms  {2 * x.numel() * 4 / ms / 1e6:8.1f} GB/s (8 B/elem)")

# Sparsify with a learnable score tensor and the mask written out (training-time path, K_AUX)
for dt in (torch.float32, torch.bfloat16):
    w = torch.randn(16384, 4096, device="cuda").to(dt)
    sc = torch.rand(16384, 4096, device="cuda")
    n = w.numel()
    ms = timeit(lambda: ops.nm_prune(w, 2, 4, -1, score=sc, return_mask=True))
    print(f"{str(dt):16s} 2:4 prune, fp32 score in, fp32 mask out   {ms:8.3f} ms  {(2 * n * w.element_size() + 8 * n) / ms / 1e6:8.1f} GB/s")
    ms = timeit(lambda: ops.nm_prune(w, 2, 4, -1, score=sc))
    print(f"{str(dt):16s} 2:4 prune, fp32 score in                  {ms:8.3f} ms  {(2 * n * w.element_size() + 4 * n) / ms / 1e6:8.1f} GB/s")
    ms = timeit(lambda: ops.nm_prune(w, 4, 8, -1))
    print(f"{str(dt):16s} 4:8 prune (score |x|)                     {ms:8.3f} ms  {(2 * n * w.element_size()) / ms / 1e6:8.1f} GB/s")

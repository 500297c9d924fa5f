"""device time of tiny casts (bias vectors) replayed from a CUDA graph (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = "cuda:0"
for dt in (torch.bfloat16, torch.float32):
    for n in (768, 3072, 65536):
        x = torch.randn(n, device=dev).to(dt)
        y = torch.empty_like(x)
        for sh in ("BFP[24|8]{1}(SN)", "FP[1|5|10,15](FN)", "BFP[8|8]{64}(SN)", "FP[1|8|22,127](_N)"):
            st = [Format.from_shorthand(sh).stage()]
            ops.cast_chain(x, st, -1, out=y)
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                ops.cast_chain(x, st, -1, out=y)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                for _ in range(50):
                    ops.cast_chain(x, st, -1, out=y)
            g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            print(f"{str(dt):15s} n={n:6d} {sh:20s} {a.elapsed_time(b) / 50 * 1e3:7.2f} us per launch", flush=True)

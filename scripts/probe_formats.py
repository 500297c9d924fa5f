"""per-format kernel throughput on 2^28-element tensors (bench.roofline_by_format), one row per line (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
peak, _ = bench.load_peak()
for r in bench.roofline_by_format(torch.device("cuda", 0), peak):
    print(r, flush=True)

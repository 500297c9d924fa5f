"""OPT-125m-shaped stack of the REFERENCE's own modules, batch 8 x 2048: unquantised twin, reference unpatched (its CUDA path),
plugin drop-in, plugin + elision (development aid; bench.py reports the same under `plugin_opt125m`)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
print(json.dumps(bench.plugin_opt125m(torch.device("cuda", 0), with_unpatched=True), indent=1))

"""torch.profiler of one forward of the reference-module OPT-125m stack under the plugin (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "refshim"))
import torch
import load_reference
ref = load_reference.load_full()
from dmx_compressor_b200 import elide, opt, plugin
dev = torch.device("cuda", 0)
dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
mode = sys.argv[2] if len(sys.argv) > 2 else "dropin"
torch.manual_seed(0)
q = opt.OPTStack(None, mods=ref.nn).to(device=dev, dtype=dt).eval()
for m in q.modules():
    if isinstance(m, ref.nn.DmxModule):
        for rule in ref.config_rules.BASIC:
            if isinstance(m, rule.module_types):
                m.configure(rule.module_config); break
ids = torch.randint(0, 50272, (8, 2048), device=dev)
plugin.install("dmx.compressor", elide=(mode == "elided"))
from torch.profiler import profile, ProfilerActivity
import contextlib
ctx = elide.enabled() if mode == "elided" else contextlib.nullcontext()
with torch.no_grad(), ctx:
    elide.materialise(q(ids)); elide.materialise(q(ids))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        elide.materialise(q(ids))
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

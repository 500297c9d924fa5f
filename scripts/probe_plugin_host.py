"""cProfile of the plugin-elided forward of the reference-module OPT-125m stack: where the host time goes (development aid)"""
import os, sys, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import torch
import load_reference
ref = load_reference.load_full()
from dmx_compressor_b200 import elide, opt, plugin
dev = torch.device("cuda", 0)
torch.manual_seed(0)
q = opt.OPTStack(None, mods=ref.nn).to(device=dev, dtype=torch.bfloat16).eval()
for m in q.modules():
    if isinstance(m, ref.nn.DmxModule):
        for rule in ref.config_rules.BASIC:
            if isinstance(m, rule.module_types):
                m.configure(rule.module_config); break
ids = torch.randint(0, 50272, (8, 2048), device=dev)
plugin.install("dmx.compressor", elide=True)
with torch.no_grad(), elide.enabled():
    for _ in range(3):
        elide.materialise(q(ids))
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        elide.materialise(q(ids))
    t1 = time.perf_counter()   # host time to ENQUEUE (no sync)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0) / 5:.2f} ms / forward, wall incl. sync {1e3 * (t2 - t0) / 5:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        elide.materialise(q(ids))
    pr.disable()
    torch.cuda.synchronize()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(32)
print("\n".join(l[:160] for l in st.getvalue().splitlines()[:60]))

"""Where the plugin-elided forward differs from the mirror-elided one (OPT-125m stack, batch 8 x 2048): launches per forward,
kernel time vs wall time, the top kernels of each (development aid).   usage: probe_elide_gap.py [bf16|fp32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import torch
import load_reference
ref = load_reference.load_full()
from dmx_compressor_b200 import _lib, elide, opt, plugin
from dmx_compressor_b200 import nn as dnn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
ids = torch.randint(0, 50272, (8, 2048), device=dev)


def build(mods, rules_owner):
    torch.manual_seed(0)
    q = opt.OPTStack(None, mods=mods).to(device=dev, dtype=dt).eval()
    if mods is dnn:
        return dnn.to_basic_mode(q)
    for m in q.modules():
        if isinstance(m, mods.DmxModule):
            for rule in rules_owner.config_rules.BASIC:
                if isinstance(m, rule.module_types):
                    m.configure(rule.module_config); break
    return q


def measure(name, q):
    with torch.no_grad(), elide.enabled():
        for _ in range(3):
            elide.materialise(q(ids))
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            elide.materialise(q(ids))
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        nl = (_lib.launch_count() - n0) // 3
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            elide.materialise(q(ids))
            torch.cuda.synchronize()
    ka = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.device_time_total for e in ka) / 1e3
    print(f"== {name}: {ms:.2f} ms / forward, {nl} dmxq launches, {sum(e.count for e in ka)} kernels, sum of kernel time {tot:.2f} ms")
    for e in sorted(ka, key=lambda e: -e.device_time_total)[:22]:
        print(f"   {e.device_time_total / 1e3:8.3f} ms  x{e.count:<4d} {e.key[:110]}")


measure("mirror + elide", build(dnn, dnn))
plugin.install("dmx.compressor", elide=True)
measure("reference modules + plugin(elide=True)", build(ref.nn, ref))
plugin.uninstall()

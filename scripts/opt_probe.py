"""OPT-125m BASIC-mode forward probe (development aid; the bench reports the same numbers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import elide, opt, _lib
dev = "cuda:0"
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
for dt in (torch.float32, torch.bfloat16):
    q, p = opt.build_pair(device=dev, dtype=dt)
    ids = torch.randint(0, 50272, (B, S), device=dev)
    def timeit(fn, n=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    with torch.no_grad():
        t_plain = timeit(lambda: p(ids))
        n0 = _lib.launch_count()
        y1 = q(ids); n1 = _lib.launch_count()
        t_basic = timeit(lambda: q(ids))
        with elide.enabled():
            y2 = q(ids)
            n2 = _lib.launch_count()
            y2 = q(ids)
            n3 = _lib.launch_count()
            t_el = timeit(lambda: q(ids))
        same = torch.equal(y1, y2)
    print(f"{dt}: plain {t_plain:.1f} ms ({B*S/t_plain*1e3:.0f} tok/s)  BASIC {t_basic:.1f} ms ({B*S/t_basic*1e3:.0f} tok/s, {n1-n0} dmxq launches, overhead {100*(t_basic-t_plain)/t_plain:.0f}%)  "
          f"BASIC+elide {t_el:.1f} ms ({B*S/t_el*1e3:.0f} tok/s, {n3-n2} launches, overhead {100*(t_el-t_plain)/t_plain:.0f}%)  elided==drop-in: {same}", flush=True)
    del q, p
    torch.cuda.empty_cache()

"""Is the whole BASIC-mode forward CUDA-graph capturable? (kernels run on the current stream, never sync,
never allocate outside torch's allocator) -- development probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import elide, opt, _lib
dev = "cuda:0"
B, S = 4, 128
q, p = opt.build_pair(device=dev, dtype=torch.float32)
ids = torch.randint(0, 50272, (B, S), device=dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    y_ref = q(ids)
    t_eager = timeit(lambda: q(ids))
    t_plain = timeit(lambda: p(ids))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): q(ids)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y_g = q(ids)
    g.replay(); torch.cuda.synchronize()
    print("graph output equals eager:", torch.equal(y_g, y_ref))
    t_graph = timeit(lambda: g.replay())
    gp = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gp):
        y_p = p(ids)
    t_pgraph = timeit(lambda: gp.replay())
print(f"B={B} S={S}: plain eager {t_plain:.2f} ms, plain graph {t_pgraph:.2f} ms, BASIC eager {t_eager:.2f} ms, BASIC graph {t_graph:.2f} ms")

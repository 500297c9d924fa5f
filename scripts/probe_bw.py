"""Quick device-side bandwidth probe of the main kernels (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format

dev = "cuda:0"

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2], ts[0]

def report(name, nbytes, fn):
    med, best = timeit(fn)
    print(f"{name:58s} med {med:8.3f} ms  {nbytes/med/1e6:8.1f} GB/s   best {nbytes/best/1e6:8.1f} GB/s", flush=True)

def main():
    n = 2**28
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(n // 4096, 4096, device=dev).to(dt)
        y = torch.empty_like(x)
        es = x.element_size()
        report(f"torch copy_ {dt}", 2*n*es, lambda: y.copy_(x))
        for sh in ("BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[8|8]{64}(_N)", "BFP[8|8]{64}(SU)", "FP[1|5|10,15](FN)", "FP[1|4|3,7](_N)", "FP[1|5|2,15](_N)", "XP[8,0](CSN)",
                   "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"):
            st = [Format.from_shorthand(sh).stage()]
            report(f"{sh} {dt} flat", 2*n*es, lambda: ops.cast_chain(x, st, -1, out=y))
        report(f"BFP16 pack -> int8+exp {dt}", n*es + n + n//64, lambda: ops.bfp_pack(x, 64, 8))
        _m, _e = ops.bfp_pack(x, 64, 8)
        report(f"BFP16 unpack {dt}", n*es + n + n//64, lambda: ops.bfp_unpack(_m, _e, 64, 8, dtype=dt))
        del _m, _e
        report(f"INT8 CastTo (device qparams) {dt}", 2*n*es, lambda: ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=torch.full((1,), 0.037, device=dev), zero_point=torch.full((1,), 3.0, device=dev), out=y))
        rnd = torch.randint(0, 2**31 - 1, x.shape, device=dev, dtype=torch.int32)
        report(f"BFP16 stochastic (ext. rand) {dt}", (2*es+4)*n, lambda: ops.cast_chain(x, [Format.from_shorthand("BFP[8|8]{64}(SS)").stage()], -1, out=y, rand=rnd))
        del rnd
        report(f"MXFP8[E4M3]{{32}} {dt}", 2*n*es, lambda: ops.cast_chain(x, [Format.from_shorthand("MXFP8[E4M3]{32}").stage()], -1, out=y))
        report(f"2:4 prune {dt}", 2*n*es, lambda: ops.cast_chain(x, [ops.nm_stage(2, 4)], -1, out=y))
        report(f"2:4 -> BFP12 fused {dt}", 2*n*es, lambda: ops.cast_chain(x, [ops.nm_stage(2, 4), Format.from_shorthand('BFP[4|8]{64}(SN)').stage()], -1, out=y))
        report(f"FLOAT16 -> BFP16 fused {dt}", 2*n*es, lambda: ops.cast_chain(x, [Format.from_shorthand('FP[1|5|10,15](FN)').stage(), Format.from_shorthand('BFP[8|8]{64}(SN)').stage()], -1, out=y))
        # non-flat rows (row stride != K)
        xs = torch.randn(n // 4096, 4096 + 64, device=dev).to(dt)[:, :4096]
        ys = torch.empty(n // 4096, 4096 + 64, device=dev, dtype=dt)[:, :4096]
        st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
        report(f"BFP16 {dt} strided rows", 2*n*es, lambda: ops.cast_chain(xs, st, -1, out=ys))
        # per-head view of a fused projection: [B, H, S, 64] view of [B, S, H*64] (two outer dims with odd strides)
        qkv = torch.randn(64, 2048, 12 * 64, device=dev).to(dt)
        qv = qkv.view(64, 2048, 12, 64).transpose(1, 2)
        qy = torch.empty(64, 12, 2048, 64, device=dev, dtype=dt)
        report(f"BFP16 {dt} per-head view [64,12,2048,64] of [64,2048,768]", 2*qv.numel()*es, lambda: ops.cast_chain(qv, st, -1, out=qy))
        del qkv, qv, qy
        # cols: [96, 2048, 64*...] along dim -2
        v = torch.randn(96*8, 2048, 64, device=dev).to(dt)
        vy = torch.empty_like(v)
        report(f"BFP16 {dt} cols [768,2048,64] d=-2", 2*v.numel()*es, lambda: ops.cast_chain(v, st, -2, out=vy))
        v2 = torch.randn(96, 2048, 2048, device=dev).to(dt)
        vy2 = torch.empty_like(v2)
        report(f"BFP16 {dt} cols [96,2048,2048] d=-2", 2*v2.numel()*es, lambda: ops.cast_chain(v2, st, -2, out=vy2))
        del v, vy, v2, vy2, xs, ys
    for e in (20, 22, 24, 26, 30):
        n2 = 2**e
        x = torch.randn(n2 // 4096, 4096, device=dev)
        y = torch.empty_like(x)
        st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
        report(f"BFP16 fp32 n=2^{e}", 8*n2, lambda: ops.cast_chain(x, st, -1, out=y))


if __name__ == "__main__":
    main()

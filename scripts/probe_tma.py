"""The TMA experiment (csrc/dmxq_tma.cu): BFP16_64 / BFP12_64 along a strided dim of [outer, K, inner] -- the production kernels
(bfp_cols16_kernel for 16-bit tensors, chain_cols_kernel for fp32) vs the TMA-tiled kernel; bit-equality first, then timing."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import _lib as L, ops
from dmx_compressor_b200.numerical import Format
fn = L.lib.dmxq_x_bfp_cols_tma
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
fn.restype = C.c_int
dev = "cuda:0"


def tma(x, wl, out, cfg=0):
    o, k, i = x.shape
    rc = fn(x.data_ptr(), out.data_ptr(), L.dtype_code(x.dtype), o, k, i, 64, wl, cfg, L.stream_ptr(x.device))
    assert rc == 0, rc
    return out


def t(f, reps=20):
    for _ in range(3):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dt in (torch.bfloat16, torch.float32):
    for shape in ((96, 2048, 64), (96, 2048, 256), (768, 2048, 64), (7, 200, 72), (64, 4096, 1024)):
        g = torch.Generator(device=dev).manual_seed(1)
        x = (torch.randn(shape, device=dev, generator=g) * torch.pow(2.0, torch.randint(-6, 7, (shape[0], 1, shape[2]), device=dev, generator=g).float())).to(dt)
        x[0, :, 1] = 0.0
        if shape[0] > 1:
            x[1, :, 2] *= 2.0 ** (-100 if dt != torch.float16 else -10)
        for wl, sh in ((8, "BFP[8|8]{64}(SN)"), (4, "BFP[4|8]{64}(SN)")):
            st = [Format.from_shorthand(sh).stage()]
            y0 = ops.cast_chain(x, st, -2)
            y1 = tma(x, wl, torch.empty_like(x))
            torch.cuda.synchronize()
            same = torch.equal(y0.view(torch.int32 if dt == torch.float32 else torch.int16), y1.view(torch.int32 if dt == torch.float32 else torch.int16))
            if x.numel() < 1 << 20:
                print(f"{str(dt):15s} {str(shape):18s} {sh:18s} bit-equal {same}")
                continue
            out0, out1 = torch.empty_like(x), torch.empty_like(x)
            t0 = t(lambda: ops.cast_chain(x, st, -2, out=out0))
            nb = 2 * x.numel() * x.element_size()
            res = []
            for cfg in range(4):
                ok = torch.equal(tma(x, wl, torch.empty_like(x), cfg), y1)
                res.append(f"cfg{cfg} {nb / t(lambda: tma(x, wl, out1, cfg)) / 1e6:5.0f}{'' if ok else ' MISMATCH'}")
            print(f"{str(dt):15s} {str(shape):18s} {sh:18s} bit-equal {same}   production {nb / t0 / 1e6:6.0f} GB/s   TMA " + "  ".join(res), flush=True)

"""ncu driver: production cols kernels vs the TMA-tiled experiment on [768,2048,64] (development aid)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import _lib as L, ops
from dmx_compressor_b200.numerical import Format
fn = L.lib.dmxq_x_bfp_cols_tma
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
fn.restype = C.c_int
st = [Format.from_shorthand("BFP[8|8]{64}(SN)").stage()]
for dt, cfg in ((torch.float32, 0), (torch.bfloat16, 2)):
    x = torch.randn(768, 2048, 64, device="cuda:0").to(dt)
    y = torch.empty_like(x)
    ops.cast_chain(x, st, -2, out=y)
    fn(x.data_ptr(), y.data_ptr(), L.dtype_code(dt), 768, 2048, 64, 64, 8, cfg, L.stream_ptr(x.device))
    torch.cuda.synchronize()

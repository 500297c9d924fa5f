"""ncu driver #2: general-path kernels (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmx_compressor_b200 import ops
from dmx_compressor_b200.numerical import Format
dev = "cuda:0"
n = 2**26
F = lambda sh: Format.from_shorthand(sh).stage()
x = torch.randn(n // 4096, 4096, device=dev); y = torch.empty_like(x)
xs = torch.randn(n // 4096, 4096 + 64, device=dev)[:, :4096]
ys = torch.empty(n // 4096, 4096 + 64, device=dev)[:, :4096]
ops.cast_chain(xs, [F("BFP[8|8]{64}(SN)")], -1, out=ys)                 # 0 rows nonflat special2
ops.cast_chain(x, [F("XP[8,0](CSN)")], -1, out=y)                       # 1 fixed
ops.cast_chain(x, [F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")], -1, out=y)   # 2 sbfp
ops.cast_chain(x, [F("FP[1|5|10,15](FN)"), F("BFP[8|8]{64}(SN)")], -1, out=y)  # 3 fused pair
ops.cast_chain(x, [ops.nm_stage(2, 4), F("BFP[4|8]{64}(SN)")], -1, out=y)      # 4 prune+bfp
xb = x.bfloat16(); yb = torch.empty_like(xb)
ops.cast_chain(xb, [F("FP[1|5|10,15](FN)")], -1, out=yb)                # 5 float16 on bf16
v = torch.randn(96, 2048, 256, device=dev).bfloat16(); vy = torch.empty_like(v)
ops.cast_chain(v, [F("BFP[8|8]{64}(SN)")], -2, out=vy)                  # 6 cols bf16
ops.cast_chain(xb, [ops.nm_stage(2, 4)], -1, out=yb)                    # 7 2:4 on bf16 (raw packed path)
ops.cast_chain(xb, [ops.nm_stage(2, 4), F("BFP[4|8]{64}(SN)")], -1, out=yb)   # 8 2:4 -> BFP12 bf16
ops.cast_chain(xb, [F("FP[1|5|10,15](FN)"), F("BFP[8|8]{64}(SN)")], -1, out=yb)  # 9 FLOAT16 -> BFP16 bf16
ops.cast_chain(xb, [F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")], -1, out=yb)   # 10 sbfp bf16
ops.cast_chain(x, [F("MXFP8[E4M3]{32}")], -1, out=y)                    # 11 mxfp fp32
ops.cast_chain(xb, [F("MXFP8[E4M3]{32}")], -1, out=yb)                  # 12 mxfp bf16
ops.histc(x, 2048, min=-7, max=9, return_minmax=True)                   # 13 histc (+ init / final helper kernels)
ops.minmax(xb)                                                          # 14 minmax bf16
ops.cast_chain(xb, [F("BFP[8|8]{64}(_N)")], -1, out=yb)                 # 15 asymmetric BFP16A bf16 (K_BFP_ASYM)
rnd = torch.randint(0, 2**31 - 1, xb.shape, device=dev, dtype=torch.int32)
ops.cast_chain(xb, [F("BFP[8|8]{64}(SS)")], -1, out=yb, rand=rnd)       # 16 stochastic BFP16 bf16 (K_BFP_STOCH)
ops.cast_chain(xb, [F("FP[1|4|3,7](_N)")], -1, out=yb)                  # 17 FP8 E4M3 on bf16 (K_FLOAT, subnormals kept)
_m, _e = ops.bfp_pack(x, 64, 8)                                         # 18 pack fp32 -> int8 + exponent bytes
ops.bfp_unpack(_m, _e, 64, 8, dtype=torch.float32)                      # 19 unpack
sc = torch.rand(4096, device=dev) * 0.05 + 0.01
ops.fixed_qdq(x, 8, 0, True, True, "nearest", scale=sc, zero_point=torch.zeros(4096, device=dev), ch_axis=1, out=y)   # 20 INT8 per column
ops.minmax(x, 1)                                                        # 21 per-column amin / amax
_st = Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()
_pm, _ps = ops.sbfp_pack(xb, _st)                                       # 22 packed SBFP12_16 storage of a bf16 tensor
ops.sbfp_unpack(_pm, _ps, _st, dtype=torch.bfloat16)                    # 23 unpack
scr = torch.rand(n // 4096, device=dev) * 0.05 + 0.01
ops.fixed_qdq(xb, 8, 0, True, True, "nearest", scale=scr, zero_point=torch.zeros(n // 4096, device=dev), ch_axis=0, out=yb)  # 24 INT8 per row, bf16
torch.cuda.synchronize()

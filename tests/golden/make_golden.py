"""Generate tests/golden/*.npz by running the reference's OWN python (unmodified, imported
from /root/reference via oracle/refshim) on seeded inputs.

Run here (CPU container):   python tests/golden/make_golden.py
The reference cannot travel to the GPU box; these small fixtures do.  Each case stores the
input and the reference output as raw bit patterns (uint32 for fp32, uint16 for bf16) so
that parity is checked bit-for-bit (NaNs compare by class, see tests/util.py).

Reference entry points exercised (all through the reference's public classes):
  dmx.compressor.numerical.CastTo(format, block_dim=...)           S/numerical/cast.py:136-306
  dmx.compressor.numerical.Format.from_shorthand(...).cast          S/numerical/format.py
  dmx.compressor.sparse.Sparsify(shape, "BTOPK{K:M,d}(U)")          S/sparse.py:245-301
Stochastic BFP/FP on the reference CPU path draw from an unseeded std::mt19937
(Q/quant_cpu/quant_cpu.cpp:32-34) and are not reproducible: no golden exists for them;
their parity is "same random tensor in => same bits out" against oracle/dmxq_oracle.c.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import load_reference  # noqa: E402

num, sp, _ = load_reference.load()


def u32(t):
    return t.detach().contiguous().float().numpy().view(np.uint32).copy()


def specials():
    v = [0.0, -0.0, 1.0, -1.0, 1.0 + 2**-7, 1.0 + 2**-6, 1 + 2**-6 + 2**-7, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5,
         2**-7 + 2**-23, 1e-40, -1e-40, 1e-45, 6.0e-5, 6.2e-5, 2**-14, 2**-15, 2**-24, 65504.0, 65520.0, 1e5,
         131008.0, 2e5, 448.0, 480.0, 500.0, 57344.0, 1e38, -1e38, 3e38, 1.9999999, -1.9999999, 1.99, -1.99,
         -1.9921875, 127.5, -127.5, 126.5, 7.5, -7.5, 6.5, 0.49999997, 8388607.5, 1.17549435e-38, 2.0**-126,
         2.0**-127, float("inf"), float("-inf"), float("nan")]
    return torch.tensor(v, dtype=torch.float32)


def rand_rows(gen, rows, cols, spread=8):
    x = torch.randn(rows, cols, generator=gen)
    k = torch.randint(-spread, spread + 1, (rows, 1), generator=gen)
    return x * torch.pow(2.0, k.float())


def main():
    g = torch.Generator().manual_seed(20261017)
    cases = {}
    meta = {}

    def add(name, x, y, **kw):
        cases[name + ".x"] = x
        cases[name + ".y"] = y
        meta[name] = kw

    # ---------------------------------------------------------------- BFP (blocked)
    bfp_formats = ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[6|8]{32}(SN)", "BFP[8|8]{16}(SN)",
                   "BFP[16|8]{64}(SN)", "BFP[4|8]{128}(SN)", "BFP[8|8]{64}(SU)", "BFP[8|8]{64}(SD)",
                   "BFP[8|8]{64}(_N)", "BFP[4|8]{32}(_N)", "BFP[6|8]{16}(_N)", "BFP[8|8]{128}(_N)", "MXINT8{32}"]
    shapes = [((24, 256), -1), ((5, 100), -1), ((3, 70, 6), 1), ((2, 128, 8), -2), ((4, 6, 5, 5), 1), ((7, 1, 3), 1)]
    ci = 0
    for fmt in bfp_formats:
        for shp, bd in shapes:
            x = rand_rows(g, int(np.prod(shp[:-1])), shp[-1]).reshape(shp)
            if ci % 3 == 0:  # sprinkle exact ties / zeros / tiny values
                flat = x.view(-1)
                idx = torch.randint(0, flat.numel(), (flat.numel() // 8,), generator=g)
                flat[idx] = torch.round(flat[idx] * 64) / 64
                flat[::17] = 0.0
                flat[5::29] *= 2.0**-30
            y = num.CastTo(fmt, block_dim=bd)(x)
            add(f"bfp/{ci}", u32(x), u32(y), kind="cast", fmt=fmt, block_dim=bd, shape=list(shp))
            ci += 1
    # adversarial block contents
    sv = specials()
    fin = sv[torch.isfinite(sv)]
    adv = []
    for base in (1.0, 1.999, 3e-39, 2.0**100, 2.0**126, 1e38):
        blk = torch.zeros(64)
        blk[0] = base
        blk[1:1 + min(63, fin.numel())] = (fin[:63] * base / 4)[:63]
        adv.append(blk)
    adv.append(torch.zeros(64))
    adv.append(torch.full((64,), 1e-40))
    b = torch.randn(64, generator=g); b[3] = float("inf"); adv.append(b)
    b = torch.randn(64, generator=g); b[9] = float("nan"); adv.append(b)
    b = torch.randn(64, generator=g); b[0] = -1.9999999; b[1] = 1.99; adv.append(b)   # asymmetric edge
    b = torch.randn(64, generator=g) * 0.1; b[0] = -1.9921875; b[1] = -1.999; b[2] = -1.99; b[3] = 1.999; adv.append(b)
    adv = torch.stack(adv)
    for fmt in ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[8|8]{64}(_N)", "BFP[8|8]{16}(SN)"]:
        y = num.CastTo(fmt, block_dim=-1)(adv)
        add(f"bfp/{ci}", u32(adv), u32(y), kind="cast", fmt=fmt, block_dim=-1, shape=list(adv.shape))
        ci += 1

    # ---------------------------------------------------------------- FP / BFP block 1 (elementwise)
    wide = torch.randn(4096, generator=g) * torch.pow(2.0, torch.randint(-30, 20, (4096,), generator=g).float())
    ew = torch.cat([sv, wide, torch.randn(1024, generator=g), torch.randn(512, generator=g) * 1e-5])
    for i, fmt in enumerate(["FP[1|5|10,15](FN)", "FP[1|5|10,15](_N)", "FP[1|8|7,127](FN)", "FP[1|4|3,7](_N)",
                             "FP[1|5|2,15](_N)", "FP[0|4|4,7](FN)", "FP[1|4|3,7](FN)", "FP[1|8|23,127](_N)",
                             "FP[0|4|4,10](FN)", "FP[1|2|1,1](_N)", "FP[1|3|2,3](_N)", "FP[1|8|22,127](_N)",
                             "BFP[24|8]{1}(SN)", "BFP[8|8]{1}(SN)", "BFP[4|8]{1}(SN)"]):
        y = num.CastTo(fmt)(ew)
        add(f"fp/{i}", u32(ew), u32(y), kind="cast", fmt=fmt, block_dim=-1, shape=list(ew.shape))

    # ---------------------------------------------------------------- XP (reference CPU => ties-to-even quirk)
    xs = torch.cat([sv[torch.isfinite(sv) & (sv.abs() < 1e6)], torch.randn(2048, generator=g) * 40,
                    torch.round(torch.randn(512, generator=g) * 20) + 0.5, torch.randn(512, generator=g)])
    for i, fmt in enumerate(["XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,0](C_N)", "XP[8,0](_SN)", "XP[8,+4](CSN)",
                             "XP[8,-2](CSN)", "XP[16,+8](CSN)", "XP[8,0](CSU)", "XP[8,0](CSD)", "XP[24,+10](CSN)"]):
        y = num.CastTo(fmt)(xs)
        add(f"xp/{i}", u32(xs), u32(y), kind="cast", fmt=fmt, block_dim=-1, shape=list(xs.shape), tie="even")
    # affine: per-tensor, per-channel, group (cast.py:228-237, 279-296)
    xa = torch.randn(6, 16, 10, generator=g) * 3
    c = num.CastTo("XP[8,0](CSN)")
    c.scale.copy_(torch.tensor([0.1])); c.zero_point.copy_(torch.tensor([3.0]))
    add("xpa/0", u32(xa), u32(c(xa)), kind="xp_affine", fmt="XP[8,0](CSN)", shape=list(xa.shape),
        scale=[0.1], zero_point=[3.0], ch_axis=None, group_size=None, tie="even")
    sc = (torch.rand(16, generator=g) * 0.2 + 0.01)
    zp = torch.round(torch.randn(16, generator=g) * 3)
    c = num.CastTo("XP[8,0](CSN)", qscheme=torch.per_channel_affine, ch_axis=1)
    c.scale = sc.clone(); c.zero_point = zp.clone()
    add("xpa/1", u32(xa), u32(c(xa)), kind="xp_affine", fmt="XP[8,0](CSN)", shape=list(xa.shape),
        scale=sc.tolist(), zero_point=zp.tolist(), ch_axis=1, group_size=None, tie="even")
    c = num.CastTo("XP[4,0](CSN)", group_size=4, ch_axis=1)
    c.scale = sc[:4].clone() * 5; c.zero_point = torch.zeros(4)
    add("xpa/2", u32(xa), u32(c(xa)), kind="xp_affine", fmt="XP[4,0](CSN)", shape=list(xa.shape),
        scale=(sc[:4] * 5).tolist(), zero_point=[0.0] * 4, ch_axis=1, group_size=4, tie="even")

    # ---------------------------------------------------------------- SBFP
    si = 0
    for fmt in ["SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[4,0](CSN)><FP[0|4|4,10](FN)>{16}",
                "SBFP<XP[8,0](CSN)><FP[0|4|4,7](FN)>{64}", "SBFP<XP[4,0](CSN)><FP[1|5|10,15](FN)>{16}",
                "SBFP<XP[4,0](CSN)><FP[0|8|7,127](_N)>{32}"]:
        for shp, bd in [((24, 256), -1), ((5, 100), -1), ((3, 40, 6), 1), ((2, 128, 8), -2)]:
            x = rand_rows(g, int(np.prod(shp[:-1])), shp[-1], spread=4).reshape(shp)
            x.view(-1)[::13] = 0.0
            if si % 2 == 0:
                x.view(-1)[: shp[-1]] = 0.0  # an all-zero block row -> passthrough branch
            y = num.CastTo(fmt, block_dim=bd)(x)
            add(f"sbfp/{si}", u32(x), u32(y), kind="cast", fmt=fmt, block_dim=bd, shape=list(shp), tie="even")
            si += 1

    # ---------------------------------------------------------------- N:M Sparsify
    ni = 0
    for sh, shp in [("BTOPK{2:4,-1}(U)", (16, 64)), ("BTOPK{4:8,-1}(U)", (16, 64)), ("BTOPK{2:8,-1}(U)", (8, 32)),
                    ("BTOPK{2:4,0}(U)", (8, 12)), ("BTOPK{4:8,1}(U)", (3, 16, 5)), ("BTOPK{1:4,-1}(U)", (4, 16)),
                    ("BTOPK{3:4,-1}(U)", (4, 16)), ("BTOPK{2:4,-1}(U)", (6, 3, 8))]:
        for variant in ("abs", "param", "ties"):
            x = torch.randn(shp, generator=g)
            if variant == "ties":  # bf16-like coarse values: many equal |x|
                x = torch.round(x * 2) / 2
            s = sp.Sparsify(shp, sh)
            s.eval()
            if variant == "param":
                with torch.no_grad():
                    s.score.copy_(torch.rand(shp, generator=g))
                score = s.score.detach().clone()
            else:
                s.configure(score_func=lambda sc_, x_: x_.abs())
                score = x.abs()
            y = s(x)
            cases[f"nm/{ni}.score"] = u32(score)
            cases[f"nm/{ni}.mask"] = u32(s.mask)
            add(f"nm/{ni}", u32(x), u32(y), kind="nm", sparseness=sh, variant=variant, shape=list(shp))
            ni += 1

    # ---------------------------------------------------------------- CastTo dtype round trip (bf16 / fp16 in -> same out)
    di = 0
    for fmt, bd in [("BFP[8|8]{64}(SN)", -1), ("BFP[4|8]{64}(SN)", -1), ("FP[1|5|10,15](FN)", -1),
                    ("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", -1), ("BFP[8|8]{64}(SN)", 0), ("FP[1|4|3,7](_N)", -1)]:
        for dt in (torch.bfloat16, torch.float16):
            x = rand_rows(g, 16, 192, spread=4).to(dt)
            y = num.CastTo(fmt, block_dim=bd)(x)
            assert y.dtype == dt
            cases[f"dt/{di}.x"] = x.view(torch.int16).numpy().view(np.uint16).copy()
            cases[f"dt/{di}.y"] = y.view(torch.int16).numpy().view(np.uint16).copy()
            meta[f"dt/{di}"] = dict(kind="cast_dtype", fmt=fmt, block_dim=bd, shape=[16, 192],
                                    dtype=str(dt).split(".")[-1], tie="even")
            di += 1

    # ---------------------------------------------------------------- weight hypernet composition (core.py:184-196)
    w = torch.randn(32, 128, generator=g) * 0.05
    s = sp.Sparsify(w.shape, "BTOPK{2:4,-1}(U)"); s.eval(); s.configure(score_func=lambda sc_, x_: x_.abs())
    y = num.CastTo("BFP[4|8]{64}(SN)")(s(w))
    add("hyper/0", u32(w), u32(y), kind="hyper", sparseness="BTOPK{2:4,-1}(U)", storage="SAME", fmt="BFP[4|8]{64}(SN)",
        shape=list(w.shape))
    s = sp.Sparsify(w.shape, "BTOPK{4:8,-1}(U)"); s.eval(); s.configure(score_func=lambda sc_, x_: x_.abs())
    y = num.CastTo("BFP[8|8]{64}(SN)")(num.CastTo("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")(s(w)))
    add("hyper/1", u32(w), u32(y), kind="hyper", sparseness="BTOPK{4:8,-1}(U)",
        storage="SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", fmt="BFP[8|8]{64}(SN)", shape=list(w.shape), tie="even")

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **cases)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, reference="dmx-compressor v0.1.11 (/root/reference)",
                       generator="tests/golden/make_golden.py", cases=meta), f, indent=1, sort_keys=True)
    print(f"wrote {len(meta)} cases, {sum(v.nbytes for v in cases.values()) / 1e6:.2f} MB raw")


if __name__ == "__main__":
    main()

"""GPU parity tests: the CUDA path (through the C ABI, libdmxq.so) against
  (a) the committed golden vectors produced by the reference's own python,
  (b) the CPU oracle (oracle/dmxq_oracle.c) on fresh seeded inputs over many layouts,
  (c) the reference's own CUDA kernels (oracle/_ref/ref_quant_cuda.so) when that binary travelled,
  (d) size-independent properties at BASELINE sizes (idempotence, block-scale equivariance).
Bit-exact everywhere (NaNs compare by class).  Run with:  pytest -m gpu
"""
import os
import sys

import numpy as np
import pytest
import torch

from util import ROOT, assert_bits_equal, bits, case, f32, golden_cases

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import dmx_compressor_b200 as dmx
    from dmx_compressor_b200 import _lib as L
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format, FixedPoint, ScaledBlockFloatingPoint

DEV = "cuda:0"
TIE = {"even": 1, "away": 0}


def fmt_from(sh, tie="away"):
    f = Format.from_shorthand(sh)
    if isinstance(f, FixedPoint):
        f.tie = TIE[tie]
    if isinstance(f, ScaledBlockFloatingPoint):
        f.block_format.tie = TIE[tie]
    return f


def gpu_cast(x, sh, block_dim=-1, tie="away"):
    """x: torch CUDA tensor (any dtype/strides) -> fp32 result of Format.cast, as numpy bits."""
    y = fmt_from(sh, tie).cast(x, block_dim)
    torch.cuda.synchronize()
    return y


def special_block_masks(x, block_dim, bs):
    """Blocks where CPU and GPU *hardware* legitimately differ for the reference's own code:
      huge:      block max has exponent field >= 253: the reference computes inf - inf and then
                 clips the NaN with *its sign*; the default NaN is negative on x86, positive on
                 the GPU.  Only the magnitude is comparable across CPU and GPU.
      nonfinite: block holds an Inf / NaN: every element becomes NaN on x86; on the GPU the
                 canonical NaN (0x7fffffff) overflows in round_bitwise and x = -Inf / NaN inputs
                 come out as +-Inf.  Only 'non-finite' is comparable.
    test_vs_reference_cuda_adversarial pins both cases bit-for-bit against the reference's own
    CUDA kernels instead."""
    a = np.abs(np.moveaxis(x, block_dim, -1)).astype(np.float32)
    K = a.shape[-1]
    huge = np.zeros_like(a, dtype=bool)
    nonfinite = np.zeros_like(a, dtype=bool)
    for k0 in range(0, K, bs):
        blk = a[..., k0:k0 + bs]
        nf = ~np.isfinite(blk).all(-1, keepdims=True)
        with np.errstate(invalid="ignore"):
            big = np.nan_to_num(blk, nan=0.0, posinf=0.0).max(-1, keepdims=True) >= 2.0**126
        nonfinite[..., k0:k0 + bs] = nf
        huge[..., k0:k0 + bs] = big & ~nf
    mv = lambda m: np.moveaxis(m, -1, block_dim % x.ndim)
    return mv(huge), mv(nonfinite)


def check(got_t, want_bits, what, x=None, fmt=None, block_dim=-1, dtype="float32"):
    got = got_t.detach().contiguous().cpu()
    if dtype == "float32":
        g = got.numpy().view(np.uint32)
    else:
        g = got.view(torch.int16).numpy().view(np.uint16)
    w = np.asarray(want_bits).reshape(g.shape)
    if x is not None and fmt is not None and fmt.startswith("BFP") and "{1}" not in fmt:
        bs = int(fmt.split("{")[1].split("}")[0])
        huge, nonfinite = special_block_masks(x, block_dim, bs)
        if huge.any():
            g = np.where(huge, g & 0x7FFFFFFF, g)
            w = np.where(huge, w & 0x7FFFFFFF, w)
        if nonfinite.any():
            if "(S" in fmt:  # (the asymmetric post-pass turns NaN blocks into int-cast garbage on the CPU)
                assert ((g[nonfinite] & 0x7F800000) == 0x7F800000).all(), f"{what}: finite value in a non-finite block"
                assert ((w[nonfinite] & 0x7F800000) == 0x7F800000).all()
            g = np.where(nonfinite, 0, g)
            w = np.where(nonfinite, 0, w)
    assert_bits_equal(g, w, what, dtype=dtype)


# =============================================================================== (a) golden vectors
@pytest.mark.parametrize("name", golden_cases(kind="cast"))
def test_golden_cast(name):
    m, d = case(name)
    x = f32(d["x"]).reshape(m["shape"])
    # the goldens were produced through the reference's CastTo module (incl. its affine wrap for
    # FixedPoint, cast.py:279-296), so they are replayed through our CastTo module
    from dmx_compressor_b200.numerical import CastTo

    c = CastTo(m["fmt"], block_dim=m["block_dim"]).to(DEV)
    c.format = fmt_from(m["fmt"], m.get("tie", "even"))
    y = c(torch.from_numpy(x).to(DEV))
    check(y, d["y"], f"{name} {m['fmt']}", x=x, fmt=m["fmt"], block_dim=m["block_dim"])
    if not m["fmt"].startswith("XP"):  # Format.cast itself (no module) gives the same bits
        y = gpu_cast(torch.from_numpy(x).to(DEV), m["fmt"], m["block_dim"], m.get("tie", "even"))
        check(y, d["y"], f"{name} {m['fmt']} Format.cast", x=x, fmt=m["fmt"], block_dim=m["block_dim"])


@pytest.mark.parametrize("name", golden_cases(kind="xp_affine"))
def test_golden_xp_affine(name):
    m, d = case(name)
    x = torch.from_numpy(f32(d["x"]).reshape(m["shape"])).to(DEV)
    f = fmt_from(m["fmt"], "even")
    sc = torch.tensor(m["scale"], dtype=torch.float32, device=DEV)
    zp = torch.tensor(m["zero_point"], dtype=torch.float32, device=DEV)
    y = ops.fixed_qdq(x, f.precision, f.fraction, f.clamp, f.symmetric, f.rounding, f.tie, sc, zp,
                      ch_axis=m["ch_axis"] if m["ch_axis"] is not None else -1, group_size=m["group_size"])
    check(y, d["y"], name)


@pytest.mark.parametrize("name", golden_cases(kind="nm"))
def test_golden_nm(name):
    m, d = case(name)
    x = torch.from_numpy(f32(d["x"]).reshape(m["shape"])).to(DEV)
    score = torch.from_numpy(f32(d["score"]).reshape(m["shape"])).to(DEV)
    k, rest = m["sparseness"][len("BTOPK{"):].split(":")
    mm, dim = rest.split("}")[0].split(",")
    y, mask = ops.nm_prune(x, int(k), int(mm), int(dim), score=score, return_mask=True)
    check(mask, d["mask"], name + " mask")
    check(y, d["y"], name)
    if m["variant"] != "param":
        check(ops.nm_prune(x, int(k), int(mm), int(dim)), d["y"], name + " implicit |x| score")


@pytest.mark.parametrize("name", golden_cases(kind="cast_dtype"))
def test_golden_cast_dtype(name):
    """CastTo.forward on bf16 / fp16 tensors: x.float() -> cast -> .to(dtype), fused in-kernel."""
    m, d = case(name)
    dt = getattr(torch, m["dtype"])
    x = torch.from_numpy(d["x"].view(np.int16).reshape(m["shape"])).view(dt).to(DEV)
    st = fmt_from(m["fmt"], "even").stage()
    y = ops.cast_chain(x, [st], m["block_dim"])
    assert y.dtype == dt
    check(y, d["y"], name, dtype=m["dtype"])


@pytest.mark.parametrize("name", golden_cases(kind="hyper"))
def test_golden_hypernet(name):
    """sparsify -> storage cast -> weight cast (DmxModule.weight_hypernet) as ONE fused kernel."""
    m, d = case(name)
    x = torch.from_numpy(f32(d["x"]).reshape(m["shape"])).to(DEV)
    k, rest = m["sparseness"][len("BTOPK{"):].split(":")
    mm, dim = rest.split("}")[0].split(",")
    stages = [ops.nm_stage(int(k), int(mm))]
    for sh in (m["storage"], m["fmt"]):
        st = fmt_from(sh, m.get("tie", "even")).stage()
        if st is not None:
            stages.append(st)
    y = ops.cast_chain(x, stages, -1)
    check(y, d["y"], name)


# =============================================================================== (b) oracle, many layouts
def _rand(shape, seed, spread=8, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    e = torch.randint(-spread, spread + 1, shape[:-1] + (1,), generator=g).float()
    x = x * torch.pow(2.0, e)
    flat = x.view(-1)
    flat[::37] = 0.0
    flat[5::101] = torch.round(flat[5::101] * 32) / 32
    return x.to(dtype)


LAYOUTS = [
    # (shape, block_dim, description)
    ((64, 4096), -1, "flat rows"),
    ((8, 33, 256), -1, "3-d contiguous"),
    ((16, 100), -1, "ragged K, vectorisable"),
    ((16, 70), -1, "ragged K, K%4!=0 -> generic"),
    ((5, 6), -1, "tiny"),
    ((12, 128, 64), -2, "cols: block along strided dim, inner 64"),
    ((3, 200, 32), 1, "cols ragged K"),
    ((4, 6, 5, 5), 1, "conv weight: cols generic (inner 25)"),
    ((2, 4, 96, 64), -1, "4-d"),
    ((1, 1, 64), -1, "leading ones"),
    ((257, 64), -1, "odd rows"),
    ((3, 1), -1, "K == 1"),
]
FORMATS = ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[8|8]{16}(SN)", "BFP[6|8]{32}(SN)", "BFP[8|8]{128}(SN)",
           "BFP[8|8]{64}(_N)", "BFP[8|8]{64}(SU)", "BFP[8|8]{64}(SD)", "BFP[16|8]{64}(SN)", "BFP[8|8]{256}(SN)",
           "BFP[8|8]{24}(SN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[8,0](CSN)><FP[0|4|4,7](FN)>{64}",
           "MXFP8[E4M3]{32}", "MXFP8[E5M2]{64}", "MXFP4[E2M1]{32}", "MXINT8{32}"]


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("shape,bd,desc", LAYOUTS)
def test_oracle_blocked_layouts(fmt, shape, bd, desc):
    x = _rand(shape, hash((fmt, shape)) % 10000)
    want = O.cast(x.numpy(), fmt, bd, tie=O.TIE_AWAY)
    y = gpu_cast(x.to(DEV), fmt, bd, "away")
    check(y, bits(want), f"{fmt} {desc}")


@pytest.mark.parametrize("seed", range(int(os.environ.get("DMXQ_FUZZ_SEEDS", "12"))))
def test_fuzz_random_views_formats_dtypes(seed):
    """seeded differential fuzzing: random rank / extents / permutation / slicing (so rows, cols and generic kernels all
    get odd strides, offsets and ragged tails), random format, block dim and source dtype, against the oracle"""
    rng = np.random.default_rng(1000 + seed)
    fmts = FORMATS + ELEMENTWISE
    for _ in range(40):
        rank = int(rng.integers(1, 5))
        shape = [int(rng.choice([1, 2, 3, 5, 8, 16, 24, 33, 64, 96, 130, 256])) for _ in range(rank)]
        while int(np.prod(shape)) > 1 << 18:
            shape[int(np.argmax(shape))] //= 2
        base = _rand(tuple(shape), int(rng.integers(1 << 30)), spread=int(rng.integers(1, 10)))
        view = base
        if rank >= 2 and rng.random() < 0.5:
            view = view.permute(*[int(i) for i in rng.permutation(rank)])
        if rng.random() < 0.5:  # slice one dim (offset + shorter extent)
            d = int(rng.integers(rank))
            n = view.shape[d]
            if n > 2:
                lo = int(rng.integers(0, n // 2))
                view = view.narrow(d, lo, int(rng.integers(1, n - lo + 1)))
        if rng.random() < 0.2 and view.shape[-1] > 3:
            view = view[..., ::2]
        fmt = fmts[int(rng.integers(len(fmts)))]
        bd = int(rng.integers(-view.dim(), view.dim()))
        if fmt.startswith("MXFP") and bd not in (-1, view.dim() - 1):
            pass  # (the oracle follows the reference's definition on any dim; the reference itself only runs dim -1)
        dt = [torch.float32, torch.bfloat16, torch.float16][int(rng.integers(3))]
        xv = view.to(dt) if dt == torch.float32 else view.to(dt)
        if dt != torch.float32:
            xv = torch.where(torch.isfinite(xv), xv, torch.zeros_like(xv))
            # (.to() of a strided view returns a dense tensor: rebuild the same striding on the 16-bit copy)
            dense = base.to(dt)
            dense = torch.where(torch.isfinite(dense), dense, torch.zeros_like(dense))
            xv = dense.as_strided(view.shape, view.stride(), view.storage_offset())
        want = O.cast(xv.float().contiguous().numpy(), fmt, bd, tie=O.TIE_AWAY)
        if dt == torch.float16 and fmt == "FP[1|5|10,15](_N)":
            want = xv.float().contiguous().numpy()  # native format of the tensor: handed back unchanged (reference format.py:209-212)
        if fmt.startswith("XP"):
            # the oracle's XP entry is CastTo.forward, i.e. with the (unit) affine wrap around the quantiser, which turns
            # a -0 input into +0 before rounding: run the same entry on the device
            f = fmt_from(fmt, "away")
            got = ops.fixed_qdq(_same_layout(xv), f.precision, f.fraction, f.clamp, f.symmetric, f.rounding, f.tie,
                                scale=torch.ones(1, device=DEV), zero_point=torch.zeros(1, device=DEV), out_dtype=torch.float32)
        else:
            got = gpu_cast(_same_layout(xv), fmt, bd, "away")
        what = f"seed {seed}: {fmt} bd={bd} shape={tuple(view.shape)} stride={view.stride()} {dt}"
        blocked = (fmt.startswith("BFP") and "{1}" not in fmt)
        # (.float(): FloatingPoint.cast hands a native fp16 / fp32 tensor back unchanged, reference format.py:209-212)
        check(got.float(), bits(want), what, x=xv.float().contiguous().numpy() if blocked else None, fmt=fmt if blocked else None, block_dim=bd)


@pytest.mark.parametrize("fmt", ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(_N)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"])
def test_oracle_strided_views(fmt):
    """Views must be consumed in place: transposes, slices, attention-head views."""
    dev = DEV
    # k.transpose(-2,-1) blocked along -2: physically contiguous blocks (rows path, no copy)
    k = _rand((6, 128, 64), 1).to(dev)
    kt = k.transpose(-2, -1)
    want = np.swapaxes(O.cast(k.cpu().numpy(), fmt, -1, tie=O.TIE_AWAY), -2, -1)
    check(gpu_cast(kt, fmt, -2), bits(np.ascontiguousarray(want)), "k^T view")
    # HF attention layout: [B,S,H*D] viewed as [B,H,S,D]
    B, S, H, D = 2, 96, 4, 64
    base = _rand((B, S, H * D), 2).to(dev)
    q = base.view(B, S, H, D).transpose(1, 2)
    assert not q.is_contiguous()
    want = O.cast(q.cpu().contiguous().numpy(), fmt, -1, tie=O.TIE_AWAY)
    check(gpu_cast(q, fmt, -1), bits(want), "[B,H,S,D] view, blocks along D")
    want = O.cast(q.cpu().contiguous().numpy(), fmt, -2, tie=O.TIE_AWAY)
    check(gpu_cast(q, fmt, -2), bits(want), "[B,H,S,D] view, blocks along S (cols)")
    # row-sliced matrix (row stride != K)
    w = _rand((32, 512), 3).to(dev)[:, 128:384]
    want = O.cast(w.cpu().contiguous().numpy(), fmt, -1, tie=O.TIE_AWAY)
    check(gpu_cast(w, fmt, -1), bits(want), "column slice")
    # misaligned slice -> generic
    w = _rand((32, 515), 4).to(dev)[:, 3:]
    want = O.cast(w.cpu().contiguous().numpy(), fmt, -1, tie=O.TIE_AWAY)
    check(gpu_cast(w, fmt, -1), bits(want), "misaligned slice")


ELEMENTWISE = ["FP[1|5|10,15](FN)", "FP[1|5|10,15](_N)", "FP[1|8|7,127](FN)", "FP[1|4|3,7](_N)", "FP[1|5|2,15](_N)",
               "FP[0|4|4,7](FN)", "BFP[24|8]{1}(SN)", "BFP[8|8]{1}(SN)", "XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,+4](CSN)",
               "XP[8,0](C_N)", "XP[8,0](_SN)", "XP[8,0](CSU)", "XP[8,0](CSD)", "XP[16,+8](CSN)"]


@pytest.mark.parametrize("fmt", ELEMENTWISE)
@pytest.mark.parametrize("tie", ["away", "even"])
def test_oracle_elementwise(fmt, tie):
    if not fmt.startswith("XP") and tie == "even":
        pytest.skip("tie mode only affects XP")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 1000, 8, generator=g) * torch.pow(2.0, torch.randint(-24, 18, (3, 1000, 1), generator=g).float())
    x.view(-1)[::7] = torch.round(x.view(-1)[::7]) + 0.5
    for xt, desc in ((x, "contiguous"), (x.transpose(0, 2), "permuted dense"), (x[:, ::2], "strided"), (x.view(-1)[:4093], "odd length")):
        want = O.cast(xt.contiguous().numpy(), fmt, -1, tie=O.TIE_AWAY if tie == "away" else O.TIE_EVEN)
        y = gpu_cast(xt.to(DEV) if xt.is_contiguous() else _same_layout(xt), fmt, -1, tie)
        check(y, bits(want), f"{fmt} {desc} tie={tie}")


def _same_layout(xt):
    """move a strided CPU view to the GPU keeping its strides"""
    base = xt._base if xt._base is not None else xt
    gb = base.to(DEV)
    return gb.as_strided(xt.shape, xt.stride(), xt.storage_offset())


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("fmt,bd", [("BFP[8|8]{64}(SN)", -1), ("BFP[4|8]{64}(SN)", -1), ("BFP[8|8]{64}(SN)", -2),
                                    ("FP[1|5|10,15](FN)", -1), ("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", -1),
                                    ("BFP[8|8]{16}(_N)", -1)])
def test_oracle_16bit_io(dt, fmt, bd):
    tdt = getattr(torch, dt)
    x = _rand((24, 128, 64), 5, spread=5, dtype=tdt)
    xf = x.float().numpy()
    want32 = O.cast(xf, fmt, bd, tie=O.TIE_AWAY)
    # Format.cast semantics: 16-bit in -> fp32 out
    check(gpu_cast(x.to(DEV), fmt, bd), bits(want32), f"{fmt} {dt}->f32")
    # CastTo.forward semantics: back to the input dtype
    y = ops.cast_chain(x.to(DEV), [fmt_from(fmt).stage()], bd)
    want16 = torch.from_numpy(want32).to(tdt).view(torch.int16).numpy().view(np.uint16)
    check(y, want16, f"{fmt} {dt}->{dt}", dtype=dt)


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("bs", [8, 16, 32, 64, 128])
@pytest.mark.parametrize("wl", [8, 4])
def test_packed_16bit_cols_kernel(dt, bs, wl):
    """bfp_cols16_kernel (strided blocks of a 16-bit tensor, tile kept packed in registers): every tiling, ragged K and
    inner extents, and columns whose block is all-zero / denormal / huge / non-finite (literal path inside the kernel)"""
    tdt = getattr(torch, dt)
    fmt = f"BFP[{wl}|8]{{{bs}}}(SN)"
    for shape in ((3, 2 * bs + bs // 2, 72), (2, 256, 200), (5, bs, 8)):
        x = _rand(shape, 31 + bs + wl, spread=6 if dt == "bfloat16" else 3, dtype=tdt)
        x[0, :, 1] = 0.0
        x[0, :, 2] = x[0, :, 2] * (2.0**-120 if dt == "bfloat16" else 2.0**-12)  # denormal-range column
        if dt == "bfloat16":
            x[1, :, 3] = x[1, :, 3] * 2.0**110  # exponent field >= 228: off the fast path
            x[1, 0, 5] = 3.0e38
        x[1, 1, 4] = float("inf")
        x[1, 2, 6] = float("nan")
        xf = x.float().numpy()
        want32 = O.cast(xf, fmt, -2)
        y = ops.cast_chain(x.to(DEV), [fmt_from(fmt).stage()], -2)
        assert y.dtype == tdt
        want16 = torch.from_numpy(want32).to(tdt).float().numpy()
        check(y.float(), bits(want16), f"{fmt} {dt} {shape}", x=xf, fmt=fmt, block_dim=-2)


@pytest.mark.parametrize("shape,bd", [((32, 256), -1), ((4, 128, 16), 1), ((16, 70), -1)])
@pytest.mark.parametrize("fmt", ["BFP[8|8]{64}(SS)", "BFP[4|8]{16}(SS)", "FP[1|4|3,7](_S)", "XP[8,0](CSS)"])
def test_oracle_stochastic_same_random_tensor(shape, bd, fmt):
    """'identical given the same random tensor': the random tensor is an explicit input."""
    x = _rand(shape, 21, spread=3)
    g = torch.Generator().manual_seed(3)
    f = fmt_from(fmt)
    if fmt.startswith("XP"):
        r = torch.rand(shape, generator=g)
        want = O.cast(x.numpy(), fmt, bd, tie=O.TIE_AWAY, rand=r.numpy())
        y = ops.fixed_qdq(x.to(DEV), f.precision, f.fraction, f.clamp, f.symmetric, "stochastic", rand=r.to(DEV))
    else:
        r = torch.randint(0, 2**31 - 1, shape, generator=g, dtype=torch.int32)
        want = O.cast(x.numpy(), fmt, bd, rand=r.numpy())
        y = ops.cast_chain(x.to(DEV), [f.stage()], bd, rand=r.to(DEV))
    check(y, bits(want), f"{fmt} {shape}")


@pytest.mark.parametrize("k,m", [(2, 4), (4, 8), (2, 8), (1, 2), (8, 16), (1, 4), (3, 4), (5, 6), (2, 32)])
@pytest.mark.parametrize("shape,bd", [((32, 96), -1), ((6, 96, 8), 1), ((96, 5), 0)])
def test_oracle_nm(k, m, shape, bd):
    if shape[bd] % m:
        pytest.skip("not divisible")
    x = _rand(shape, 31)
    x = torch.round(x * 4) / 4  # plenty of ties
    want, wmask = O.nm_prune(x.numpy(), k, m, bd, return_mask=True)
    y, mask = ops.nm_prune(x.to(DEV), k, m, bd, return_mask=True)
    check(mask, bits(wmask), f"{k}:{m} mask")
    check(y, bits(want), f"{k}:{m}")
    g = torch.Generator().manual_seed(1)
    score = torch.rand(shape, generator=g)
    want = O.nm_prune(x.numpy(), k, m, bd, score=score.numpy())
    check(ops.nm_prune(x.to(DEV), k, m, bd, score=score.to(DEV)), bits(want), f"{k}:{m} explicit score")
    xb = x.to(torch.bfloat16)
    want = O.nm_prune(xb.float().numpy(), k, m, bd)
    check(ops.nm_prune(xb.to(DEV), k, m, bd, out_dtype=torch.float32), bits(want), f"{k}:{m} bf16 in")


def test_nm_not_divisible_raises():
    with pytest.raises(AssertionError):
        ops.nm_prune(torch.randn(4, 10, device=DEV), 2, 4)


def test_fused_chain_matches_sequential():
    x = _rand((64, 512), 41, spread=2) * 0.05
    stages = [ops.nm_stage(2, 4), fmt_from("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage(), fmt_from("BFP[8|8]{64}(SN)").stage()]
    want = O.cast(O.cast(O.nm_prune(x.numpy(), 2, 4), "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", tie=O.TIE_AWAY), "BFP[8|8]{64}(SN)")
    check(ops.cast_chain(x.to(DEV), stages, -1), bits(want), "prune->sbfp->bfp fused")
    # FLOAT16 output cast followed by BFP16 input cast (the BASIC-mode pair) in one pass
    stages = [fmt_from("FP[1|5|10,15](FN)").stage(), fmt_from("BFP[8|8]{64}(SN)").stage()]
    want = O.cast(O.cast(x.numpy(), "FP[1|5|10,15](FN)"), "BFP[8|8]{64}(SN)")
    check(ops.cast_chain(x.to(DEV), stages, -1), bits(want), "float16->bfp16 fused")
    # bf16 tensor: intermediates are rounded to bf16 like consecutive CastTo.forward calls
    xb = x.to(torch.bfloat16)
    t = torch.from_numpy(O.cast(xb.float().numpy(), "FP[1|5|10,15](FN)")).to(torch.bfloat16)
    t = torch.from_numpy(O.cast(t.float().numpy(), "BFP[8|8]{64}(SN)")).to(torch.bfloat16)
    y = ops.cast_chain(xb.to(DEV), stages, -1)
    check(y, t.view(torch.int16).numpy().view(np.uint16), "bf16 chain", dtype="bfloat16")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sh", ["BFP[4|8]{64}(SN)", "BFP[8|8]{64}(SN)", "BFP[4|8]{128}(SN)", "BFP[8|8]{16}(SN)", "BFP[6|8]{8}(SN)"])
def test_nm24_bfp_specialisation_on_16bit_tensors(dt, sh):
    """2:4 -> BFP on a bf16 / fp16 tensor runs the straight-line K_NM24_BFP kernel (one sorting network for the mask and
    the block max, BFP before the zeroing).  Checked (a) against the oracle applied stage by stage on finite data with
    ties in |x|, signed zeros, equal groups, blocks whose max sits in the clip quantum, denormal and near-maximal
    blocks; (b) against the general K_NM_BFP kernel -- the same chain with an fp32 output -- on rows that also hold
    Inf / NaN blocks, which must come out identical once rounded to the tensor dtype; (c) on a row-strided view."""
    g = torch.Generator().manual_seed(sum(map(ord, sh)) * 2 + (dt == torch.bfloat16))
    R, K = 96, 1024
    x = torch.randn(R, K, generator=g) * torch.pow(2.0, torch.randint(-6, 5, (R, 1), generator=g).float())
    x = x.to(dt)
    x[1] = x[1, :4].repeat(K // 4)                      # every group the same four values
    x[2] = x[2].abs()[0]                                # all magnitudes equal: the stable order decides
    x[2, 1::2] *= -1
    x[3] = 0.0
    x[3, ::3] = -0.0
    x[4, :64] = torch.tensor(1.9375 if dt == torch.bfloat16 else 1.9990234375).to(dt)   # block max in the top quantum
    x[4, 1:64:2] *= -0.5
    tiny = 2.0**-130 if dt == torch.bfloat16 else 2.0**-22
    x[5] = (torch.randn(K, generator=g) * tiny).to(dt)  # denormal blocks
    big = 2.0**126 if dt == torch.bfloat16 else 30000.0
    x[6] = (torch.sign(torch.randn(K, generator=g)) * big * (1 + torch.rand(K, generator=g))).to(dt)  # near the dtype's maximum
    x.view(-1)[11::97] = x.view(-1)[10::97][: x.view(-1)[11::97].numel()]   # exact ties between neighbours
    assert torch.isfinite(x).all()
    stages = [ops.nm_stage(2, 4), fmt_from(sh).stage()]
    v = torch.int16
    xd = x.to(DEV)
    got = ops.cast_chain(xd, stages, -1)
    assert got.dtype == dt
    want = O.cast(O.nm_prune(x.float().numpy(), 2, 4), sh)
    hugemask = np.zeros((R, K), bool)
    hugemask[6] = dt == torch.bfloat16      # blocks with max >= 2^126: hardware NaN sign differs between x86 and the GPU (see special_block_masks)
    w16 = torch.from_numpy(want).to(dt)
    ok = torch.from_numpy(~hugemask)
    assert torch.equal(got.cpu().view(v)[ok], w16.view(v)[ok])
    # (b) the general kernel on the same GPU, special values included
    xs = x.clone()
    xs[7, 5] = float("inf")
    xs[7, 300] = float("-inf")
    xs[8, 70] = float("nan")
    xs[9, :8] = float("inf")                # three Infs in a group: a pruned Inf times zero
    xsd = xs.to(DEV)
    a = ops.cast_chain(xsd, stages, -1)                                   # K_NM24_BFP
    b = ops.cast_chain(xsd, stages, -1, out_dtype=torch.float32).to(dt)   # K_NM_BFP (16-bit in, fp32 out)
    an, bn = torch.isnan(a), torch.isnan(b)
    assert torch.equal(an, bn)
    assert torch.equal(a.view(v)[~an], b.view(v)[~bn])
    # (c) rows of a wider matrix (not FLAT)
    wide = torch.zeros(R, K + 64, dtype=dt, device=DEV)
    wide[:, :K] = xd
    outw = torch.empty_like(wide)
    ops.cast_chain(wide[:, :K], stages, -1, out=outw[:, :K])
    assert torch.equal(outw[:, :K].contiguous().view(v), got.view(v))


@pytest.mark.parametrize("seed", range(int(os.environ.get("DMXQ_FUZZ_SEEDS", "12")) // 2))
def test_fuzz_fused_chains(seed):
    """seeded fuzzing of the fused kernels (N:M -> BFP, N:M alone, FLOAT -> BFP and three-stage chains) on fp32 / bf16 /
    fp16 tensors with ties in |x| (stable N:M order), zeros, negative zeros and out-of-range values, against the
    oracle applied stage by stage with the tensor dtype's rounding in between (consecutive CastTo.forward calls)"""
    rng = np.random.default_rng(2000 + seed)
    floats = ["FP[1|5|10,15](FN)", "FP[1|8|7,127](FN)", "FP[1|4|3,7](_N)"]
    bfps = ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[8|8]{16}(SN)", "BFP[4|8]{128}(SN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "MXFP8[E4M3]{32}"]
    for _ in range(24):
        rows, K = int(rng.choice([1, 3, 64, 257])), int(rng.choice([128, 256, 384, 1024]))
        dt = [torch.float32, torch.bfloat16, torch.float16][int(rng.integers(3))]
        x = _rand((rows, K), int(rng.integers(1 << 30)), spread=int(rng.integers(1, 8)))
        x = torch.round(x * 8) / 8 if rng.random() < 0.5 else x  # many equal magnitudes
        flat = x.view(-1)
        flat[1::13] = -flat[0::13][: flat[1::13].numel()]  # +-pairs: ties in |x|
        flat[2::29] = -0.0
        if rng.random() < 0.3:
            flat[3::31] *= 2.0**-20  # below the FLOAT16 flush threshold
        x = x.to(dt)
        x = torch.where(torch.isfinite(x), x, torch.zeros_like(x))
        kind = int(rng.integers(4))
        m = int(rng.choice([2, 4, 8]))
        nk = int(rng.integers(1, m))
        if kind == 0:
            chain = [("nm", nk, m), bfps[int(rng.integers(len(bfps)))]]
        elif kind == 1:
            chain = [("nm", nk, m)]
        elif kind == 2:
            chain = [floats[int(rng.integers(len(floats)))], bfps[int(rng.integers(4))]]
        else:
            chain = [("nm", nk, m), bfps[int(rng.integers(len(bfps)))], "BFP[8|8]{64}(SN)"]
        stages, t = [], x
        for st in chain:
            if isinstance(st, tuple):
                stages.append(ops.nm_stage(st[1], st[2]))
                t = torch.from_numpy(O.nm_prune(t.float().numpy(), st[1], st[2])).to(dt)
            else:
                stages.append(fmt_from(st).stage())
                t = torch.from_numpy(O.cast(t.float().numpy(), st, -1, tie=O.TIE_AWAY)).to(dt)
        y = ops.cast_chain(x.to(DEV), stages, -1)
        assert y.dtype == dt
        check(y.float(), bits(t.float().numpy()), f"seed {seed}: {chain} [{rows},{K}] {dt}")


def test_minmax_exact():
    x = _rand((7, 33, 50), 51)
    mn, mx = ops.minmax(x.to(DEV))
    assert mn.item() == x.min().item() and mx.item() == x.max().item()
    mn, mx = ops.minmax(x.to(DEV), ch_axis=1)
    assert torch.equal(mn.cpu(), x.amin((0, 2))) and torch.equal(mx.cpu(), x.amax((0, 2)))
    xb = x.to(torch.bfloat16)
    mn, mx = ops.minmax(xb.to(DEV), ch_axis=0)
    assert torch.equal(mn.cpu(), xb.float().amin((1, 2))) and torch.equal(mx.cpu(), xb.float().amax((1, 2)))
    x[2, 3, 4] = float("nan")
    mn, mx = ops.minmax(x.to(DEV))
    assert torch.isnan(mn).all() and torch.isnan(mx).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,axis", [((96, 4096), 0), ((96, 4096), 1), ((513, 1032), 1), ((4, 24, 2048), 1), ((3, 7, 64), 2),
                                        ((8, 100, 768), 2), ((2000, 8), 1), ((5, 6, 49), 1), ((70000, 8), 0), ((16, 40), 1)])
def test_minmax_per_channel_kernels(dt, shape, axis):
    """vectorised per-channel statistics (channel = contiguous run / channel = column) and the scalar fallback, exact,
    NaN confined to its own channel"""
    x = _rand(shape, 57, spread=4).to(dt)
    dims = tuple(d for d in range(len(shape)) if d != axis)
    mn, mx = ops.minmax(x.to(DEV), ch_axis=axis)
    assert torch.equal(mn.cpu(), x.float().amin(dims)) and torch.equal(mx.cpu(), x.float().amax(dims))
    idx = [0] * len(shape)
    idx[axis] = shape[axis] // 2
    x[tuple(idx)] = float("nan")
    mn, mx = ops.minmax(x.to(DEV), ch_axis=axis)
    want_mn, want_mx = x.float().amin(dims), x.float().amax(dims)
    assert torch.isnan(mn[shape[axis] // 2]) and torch.isnan(mx[shape[axis] // 2])
    keep = torch.ones(shape[axis], dtype=torch.bool)
    keep[shape[axis] // 2] = False
    assert torch.equal(mn.cpu()[keep], want_mn[keep]) and torch.equal(mx.cpu()[keep], want_mx[keep])


@pytest.mark.parametrize("dim", [-1, 0, 1, 2])
@pytest.mark.parametrize("mode", ["nearest", "up", "down", "stochastic"])
@pytest.mark.parametrize("symmetric", [True, False])
def test_l1_block_quantize(dim, mode, symmetric):
    """L1 mirror of quant_cuda.block_quantize_* (dim semantics of Q/quant_cuda/quant.cu:14-34)."""
    x = _rand((12, 20, 16), 61, spread=2)
    x[3, 4, 5] = -x[3].abs().max() * 1.0  # exercise the !symmetric branch candidates
    g = torch.Generator().manual_seed(2)
    r = torch.randint(0, 2**31 - 1, x.shape, generator=g, dtype=torch.int32)
    xn = x.numpy()
    if dim == -1:
        rows = xn.reshape(1, -1)
        rr = r.numpy().reshape(1, -1)
        want = O.block_quantize_rows(rows, 8, symmetric, mode, rand=rr).reshape(xn.shape)
    elif dim == 0:
        want = O.block_quantize_rows(xn.reshape(12, -1), 8, symmetric, mode, rand=r.numpy().reshape(12, -1)).reshape(xn.shape)
    elif dim == 1:
        xt = np.ascontiguousarray(np.swapaxes(xn, 0, 1)).reshape(20, -1)
        rt = np.ascontiguousarray(np.swapaxes(r.numpy(), 0, 1)).reshape(20, -1)
        want = np.swapaxes(O.block_quantize_rows(xt, 8, symmetric, mode, rand=rt).reshape(20, 12, 16), 0, 1)
    else:  # slices along the contiguous dim
        xt = np.ascontiguousarray(np.moveaxis(xn, 2, 0)).reshape(16, -1)
        rt = np.ascontiguousarray(np.moveaxis(r.numpy(), 2, 0)).reshape(16, -1)
        want = np.moveaxis(O.block_quantize_rows(xt, 8, symmetric, mode, rand=rt).reshape(16, 12, 20), 0, 2)
    y = ops.block_quantize_l1(x.to(DEV), 8, dim, symmetric, mode, rand=r.to(DEV))
    check(y, bits(np.ascontiguousarray(want)), f"block_quantize dim={dim} {mode} sym={symmetric}")


# =============================================================================== (c) reference CUDA kernels
@pytest.fixture(scope="module")
def ref_cuda():
    import build_ref

    try:
        mod = build_ref.load("ref_quant_cuda")
    except Exception as e:  # pragma: no cover
        pytest.skip(f"ref_quant_cuda.so not loadable: {e}")
    if mod is None:
        pytest.skip("oracle/_ref/ref_quant_cuda.so did not travel")
    return mod


def test_vs_reference_cuda_float_fixed_block(ref_cuda):
    x = _rand((256, 64), 71, spread=12).to(DEV)
    x.view(-1)[::5] = torch.round(x.view(-1)[::5]) + 0.5  # exact .5 ties: pins half-away
    for man, exp, bias, flush in [(10, 5, 15, True), (3, 4, 7, False), (2, 5, 15, False), (7, 8, 127, True)]:
        want = ref_cuda.float_quantize_nearest(x, man, exp, bias, flush)
        got = ops.float_qdq(x, man, exp, bias, flush)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"float m{man}e{exp}"
    for wl, fl, clamp, sym in [(8, 0, True, True), (4, 0, True, True), (8, 4, True, False), (8, 0, False, True)]:
        for mode in ("nearest", "up", "down"):
            want = getattr(ref_cuda, f"fixed_point_quantize_{mode}")(x, wl, fl, clamp, sym)
            got = ops.fixed_qdq(x, wl, fl, clamp, sym, mode, tie=L.TIE_AWAY)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"fixed {wl},{fl} {mode}"
    for wl in (4, 8, 16):
        for mode in ("nearest", "up", "down"):
            want = getattr(ref_cuda, f"block_quantize_{mode}")(x, wl, 0, True)
            got = ops.bfp_qdq(x, -1, 64, wl, True, mode)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"block wl={wl} {mode}"
            got = ops.block_quantize_l1(x, wl, 0, True, mode)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"L1 block wl={wl} {mode}"


def test_vs_reference_cuda_adversarial(ref_cuda):
    """Inf / NaN / huge / denormal / zero blocks: bit-identical (NaN payloads included) to the
    reference's own CUDA kernels, which run the same IEEE ops on the same hardware."""
    rows = []
    g = torch.Generator().manual_seed(5)
    for base in (1.0, 3e-39, 1e-42, 2.0**100, 2.0**125, 2.0**126, 2.0**127, 3e38):
        b = torch.randn(64, generator=g) * base
        b[0] = base
        rows.append(b)
    for special in (float("inf"), float("-inf"), float("nan")):
        b = torch.randn(64, generator=g)
        b[7] = special
        b[9] = -special
        rows.append(b)
    rows.append(torch.zeros(64))
    rows.append(-torch.zeros(64))
    x = torch.stack(rows).to(DEV)
    for wl in (4, 8):
        for mode in ("nearest", "up", "down"):
            want = getattr(ref_cuda, f"block_quantize_{mode}")(x, wl, 0, True)
            got = ops.bfp_qdq(x, -1, 64, wl, True, mode)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"adversarial block wl={wl} {mode}"
    xe = torch.cat([x.view(-1), torch.tensor([1e-45, -1e-45, 65504.0, 65520.0, 1e30, -1e30, 6.1e-5, 5.9e-8], device=DEV)])
    for man, exp, bias, flush in [(10, 5, 15, True), (10, 5, 15, False), (3, 4, 7, False), (22, 8, 127, False), (7, 8, 127, True)]:
        want = ref_cuda.float_quantize_nearest(xe, man, exp, bias, flush)
        got = ops.float_qdq(xe, man, exp, bias, flush)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"adversarial float m{man}e{exp}"
    want = ref_cuda.fixed_point_quantize_nearest(xe, 8, 0, True, True)
    got = ops.fixed_qdq(xe, 8, 0, True, True, "nearest")
    assert torch.equal(got.view(torch.int32), want.view(torch.int32)), "adversarial fixed"


def test_vs_reference_cuda_stochastic_stream(ref_cuda):
    """The reference draws randint_like(a, INT_MAX) / rand_like(a) internally (quant.cu:40,160,244):
    re-seeding torch and drawing the same tensor ourselves must reproduce its output bit for bit."""
    x = _rand((128, 64), 81, spread=4).to(DEV)
    torch.manual_seed(123)
    want = ref_cuda.block_quantize_stochastic(x, 8, 0, True)
    torch.manual_seed(123)
    zeros = torch.zeros_like(x)  # the reference allocates its output before drawing
    r = torch.randint_like(x, 2**31 - 1, dtype=torch.int32)
    got = ops.bfp_qdq(x, -1, 64, 8, True, "stochastic", rand=r)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    torch.manual_seed(7)
    want = ref_cuda.fixed_point_quantize_stochastic(x, 8, 0, True, True)
    torch.manual_seed(7)
    r = torch.rand_like(x)
    got = ops.fixed_qdq(x, 8, 0, True, True, "stochastic", rand=r)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_vs_torch_cuda_argsort():
    """BlockTopK on CUDA uses torch.argsort (S/sparse.py:172), whose default CUDA sort for rows
    <= 32 is an *unstable* bitonic network: on tie-free scores our mask equals it exactly; on
    tied scores the reference's own CPU and CUDA back ends disagree with each other, and we
    follow the stable (CPU) order, i.e. torch.argsort(stable=True)."""
    for m, k in ((4, 2), (8, 4), (8, 2)):
        x = torch.randn(1 << 16, m, device=DEV)  # continuous: no ties
        score = x.abs()
        idx = torch.argsort(score, dim=1)[:, : m - k]
        want = x * torch.ones_like(score).scatter_(dim=1, index=idx, value=0)
        got = ops.nm_prune(x, k, m, -1)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"{k}:{m} tie-free"
        x = torch.round(torch.randn(1 << 16, m, device=DEV) * 2) / 2  # tie-heavy
        score = x.abs()
        idx = torch.argsort(score, dim=1, stable=True)[:, : m - k]
        want = x * torch.ones_like(score).scatter_(dim=1, index=idx, value=0)
        got = ops.nm_prune(x, k, m, -1)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"{k}:{m} ties, stable order"
        # either way the pruned set has the same multiset of scores: the kept |x| sum is identical
        idx = torch.argsort(score, dim=1)[:, : m - k]
        unstable = x * torch.ones_like(score).scatter_(dim=1, index=idx, value=0)
        assert torch.equal(unstable.abs().sum(1), got.abs().sum(1))


def test_nm_torch_cuda_tie_order():
    """nm_order = NM_TORCH_CUDA: the pruned set is exactly the first M-K entries of torch.argsort's default (unstable) CUDA
    order -- checked against torch.argsort itself on this GPU and against the oracle's restatement of the bitonic network --
    for in-thread groups (rows kernel), groups wider than a thread's vector / not a power of two (generic kernel), score
    tensors, mask output, strided block dims, 16-bit sources and the fused sparsify -> BFP chain"""
    g = torch.Generator(device=DEV).manual_seed(123)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        for m, k in ((2, 1), (4, 2), (4, 1), (4, 3), (8, 4), (8, 2), (16, 8), (32, 16), (6, 3)):
            x = (torch.randint(-4, 5, (512, 2 * 96), device=DEV, generator=g).float() / 4).to(dt)  # tie-heavy
            score = x.abs().float()
            idx = torch.argsort(score.reshape(-1, m), dim=1)[:, : m - k]
            mask = torch.ones_like(score).reshape(-1, m).scatter_(dim=1, index=idx, value=0).reshape(x.shape)
            want = x * mask.to(dt)
            got, gmask = ops.nm_prune(x, k, m, -1, return_mask=True, nm_order=L.NM_TORCH_CUDA)
            assert torch.equal(gmask, mask), f"{k}:{m} {dt} mask vs torch.argsort"
            assert torch.equal(got.view(torch.int16 if dt != torch.float32 else torch.int32),
                               want.view(torch.int16 if dt != torch.float32 else torch.int32)), f"{k}:{m} {dt}"
            got2 = ops.nm_prune(x, k, m, -1, nm_order=L.NM_TORCH_CUDA)  # no mask output: another kernel specialisation
            assert torch.equal(got2, got)
            ow = O.nm_prune(x.float().cpu().numpy(), k, m, -1, nm_order=O.NM_TORCH_CUDA)
            assert (got.float().cpu().numpy().view(np.uint32) == ow.view(np.uint32)).all(), f"{k}:{m} {dt} vs oracle"
            if m <= 8:
                st = ops.nm_prune(x, k, m, -1)
                assert not torch.equal(st, got), "tie-heavy data must tell the two orders apart"
    # explicit score tensor (the plugin's path), block dim 0 (strided groups), and the sparsify -> BFP12 chain
    x = torch.randn(64, 256, device=DEV, generator=g)
    sc = torch.randint(0, 3, (64, 256), device=DEV, generator=g).float()
    for bd, m, k in ((-1, 4, 2), (0, 4, 2), (0, 8, 4), (-1, 8, 4)):
        s2 = sc.transpose(bd, -1).reshape(-1, m)
        idx = torch.argsort(s2, dim=1)[:, : m - k]
        mask = torch.ones_like(s2).scatter_(dim=1, index=idx, value=0).reshape(sc.transpose(bd, -1).shape).transpose(bd, -1)
        got, gmask = ops.nm_prune(x, k, m, bd, score=sc, return_mask=True, nm_order=L.NM_TORCH_CUDA)
        assert torch.equal(gmask, mask) and torch.equal(got, x * mask), f"score tensor, block_dim {bd}, {k}:{m}"
    for dt in (torch.float32, torch.bfloat16):
        w = (torch.randint(-8, 9, (128, 512), device=DEV, generator=g).float() / 8).to(dt)
        f12 = fmt_from("BFP[4|8]{64}(SN)").stage()
        fused = ops.cast_chain(w, [ops.nm_stage(2, 4, L.NM_TORCH_CUDA), f12], -1)
        two = ops.cast_chain(ops.nm_prune(w, 2, 4, -1, nm_order=L.NM_TORCH_CUDA), [f12], -1)
        assert torch.equal(fused, two), f"fused 2:4 -> BFP12 in torch order, {dt}"
        assert not torch.equal(fused, ops.cast_chain(w, [ops.nm_stage(2, 4), f12], -1))


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_sbfp_scale_modes(dt):
    """the SBFP block scale as the reference computes it on CPU tensors (max / 7) and on CUDA tensors (max * fp32(1/7):
    torch's division of a CUDA tensor by a python scalar), each with both XP tie rules, against the oracle; rows, cols and
    generic kernels; packed storage follows the same rule"""
    x = _rand((256, 1024), 321, spread=10).to(dt)
    xn = x.float().numpy()
    for sh in ("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[8,0](CSN)><FP[0|4|4,5](FN)>{32}"):
        for tie in ("away", "even"):
            for mode in (L.SCALE_DIV, L.SCALE_RECIP):
                f = fmt_from(sh, tie)
                f.scale_mode = mode
                for bd in (-1, 0):
                    got = f.cast(x.to(DEV), bd)
                    m = O._RX_SBFP.match(sh)
                    xp, fp = O._RX_XP.match(m[1]), O._RX_FP.match(m[2])
                    want = O.sbfp_cast(xn, bd, int(m[3]), int(xp[1]), True, "nearest", TIE[tie], int(fp[3]), int(fp[2]), int(fp[4]),
                                       True, True, "nearest", scale_recip=bool(mode))
                    check(got, bits(want), f"{sh} tie={tie} scale_mode={mode} bd={bd} {dt}")
    # default: the scale rule goes with the tie rule
    sh = "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"
    a = fmt_from(sh, "away").cast(x.to(DEV), -1)
    f = fmt_from(sh, "away"); f.scale_mode = L.SCALE_RECIP
    assert torch.equal(a, f.cast(x.to(DEV), -1))
    f = fmt_from(sh, "away"); f.scale_mode = L.SCALE_DIV
    if dt != torch.float32:
        assert not torch.equal(a, f.cast(x.to(DEV), -1)), "16-bit data: the two scale rules must differ on some tie"


# =============================================================================== (d) properties at full size
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fmt", ["BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)"])
def test_full_size_properties(dt, fmt):
    n = 2**28
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n // 4096, 4096, device=DEV, generator=g)
    x *= torch.pow(2.0, torch.randint(-8, 9, (n // 4096, 1), device=DEV, generator=g).float())
    x = x.to(dt)
    st = [fmt_from(fmt).stage()]
    y = ops.cast_chain(x, st, -1)
    # idempotence: cast(cast(x)) == cast(x)
    y2 = ops.cast_chain(y, st, -1)
    assert torch.equal(y.view(torch.int16 if dt == torch.bfloat16 else torch.int32), y2.view(torch.int16 if dt == torch.bfloat16 else torch.int32))
    # block-scale equivariance: cast(x * 2^k) == cast(x) * 2^k (exact, no overflow at these scales)
    y3 = ops.cast_chain(x * 8.0, st, -1)
    assert torch.equal(y3, y * 8.0)
    # error bound: |y - x| <= one quantum of the block (half a quantum from rounding, up to one at
    # the clipped top of the range); every output is an integer multiple of the quantum
    wl = int(fmt[4:].split("|")[0])
    xf, yf = x.float().view(-1, 64), y.float().view(-1, 64)
    e = torch.floor(torch.log2(xf.abs().amax(-1, keepdim=True).clamp_min(1e-30)))
    quantum = torch.pow(2.0, e + 2 - wl)
    assert ((yf - xf).abs() <= quantum).all()
    assert torch.equal(torch.round(yf / quantum) * quantum, yf)
    assert (yf.abs() <= (2 ** (wl - 1) - 1) * quantum).all()
    # a sampled slab equals the oracle bit for bit
    rows = slice(1000, 1016)
    want = O.cast(x[rows].float().cpu().numpy(), fmt, -1)
    if dt == torch.float32:
        check(y[rows], bits(want), "sampled slab")
    else:
        check(y[rows], torch.from_numpy(want).to(dt).view(torch.int16).numpy().view(np.uint16), "sampled slab", dtype="bfloat16")


def test_host_entry_matches_device():
    x = _rand((3000, 4096), 91)
    st = [fmt_from("BFP[8|8]{64}(SN)").stage()]
    xh = x.pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    ops.cast_chain_host(xh, yh, st, 0)
    want = ops.cast_chain(x.to(DEV), st, -1).cpu()
    assert torch.equal(yh.view(torch.int32), want.view(torch.int32))


def test_sbfp_on_adversarial_ties():
    """SBFP rounds fl(x / cmax) half-away: hammer exactly the boundaries (k + 0.5) * cmax +- a few ulps,
    where any shortcut around the IEEE division would show."""
    g = torch.Generator().manual_seed(77)
    rows = []
    for _ in range(512):
        m = (torch.rand(1, generator=g) * 4 + 0.1) * 2.0 ** int(torch.randint(-20, 20, (1,), generator=g))
        cmax = (m / 7.0).float()
        k = torch.randint(0, 7, (15,), generator=g).float() + 0.5
        x = (k * cmax).float()
        ulps = torch.randint(-3, 4, (15,), generator=g)
        x = (x.view(torch.int32) + ulps.int()).view(torch.float32)
        x = x * (torch.randint(0, 2, (15,), generator=g).float() * 2 - 1)
        rows.append(torch.cat([m.float(), x]))
    x = torch.stack(rows)
    for sh in ("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[8,0](CSN)><FP[0|4|4,12](FN)>{16}", "SBFP<XP[4,0](_SN)><FP[1|5|10,15](FN)>{16}"):
        want = O.cast(x.numpy(), sh, -1, tie=O.TIE_AWAY)
        check(gpu_cast(x.to(DEV), sh, -1), bits(want), sh)
        xb = x.to(torch.bfloat16)
        want = O.cast(xb.float().numpy(), sh, -1, tie=O.TIE_AWAY)
        check(gpu_cast(xb.to(DEV), sh, -1), bits(want), sh + " bf16")


def test_sbfp_division_free_quotient_at_scale():
    """the kernel's x / cmax is a reciprocal with two FMA refinements: 2^23 tie-adjacent quotients (vectorised version of
    the test above, every block its own scale) and 2^22 ordinary values must match the dividing oracle bit for bit,
    including blocks whose cmax forces the IEEE-division path (significand all ones, tiny, huge)"""
    g = torch.Generator().manual_seed(78)
    nblk = 1 << 19
    m = (torch.rand(nblk, 1, generator=g) * 4 + 0.1) * torch.pow(2.0, torch.randint(-30, 30, (nblk, 1), generator=g).float())
    m[0::1001] = (7.0 * (2.0 - 2.0**-23))  # cmax = 2 - ulp: all-ones significand
    m[1::1001] *= 2.0**-90
    m[2::1001] *= 2.0**80
    cmax = m / 7.0
    k = torch.randint(0, 7, (nblk, 15), generator=g).float() + 0.5
    x = k * cmax
    x = (x.view(torch.int32) + torch.randint(-2, 3, (nblk, 15), generator=g, dtype=torch.int32)).view(torch.float32)
    x = x * (torch.randint(0, 2, (nblk, 15), generator=g).float() * 2 - 1)
    x = torch.cat([m, x], 1).reshape(-1, 4096)
    for sh in ("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[8,0](CSN)><FP[0|4|4,12](FN)>{16}"):
        check(gpu_cast(x.to(DEV), sh, -1), bits(O.cast(x.numpy(), sh, -1, tie=O.TIE_AWAY)), sh)
    y = _rand((1024, 4096), 79, spread=20)
    sh = "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"
    check(gpu_cast(y.to(DEV), sh, -1), bits(O.cast(y.numpy(), sh, -1, tie=O.TIE_AWAY)), sh)
    yb = y.to(torch.bfloat16)
    check(gpu_cast(yb.to(DEV), sh, -1), bits(O.cast(yb.float().numpy(), sh, -1, tie=O.TIE_AWAY)), sh + " bf16")
    check(gpu_cast(y.to(DEV), sh, 0), bits(O.cast(y.numpy(), sh, 0, tie=O.TIE_AWAY)), sh + " cols")


def test_int8_device_qparams_fast_path():
    """CastTo(INT8/INT4) per-tensor: vectorised kernel reading scale / zero-point from device memory"""
    x = _rand((64, 1024), 91, spread=3) * 20
    for sc, zp in ((1.0, 0.0), (0.05, 3.0), (0.37, -2.0)):
        for fmt in ("XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,+2](C_N)"):
            f = fmt_from(fmt)
            want = O.cast(x.numpy(), fmt, tie=O.TIE_AWAY, scale=[sc], zero_point=[zp])
            y = ops.fixed_qdq(x.to(DEV), f.precision, f.fraction, f.clamp, f.symmetric, "nearest", scale=torch.tensor([sc], device=DEV),
                              zero_point=torch.tensor([zp], device=DEV))
            check(y, bits(want), f"{fmt} sc={sc} zp={zp}")
            xb = x.to(torch.bfloat16)
            want = torch.from_numpy(O.cast(xb.float().numpy(), fmt, tie=O.TIE_AWAY, scale=[sc], zero_point=[zp])).to(torch.bfloat16)
            y = ops.fixed_qdq(xb.to(DEV), f.precision, f.fraction, f.clamp, f.symmetric, "nearest", scale=torch.tensor([sc], device=DEV),
                              zero_point=torch.tensor([zp], device=DEV))
            check(y, want.view(torch.int16).numpy().view(np.uint16), f"{fmt} bf16 sc={sc}", dtype="bfloat16")


def test_more_than_2_31_elements():
    """The reference indexes with `int` and breaks above 2^31 elements; we index
    with int64: a 2^31 + 2^20 element bf16 tensor, checked against the oracle at both ends."""
    n = (1 << 31) + (1 << 20)
    x = torch.empty(n // 4096, 4096, device=DEV, dtype=torch.bfloat16)
    x.normal_()
    st = [fmt_from("BFP[8|8]{64}(SN)").stage()]
    y = ops.cast_chain(x, st, -1)
    for rows in (slice(0, 8), slice(n // 4096 - 8, n // 4096), slice((1 << 31) // 4096 - 4, (1 << 31) // 4096 + 4)):
        want = torch.from_numpy(O.cast(x[rows].float().cpu().numpy(), "BFP[8|8]{64}(SN)")).to(torch.bfloat16)
        check(y[rows], want.view(torch.int16).numpy().view(np.uint16), f"rows {rows}", dtype="bfloat16")
    # strided (non-flat) addressing beyond 2^31 as well
    xs = x[:, :2048]
    ys = ops.cast_chain(xs, st, -1)
    rows = slice(n // 4096 - 8, n // 4096)
    want = torch.from_numpy(O.cast(xs[rows].float().cpu().numpy(), "BFP[8|8]{64}(SN)")).to(torch.bfloat16)
    check(ys[rows], want.view(torch.int16).numpy().view(np.uint16), "strided tail", dtype="bfloat16")


def test_degenerate_shapes():
    for shape in ((0, 64), (4, 0), (), (1,), (1, 1, 1)):
        x = torch.randn(shape, device=DEV)
        for sh in ("FP[1|5|10,15](FN)", "XP[8,0](CSN)"):
            y = gpu_cast(x, sh)
            want = O.cast(x.cpu().numpy().reshape(shape), sh, tie=O.TIE_AWAY)
            assert y.shape == x.shape
            if x.numel():
                check(y, bits(want), f"{sh} {shape}")
        if len(shape) >= 1:
            y = gpu_cast(x, "BFP[8|8]{64}(SN)", -1)
            assert y.shape == x.shape
            if x.numel():
                check(y, bits(O.cast(x.cpu().numpy(), "BFP[8|8]{64}(SN)", -1)), f"bfp {shape}")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_per_channel_and_group_fixed_point_vectorised(dt):
    """per-channel (ch_axis 0 and 1) and group qparams on shapes that take the 16-byte-vector kernel"""
    g = torch.Generator().manual_seed(13)
    x = (torch.randn(48, 64, 32, generator=g) * 30).to(dt)
    xf = x.float().numpy()
    for ch_axis, group in ((0, None), (1, None), (1, 16), (0, 5)):
        C = x.shape[ch_axis]
        nq = C if group is None else -(-C // group)
        sc = torch.rand(nq, generator=g) * 0.5 + 0.02
        zp = torch.round(torch.randn(nq, generator=g) * 4)
        want = O.cast(xf, "XP[8,0](CSN)", tie=O.TIE_AWAY, scale=sc.tolist(), zero_point=zp.tolist(), ch_axis=ch_axis, group_size=group)
        y = ops.fixed_qdq(x.to(DEV), 8, 0, True, True, "nearest", scale=sc.to(DEV), zero_point=zp.to(DEV), ch_axis=ch_axis, group_size=group)
        if dt == torch.float32:
            check(y, bits(want), f"ch_axis={ch_axis} group={group}")
        else:
            check(y, torch.from_numpy(want).to(dt).view(torch.int16).numpy().view(np.uint16), f"ch_axis={ch_axis} group={group}", dtype="bfloat16")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_fixed_point_qparams_along_contiguous_dim(dt):
    """per-channel / group qparams where the channel runs along the contiguous dim (activations per feature): the
    column kernel (any group size), the vector-uniform kernel (groups of whole vectors), ragged column counts, quotient
    rounding ties (x / scale is reciprocal-based), non-finite and huge inputs (IEEE-division path), other fraction bits"""
    g = torch.Generator().manual_seed(29)
    for shape, group in (((257, 328), None), ((64, 512), 128), ((64, 512), 3), ((33, 40), 8), ((1000, 8), None), ((4, 9, 264), None), ((16, 96), 24)):
        C = shape[-1]
        nq = C if group is None else -(-C // group)
        sc = torch.rand(nq, generator=g) * 0.5 + 0.02
        sc[0] = float(np.float32(2.0) - np.float32(2.0**-23)) / 16  # significand all ones: must divide
        zp = torch.round(torch.randn(nq, generator=g) * 4)
        x = torch.randn(shape, generator=g) * 20
        q = sc.repeat_interleave(group or 1)[:C]
        ties = (torch.randint(-120, 120, shape, generator=g).float() + 0.5) * q  # x / scale on (and an ulp around) a tie
        ties = (ties.view(torch.int32) + torch.randint(-1, 2, shape, generator=g, dtype=torch.int32)).view(torch.float32)
        x = torch.where(torch.rand(shape, generator=g) < 0.3, ties, x).to(dt)
        if dt != torch.float16:
            x.view(-1)[5] = 3.0e38
        x.view(-1)[6] = float("inf")
        x.view(-1)[7] = float("nan")
        x.view(-1)[8] = -0.0
        xf = x.float().numpy()
        for fmt, (wl, fl, sym) in (("XP[8,0](CSN)", (8, 0, True)), ("XP[4,0](C_N)", (4, 0, False)), ("XP[8,+2](CSN)", (8, 2, True))):
            want = O.cast(xf, fmt, tie=O.TIE_AWAY, scale=sc.tolist(), zero_point=zp.tolist(), ch_axis=len(shape) - 1, group_size=group)
            y = ops.fixed_qdq(x.to(DEV), wl, fl, True, sym, "nearest", scale=sc.to(DEV), zero_point=zp.to(DEV), ch_axis=-1, group_size=group)
            if dt == torch.float32:
                check(y, bits(want), f"{fmt} {shape} group={group}")
            else:
                check(y.float(), bits(torch.from_numpy(want).to(dt).float().numpy()), f"{fmt} {shape} group={group} {dt}")


def test_asymmetric_bfp_edge_blocks():
    """BFP16A / BFP12A fast path: blocks whose most negative element sits on / next to the -(2^(wl-1)-1)
    mantissa, with the block max just below, at and above the clamp threshold"""
    g = torch.Generator().manual_seed(17)
    rows = []
    for wl in (8, 4):
        q = 2.0 ** (2 - wl)
        for top in (2 - q, 2 - 1.5 * q, 2 - 0.75 * q, 2 - 0.5 * q, 2 - 0.25 * q, 1.9999999, 2 - 2 * q):
            for sign in (1.0, -1.0):
                b = torch.randn(64, generator=g) * 0.3
                b[0] = sign * top
                b[1] = -(2 - q)
                b[2] = -(2 - 0.5 * q)
                b[3] = -(2 - 1.49 * q)
                b[4] = -(2 - 0.51 * q)
                rows.append(b * 2.0 ** int(torch.randint(-10, 10, (1,), generator=g)))
    x = torch.stack(rows)
    for sh in ("BFP[8|8]{64}(_N)", "BFP[4|8]{64}(_N)", "BFP[6|8]{16}(_N)"):
        want = O.cast(x.numpy(), sh, -1)
        check(gpu_cast(x.to(DEV), sh, -1), bits(want), sh)
        xb = x.to(torch.bfloat16)
        want = O.cast(xb.float().numpy(), sh, -1)
        check(gpu_cast(xb.to(DEV), sh, -1), bits(want), sh + " bf16")
        want = O.cast(x.numpy().T.copy(), sh, 0)
        check(gpu_cast(x.t().contiguous().to(DEV), sh, 0), bits(want), sh + " cols")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("wl,bs", [(8, 64), (4, 64), (8, 16), (6, 32), (4, 128), (2, 64), (8, 128)])
def test_packed_bfp_storage_round_trip(dt, wl, bs):
    """dmxq_bfp_unpack(dmxq_bfp_pack(x)) == the QDQ cast, bit for bit; packed size == Format.bytes_per_elem"""
    from dmx_compressor_b200.numerical import BlockFloatingPoint

    if dt != torch.float32 and bs // 8 > 32:
        pytest.skip("block too wide for one warp")
    x = _rand((96, 1024), 100 + wl + bs, spread=12 if dt != torch.float16 else 5).to(dt)  # (fp16 would overflow to inf)
    x.view(-1)[::5] = torch.round(x.view(-1)[::5].float() * 16).to(dt) / 16
    x[5] = 0
    assert torch.isfinite(x).all()
    xd = x.to(DEV)
    f = BlockFloatingPoint(precision=wl, block_size=bs)
    want = ops.cast_chain(xd, [f.stage()], -1)
    mant, exps = ops.bfp_pack(xd, bs, wl)
    stored_bits = 8 if wl > 4 else 4  # mantissas are byte- or nibble-aligned
    assert mant.numel() * mant.element_size() + exps.numel() == x.numel() * stored_bits // 8 + x.numel() // bs
    if wl in (8, 4):
        assert mant.numel() * mant.element_size() + exps.numel() == int(f.bytes_per_elem * x.numel())
    got = ops.bfp_unpack(mant, exps, bs, wl, dtype=dt)
    v = torch.int32 if dt == torch.float32 else torch.int16
    assert torch.equal(got.view(v), want.view(v))
    # mantissas really are wl-bit integers and the exponent byte is the block exponent
    if wl > 4:
        assert int(mant.abs().max()) <= 2 ** (wl - 1) - 1
    e = torch.floor(torch.log2(x.float().abs().view(96, -1, bs).amax(-1).clamp_min(1e-38))) + 127
    nz = x.float().abs().view(96, -1, bs).amax(-1) > 0
    assert torch.equal(exps.cpu().float()[nz], e[nz])
    # and against the oracle
    if dt == torch.float32:
        check(got, bits(O.cast(x.numpy(), f"BFP[{wl}|8]{{{bs}}}(SN)", -1)), "packed vs oracle")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("prec,bias,bs", [(4, 7, 16), (4, 4, 16), (4, 10, 16), (8, 7, 16), (4, 7, 64), (6, 7, 8), (4, 7, 128), (2, 7, 32)])
def test_packed_sbfp_storage_round_trip(dt, prec, bias, bs):
    """dmxq_sbfp_unpack(dmxq_sbfp_pack(x)) == the SBFP QDQ cast bit for bit (signs of zeros included); the bytes decode
    to the oracle's SBFP cast with plain numpy; SBFP12_16 is 0.5625 B per element"""
    sh = f"SBFP<XP[{prec},0](CSN)><FP[0|4|4,{bias}](FN)>{{{bs}}}"
    f = fmt_from(sh)
    # magnitudes inside the byte's scaler range for every bias here (bias 10: max scaler exponent 15 - 10 = 5)
    x = _rand((96, 1024), 300 + prec + bias + bs, spread=3).to(dt)
    x.view(-1)[::5] = torch.round(x.view(-1)[::5].float() * 16).to(dt) / 16
    x[5] = 0
    x[6, ::2] = -0.0
    x[7] = x[7] * 2.0**-12  # scalers below the smallest normal of the scaler format: flushed to zero, signs survive
    xd = x.to(DEV)
    want = ops.cast_chain(xd, [f.stage()], -1)
    mant, scal, bad = f.pack(xd, return_inexact=True)
    assert int(bad) == 0
    nbytes = mant.numel() + scal.numel()
    assert nbytes == x.numel() * (4 if prec <= 4 else 8) // 8 + x.numel() // bs
    # (the reference's bytes_per_elem counts a sign bit for the unsigned scaler too: 9 bits where the byte holds all of it)
    if prec in (4, 8):
        assert nbytes <= f.bytes_per_elem * x.numel() < nbytes + x.numel() // bs
    got = f.unpack(mant, scal, dtype=dt)
    v = torch.int32 if dt == torch.float32 else torch.int16
    assert torch.equal(got.view(v), want.view(v))
    # the bytes themselves are the oracle's, and they decode (plain numpy) to the oracle's SBFP cast
    om, os_, obad = O.sbfp_pack(x.float().numpy(), bs, prec, O.TIE_AWAY, 4, 4, bias)
    assert obad == 0
    assert np.array_equal(mant.cpu().numpy(), om) and np.array_equal(scal.cpu().numpy(), os_)
    dec = O.sbfp_unpack(mant.cpu().numpy(), scal.cpu().numpy(), bs, prec, 4, bias)
    assert_bits_equal(bits(dec), bits(O.cast(x.float().numpy(), sh, -1, tie=O.TIE_AWAY)), "packed bytes vs oracle")
    if prec <= 4:  # mantissa magnitudes really are (prec-1)-bit integers
        m = mant.cpu()
        assert int(torch.maximum(m & 7, (m >> 4) & 7).max()) <= 2 ** (prec - 1) - 1


def test_packed_sbfp_reports_blocks_the_bytes_cannot_hold():
    """non-finite blocks, denormal blocks whose max / 7 underflows, and scalers above the byte's exponent range (bias > 7:
    the simulated scaler saturates at 2^8, the real E4M4 byte at 2^(15 - bias)) are counted, everything else is exact"""
    f = fmt_from("SBFP<XP[4,0](CSN)><FP[0|4|4,12](FN)>{16}")
    x = _rand((8, 256), 77, spread=0) * 2.0**-6
    x[1, 16:32] *= 2.0**14      # scaler exponent >= 4 > 15 - 12
    x[2, 3] = float("nan")
    x[3, 40] = float("inf")
    x[4, 64:80] = 0
    x[4, 70] = 1e-45            # max / 7 == 0: the cast passes the denormal through
    xd = x.to(DEV)
    mant, scal, bad = f.pack(xd, return_inexact=True)
    assert int(bad) == 4
    om, os_, obad = O.sbfp_pack(x.numpy(), 16, 4, O.TIE_AWAY, 4, 4, 12)
    assert obad == 4 and np.array_equal(scal.cpu().numpy(), os_)
    got = f.unpack(mant, scal).cpu()
    want = ops.cast_chain(xd, [f.stage()], -1).cpu()
    ok = torch.ones(8, 16, dtype=torch.bool)
    ok[1, 1] = ok[2, 0] = ok[3, 2] = ok[4, 4] = False
    ok = ok.repeat_interleave(16, 1)
    assert torch.equal(got.view(torch.int32)[ok], want.view(torch.int32)[ok])
    assert torch.equal(got[2, :16], torch.zeros(16)) and torch.equal(got[4, 64:80], torch.zeros(16))
    assert int(scal[1, 1]) == 255 and torch.isfinite(got).all()
    with pytest.raises(RuntimeError, match="unsupported"):
        fmt_from("SBFP<XP[4,0](CSN)><FP[0|4|4,7](_N)>{16}").pack(xd)       # subnormal-keeping scaler
    with pytest.raises(RuntimeError, match="unsupported"):
        fmt_from("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", tie="even").pack(xd)


def test_golden_mxfp():
    """MXFP: the reference's own MXFP.cast on CPU vs the CUDA path"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "mxfp_reference.npz"))
    for n in [str(t) for t in z["names"]]:
        i, sh, shp = n.split("|")
        shp = tuple(int(t) for t in shp.split(","))
        x = f32(z[f"{i}.x"]).reshape(shp)
        check(gpu_cast(torch.from_numpy(x).to(DEV), sh, -1), z[f"{i}.y"].reshape(shp), n)
        xb = torch.from_numpy(x).to(torch.bfloat16)
        want = O.cast(xb.float().numpy(), sh, -1)
        check(gpu_cast(xb.to(DEV), sh, -1), bits(want), n + " bf16")


@pytest.mark.parametrize("sh", ["MXFP8[E4M3]{32}", "MXFP8[E5M2]{32}", "MXFP6[E2M3]{32}", "MXFP6[E3M2]{64}", "MXFP4[E2M1]{32}", "MXFP8[E4M3]{128}"])
def test_mxfp_vs_oracle_at_scale(sh):
    """the branch-free element rounding (shared op sequence for subnormal and normal values) against the oracle on
    2^21 values whose in-block spread reaches deep into the element format's subnormals, plus exact rounding ties,
    for rows, strided-rows and cols layouts and 16-bit sources"""
    g = torch.Generator().manual_seed(97)
    x = torch.randn(512, 4096, generator=g) * torch.pow(2.0, torch.randint(-24, 3, (512, 4096), generator=g).float())
    x = x * torch.pow(2.0, torch.randint(-20, 20, (512, 1), generator=g).float())
    x.view(-1)[::5] = torch.round(x.view(-1)[::5] * 16) / 16  # coarse values: many exact ties
    x.view(-1)[3::17] = 0.0
    x.view(-1)[4::17] *= -1
    x[7] = 0.0  # an all-zero row: 0 / 0 -> NaN blocks, as in the reference
    check(gpu_cast(x.to(DEV), sh, -1), bits(O.cast(x.numpy(), sh, -1)), sh)
    for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        xh = x.to(dt)
        xh = torch.where(torch.isfinite(xh), xh, torch.zeros_like(xh))
        check(gpu_cast(xh.to(DEV), sh, -1), bits(O.cast(xh.float().numpy(), sh, -1)), f"{sh} {name}")
    xs = x.to(DEV)[:, :2048]  # strided rows
    check(gpu_cast(xs, sh, -1), bits(O.cast(x[:, :2048].contiguous().numpy(), sh, -1)), sh + " strided")
    xc = x[:256].to(DEV)  # blocks along dim 0
    check(gpu_cast(xc, sh, 0), bits(O.cast(x[:256].contiguous().numpy(), sh, 0)), sh + " cols")


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("sh", ["FP[1|4|3,7](_N)", "FP[1|5|2,15](_N)", "FP[1|2|3,1](_N)", "FP[1|3|2,3](_N)", "FP[1|2|1,1](_N)", "FP[1|4|5,7](_N)",
                                "FP[1|5|6,15](_N)", "FP[1|8|3,127](_N)", "MXFP8[E4M3]{32}", "MXFP8[E5M2]{32}", "MXFP6[E2M3]{32}", "MXFP6[E3M2]{64}",
                                "MXFP4[E2M1]{32}", "MXFP8[E4M3]{128}"])
def test_packed16_low_bit_float(dt, sh):
    """K_FLOAT (subnormals kept) / K_MXFP on a 16-bit tensor, same dtype out: the packed form (two elements per instruction,
    rounding in the source's own arithmetic, dmxq_stages.cuh sub16_pair) against the oracle -- values deep in the element
    format's subnormals, every rounding tie of the format, +-0, results that round to zero from either side, saturation, rows
    whose magnitudes leave no room for the rounding constant / Inf / NaN (literal path), power-of-two and zero block maxima"""
    tdt = getattr(torch, dt)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(384, 2048, generator=g) * torch.pow(2.0, torch.randint(-14, 3, (384, 2048), generator=g).float())
    x = x * torch.pow(2.0, torch.randint(-10, 10, (384, 1), generator=g).float())
    flat = x.view(-1)
    flat[::5] = torch.round(flat[::5] * 64) / 64          # coarse values: exact ties of 2..5-bit mantissas
    flat[1::11] = torch.round(flat[1::11] * 4) / 4
    flat[3::17] = 0.0
    flat[4::17] = -0.0
    flat[6::17] *= -1
    x[7] = 0.0                                              # all-zero blocks
    x[9, ::3] = 0.5                                         # power-of-two block maxima
    x[9, 1::3] = 0.25
    x[9, 2::3] = -0.5
    x[11] = x[11] * 2.0**9                                  # saturation
    if dt == "bfloat16":
        x[13] = x[13] * 2.0**105                            # exponent fields above the packed path's limit, still finite
        x[14] = x[14] * 2.0**-120                           # bf16 denormals
    else:
        x[13] = (x[13] * 2.0**12).clamp(-60000.0, 60000.0)
        x[14] = x[14] * 2.0**-14                            # fp16 denormals
    xh = x.to(tdt)
    assert bool(torch.isfinite(xh).all())
    xh[15, 5] = float("inf"); xh[15, 77] = float("-inf"); xh[16, 9] = float("nan")
    finite_rows = np.ones(384, dtype=bool)
    finite_rows[15:17] = False
    want32 = O.cast(xh.float().numpy(), sh, -1)
    y = ops.cast_chain(xh.to(DEV), [fmt_from(sh).stage()], -1)
    assert y.dtype == tdt
    got = y.cpu().view(torch.int16).numpy().view(np.uint16)
    want = torch.from_numpy(want32).to(tdt).view(torch.int16).numpy().view(np.uint16)
    # rows of finite values: the oracle, bit for bit (MX all-zero blocks are NaN = 0 / 0 on both sides)
    nan_w = np.isnan(want32)
    nan_g = torch.isnan(y.float()).cpu().numpy()
    assert np.array_equal(nan_w[finite_rows], nan_g[finite_rows]), f"{sh} {dt}: NaN positions differ"
    bad = (got != want) & ~nan_w
    bad[~finite_rows] = False
    assert not bad.any(), f"{sh} {dt}: {int(bad.sum())} mismatches, first at {np.argwhere(bad)[0]}: x={xh.float().numpy()[tuple(np.argwhere(bad)[0])]!r} got={got[bad][0]:#06x} want={want[bad][0]:#06x}"
    # every row, Inf / NaN included (where x86 and the GPU legitimately produce different NaNs, see special_block_masks): the
    # fp32-out kernels -- the unpacked path, pinned to the reference's own CUDA kernels elsewhere -- narrowed like CastTo does
    y32 = gpu_cast(xh.to(DEV), sh, -1)
    assert torch.equal(y32.to(tdt).view(torch.int16), y.view(torch.int16)), f"{sh} {dt}: packed path != unpacked path"
    b32 = (y32.cpu().numpy().view(np.uint32) != bits(want32)) & ~nan_w
    b32[~finite_rows] = False
    assert not b32.any()
    if dt == "bfloat16":  # the fp32 kernel on the unrounded values (incl. magnitudes whose rounding constant would overflow)
        w = O.cast(x.numpy(), sh, -1)
        yf = gpu_cast(x.to(DEV), sh, -1).cpu().numpy()
        assert np.array_equal(np.isnan(w), np.isnan(yf))
        assert not ((yf.view(np.uint32) != bits(w)) & ~np.isnan(w)).any(), f"{sh} fp32"


# ---- the TMA experiment (csrc/dmxq_tma.cu): not on the product path, but it ships in the library, so it is held to the same parity
@pytest.mark.parametrize("dt", ["float32", "bfloat16", "float16"])
@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_tma_tiled_cols_kernel_equals_production(dt, cfg):
    import ctypes as C

    fn = L.lib.dmxq_x_bfp_cols_tma
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    fn.restype = C.c_int
    tdt = getattr(torch, dt)
    for shape in ((5, 256, 64), (7, 200, 72), (3, 64, 8), (2, 130, 520)):
        x = _rand(shape, 7, spread=5 if dt != "float16" else 3, dtype=tdt).to(DEV)
        x[0, :, 1] = 0.0
        x[1, :, 2] = x[1, :, 2] * (2.0**-100 if dt != "float16" else 2.0**-12)  # denormal-range column: literal path
        x[1, 3, 4] = float("inf")
        for wl in (8, 4):
            want = ops.cast_chain(x, [fmt_from(f"BFP[{wl}|8]{{64}}(SN)").stage()], -2)
            got = torch.empty_like(x)
            assert fn(x.data_ptr(), got.data_ptr(), L.dtype_code(tdt), *shape, 64, wl, cfg, L.stream_ptr(x.device)) == 0
            torch.cuda.synchronize()
            it = torch.int32 if dt == "float32" else torch.int16
            assert torch.equal(got.view(it), want.view(it)), (dt, cfg, shape, wl)


# ---- stochastic rounding with in-kernel random words (dmxq_cast_chain_philox) -----------------------------------------
def test_philox_fill_equals_numpy_restatement():
    for n, seed, sid in ((1, 0, 0), (4, 0, 0), (1003, 0x123456789ABCDEF, 5), (1 << 20, 2**63 + 11, 2**40 + 3)):
        got = ops.philox_fill((n,), seed, sid).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, O.philox_words(n, seed, sid)), (n, seed, sid)
        gf = ops.philox_fill((n,), seed, sid, as_float=True).cpu().numpy()
        assert np.array_equal(gf.view(np.uint32), O.philox_unit_floats(n, seed, sid).view(np.uint32))


@pytest.mark.parametrize("dt", ["float32", "bfloat16"])
@pytest.mark.parametrize("fmt", ["BFP[8|8]{64}(SS)", "BFP[4|8]{16}(SS)", "FP[1|4|3,7](_S)", "FP[1|5|10,15](FS)", "XP[8,0](CSS)", "XP[8,+4](CSS)"])
def test_philox_in_kernel_equals_external_tensor(dt, fmt):
    """cast(philox = (seed, stream)) == cast(rand = philox_fill(shape, seed, stream)) bit for bit, for every layout: flat rows (words
    computed in registers, four elements per Philox call), strided rows, misaligned / column layouts (filled tensor behind the scenes).
    The external-tensor path is the one pinned to the oracle and to the reference's kernels (test_oracle_stochastic_same_random_tensor)."""
    tdt = getattr(torch, dt)
    seed, sid = 0xC0FFEE12345, 17
    st = [fmt_from(fmt).stage()]
    fixed = fmt.startswith("XP")
    g = torch.Generator().manual_seed(3)
    base = (torch.randn(48, 6, 512, generator=g) * torch.pow(2.0, torch.randint(-6, 7, (48, 6, 1), generator=g).float())).to(tdt).to(DEV)
    views = [("flat", base, -1), ("3-d rows, 2-d view", base[:, 0], -1), ("strided rows", base[:, :, 128:384], -1), ("blocks along dim 1", base, 1),
             ("odd length", base.reshape(-1)[:4093], -1), ("misaligned", base.reshape(-1)[3:4099], -1)]
    n0 = _lib_launches()
    for name, x, bd in views:
        r = ops.philox_fill(x.shape, seed, sid, as_float=fixed)
        want = ops.cast_chain(x, st, bd, rand=r)
        got = ops.cast_chain(x, st, bd, philox=(seed, sid))
        assert torch.equal(got.view(torch.int32 if dt == "float32" else torch.int16), want.view(torch.int32 if dt == "float32" else torch.int16)), (fmt, dt, name)
    # and it is random: another stream id moves results, the same one reproduces them
    a = ops.cast_chain(base, st, -1, philox=(seed, sid))
    b = ops.cast_chain(base, st, -1, philox=(seed, sid + 1))
    c = ops.cast_chain(base, st, -1, philox=(seed, sid))
    assert torch.equal(a, c)
    if not (dt == "bfloat16" and fmt.startswith("FP[1|5|10")):  # (FLOAT16 keeps every bf16 value: nothing to round)
        assert not torch.equal(a, b)


def test_stochastic_source_switch():
    """ops.stochastic_source("philox"): Format.cast with stochastic rounding draws no tensor; reproducible per (seed, call order)"""
    x = torch.randn(64, 256, device=DEV)
    f = fmt_from("BFP[8|8]{64}(SS)")
    try:
        ops.stochastic_source("philox", seed=42)
        a1, a2 = f.cast(x, -1), f.cast(x, -1)
        ops.stochastic_source("philox", seed=42)
        b1, b2 = f.cast(x, -1), f.cast(x, -1)
        assert torch.equal(a1, b1) and torch.equal(a2, b2) and not torch.equal(a1, a2)
        want = ops.cast_chain(x, [f.stage()], -1, rand=ops.philox_fill(x.shape, 42, 1), out_dtype=torch.float32)
        assert torch.equal(a1, want)
    finally:
        ops.stochastic_source("torch")


def _lib_launches():
    return L.lib.dmxq_launch_count()


def test_philox_flat_path_launches_one_kernel_and_is_unbiased():
    """the flat rows path computes the words in the cast kernel (ONE launch, no fill); stochastic rounding stays unbiased"""
    x = torch.full((1 << 14, 64), 0.3, device=DEV)
    x[:, 0] = 1.0  # block max 1.0: BFP[4|8] grid step 0.25 -> 0.3 rounds to 0.25 (p = 0.8) or 0.5 (p = 0.2)
    st = [fmt_from("BFP[4|8]{64}(SS)").stage()]
    n0 = _lib_launches()
    y = ops.cast_chain(x, st, -1, philox=(99, 0))
    assert _lib_launches() - n0 == 1
    v = y[:, 1:].reshape(-1)
    assert set(np.unique(v.cpu().numpy()).tolist()) == {0.25, 0.5}
    assert abs(float(v.mean()) - 0.3) < 2e-3


# ---- (f3) calibration histogram: dmxq_histc / HistogramObserver ---------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bins,lo,hi", [(2048, -3, 4), (1000, -7, 10), (7, -1, 1), (2048, 0, 0), (12288, -40, 41), (1, -2, 2)])
def test_histc_vs_oracle(dtype, bins, lo, hi):
    """torch.histc semantics bit for bit: oracle (pinned to torch's CPU histc) and torch's own CUDA histc."""
    x = (_rand((37, 1031), 71, spread=2)).to(dtype)  # 38147 elements: ragged vector tail
    x.view(-1)[5] = float("nan")
    x.view(-1)[6] = float("inf")
    x.view(-1)[7] = lo if lo != hi else 0.25  # left edge
    x.view(-1)[8] = hi if lo != hi else 0.5  # right edge -> last bin
    if lo == hi:  # "use the data's range" needs finite data
        x.view(-1)[5] = 0.0
        x.view(-1)[6] = 0.0
    got, mn, mx = ops.histc(x.to(DEV), bins, min=lo, max=hi, return_minmax=True)
    want = O.histc(x.float().numpy(), bins, lo, hi)
    assert np.array_equal(got.cpu().numpy(), want)
    if lo != hi:
        assert torch.isnan(mn) and torch.isnan(mx)
        assert torch.equal(got, torch.histc(x.float().to(DEV), bins, min=lo, max=hi))
    else:
        assert mn.item() == x.float().min().item() and mx.item() == x.float().max().item()
    assert torch.equal(ops.histc(x.to(DEV), bins, min=lo, max=hi), got)  # without the fused min/max


def test_histc_errors_and_degenerate():
    x = torch.randn(1000, device=DEV)
    with pytest.raises(RuntimeError, match="larger"):
        ops.histc(x, 16, min=2, max=1)
    with pytest.raises(RuntimeError, match="finite"):
        ops.histc(x, 16, min=0, max=float("inf"))
    with pytest.raises(RuntimeError, match="bins"):
        ops.histc(x, 100000, min=-1, max=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.histc(x.cpu(), 16, min=-1, max=1)
    c = torch.full((64,), 2.5, device=DEV)  # constant data and no range: [1.5, 3.5]
    assert torch.equal(ops.histc(c, 8), torch.histc(c, 8))
    assert ops.histc(torch.empty(0, device=DEV), 8, min=-1, max=1).sum().item() == 0


def test_histc_full_size():
    """2^28 elements: every in-range value lands in exactly one bin, and the result equals torch's CUDA histc."""
    n = 2**28
    x = torch.randn(n, device=DEV) * 3
    got, mn, mx = ops.histc(x, 2048, min=-7, max=9, return_minmax=True)
    assert got.double().sum().item() == ((x >= -7) & (x <= 9)).sum().item()
    assert torch.equal(got, torch.histc(x, 2048, min=-7, max=9))
    assert mn.item() == x.min().item() and mx.item() == x.max().item()
    xb = x.to(torch.bfloat16)
    assert torch.equal(ops.histc(xb, 2048, min=-7, max=9), torch.histc(xb.float(), 2048, min=-7, max=9))


@pytest.mark.parametrize("name", ["widening", "steady", "unit_interval", "one_sided", "int4_sym", "bins_1000", "constant_then_data"])
def test_histogram_observer_gpu(name):
    """the reference's golden calibration sequences through dmxq_histc / dmxq_minmax on the device"""
    from test_observer_cpu import check_sequence
    from dmx_compressor_b200.numerical import HistogramObserver

    obs = check_sequence(HistogramObserver, name, device=DEV, exact_search=False)
    if name == "steady":
        assert obs.stats == {"fused_steps": 3, "rebinned_steps": 0}


def test_castto_calibration_default_observer():
    """CastTo.enable_calibration() -> HistogramObserver (reference cast.py:308-340), then quantise with its qparams"""
    from dmx_compressor_b200.numerical import CastTo, HistogramObserver

    c = CastTo("XP[8,0](CSN)").to(DEV)
    c.enable_calibration(True)
    assert isinstance(c.activation_post_process, HistogramObserver)
    x = torch.randn(256, 512, device=DEV) * 4
    for _ in range(3):
        assert torch.equal(c(x), x)  # observing only
    c.enable_calibration(False)
    sc, zp = c.scale.item(), c.zero_point.item()
    assert 0.02 < sc < 0.2
    y = c(x)
    want = O.cast(x.cpu().numpy(), "XP[8,0](CSN)", -1, tie=O.TIE_AWAY, scale=np.array([sc], np.float32), zero_point=np.array([zp], np.float32))
    assert_bits_equal(bits(y.cpu().numpy()), bits(want))


def test_histc_division_free_bins_are_exact():
    """the kernel replaces (v - lo) * bins / (hi - lo) by a reciprocal + two FMA refinements: it must truncate like the
    IEEE division for values on and next to every bin edge, for awkward widths and bin counts (torch's CUDA histc divides)"""
    g = torch.Generator().manual_seed(5)
    for bins, lo, hi in [(2048, -7, 9), (2048, -3, 4), (1000, -7, 10), (12288, -41, 45), (7, 0, 3), (2047, -1, 2), (4096, -100, 27)]:
        k = torch.arange(0, bins + 1, dtype=torch.float64)
        edges = (lo + k * (hi - lo) / bins).float()
        near = torch.cat([edges, torch.nextafter(edges, torch.tensor(float("inf"))), torch.nextafter(edges, torch.tensor(float("-inf")))])
        x = torch.cat([near.repeat(40), torch.rand(1 << 20, generator=g) * (hi - lo) + lo]).to(DEV)
        assert torch.equal(ops.histc(x, bins, min=lo, max=hi), torch.histc(x, bins, min=lo, max=hi)), (bins, lo, hi)
        xb = x.to(torch.bfloat16)  # coarse values: very many sit exactly on a bin edge
        assert torch.equal(ops.histc(xb, bins, min=lo, max=hi), torch.histc(xb.float(), bins, min=lo, max=hi)), (bins, lo, hi)
    # a width whose significand is all ones, a huge and a tiny one: the IEEE-division path
    for lo, hi in [(0.0, float(np.float32(2.0) - np.float32(2.0**-23))), (-1e30, 1e30), (0.0, 1e-20)]:
        x = (torch.rand(1 << 18, generator=g) * (hi - lo) + lo).to(DEV)
        assert torch.equal(ops.histc(x, 100, min=lo, max=hi), torch.histc(x, 100, min=lo, max=hi))


# =============================================================================== many tensors, one launch
def test_cast_chain_multi_equals_per_tensor():
    """dmxq_cast_chain_multi == dmxq_cast_chain tensor by tensor, in far fewer launches: the whole-model weight-cast kinds,
    fp32 / bf16 / fp16, more tensors than one table holds, ragged sizes, an empty tensor, a layout the batched kernel does
    not take (row-strided view) mixed in"""
    g = torch.Generator(device=DEV).manual_seed(2024)
    F = lambda sh: fmt_from(sh).stage()
    chains = {
        "bfp16": [F("BFP[8|8]{64}(SN)")], "bfp12": [F("BFP[4|8]{64}(SN)")], "sbfp": [F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")],
        "24_bfp12": [ops.nm_stage(2, 4), F("BFP[4|8]{64}(SN)")], "24": [ops.nm_stage(2, 4)], "float16": [F("FP[1|5|10,15](FN)")],
        "sbfp_bfp": [F("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"), F("BFP[8|8]{64}(SN)")],  # runtime chain
        "24_torch_order": [ops.nm_stage(2, 4, L.NM_TORCH_CUDA), F("BFP[4|8]{64}(SN)")],
    }
    shapes = [(37, 256), (1, 64), (128, 1024), (5, 4096), (0, 64), (1000, 64), (3, 7, 128), (64, 64)] + [(2 + i, 128) for i in range(70)]
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        xs = [(torch.randn(s, device=DEV, generator=g) * 3).to(dt) for s in shapes]
        big = (torch.randn(40, 512, device=DEV, generator=g)).to(dt)
        xs.insert(3, big[:, :256])  # row-strided view: launched on its own inside the same call
        for name, st in chains.items():
            want = [ops.cast_chain(x, st, -1) for x in xs]
            n0 = L.launch_count()
            got = ops.cast_chain_multi(xs, st, -1)
            used = L.launch_count() - n0
            assert used <= 6, f"{name} {dt}: {used} launches for {len(xs)} tensors"
            for i, (a, b) in enumerate(zip(got, want)):
                assert a.shape == b.shape and torch.equal(a.view(torch.int16 if dt != torch.float32 else torch.int32),
                                                          b.view(torch.int16 if dt != torch.float32 else torch.int32)), f"{name} {dt} tensor {i}"
        # preallocated outputs
        outs = [torch.empty_like(x) for x in xs]
        r = ops.cast_chain_multi(xs, chains["bfp12"], -1, outs=outs)
        assert all(a is b for a, b in zip(r, outs))


def test_cast_chain_multi_amax_drives_sbfp_bias_on_device():
    """amax= : the SBFP scaler bias is derived inside the kernel from a device-resident tensor-wide amax and equals the host
    rule (parallel.sbfp_scaler_bias_from_amax) for every tensor -- amax at exact power-of-two boundaries of amax / 7, one ulp
    either side, tiny, huge, zero, inf"""
    from dmx_compressor_b200 import parallel as P

    g = torch.Generator(device=DEV).manual_seed(77)
    amaxs = [1.0, 0.02, 7.0, 7.0 * 2.0**-5, float(np.nextafter(np.float32(7.0 * 2.0**-5), np.float32(0))), float(np.nextafter(np.float32(7.0 * 2.0**3), np.float32(1e9))),
             3.5, 1e-30, 1e30, 0.0, float("inf"), 0.4375, 14.0 - 2.0**-20, 6.9999995]
    for dt in (torch.bfloat16, torch.float32):
        xs = [(torch.randn(16 + i, 256, device=DEV, generator=g) * (a if 0 < a < 1e20 else 1.0) / 4).to(dt) for i, a in enumerate(amaxs)]
        amax = torch.tensor(amaxs, dtype=torch.float32, device=DEV)
        sh0 = fmt_from("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}").stage()
        got = ops.cast_chain_multi(xs, [sh0], -1, amax=amax)
        for i, (x, a) in enumerate(zip(xs, amaxs)):
            b = P.sbfp_scaler_bias_from_amax(float(np.float32(a)))
            want = ops.cast_chain(x, [fmt_from(f"SBFP<XP[4,0](CSN)><FP[0|4|4,{b}](FN)>{{16}}").stage()], -1)
            assert torch.equal(got[i].view(torch.int16 if dt != torch.float32 else torch.int32),
                               want.view(torch.int16 if dt != torch.float32 else torch.int32)), f"tensor {i}: amax {a} -> bias {b}, {dt}"
    # shard_amax feeds it: whole tensors on one GPU
    shapes = {f"t{i}": (32 + 8 * i, 512) for i in range(6)}
    plan = P.plan_shards(shapes, 1)
    ws = [torch.randn(shapes[sh.name], device=DEV, generator=g).mul_(0.02 * (1 + i)).to(torch.bfloat16) for i, sh in enumerate(plan[0])]
    amax = P.shard_amax(plan, 0, ws)
    assert torch.equal(amax, torch.stack([w.float().abs().max() for w in ws]))
    got = ops.cast_chain_multi(ws, [sh0], -1, amax=amax)
    for w, y, a in zip(ws, got, amax.tolist()):
        b = P.sbfp_scaler_bias_from_amax(a)
        want = ops.cast_chain(w, [fmt_from(f"SBFP<XP[4,0](CSN)><FP[0|4|4,{b}](FN)>{{16}}").stage()], -1)
        assert torch.equal(y.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
def test_int8_affine_quotient_ties(dt):
    """calibrated INT8: x / scale + zp lands on (and one grid step either side of) the rounding ties k + 0.5, the clamp edges
    +-127.5 and beyond, for many scales -- per tensor (K_FIXED rows kernel), per row and per column (fixed_chan kernels).  On
    16-bit tensors the quotient is the two-operation form (div_by_recip16) and the rounding is clamp-first: both must equal
    the oracle's divide / roundf / clamp sequence bit for bit"""
    g = torch.Generator().manual_seed(4242)
    R, Cn = 256, 512
    sc_row = (torch.rand(R, generator=g) * 0.2 + 0.003) * torch.pow(2.0, torch.randint(-6, 7, (R,), generator=g).float())
    zp_row = torch.randint(-5, 6, (R,), generator=g).float()
    k = torch.randint(-140, 141, (R, Cn), generator=g).float() + 0.5
    x = (k - zp_row[:, None]) * sc_row[:, None]
    x = x.to(dt)  # onto the dtype's grid ...
    it = torch.int32 if dt == torch.float32 else torch.int16
    x = (x.view(it) + torch.randint(-1, 2, (R, Cn), generator=g, dtype=torch.int32).to(it)).view(dt)  # ... and its neighbours
    x[0, :8] = torch.tensor([0.0, -0.0, 1e-30, -1e-30, 3e4, -3e4, 1.0, -1.0]).to(dt)
    xn = x.float().numpy()

    def cmp(got, want, what):
        want = torch.from_numpy(want).to(dt)
        assert torch.equal(got.cpu().view(it), want.view(it)), what

    for fmt in ("XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,0](C_N)"):
        f = fmt_from(fmt)
        args = (f.precision, f.fraction, f.clamp, f.symmetric, "nearest")
        for r in (0, 7, 100):  # per tensor
            sc, zp = sc_row[r:r + 1], zp_row[r:r + 1]
            want = O.fixed_cast(xn, *args, tie=O.TIE_AWAY, scale=sc.numpy(), zero_point=zp.numpy())
            cmp(ops.fixed_qdq(x.to(DEV), *args, scale=sc.to(DEV), zero_point=zp.to(DEV)), want, f"{fmt} per tensor {dt} row-scale {r}")
        want = O.fixed_cast(xn, *args, tie=O.TIE_AWAY, scale=sc_row.numpy(), zero_point=zp_row.numpy(), ch_axis=0)
        cmp(ops.fixed_qdq(x.to(DEV), *args, scale=sc_row.to(DEV), zero_point=zp_row.to(DEV), ch_axis=0), want, f"{fmt} per row {dt}")
        xt = x.t().contiguous()
        want = O.fixed_cast(xt.float().numpy(), *args, tie=O.TIE_AWAY, scale=sc_row.numpy(), zero_point=zp_row.numpy(), ch_axis=1)
        cmp(ops.fixed_qdq(xt.to(DEV), *args, scale=sc_row.to(DEV), zero_point=zp_row.to(DEV), ch_axis=1), want, f"{fmt} per column {dt}")
        # group quantisation along the contiguous dim (groups of 16 columns)
        scg, zpg = sc_row[: Cn // 16], zp_row[: Cn // 16]
        want = O.fixed_cast(xn, *args, tie=O.TIE_AWAY, scale=scg.numpy(), zero_point=zpg.numpy(), ch_axis=1, group_size=16)
        cmp(ops.fixed_qdq(x.to(DEV), *args, scale=scg.to(DEV), zero_point=zpg.to(DEV), ch_axis=1, group_size=16), want, f"{fmt} groups {dt}")

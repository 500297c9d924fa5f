"""The drop-in, on the GPU, against the reference package itself.

The reference's own python (``oracle/_ref/pysrc``, staged verbatim by ``oracle/build_ref.py stage_python``) and its own
compiled CUDA kernels (``oracle/_ref/ref_quant_cuda.so``) are imported UNMODIFIED.  Every test runs the same reference
call twice on the same CUDA tensors:

  unpatched   the reference's real CUDA path (python split / cat loop + quant_cuda kernels + torch ops)
  patched     after ``dmx_compressor_b200.plugin.install()``  (libdmxq through the C ABI)

and requires bit-equal results, plus evidence that the patched run really launched libdmxq kernels and the unpatched run
launched none.  Covers: ``CastTo`` for every BASELINE format x fp32 / bf16 / fp16 x block_dim {-1, 1, -2} (incl. ragged K),
the STE backward in training mode, calibrated FixedPoint (per tensor / per channel / group), ``Sparsify`` with
``plastic`` / ``score_func``, the reference's ``dmx.compressor.nn`` modules configured with ``config_rules.BASIC``
(``DmxModule.forward``, S/modeling/nn/core.py:215-264), ``weight_hypernet`` / ``fold_weight_and_bias``
(core.py:146-213) and ``DmxModel.from_torch`` + BASIC (S/modeling/model.py:575-645).
"""
import os
import sys

import pytest
import torch

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import load_reference  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not load_reference.available(), reason="reference python not staged (oracle/_ref/pysrc)")]

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    pkg = load_reference.load_full()
    assert "quant_cuda" in repr(sys.modules["dmx.compressor.quant.quant_function"].quant_cuda), \
        "the reference must run its own CUDA extension when unpatched"
    return pkg


@pytest.fixture()
def plugin():
    from dmx_compressor_b200 import plugin as P

    P.uninstall()
    yield P
    P.uninstall()


def launches():
    from dmx_compressor_b200 import _lib

    return _lib.launch_count()


def both(plugin, fn, expect_launches=True, **install_kw):
    """fn() unpatched, then patched -> (unpatched, patched); asserts who launched what"""
    plugin.uninstall()
    n0 = launches()
    a = fn()
    torch.cuda.synchronize()
    assert launches() == n0, "unpatched reference run launched libdmxq kernels"
    plugin.install("dmx.compressor", **install_kw)
    try:
        n0 = launches()
        b = fn()
        torch.cuda.synchronize()
        if expect_launches:
            assert launches() > n0, "patched run did not reach libdmxq"
    finally:
        plugin.uninstall()
    return a, b


def same_bits(a, b, what=""):
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a, b = a.detach().contiguous(), b.detach().contiguous()
    it = {4: torch.int32, 2: torch.int16}[a.element_size()]
    ne = a.view(it) != b.view(it)
    if ne.any():
        idx = ne.nonzero()[:6].tolist()
        raise AssertionError(f"{what}: {int(ne.sum())}/{ne.numel()} elements differ, e.g. " +
                             ", ".join(f"{tuple(i)}: {a[tuple(i)].item()!r} vs {b[tuple(i)].item()!r}" for i in idx))


def make_x(shape, dtype, seed=0, specials=True):
    """finite test data: per-row scales 2^-12..2^12 (2^-6..2^5 for fp16, whose range ends at 65504), zeros of both signs,
    denormals and a few outliers.  Non-finite values are left to tests/test_parity_gpu.py (kernel against kernel): through
    the python layers the reference's behaviour on Inf / NaN blocks depends on tensor-wide state (e.g.
    make_mantissa_asymmetric rewrites every block of a chunk through int() once ANY block holds an edge mantissa,
    S/numerical/format.py:349-372) and is not a per-block contract."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(shape, device=DEV, generator=g)
    lo, hi = (-6, 6) if dtype == torch.float16 else (-12, 13)
    x = x * torch.pow(2.0, torch.randint(lo, hi, shape[:-1] + (1,), device=DEV, generator=g).float())
    if specials:
        f = x.view(-1)
        f[::97] = 0.0
        f[5::131] = -0.0
        f[7::211] = 1e-41  # denormals
        f[11::223] = -3e-39
        f[13::301] *= 64.0
    return x.to(dtype)


FORMATS = [
    "BFP[8|8]{64}(SN)", "BFP[4|8]{64}(SN)", "BFP[8|8]{128}(SN)", "BFP[6|8]{16}(SN)", "BFP[16|8]{32}(SN)",
    "BFP[8|8]{64}(_N)", "BFP[4|8]{128}(_N)", "BFP[8|8]{64}(SU)", "BFP[8|8]{64}(SD)", "BFP[24|8]{1}(SN)",
    "FP[1|5|10,15](FN)", "FP[1|8|7,127](FN)", "FP[1|4|3,7](_N)", "FP[1|5|2,15](_N)", "FP[0|4|4,7](FN)",
    "XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,4](CUN)", "XP[8,0](CSU)",
    "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "SBFP<XP[4,0](CSN)><FP[0|4|4,5](FN)>{16}",
    "MXINT8{64}", "MXINT4{32}",
]
DTYPES = [torch.float32, torch.bfloat16, torch.float16]


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: str(d).split(".")[-1])
@pytest.mark.parametrize("sh", FORMATS)
def test_castto_patched_equals_reference_cuda(ref, plugin, sh, dtype):
    """ref.CastTo(fmt, block_dim)(x_cuda): the reference's CUDA path vs the plugin, numerical/cast.py:261-306"""
    num = ref.numerical
    x4 = make_x((2, 128, 64, 192), dtype, seed=1)
    for bd in (-1, 1, -2):
        c = num.CastTo(sh, block_dim=bd).to(DEV).eval()
        a, b = both(plugin, lambda: c(x4))
        same_bits(a, b, f"{sh} block_dim={bd} {dtype}")
        assert a.dtype == dtype
    # ragged K (last block short), 2-D and 3-D, and a transposed view as the input
    xr = make_x((37, 100), dtype, seed=2)
    c = num.CastTo(sh).to(DEV).eval()
    if not sh.startswith("SBFP"):  # torch.split gives a short last chunk; SBFP's reference path handles it too, but
        a, b = both(plugin, lambda: c(xr))  # K % 16 != 0 leaves a 4-wide block: keep SBFP on whole blocks below
        same_bits(a, b, f"{sh} ragged {dtype}")
    xt = make_x((6, 96, 256), dtype, seed=3).transpose(-2, -1)
    a, b = both(plugin, lambda: c(xt))
    same_bits(a, b, f"{sh} transposed view {dtype}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: str(d).split(".")[-1])
@pytest.mark.parametrize("sh", ["BFP[8|8]{64}(SN)", "FP[1|5|10,15](FN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "XP[8,0](CSN)"])
def test_castto_training_ste_backward(ref, plugin, sh, dtype):
    """training mode, autograd on: forward bits and the straight-through gradient (CastToFormat.backward, cast.py:29-32)"""
    num = ref.numerical
    c = num.CastTo(sh).to(DEV).train()
    x0 = make_x((16, 256), dtype, seed=4)
    g = make_x((16, 256), dtype, seed=5, specials=False)

    def run():
        x = x0.clone().requires_grad_(True)
        y = c(x)
        y.backward(g)
        return y.detach(), x.grad

    (ya, ga), (yb, gb) = both(plugin, run)
    same_bits(ya, yb, f"{sh} forward")
    same_bits(ga, gb, f"{sh} grad")
    if not sh.startswith("XP"):
        same_bits(gb, g, f"{sh} STE")


def test_castto_without_fusion_goes_through_format_cast(ref, plugin):
    """install(fuse_castto=False): only Format.cast is patched; CastTo.forward / CastToFormat stay the reference's"""
    num = ref.numerical
    for sh in ("BFP[8|8]{64}(SN)", "FP[1|5|10,15](FN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "XP[8,0](CSN)"):
        for dtype in DTYPES:
            c = num.CastTo(sh).to(DEV).eval()
            x = make_x((8, 64, 128), dtype, seed=6)
            a, b = both(plugin, lambda: c(x), fuse_castto=False)
            same_bits(a, b, f"{sh} {dtype} unfused")


def test_format_cast_direct(ref, plugin):
    """Format.cast(x, block_dim) itself (the plugin point of SURVEY section 8b): fp32 result of the input's shape"""
    num = ref.numerical
    for sh in FORMATS:
        f = num.Format.from_shorthand(sh)
        for dtype in (torch.float32, torch.bfloat16):
            x = make_x((4, 64, 128), dtype, seed=7)
            if dtype != torch.float32 and sh.startswith(("FP", "XP", "BFP[24|8]{1}")):
                continue  # the reference's elementwise kernels only take fp32 ("expected scalar type Float")
            for bd in ((-1, 1) if not sh.startswith(("FP", "XP")) else (-1,)):
                a, b = both(plugin, lambda: f.cast(x, bd))
                assert b.shape == x.shape
                same_bits(a.to(torch.float32), b.to(torch.float32), f"{sh} cast {dtype} bd={bd}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["float32", "bfloat16"])
def test_fixedpoint_calibrated(ref, plugin, dtype):
    """CastTo with a MinMaxObserver: observer statistics -> qparams -> affine wrap (cast.py:179-237, 279-296),
    per tensor, per channel and with group_size"""
    num = ref.numerical
    x = make_x((64, 128), dtype, seed=8, specials=False)

    from dmx.compressor.numerical.observer import MinMaxObserver

    def run(qscheme, ch_axis=None, group_size=None):
        def f():
            kw = dict(qscheme=qscheme)
            if ch_axis is not None:
                kw["ch_axis"] = ch_axis
            c = num.CastTo("XP[8,0](CSN)", observer=MinMaxObserver, group_size=group_size, **kw).to(DEV)
            c.enable_calibration(True, MinMaxObserver, qscheme_to_overload=qscheme, group_size=group_size, ch_axis=ch_axis)
            c.to(DEV)
            c(x)
            c.enable_calibration(False)
            c.eval()
            return c(x), c.scale.clone().float(), c.zero_point.clone().float()
        return f

    cases = [("per_tensor_symmetric", run(torch.per_tensor_symmetric)), ("per_tensor_affine", run(torch.per_tensor_affine)),
             ("per_channel_symmetric_0", run(torch.per_channel_symmetric, 0)), ("per_channel_affine_1", run(torch.per_channel_affine, 1)),
             ("group_16_axis_1", run(torch.per_tensor_symmetric, 1, 16)), ("group_2_axis_0", run(torch.per_tensor_symmetric, 0, 2))]
    for name, f in cases:
        try:
            a = f()
        except Exception as e:  # a configuration the reference itself rejects is not a parity case
            pytest.skip(f"reference rejects {name}: {e!r}")
        (ya, sa, za), (yb, sb, zb) = both(plugin, f)
        same_bits(sa, sb, f"{name} scale")
        same_bits(za, zb, f"{name} zero_point")
        same_bits(ya, yb, f"{name} {dtype}")


def _distinct_scores(shape, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    n = 1
    for s in shape:
        n *= s
    return (torch.randperm(n, device=DEV, generator=g).float() / n).reshape(shape)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: str(d).split(".")[-1])
@pytest.mark.parametrize("sp", ["BTOPK{2:4,-1}(U)", "BTOPK{4:8,-1}(U)", "BTOPK{2:8,0}(U)", "BTOPK{1:4,1}(U)", "BTOPK{2:4,-1}(M)"])
def test_sparsify_patched_equals_reference_cuda(ref, plugin, sp, dtype):
    """Sparsify.forward (sparse.py:287-301): plastic with score_func (first forward), then the stored score; y and .mask.
    The |w| scores of 16-bit weights are full of ties: the plugin's default tie order is torch's CUDA argsort order."""
    S = ref.sparse
    shape = (64, 32, 16)
    x = make_x(shape, dtype, seed=9, specials=False)

    def run():
        s = S.Sparsify(shape, sp).to(DEV).eval()
        with torch.no_grad():
            s.score.copy_(_distinct_scores(shape, 10))
        s.configure(score_func=lambda score, w: w.abs().float())
        assert s.plastic
        y1 = s(x)
        m1 = s.mask.clone()
        assert not s.plastic
        y2 = s(x)  # second forward: the learnable score tensor decides
        return y1, m1, y2, s.mask.clone()

    a, b = both(plugin, run)
    for u, v, what in zip(a, b, ("y plastic", "mask plastic", "y score", "mask score")):
        same_bits(u, v, f"{sp} {dtype} {what}")


def test_sparsify_tie_heavy_and_stable_option(ref, plugin):
    """scores drawn from a handful of values (most groups tied): default install == the unpatched reference on CUDA;
    install(tie_order="stable") == the reference's CPU result for the same tensors"""
    S = ref.sparse
    for sp, shape in (("BTOPK{2:4,-1}(U)", (128, 64)), ("BTOPK{4:8,-1}(U)", (128, 64)), ("BTOPK{2:8,0}(U)", (64, 48)),
                      ("BTOPK{8:16,-1}(U)", (32, 64)), ("BTOPK{16:32,-1}(U)", (32, 64)), ("BTOPK{3:6,-1}(U)", (32, 36))):
        g = torch.Generator(device=DEV).manual_seed(31)
        x = (torch.randint(-3, 4, shape, device=DEV, generator=g).float() / 2)
        sc = (torch.randint(0, 3, shape, device=DEV, generator=g).float())

        def run(dev=DEV):
            s = S.Sparsify(shape, sp).to(dev).eval()
            with torch.no_grad():
                s.score.copy_(sc.to(dev))
            s.configure(score_func=lambda score, w: w.abs())
            y1 = s(x.to(dev))
            y2 = s(x.to(dev))
            return y1, s.mask.clone(), y2

        a, b = both(plugin, run)
        for u, v, what in zip(a, b, ("y plastic", "mask", "y score")):
            same_bits(u, v, f"{sp} {what}")
        if sp.startswith("BTOPK{16:32"):
            continue  # torch's CPU sort of 32-key rows is not the stable order either (unspecified; stable is what we document)
        cpu = run("cpu")
        _, c = both(plugin, run, tie_order="stable")
        for u, v, what in zip(cpu, c, ("y plastic", "mask", "y score")):
            same_bits(u.to(DEV), v, f"{sp} stable == reference on CPU: {what}")


def test_sparsify_training_mode_keeps_reference_autograd(ref, plugin):
    """training + grad: the plugin leaves the reference's own autograd path in place (mask gradients, STE)"""
    S = ref.sparse
    shape = (32, 64)
    x0 = make_x(shape, torch.float32, seed=11, specials=False)

    def run():
        s = S.Sparsify(shape, "BTOPK{2:4,-1}(M)", backward_mode="joint").to(DEV).train()
        with torch.no_grad():
            s.score.copy_(_distinct_scores(shape, 12))
        x = x0.clone().requires_grad_(True)
        y = s(x)
        y.sum().backward()
        return y.detach(), x.grad, s.score.grad

    a, b = both(plugin, run, expect_launches=False)
    for u, v, what in zip(a, b, ("y", "x.grad", "score.grad")):
        same_bits(u, v, what)


def _basic_config(ref, module):
    for rule in ref.config_rules.BASIC:
        if isinstance(module, rule.module_types):
            return rule.module_config
    raise KeyError(type(module))


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: str(d).split(".")[-1])
def test_reference_nn_modules_basic(ref, plugin, dtype):
    """the reference's own dmx.compressor.nn modules, configured with config_rules.BASIC's module configs
    (S/__init__.py:306-469), forward on CUDA: DmxModule.forward core.py:215-264"""
    nn = ref.nn
    torch.manual_seed(0)
    B, S, D, H = 2, 128, 256, 4
    x = make_x((B, S, D), dtype, seed=13, specials=False)
    q = make_x((B, H, S, 64), dtype, seed=14, specials=False)
    k = make_x((B, H, S, 64), dtype, seed=15, specials=False)

    mods = {
        "Linear": (nn.Linear(D, 512), lambda m: m(x)),
        "LinearNoBias": (nn.Linear(D, 128, bias=False), lambda m: m(x)),
        "ResAdd": (nn.ResAdd(), lambda m: m(x, x.flip(0))),
        "ActActMatMul_qkT": (nn.ActActMatMul(), lambda m: m(q, k.transpose(-2, -1))),
        "ActActMatMul_pv": (nn.ActActMatMul(), lambda m: m(torch.softmax(q @ k.transpose(-2, -1) / 8, -1), k)),
        "Softmax": (nn.Softmax(dim=-1), lambda m: m(q @ k.transpose(-2, -1))),
        "LayerNorm": (nn.LayerNorm(D), lambda m: m(x)),
        "ReLU": (nn.ReLU(), lambda m: m(x)),
        "GELU": (nn.GELU(), lambda m: m(x)),
        "Embedding": (nn.Embedding(1000, D), lambda m: m(torch.arange(0, 512, device=DEV).view(2, 256))),
        "Conv2d": (nn.Conv2d(64, 32, 3, padding=1), lambda m: m(make_x((2, 64, 16, 16), dtype, seed=16, specials=False))),
    }
    for name, (m, call) in mods.items():
        m = m.to(DEV).to(dtype).eval()
        m.configure(_basic_config(ref, m))
        with torch.no_grad():
            a, b = both(plugin, lambda: call(m))
        same_bits(a, b, f"{name} {dtype}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["float32", "bfloat16"])
def test_weight_hypernet_and_fold(ref, plugin, dtype):
    """DmxModule._weight = sparsify -> storage cast -> weight cast (core.py:178-205), and fold_weight_and_bias (:146-176)"""
    nn = ref.nn
    torch.manual_seed(1)

    def build():
        torch.manual_seed(1)
        m = nn.Linear(256, 128).to(DEV).to(dtype).eval()
        with torch.no_grad():
            m.weight.copy_(make_x((128, 256), torch.float32, seed=17, specials=False).to(dtype))
        m.configure(dict(input_formats=["BFP[8|8]{64}(SN)"], weight_format="BFP[4|8]{64}(SN)",
                         weight_storage_format="SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}",
                         weight_sparseness="BTOPK{2:4,-1}(U)", bias_format="BFP[24|8]{1}(SN)", output_formats=["FP[1|5|10,15](FN)"]))
        m.weight_sparsifier.configure(score_func=lambda s, w: w.abs().float())
        return m

    x = make_x((4, 256), dtype, seed=18, specials=False)

    def run():
        m = build()
        with torch.no_grad():
            w = m._weight.clone()
            y = m(x)
            m.weight_sparsifier.configure(score_func=lambda s, w: w.abs().float())
            m.fold_weight_and_bias()
            return w, y, m.weight.data.clone(), m.bias.data.clone(), m(x)

    a, b = both(plugin, run)
    for u, v, what in zip(a, b, ("_weight", "forward", "folded weight", "folded bias", "forward after fold")):
        same_bits(u, v, f"{what} {dtype}")


class _MLP(torch.nn.Module):
    def __init__(self, d=128):
        super().__init__()
        self.fc1 = torch.nn.Linear(d, 4 * d)
        self.act = torch.nn.ReLU()
        self.fc2 = torch.nn.Linear(4 * d, d)
        self.ln = torch.nn.LayerNorm(d)

    def forward(self, x):
        return self.ln(x + self.fc2(self.act(self.fc1(x))))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["float32", "bfloat16"])
def test_dmxmodel_from_torch_basic(ref, plugin, dtype):
    """DmxModel.from_torch + config_rules.BASIC (model.py:575-645; the north star's user-facing flow) on CUDA"""
    torch.manual_seed(2)
    net = _MLP().to(DEV).to(dtype).eval()
    x = make_x((8, 16, 128), dtype, seed=19, specials=False)
    model = ref.DmxModel.from_torch(net)
    with torch.no_grad():
        want_plain = net(x)
        y0 = model(x)  # first call traces; BASELINE mode == the torch model
        same_bits(y0, want_plain, "BASELINE mode")
        model.to_basic_mode()
        n_dmx = len(list(model.named_dmx_modules()))
        assert n_dmx >= 4
        a, b = both(plugin, lambda: model(x))
    same_bits(a, b, f"DmxModel BASIC {dtype}")
    assert not torch.equal(a, want_plain)


def test_l1_functions_routed(ref, plugin):
    """dmx.compressor.quant.{fixed_point,block,float}_quantize on CUDA tensors: reference quant_cuda vs the L1 mirrors
    (nearest: bit-equal; stochastic draws its own random tensor on both sides, so only the error bound is checked)"""
    Q = sys.modules["dmx.compressor.quant"]
    x = make_x((64, 256), torch.float32, seed=20)
    for fn in (lambda: Q.block_quantize(x, wl=8, dim=0, rounding="nearest"),
               lambda: Q.block_quantize(x, wl=8, dim=-1, rounding="nearest"),
               lambda: Q.block_quantize(x, wl=4, dim=1, rounding="nearest"),
               lambda: Q.float_quantize(x, 5, 10, rounding="nearest"),
               lambda: Q.float_quantize(x, 4, 3, bias=7, flush_subnormal=False, rounding="nearest"),
               lambda: Q.fixed_point_quantize(x, 8, 4, rounding="nearest"),
               lambda: Q.fixed_point_quantize(x, 8, 0, clamp=True, symmetric=True, rounding="nearest")):
        a, b = both(plugin, fn)
        same_bits(a, b, "L1")


def test_histogram_observer_on_gpu(ref, plugin):
    """HistogramObserver.forward (observer.py:454-499) on CUDA tensors: histogram / min / max state and qparams"""
    from dmx.compressor.numerical.observer import HistogramObserver

    num = ref.numerical
    xs = [make_x((256, 512), torch.float32, seed=21 + i, specials=False) * (1 + i) for i in range(3)]

    def run():
        o = HistogramObserver(bins=2048, dtype=num.Format.from_shorthand("XP[8,0](CSN)"), qscheme=torch.per_tensor_symmetric).to(DEV)
        for x in xs:
            o(x)
        sc, zp = o.calculate_qparams()
        return o.histogram.clone(), o.min_val.reshape(1).clone(), o.max_val.reshape(1).clone(), sc.float().reshape(-1), zp.float().reshape(-1)

    a, b = both(plugin, run)
    for u, v, what in zip(a, b, ("histogram", "min", "max", "scale", "zero_point")):
        same_bits(u, v, what)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["float32", "bfloat16"])
def test_reference_linear_with_smoothquant(ref, plugin, dtype):
    """the reference's own Linear with an enabled ActivationWeightSmoothQuant (core.py:184-196, 227-230; smoothquant.py:253-283, 475-497):
    unpatched == plugin drop-in == plugin with elision, where `input / scale` + input cast and `weight * scale` + weight cast are one
    kernel each (DMXQ_STAGE_SCALE)"""
    from dmx_compressor_b200 import elide

    torch.manual_seed(4)
    lin = ref.nn.Linear(256, 96).to(device=DEV, dtype=dtype)
    lin.configure(_basic_config(ref, lin))
    x = (torch.randn(2, 33, 256, device=DEV) * torch.rand(256, device=DEV).mul(6).exp2()).to(dtype)
    sq = lin.smoothquant
    assert sq is not None
    with torch.no_grad():
        sq.calibrating = True
        lin(x)
        sq.calibrating = False
        sq.enable()
        assert sq.scale.numel() == 256
        a, b = both(plugin, lambda: lin(x))
        same_bits(a, b, f"smoothquant drop-in {dtype}")
        plugin.install("dmx.compressor", elide=True)
        try:
            n0 = launches()
            with elide.enabled():
                c = elide.materialise(lin(x))
            n1 = launches()
        finally:
            plugin.uninstall()
    same_bits(a, c, f"smoothquant elided {dtype}")
    assert n1 - n0 <= 5, n1 - n0  # input chain, weight chain, bias cast, output cast (+ one the reference's forward adds); no separate scale passes


def _ref_opt_stack(ref, cfg, dtype):
    """an OPT-shaped decoder assembled from the REFERENCE's own dmx.compressor.nn modules (the graph DmxModel.from_torch would
    produce for OPTForCausalLM), every module configured by the reference's config_rules.BASIC"""
    from dmx_compressor_b200 import opt

    torch.manual_seed(0)
    net = opt.OPTStack(cfg, mods=ref.nn).to(DEV).to(dtype).eval()
    for m in net.modules():
        if isinstance(m, ref.nn.DmxModule):
            for rule in ref.config_rules.BASIC:
                if isinstance(m, rule.module_types):
                    m.configure(rule.module_config)
                    break
    return net


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["float32", "bfloat16"])
def test_reference_opt_stack_dropin_and_elided(ref, plugin, dtype):
    """the reference's own module stack in BASIC mode: unpatched (its CUDA path) == plugin drop-in == plugin with elision
    (install(elide=True) + elide.enabled()), logits bit for bit; the elided run launches far fewer kernels"""
    from dmx_compressor_b200 import elide

    cfg = dict(vocab_size=512, max_position_embeddings=128, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2, dropout=0.0)
    net = _ref_opt_stack(ref, cfg, dtype)
    ids = torch.randint(0, 512, (2, 64), device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    with torch.no_grad():
        a, b = both(plugin, lambda: net(ids))
        same_bits(a, b, f"drop-in {dtype}")
        plugin.install("dmx.compressor", elide=True)
        try:
            n0 = launches()
            c0 = net(ids)  # installed with elide=True but outside the context: plain drop-in
            n1 = launches()
            with elide.enabled():
                c1 = elide.materialise(net(ids))
                n2 = launches()
                c2 = elide.materialise(net(ids))  # second forward: weights come from the cache
                n3 = launches()
        finally:
            plugin.uninstall()
    same_bits(a, c0, f"elide installed, context off {dtype}")
    same_bits(a, c1, f"elided {dtype}")
    same_bits(a, c2, f"elided, cached weights {dtype}")
    assert (n3 - n2) <= (n2 - n1) < (n1 - n0), (n1 - n0, n2 - n1, n3 - n2)

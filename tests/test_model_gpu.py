"""GPU tests of the callers of the path: the DmxModule mirror, the BASIC
rule set, weight folding, calibration, cast elision and the OPT-125m-shaped stack.  They read
like the reference's own tests (tests/test_flexible_quant.py, test_fold_weights_and_biases.py,
test_group_quant.py, test_sparse.py) and hold the same bit-equality bar."""
import os
import sys

import numpy as np
import pytest
import torch

from util import ROOT, bits

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from dmx_compressor_b200 import _lib, elide, opt
    from dmx_compressor_b200 import nn as dmxnn
    from dmx_compressor_b200.numerical import CastTo, MinMaxObserver
    from dmx_compressor_b200.sparse import Sparsify, abs_score

    fmt = dmxnn.format


def test_linear_equals_hand_composed_casts():
    """reference tests/test_flexible_quant.py:28-86: a configured module == the CastTo chain, bit for bit"""
    torch.manual_seed(0)
    lin = dmxnn.Linear(256, 128).to(DEV)
    lin.configure(dict(input_formats=[fmt.BFP16A_64], weight_format=fmt.BFP16_64, bias_format=fmt.BFP32_1, output_formats=[fmt.FLOAT16]))
    x = torch.randn(4, 17, 256, device=DEV)
    y = lin(x)
    xi = CastTo(fmt.BFP16A_64, block_dim=-1).to(DEV)(x)
    w = CastTo(fmt.BFP16_64, block_dim=-1).to(DEV)(lin.weight)
    b = CastTo(fmt.BFP32_1).to(DEV)(lin.bias)
    want = CastTo(fmt.FLOAT16).to(DEV)(torch.nn.functional.linear(xi, w, b))
    assert torch.equal(y, want)
    # and the casts inside equal the oracle
    assert (bits(xi.cpu().numpy()) == bits(O.cast(x.cpu().numpy(), "BFP[8|8]{64}(_N)"))).all()
    assert (bits(w.detach().cpu().numpy()) == bits(O.cast(lin.weight.detach().cpu().numpy(), "BFP[8|8]{64}(SN)"))).all()
    assert (bits(b.detach().cpu().numpy()) == bits(O.cast(lin.bias.detach().cpu().numpy(), "BFP[24|8]{1}(SN)"))).all()


def test_int8_per_tensor_cast_module():
    """reference tests/test_flexible_quant.py INT8 case: default scale 1 / zero-point 0 affine wrap"""
    x = torch.randn(8, 64, device=DEV) * 40
    y = CastTo(fmt.INT8).to(DEV)(x)
    want = O.cast(x.cpu().numpy(), "XP[8,0](CSN)", tie=O.TIE_AWAY, scale=[1.0], zero_point=[0.0])
    assert (bits(y.cpu().numpy()) == bits(want)).all()


def test_group_quant_kat_with_minmax_calibration():
    """reference tests/test_group_quant.py:49-63 (INT4, groups of 2 rows, symmetric MinMax)"""
    cast = CastTo(format=fmt.INT4, observer=MinMaxObserver, group_size=2, qscheme=torch.per_tensor_symmetric, ch_axis=0).to(DEV)
    cast.enable_observer()
    x = torch.tensor([[0, 1], [3, 7], [5.1, 8], [10, 14], [0.1, 0.7]], device=DEV)
    y = torch.tensor([[0, 1], [3, 7], [6, 8], [10, 14], [0.1, 0.7]], device=DEV)
    assert torch.allclose(cast(x), y, rtol=0.0, atol=1e-6)


def test_per_channel_calibration_equals_per_tensor_on_single_channel():
    """reference tests/test_group_quant.py:144-369 style equivalence: per-channel on a 1-channel view"""
    x = torch.randn(1, 4096, device=DEV) * 3
    a = CastTo(format=fmt.INT8, observer=MinMaxObserver, qscheme=torch.per_tensor_symmetric).to(DEV)
    b = CastTo(format=fmt.INT8, observer=MinMaxObserver, qscheme=torch.per_channel_symmetric, ch_axis=0).to(DEV)
    a.enable_observer(); b.enable_observer()
    assert torch.equal(a(x), b(x))
    assert torch.equal(a.scale.view(-1), b.scale.view(-1))


def test_sparsify_gradients_and_mask():
    """reference tests/test_sparse.py:12-56: gradient routing for the backward modes"""
    for mode, wg, mg in (("STE", True, False), ("supermask", False, True), ("joint", True, True)):
        sp = Sparsify((64, 64), "BTOPK{4:8,-1}(U)", backward_mode=mode).to(DEV).train()
        x = torch.randn(64, 64, device=DEV, requires_grad=True)
        y = sp(x)
        y.sum().backward()
        assert (x.grad is not None) == wg
        assert (sp.score.grad is not None) == mg
        assert sp.mask.sum().item() == 64 * 64 // 2
        if wg:
            assert torch.equal(x.grad, sp.mask)
    sp = Sparsify((8, 16), "BTOPK{2:4,-1}(U)").to(DEV).eval()
    sp.configure(score_func=abs_score)
    x = torch.randn(8, 16, device=DEV)
    want = O.nm_prune(x.cpu().numpy(), 2, 4)
    assert (bits(sp(x).cpu().numpy()) == bits(want)).all()


def test_fold_weights_and_biases_is_a_fixed_point():
    """reference tests/test_fold_weights_and_biases.py:139-148: folding leaves the output bit-identical"""
    torch.manual_seed(1)
    model = torch.nn.Sequential(dmxnn.Linear(256, 512), dmxnn.ReLU(), dmxnn.Linear(512, 64)).to(DEV).eval()
    dmxnn.to_basic_mode(model)
    for m in (model[0], model[2]):
        m.configure(dict(weight_format=fmt.BFP12_128, weight_sparseness="BTOPK{2:4,-1}(U)", weight_score_func=abs_score))
    x = torch.randn(32, 256, device=DEV)
    with torch.no_grad():
        m0 = model[0]
        m0.weight_sparsifier.plastic = True
        before_w = m0._weight.clone()
        m0.weight_sparsifier.plastic = True
        want = O.cast(O.nm_prune(m0.weight.detach().cpu().numpy(), 2, 4), "BFP[4|8]{128}(SN)")
        assert (bits(before_w.cpu().numpy()) == bits(want)).all()
        for m in (model[0], model[2]):
            m.weight_sparsifier.plastic = True
        y0 = model(x)
        for m in (model[0], model[2]):
            m.weight_sparsifier.plastic = True
        dmxnn.fold_weights_and_biases(model)
        y1 = model(x)
    assert torch.equal(y0, y1)
    assert repr(model[0].weight_format) == "SAME" and repr(model[0].weight_sparseness) == "DENSE"


TINY = dict(vocab_size=512, max_position_embeddings=128, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_opt_stack_basic_mode_elision_is_value_identical(dt):
    q, p = opt.build_pair(TINY, device=DEV, dtype=dt)
    ids = torch.randint(0, 512, (2, 96), device=DEV)
    with torch.no_grad():
        n0 = _lib.launch_count()
        y1 = q(ids)
        n1 = _lib.launch_count()
        with elide.enabled():
            y2 = q(ids)
            n2 = _lib.launch_count()
            y3 = q(ids)  # second pass: weights come from the cache
            n3 = _lib.launch_count()
        yp = p(ids)
    assert torch.equal(y1, y2) and torch.equal(y1, y3)
    assert n3 - n2 < n2 - n1 < n1 - n0
    # BASIC-mode logits stay close to the unquantised twin (BFP16 / FLOAT16 noise only)
    assert torch.isfinite(y1).all()
    rel = (y1.float() - yp.float()).norm() / yp.float().norm()
    assert rel < 0.1, rel


def test_opt_layer_casts_match_oracle():
    """first Linear of the stack, module boundary by module boundary against the CPU oracle"""
    q, _ = opt.build_pair(TINY, device=DEV, dtype=torch.float32)
    lin = q.layers[0].q_proj
    x = torch.randn(2, 96, 128, device=DEV)
    with torch.no_grad():
        y = lin(x)
    xi = O.cast(x.cpu().numpy(), "BFP[8|8]{64}(SN)")
    w = O.cast(lin.weight.detach().cpu().numpy(), "BFP[8|8]{64}(SN)")
    b = O.cast(lin.bias.detach().cpu().numpy(), "BFP[24|8]{1}(SN)")
    pre = torch.nn.functional.linear(torch.from_numpy(xi).to(DEV), torch.from_numpy(w).to(DEV), torch.from_numpy(b).to(DEV))
    want = O.cast(pre.cpu().numpy(), "FP[1|5|10,15](FN)")
    assert (bits(y.cpu().numpy()) == bits(want)).all()


def test_fused_hypernet_equals_module_by_module():
    torch.manual_seed(3)
    lin = dmxnn.Linear(512, 256).to(DEV).eval()
    lin.configure(dict(weight_format=fmt.BFP16_64, weight_storage_format=fmt.SBFP12_16, weight_sparseness="BTOPK{4:8,-1}(U)",
                       weight_score_func=abs_score))
    with torch.no_grad():
        a = lin._weight.clone()
        lin.weight_sparsifier.plastic = True
        n0 = _lib.launch_count()
        with elide.enabled():
            b = lin._weight
        assert _lib.launch_count() - n0 == 1  # sparsify -> storage cast -> weight cast: ONE kernel
    assert torch.equal(a, b)
    want = O.cast(O.cast(O.nm_prune(lin.weight.detach().cpu().numpy(), 4, 8), "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", tie=O.TIE_AWAY),
                  "BFP[8|8]{64}(SN)")
    assert (bits(a.cpu().numpy()) == bits(want)).all()


def test_lenet5_example_config_against_reference_golden():
    """BASELINE config #1: LeNet-5 with the example yaml's per-module config; every cast bit-equal to
    the reference's CPU path given the same layer input, logits within conv/GEMM-order tolerance."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "lenet5_reference.npz"))
    conv1, conv2 = dmxnn.Conv2d(1, 6, 5), dmxnn.Conv2d(6, 16, 5)
    fc1, fc2, fc3 = dmxnn.Linear(400, 120), dmxnn.Linear(120, 84), dmxnn.Linear(84, 10)
    mods = dict(conv1=conv1, conv2=conv2, fc1=fc1, fc2=fc2, fc3=fc3)
    cfg = dict(input_formats=[fmt.BFP16_64], weight_format=fmt.BFP16_64, bias_format=fmt.SAME, output_formats=[fmt.FLOAT16])
    for n, m in mods.items():
        with torch.no_grad():
            m.weight.copy_(torch.from_numpy(z[n + ".weight"]))
            m.bias.copy_(torch.from_numpy(z[n + ".bias"]))
        m.to(DEV).eval()
        m.configure(cfg)
    F = torch.nn.functional
    with torch.no_grad():
        for n, m in mods.items():
            h = torch.from_numpy(z[n + ".in"]).to(DEV)
            hi = m.input_casts.input_cast(h)
            assert (bits(hi.cpu().numpy()) == z[n + ".in_cast"]).all(), f"{n} input cast"
            assert (bits(m._weight.cpu().numpy()) == z[n + ".w_cast"]).all(), f"{n} weight cast"
            pre = torch.from_numpy(z[n + ".pre"]).to(DEV)
            assert (bits(m.output_casts.output_cast(pre).cpu().numpy()) == z[n + ".out"]).all(), f"{n} output cast"
        x = torch.from_numpy(z["x"]).to(DEV)
        h = F.max_pool2d(F.relu(conv1(x)), (2, 2))
        h = F.max_pool2d(F.relu(conv2(h)), 2)
        h = torch.flatten(h, 1)
        h = fc3(F.relu(fc2(F.relu(fc1(h)))))
    np.testing.assert_allclose(h.cpu().numpy(), z["logits"], rtol=0, atol=2e-3)


def test_basic_forward_is_cuda_graph_capturable():
    """kernels run on the current stream, never sync, never allocate: capture + replay == eager, bit for bit"""
    from dmx_compressor_b200 import graph

    q, _ = opt.build_pair(TINY, device=DEV, dtype=torch.float32)
    ids = torch.randint(0, 512, (2, 64), device=DEV)
    with torch.no_grad():
        want = q(ids)
    fwd = graph.capture(q, ids)
    assert torch.equal(fwd(ids), want)
    ids2 = torch.randint(0, 512, (2, 64), device=DEV)
    with torch.no_grad():
        want2 = q(ids2)
    assert torch.equal(fwd(ids2), want2)
    fwd_e = graph.capture(q, ids, elide_casts=True)
    assert torch.equal(fwd_e(ids2), want2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_resadd_equals_module_by_module(dt):
    """ResAdd under elision = input casts + add + output cast in ONE kernel, incl. broadcast masks"""
    from dmx_compressor_b200 import ops

    torch.manual_seed(5)
    add = dmxnn.ResAdd().to(DEV)
    add.configure(dict(input_formats=[fmt.FLOAT16, fmt.FLOAT16], output_formats=[fmt.FLOAT16]))
    a = (torch.randn(2, 3, 64, 64, device=DEV) * 300).to(dt)
    a.view(-1)[::11] = 1e-6
    S = 64
    mask = torch.full((S, S), torch.finfo(dt).min, device=DEV, dtype=dt).triu(1)[None, None].expand(2, 1, S, S)
    cases = ((mask, 2),  # broadcast operand: its un-expanded base is cast once (and remembered), then ONE fused add
             ((torch.randn(2, 3, 64, 64, device=DEV) * 5).to(dt), 1), (torch.randn(64, device=DEV).to(dt), 1),
             (torch.randn(3, 1, 64, device=DEV).to(dt), 3))  # last: 3 broadcast runs -> falls back to cast, cast, add+cast
    for b, launches in cases:
        with torch.no_grad():
            want = add(a, b)
            n0 = _lib.launch_count()
            with elide.enabled():
                got = elide.materialise(add(a, b))  # (an output cast may come back deferred)
            assert _lib.launch_count() - n0 == launches
        assert torch.equal(got.view(torch.int16 if dt != torch.float32 else torch.int32), want.view(torch.int16 if dt != torch.float32 else torch.int32))
    # oracle composition for the fp32 case
    if dt == torch.float32:
        b = torch.randn(2, 3, 64, 64, device=DEV) * 5
        with torch.no_grad(), elide.enabled():
            got = elide.materialise(add(a, b))
        F16 = "FP[1|5|10,15](FN)"
        want = O.cast(O.cast(a.cpu().numpy(), F16) + O.cast(b.cpu().numpy(), F16), F16)
        assert (bits(got.cpu().numpy()) == bits(want)).all()


def test_deferred_output_cast_is_safe_and_fuses():
    """elide.Lazy: a deferred FLOAT16 output cast (a) fuses with a BFP16 consumer into ONE kernel, (b) is
    materialised by any plain torch op, so arbitrary code between modules sees the eager values"""
    lin = dmxnn.Linear(128, 128).to(DEV).eval()
    nxt = dmxnn.Linear(128, 64).to(DEV).eval()
    for m in (lin, nxt):
        m.configure(dmxnn.config_rules.BASIC[0].module_config)
    x = torch.randn(4, 33, 128, device=DEV)
    with torch.no_grad():
        h = lin(x)
        want_next = nxt(h)
        want_scaled = h * 0.125
        with elide.enabled():
            hl = lin(x)
            assert isinstance(hl, elide.Lazy) and hl.shape == h.shape and hl.dtype == h.dtype  # metadata does not force it
            assert hl._real is None
            n0 = _lib.launch_count()
            got_next = elide.materialise(nxt(hl))          # FLOAT16 -> BFP16 fused: one kernel for both casts (+ cached weights)
            fused_launches = _lib.launch_count() - n0
            assert hl._real is None                         # the consumer never needed the stand-alone FLOAT16 tensor
            got_scaled = hl * 0.125                         # a plain torch op forces the cast
            assert hl._real is not None and type(got_scaled) is torch.Tensor
            assert torch.equal(hl.view(-1), h.view(-1))
    assert torch.equal(got_next, want_next) and torch.equal(got_scaled, want_scaled)
    assert fused_launches <= 4  # input chain, weight, bias, (deferred) output cast


def test_smoothquant_gpu_matches_reference_golden():
    """maxabs through dmxq_minmax on the device; scale within powf's CPU/GPU ulp difference of the reference's"""
    from test_observer_cpu import replay_smoothquant
    from dmx_compressor_b200 import ops

    replay_smoothquant(ops.minmax, device=DEV, ulps=8)


def test_linear_with_smoothquant():
    """reference core.py:184-196,227-230: input / s before the input cast, weight * s between sparsifier and storage cast"""
    torch.manual_seed(3)
    lin = dmxnn.Linear(256, 96).to(DEV)
    lin.configure(dict(input_formats=[fmt.BFP16_64], weight_format=fmt.BFP16_64, output_formats=[fmt.FLOAT16]))
    x = torch.randn(2, 33, 256, device=DEV) * torch.rand(256, device=DEV).mul(6).exp2()
    y_plain = lin(x)
    sq = lin.smoothquant
    sq.calibrating = True
    assert torch.equal(lin(x), y_plain)  # calibrating: statistics only, nothing scaled yet
    sq.calibrating = False
    s = sq.scale
    assert s.shape == (256,)
    mx_a = x.abs().amax((0, 1))
    mx_w = lin.weight.abs().amax(0).clamp(min=1e-5)
    torch.testing.assert_close(s, (mx_a**0.5 / mx_w**0.5).clamp(min=1e-5), rtol=1e-6, atol=0)
    sq.enable()
    y = lin(x)
    xi = CastTo(fmt.BFP16_64, block_dim=-1).to(DEV)(x / s)
    w = CastTo(fmt.BFP16_64, block_dim=-1).to(DEV)(lin.weight * s)
    want = CastTo(fmt.FLOAT16).to(DEV)(torch.nn.functional.linear(xi, w, lin.bias))
    assert torch.equal(y, want)
    assert not torch.equal(y, y_plain)
    with torch.no_grad(), elide.enabled():
        assert torch.equal(elide.materialise(lin(x)), want)
    lin.fold_weight_and_bias()
    assert sq.fused_to_weight.item() == 1
    assert torch.equal(lin.weight, w)
    assert torch.equal(lin(x), want)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_smoothquant_scale_stage_equals_torch_ops(dt):
    """DMXQ_STAGE_SCALE: `a / scale` (promoted to fp32) and `(b * scale).to(b.dtype)` of numerical/smoothquant.py:253-283 as the
    first stage of a cast chain == the torch ops followed by the same casts, bit for bit"""
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format

    g = torch.Generator(device=DEV).manual_seed(4)
    x = (torch.randn(3, 40, 512, device=DEV, generator=g) * 4).to(dt)
    sc = (torch.rand(512, device=DEV, generator=g) * 3 + 0.01)
    sc[5] = 1.0; sc[6] = 2.0 ** -20; sc[7] = 3.0e4
    for sh in ("BFP[8|8]{64}(SN)", "FP[1|5|10,15](FN)", "BFP[4|8]{128}(SN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"):
        st = Format.from_shorthand(sh).stage()
        # input side: fp32 result
        want = ops.cast_chain((x / sc), [st], -1)
        got = ops.cast_chain(x, [ops.scale_stage(sc), st], -1, out_dtype=torch.float32)
        assert got.dtype == torch.float32 and torch.equal(got.view(torch.int32), want.view(torch.int32)), (sh, "divide")
        # weight side: product rounded to the tensor dtype, then the casts
        w = x[0]
        want = ops.cast_chain((w * sc).to(dt), [st], -1)
        got = ops.cast_chain(w, [ops.scale_stage(sc, multiply=True), st], -1)
        assert got.dtype == dt and torch.equal(got.view(torch.int32 if dt == torch.float32 else torch.int16), want.view(torch.int32 if dt == torch.float32 else torch.int16)), (sh, "multiply")
    # strided rows are taken too; a channel axis that is not the contiguous dim is refused (callers keep the torch ops)
    xs = x[:, :, 128:384]
    st = Format.from_shorthand("BFP[8|8]{64}(SN)").stage()
    assert torch.equal(ops.cast_chain(xs, [ops.scale_stage(sc[128:384].contiguous()), st], -1, out_dtype=torch.float32), ops.cast_chain(xs / sc[128:384], [st], -1))
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.cast_chain(x.transpose(1, 2), [ops.scale_stage(sc[:40].contiguous()), st], -1, out_dtype=torch.float32)
    with pytest.raises(RuntimeError):
        ops.cast_chain(x, [ops.scale_stage(sc[:100].contiguous()), st], -1, out_dtype=torch.float32)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_linear_with_smoothquant_fused_under_elision(dt):
    """under elision an enabled SmoothQuant costs no extra pass: input / s + input cast = one kernel, weight * s + storage + weight
    cast = one kernel; logits bit-identical to the module-by-module forward"""
    torch.manual_seed(3)
    lin = dmxnn.Linear(256, 96).to(device=DEV, dtype=dt)
    lin.configure(dict(input_formats=[fmt.BFP16_64], weight_format=fmt.BFP16_64, output_formats=[fmt.FLOAT16]))
    x = (torch.randn(2, 33, 256, device=DEV) * torch.rand(256, device=DEV).mul(6).exp2()).to(dt)
    sq = lin.smoothquant
    with torch.no_grad():
        sq.calibrating = True
        lin(x)
        sq.calibrating = False
        sq.enable()
        want = lin(x)
        n0 = _lib.launch_count()
        with elide.enabled():
            got = elide.materialise(lin(x))
        n1 = _lib.launch_count()
    assert torch.equal(got, want)
    assert n1 - n0 == 3  # input chain, weight chain, output cast


class _FloatMLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.act, self.fc2 = dmxnn.Linear(64, 128), dmxnn.ReLU(), dmxnn.Linear(128, 32)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def test_graph_capture_with_elision_records_every_cast_of_a_float_input_model():
    """a small float-input model makes fewer cast-memo entries than the memo holds: the first input cast hit the warm-up's memo
    entry during capture and was left out of the graph.  Capture now starts from an empty memo / weight cache: replay with a
    DIFFERENT input must equal eager, and the caches hold nothing from inside the graph afterwards"""
    from dmx_compressor_b200 import graph

    torch.manual_seed(9)
    net = _FloatMLP().to(DEV).eval()
    for m in net.modules():
        if isinstance(m, dmxnn.Linear):
            m.configure(dict(input_formats=[fmt.BFP16_64], weight_format=fmt.BFP16_64, bias_format=fmt.BFP32_1, output_formats=[fmt.FLOAT16]))
        elif isinstance(m, dmxnn.ReLU):
            m.configure(dict(input_formats=[fmt.FLOAT16], output_formats=[fmt.FLOAT16]))
    x1, x2 = torch.randn(8, 64, device=DEV), torch.randn(8, 64, device=DEV) * 3
    with torch.no_grad():
        want1, want2 = net(x1), net(x2)
    fwd = graph.capture(net, x1, elide_casts=True)
    assert torch.equal(fwd(x2), want2)
    assert torch.equal(fwd(x1), want1)
    assert all(getattr(m, "_wcache", None) is None for m in net.modules())
    with torch.no_grad(), elide.enabled():  # eager elided forwards after the capture are unaffected
        assert torch.equal(net(x2), want2)


def test_weight_cache_follows_cast_state():
    """under elision the cast weight is cached; every switch that changes what `_weight` means must miss the cache:
    fake-quant off / on, calibration (observer must see each forward), new qparams, a re-armed sparsifier"""
    torch.manual_seed(10)
    lin = dmxnn.Linear(64, 32).to(DEV).eval()
    lin.configure(dict(weight_format=fmt.INT8))
    x = torch.randn(4, 64, device=DEV)
    with torch.no_grad(), elide.enabled():
        wq = lin._weight.clone()
        assert lin._weight is lin._weight  # cached
        lin.weight_cast.disable_fake_quant()
        assert torch.equal(lin._weight, lin.weight)  # not the stale quantised tensor
        lin.weight_cast.enable_fake_quant()
        assert torch.equal(lin._weight, wq)
        lin.weight_cast.enable_calibration(True, MinMaxObserver, qscheme_to_overload=torch.per_tensor_symmetric)
        lin.weight_cast.to(DEV)
        lin(x)
        lin(x)  # observer runs on every forward while calibrating
        lin.weight_cast.enable_calibration(False)
        wc = lin._weight.clone()
        assert not torch.equal(wc, wq)  # calibrated scale, not the unit scale
        lin.weight_cast.scale.mul_(2.0)
        assert not torch.equal(lin._weight, wc)
    want = CastTo(fmt.INT8).to(DEV)
    with torch.no_grad():
        want.scale.copy_(lin.weight_cast.scale); want.zero_point.copy_(lin.weight_cast.zero_point)
        with elide.enabled():
            assert torch.equal(lin._weight, want(lin.weight))


def test_fused_hypernet_leaves_the_sparsifier_state_of_the_module_path():
    """fused sparsify -> cast under elision: mask stored, lazy score created, `plastic` consumed; the NEXT forward (no longer
    plastic: the score parameter decides, as in the reference) equals the non-elided module path"""
    def build():
        torch.manual_seed(11)
        lin = dmxnn.Linear(256, 64).to(DEV).eval()
        lin.configure(dict(weight_sparseness="BTOPK{2:4,-1}(U)", weight_format=fmt.BFP12_128))
        lin.weight_sparsifier.configure(score_func=abs_score)
        return lin

    a, b = build(), build()
    with torch.no_grad():
        torch.manual_seed(12)
        w1 = a._weight.clone()
        w2 = a._weight.clone()
        with elide.enabled():
            torch.manual_seed(12)
            n0 = _lib.launch_count()
            v1 = b._weight.clone()
            assert _lib.launch_count() - n0 == 1
            v2 = b._weight.clone()
    assert torch.equal(w1, v1) and torch.equal(w2, v2)
    assert not torch.equal(w1, w2)  # the second forward prunes with the (random) score parameter
    assert torch.equal(a.weight_sparsifier.mask, b.weight_sparsifier.mask) and not b.weight_sparsifier.plastic
    assert torch.equal(a.weight_sparsifier.score, b.weight_sparsifier.score)
    assert float(b.weight_sparsifier.mask.mean()) == 0.5

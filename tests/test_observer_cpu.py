"""Host logic of the HistogramObserver (re-binning, clipping search, qparams) against golden sequences produced by
the reference's own observer (tests/golden/make_golden_hist.py).  The two kernel entry points of the observer are
swapped for the oracle here, so this runs without a GPU; tests/test_parity_gpu.py runs the same sequences through
dmxq_histc / dmxq_minmax."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from dmx_compressor_b200.numerical import Format, HistogramObserver
from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_hist", os.path.join(HERE, "golden", "make_golden_hist.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "hist_reference.npz"))


def oracle_histc(x, bins=100, min=0, max=0, return_minmax=False):
    xn = x.detach().float().numpy()
    h = torch.from_numpy(O.histc(xn, bins, min, max))
    if return_minmax:
        mn, mx = O.minmax(xn)
        return h, torch.tensor(mn[0]), torch.tensor(mx[0])
    return h


def oracle_minmax(x, ch_axis=None):
    mn, mx = O.minmax(x.detach().float().numpy(), ch_axis)
    return torch.from_numpy(mn), torch.from_numpy(mx)


class OracleBackedObserver(HistogramObserver):
    _histc = staticmethod(oracle_histc)
    _minmax = staticmethod(oracle_minmax)


def replay(cls, name, device="cpu", dtype=torch.float32):
    fmt, qs, bins, recipe = G.SEQUENCES[name]
    obs = cls(bins=bins, dtype=Format.from_shorthand(fmt), qscheme=G.QS[qs]).to(device)
    for i, x in enumerate(G.batches(recipe)):
        obs(torch.from_numpy(x).to(device=device, dtype=dtype))
        yield i, obs
    yield None, obs


def check_sequence(cls, name, device="cpu", exact_search=True):
    for i, obs in replay(cls, name, device):
        if i is None:
            break
        want_h = GOLD[f"{name}.{i}.hist"].view(np.float32)
        want_mm = GOLD[f"{name}.{i}.minmax"].view(np.float32)
        got_mm = np.array([obs.min_val.item(), obs.max_val.item()], np.float32)
        got_h = obs.histogram.cpu().numpy()
        if device == "cpu":
            assert np.array_equal(got_mm.view(np.uint32), want_mm.view(np.uint32)), (name, i, got_mm, want_mm)
            assert np.array_equal(got_h.view(np.uint32), want_h.view(np.uint32)), (name, i, np.abs(got_h - want_h).max())
        else:
            # torch's own CPU/CUDA differences, which the reference has too when it runs on the device: `t / python_number`
            # is t * (1 / number) on CUDA (an ulp off for non-power-of-two bin counts, which moves the widened maximum
            # and with it a few edge values between neighbouring bins), and the double-precision prefix sums of the
            # re-binning are accumulated in a different order
            np.testing.assert_allclose(got_mm, want_mm, rtol=1e-6)
            assert np.abs(got_h - want_h).sum() <= 1e-4 * want_h.sum(), (name, i)
            np.testing.assert_allclose(got_h.sum(), want_h.sum(), rtol=1e-6)
    lo, hi = obs._non_linear_param_search()
    sc, zp = obs.calculate_qparams()
    want_clip = GOLD[f"{name}.clip"].view(np.float32)
    if exact_search:
        assert np.array_equal(np.array([lo.item(), hi.item()], np.float32).view(np.uint32), GOLD[f"{name}.clip"])
        assert np.array_equal(sc.cpu().numpy().astype(np.float32).view(np.uint32), GOLD[f"{name}.scale"])
        assert np.array_equal(zp.cpu().numpy().astype(np.int64), GOLD[f"{name}.zero_point"])
    else:  # reduction order of the error estimate differs: the greedy search may stop a step earlier or later
        width = obs.max_val.item() - obs.min_val.item()
        assert abs(lo.item() - want_clip[0]) <= 0.02 * width and abs(hi.item() - want_clip[1]) <= 0.02 * width
    return obs


@pytest.mark.parametrize("name", list(G.SEQUENCES))
def test_histogram_observer_matches_reference(name):
    obs = check_sequence(OracleBackedObserver, name)
    if name == "steady":
        assert obs.stats == {"fused_steps": 3, "rebinned_steps": 0}
    if name == "widening":
        assert obs.stats == {"fused_steps": 1, "rebinned_steps": 3}  # the third batch is narrower than the second


def test_histogram_observer_rejects_per_channel():
    with pytest.raises(NotImplementedError):
        HistogramObserver(qscheme=torch.per_channel_affine)


def test_histogram_observer_uninitialised_qparams():
    sc, zp = HistogramObserver().calculate_qparams()
    assert sc.tolist() == [1.0] and zp.tolist() == [0]


def test_product_observer_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        HistogramObserver()(torch.randn(8, 8))


# ---- SmoothQuant host logic against the reference's golden cases -------------------------------------------------
_spec2 = importlib.util.spec_from_file_location("make_golden_smoothquant", os.path.join(HERE, "golden", "make_golden_smoothquant.py"))
GS = importlib.util.module_from_spec(_spec2)
_spec2.loader.exec_module(GS)
GOLD_SQ = np.load(os.path.join(HERE, "golden", "smoothquant_reference.npz"))


def smoothquant_classes(minmax):
    from dmx_compressor_b200.numerical.smoothquant import ActivationWeightSmoothQuant, SmoothQuant

    class SQ(SmoothQuant):
        _minmax = staticmethod(minmax)

    class AWSQ(ActivationWeightSmoothQuant):
        _minmax = staticmethod(minmax)

    return SQ, AWSQ


def replay_smoothquant(minmax, device="cpu", ulps=0):
    SQ, AWSQ = smoothquant_classes(minmax)
    _from_numpy = torch.from_numpy
    seen = []

    def check(key, value):
        want = GOLD_SQ[key].view(np.float32)
        got = value.detach().float().cpu().numpy()
        assert got.shape == want.shape, key
        if ulps == 0:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), key
        else:  # powf on the device differs from the host's by an ulp
            np.testing.assert_allclose(got, want, rtol=ulps * 2.0**-23, atol=0, err_msg=key)
        seen.append(key)

    class _Torch:  # the golden recipe calls torch.from_numpy: route it to the device under test
        @staticmethod
        def from_numpy(a):
            return _from_numpy(a).to(device)

    GS.torch, keep = _Torch, GS.torch
    try:
        GS.run(lambda *a, **k: SQ(*a, **k).to(device), lambda *a, **k: AWSQ(*a, **k).to(device), check)
    finally:
        GS.torch = keep
    assert len(seen) == len(GOLD_SQ.files)


def test_smoothquant_matches_reference():
    replay_smoothquant(oracle_minmax)


def test_smoothquant_validation():
    from dmx_compressor_b200.numerical.smoothquant import ActivationWeightSmoothQuant

    sq = ActivationWeightSmoothQuant(-1, -1)
    with pytest.raises(ValueError):
        sq.set_migration_strength(1.5)
    sq.fused_to_weight[0] = 1
    with pytest.raises(RuntimeError):
        sq.set_dynamic(True)

"""Drop-in check against the reference package itself (only where /root/reference exists, i.e. in
the authoring container): installing the plugin leaves every CPU result of the reference
bit-identical (CPU tensors keep running the reference's own extension) and patches exactly the
documented plugin points; uninstall restores them."""
import os
import sys

import pytest
import torch

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import load_reference  # noqa: E402

pytestmark = pytest.mark.skipif(not load_reference.available(), reason="reference sources not present (GPU box)")


def test_install_patches_plugin_points_and_keeps_cpu_path():
    num, sp, q = load_reference.load()
    from dmx_compressor_b200 import plugin

    fmt = sys.modules["dmx.compressor.numerical.format"]
    x = torch.randn(8, 128)
    before = {sh: num.CastTo(sh)(x).clone() for sh in ("BFP[8|8]{64}(SN)", "FP[1|5|10,15](FN)", "XP[8,0](CSN)",
                                                        "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")}
    s = sp.Sparsify(x.shape, "BTOPK{2:4,-1}(U)")
    s.eval()
    s_before = s(x).clone()
    orig = (fmt.BlockFloatingPoint.cast, fmt.FloatingPoint.cast, fmt.FixedPoint.cast, fmt.ScaledBlockFloatingPoint.cast,
            sp.Sparsify.forward, q.block_quantize)
    plugin.install("dmx.compressor")
    try:
        assert plugin.installed()
        now = (fmt.BlockFloatingPoint.cast, fmt.FloatingPoint.cast, fmt.FixedPoint.cast, fmt.ScaledBlockFloatingPoint.cast,
               sp.Sparsify.forward, q.block_quantize)
        assert all(a is not b for a, b in zip(orig, now))
        assert fmt.MXINT.cast is fmt.BlockFloatingPoint.cast  # inherits the patched method
        for sh, want in before.items():
            got = num.CastTo(sh)(x)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), sh
        assert torch.equal(s(x), s_before)
        assert torch.equal(q.float_quantize(x, 5, 10, rounding="nearest"), fmt.float_quantize(x, 5, 10, rounding="nearest"))
        # packed storage arrives as NEW methods on the reference's format classes (CUDA only: no host encoder)
        for sh in ("BFP[8|8]{64}(SN)", "MXINT8{64}", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"):
            f = num.Format.from_shorthand(sh)
            assert callable(f.pack) and callable(f.unpack)
            with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
                f.pack(x)
    finally:
        plugin.uninstall()
    assert fmt.BlockFloatingPoint.cast is orig[0] and sp.Sparsify.forward is orig[4] and not plugin.installed()
    assert not hasattr(fmt.BlockFloatingPoint, "pack") and not hasattr(fmt.ScaledBlockFloatingPoint, "unpack")


def test_histogram_step_drives_the_reference_observer_class():
    """plugin.install() binds numerical.observer.histogram_step onto the reference's HistogramObserver for CUDA tensors;
    here the same function drives a *reference* observer instance with the oracle standing in for the two kernels, and
    must leave exactly the state (and qparams) the reference's own forward leaves"""
    num, _, _ = load_reference.load()
    from dmx.compressor.numerical.observer import HistogramObserver as RefObserver

    from dmx_compressor_b200 import plugin
    from dmx_compressor_b200.numerical.observer import histogram_step
    from test_observer_cpu import G, oracle_histc, oracle_minmax

    plugin.install("dmx.compressor")
    try:
        for name in ("widening", "steady", "unit_interval"):
            fmt, qs, bins, recipe = G.SEQUENCES[name]
            ours = RefObserver(bins=bins, dtype=num.Format.from_shorthand(fmt), qscheme=G.QS[qs])
            theirs = RefObserver(bins=bins, dtype=num.Format.from_shorthand(fmt), qscheme=G.QS[qs])
            for x in G.batches(recipe):
                histogram_step(ours, torch.from_numpy(x), oracle_histc, oracle_minmax)
                theirs(torch.from_numpy(x))  # CPU tensor: the patched forward falls through to the reference's own
                assert torch.equal(ours.histogram, theirs.histogram)
                assert ours.min_val.item() == theirs.min_val.item() and ours.max_val.item() == theirs.max_val.item()
            a, b = ours.calculate_qparams(), theirs.calculate_qparams()
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    finally:
        plugin.uninstall()

"""HistogramObserver golden sequences from the reference's own python (S/numerical/observer.py:213-582), CPU.
Run here:  python tests/golden/make_golden_hist.py  ->  tests/golden/hist_reference.npz
Each sequence feeds a few batches to one observer; after every batch the histogram / min_val / max_val are
recorded, and at the end the clipping range of the search and the resulting scale / zero-point.  The batches are
regenerated in the tests from `batches()` below (numpy Generator streams are stable across machines)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

SEQUENCES = {
    # name: (format, qscheme, bins, recipe)
    "widening": ("XP[8,0](CSN)", "affine", 2048, "widening"),
    "steady": ("XP[8,0](CSN)", "affine", 2048, "steady"),
    "unit_interval": ("XP[8,0](CSN)", "symmetric", 2048, "unit"),
    "one_sided": ("XP[8,0](CSN)", "affine", 2048, "relu"),
    "int4_sym": ("XP[4,0](CSN)", "symmetric", 2048, "widening"),
    "bins_1000": ("XP[8,0](CSN)", "affine", 1000, "widening"),
    "constant_then_data": ("XP[8,0](CSN)", "affine", 2048, "constant"),
}


def batches(recipe, seed=7):
    """list of fp32 arrays [64, 512]"""
    rng = np.random.default_rng(seed)
    n = (64, 512)
    if recipe == "widening":  # every batch is wider than the last one: the stored histogram is re-binned each step
        return [(rng.standard_normal(n) * s).astype(np.float32) for s in (1.5, 2.5, 2.0, 6.0, 11.0)]
    if recipe == "steady":  # same extremes in every batch: steady-state accumulation
        out = []
        for _ in range(4):
            x = np.clip(rng.standard_normal(n) * 2.0, -7.25, 9.5).astype(np.float32)
            x[0, 0], x[0, 1] = -7.25, 9.5
            out.append(x)
        return out
    if recipe == "unit":  # everything inside (-1, 1): int(min) == int(max) == 0 -> histc uses each batch's own range
        return [(rng.random(n) * 1.6 - 0.8).astype(np.float32) * np.float32(s) for s in (1.0, 0.5, 1.1)]
    if recipe == "relu":
        return [np.maximum(rng.standard_normal(n) * s, 0).astype(np.float32) for s in (3.0, 5.0, 4.0)]
    if recipe == "constant":
        return [np.full(n, 2.5, np.float32), (rng.standard_normal(n) * 3).astype(np.float32),
                (rng.standard_normal(n) * 4).astype(np.float32)]
    raise KeyError(recipe)


QS = {"affine": torch.per_tensor_affine, "symmetric": torch.per_tensor_symmetric}


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    import load_reference

    num, _, _ = load_reference.load()
    from dmx.compressor.numerical.observer import HistogramObserver

    out = {}
    for name, (fmt, qs, bins, recipe) in SEQUENCES.items():
        obs = HistogramObserver(bins=bins, dtype=num.Format.from_shorthand(fmt), qscheme=QS[qs])
        for i, x in enumerate(batches(recipe)):
            obs(torch.from_numpy(x))
            out[f"{name}.{i}.hist"] = obs.histogram.numpy().view(np.uint32).copy()
            out[f"{name}.{i}.minmax"] = np.array([obs.min_val.item(), obs.max_val.item()], np.float32).view(np.uint32)
        lo, hi = obs._non_linear_param_search()
        sc, zp = obs.calculate_qparams()
        out[f"{name}.clip"] = np.array([lo.item(), hi.item()], np.float32).view(np.uint32)
        out[f"{name}.scale"] = sc.numpy().astype(np.float32).view(np.uint32)
        out[f"{name}.zero_point"] = zp.numpy().astype(np.int64)
        print(name, "steps", i + 1, "range", obs.min_val.item(), obs.max_val.item(), "clip", lo.item(), hi.item(), "scale", sc.item(), "zp", zp.item())
    np.savez_compressed(os.path.join(HERE, "hist_reference.npz"), **out)


if __name__ == "__main__":
    main()

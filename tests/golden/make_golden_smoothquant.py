"""SmoothQuant golden cases from the reference's own python (S/numerical/smoothquant.py), CPU.
Run here:  python tests/golden/make_golden_smoothquant.py  ->  tests/golden/smoothquant_reference.npz
Inputs are regenerated in the tests from `inputs()` below."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def inputs(seed=11):
    rng = np.random.default_rng(seed)

    def t(*shape, scale=1.0):
        x = rng.standard_normal(shape) * scale * np.exp2(rng.integers(-3, 4, shape[-1:]))
        return x.astype(np.float32)

    return {
        # activation x weight (Linear): channel axis -1 on both
        "aw": dict(inp=[t(4, 16, 32), t(4, 16, 32, scale=3.0)], wgt=t(24, 32, scale=0.1)),
        # activation x activation: running maxima over calls with shrinking / growing channel counts
        "aa": dict(a=[t(2, 3, 8, 16), t(2, 3, 8, 12, scale=4.0), t(2, 3, 8, 16, scale=0.5)],
                   b=[t(2, 3, 16, 8), t(2, 3, 12, 8), t(2, 3, 16, 8, scale=2.0)]),
    }


def run(SmoothQuant, ActivationWeightSmoothQuant, out):
    data = inputs()
    for m in (0.5, 0.8):
        sq = ActivationWeightSmoothQuant(-1, -1, migration_strength=m)
        w = torch.from_numpy(data["aw"]["wgt"])
        for i, x in enumerate(data["aw"]["inp"]):
            sq(torch.from_numpy(x), w)
            out(f"aw.{m}.{i}.scale", sq.scale)
        sq.enable()
        out(f"aw.{m}.scaled_input", sq.scale_input(torch.from_numpy(data["aw"]["inp"][0])))
        out(f"aw.{m}.scaled_weight", sq.scale_weight(w))
    for dyn in (False, True):
        sq = SmoothQuant(a_ch_axis=-1, b_ch_axis=-2, a_dynamic=dyn, b_dynamic=dyn, migration_strength=0.4)
        sq.enable()
        for i, (a, b) in enumerate(zip(data["aa"]["a"], data["aa"]["b"])):
            ya, yb = sq(torch.from_numpy(a), torch.from_numpy(b))
            out(f"aa.{int(dyn)}.{i}.scale", sq.scale)
            out(f"aa.{int(dyn)}.{i}.a", ya)
            out(f"aa.{int(dyn)}.{i}.b", yb)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    import load_reference

    load_reference.load()
    from dmx.compressor.numerical.smoothquant import ActivationWeightSmoothQuant, SmoothQuant

    store = {}
    run(SmoothQuant, ActivationWeightSmoothQuant, lambda k, v: store.__setitem__(k, v.detach().numpy().astype(np.float32).view(np.uint32).copy()))
    np.savez_compressed(os.path.join(HERE, "smoothquant_reference.npz"), **store)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()

"""MXFP golden vectors from the reference's own python (MXFP.cast, S/numerical/format.py:514-602), CPU.
Run here:  python tests/golden/make_golden_mxfp.py  ->  tests/golden/mxfp_reference.npz
The reference concatenates along `block_dim` (format.py:560), so only block_dim = -1 works there."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import load_reference  # noqa: E402

num, _, _ = load_reference.load()


def main():
    g = torch.Generator().manual_seed(4242)
    out = {}
    names = []
    i = 0
    for sh in ("MXFP8[E4M3]{32}", "MXFP8[E5M2]{32}", "MXFP6[E2M3]{32}", "MXFP6[E3M2]{64}", "MXFP4[E2M1]{32}", "MXFP8[E4M3]{128}"):
        for shp in ((16, 256), (5, 96), (3, 4, 128)):
            x = torch.randn(shp, generator=g) * torch.pow(2.0, torch.randint(-10, 11, shp[:-1] + (1,), generator=g).float())
            x.view(-1)[::7] = torch.round(x.view(-1)[::7] * 8) / 8
            x.view(-1)[3::31] = torch.pow(2.0, torch.randint(-6, 7, x.view(-1)[3::31].shape, generator=g).float())  # exact powers of two
            y = num.CastTo(sh, block_dim=-1)(x)
            out[f"{i}.x"] = x.numpy().view(np.uint32).copy()
            out[f"{i}.y"] = y.numpy().view(np.uint32).copy()
            names.append(f"{i}|{sh}|{','.join(map(str, shp))}")
            i += 1
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mxfp_reference.npz"), **out)
    print("wrote", i, "cases")


if __name__ == "__main__":
    main()

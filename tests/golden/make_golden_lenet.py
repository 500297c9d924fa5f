"""BASELINE config #1 golden: LeNet-5 through the reference's own CastTo modules on CPU, with the
per-module config of configs/dmx_example_config_lenet5.yaml translated to the current API
("BFP[8|8]{64,d}(SN)" => format BFP[8|8]{64}(SN) + block_dim=d; bias SAME;
output FP[1|5|10,15](FN); accum SAME).  The composition per layer is DmxModule.forward's
(reference modeling/nn/core.py:215-264, torch_modules.py:346-360, :679-688):
    out_cast( op( in_cast(x), weight_cast(w), bias ) )
Run here:  python tests/golden/make_golden_lenet.py   -> tests/golden/lenet5_reference.npz"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import load_reference  # noqa: E402

num, _, _ = load_reference.load()
BFP, F16 = "BFP[8|8]{64}(SN)", "FP[1|5|10,15](FN)"


def main():
    torch.manual_seed(0)
    conv1, conv2 = torch.nn.Conv2d(1, 6, 5), torch.nn.Conv2d(6, 16, 5)
    fc1, fc2, fc3 = torch.nn.Linear(400, 120), torch.nn.Linear(120, 84), torch.nn.Linear(84, 10)
    x = torch.randn(4, 1, 32, 32)
    out = {"x": x.numpy()}
    for n, m in (("conv1", conv1), ("conv2", conv2), ("fc1", fc1), ("fc2", fc2), ("fc3", fc3)):
        out[n + ".weight"] = m.weight.detach().numpy()
        out[n + ".bias"] = m.bias.detach().numpy()

    def layer(name, m, h, conv):
        bd = 1 if conv else -1
        hi = num.CastTo(BFP, block_dim=bd)(h)
        w = num.CastTo(BFP, block_dim=bd)(m.weight.detach())
        if conv:
            pre = F.conv2d(hi, w, None) + m.bias.detach().unsqueeze(-1).unsqueeze(-1)   # torch_modules.py:679-688
        else:
            pre = F.linear(hi, w, m.bias.detach())                                       # torch_modules.py:346-350
        y = num.CastTo(F16)(pre)
        out[name + ".in"] = h.numpy().copy()
        out[name + ".in_cast"] = hi.numpy().view(np.uint32).copy()
        out[name + ".w_cast"] = w.numpy().view(np.uint32).copy()
        out[name + ".pre"] = pre.numpy().copy()
        out[name + ".out"] = y.numpy().view(np.uint32).copy()
        return y

    with torch.no_grad():
        h = F.max_pool2d(F.relu(layer("conv1", conv1, x, True)), (2, 2))
        h = F.max_pool2d(F.relu(layer("conv2", conv2, h, True)), 2)
        h = torch.flatten(h, 1)
        h = F.relu(layer("fc1", fc1, h, False))
        h = F.relu(layer("fc2", fc2, h, False))
        h = layer("fc3", fc3, h, False)
    out["logits"] = h.numpy()
    np.savez_compressed(os.path.join(HERE, "lenet5_reference.npz"), **out)
    print("wrote lenet5_reference.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()

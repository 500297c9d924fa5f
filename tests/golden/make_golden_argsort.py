"""Golden generator (needs a GPU: `gpurun -- python tests/golden/make_golden_argsort.py`): the order torch.argsort
(default, unstable -- the call BlockTopK makes, reference S/sparse.py:172) gives tied keys in rows of M <= 32 on CUDA.
Model under test: the 32-slot bitonic network of ATen's bitonicSortKVInPlace (SortUtils.cuh) with `M` valid slots, for
both candidate comparators.  Prints which model reproduces torch on exhaustive / random tie patterns (measured on the
B200, torch 2.11: the LT model, for every M, batch size, stride and key dtype) and dumps the raw results to
gpurun_out/argsort_probe.npz; tests/golden/argsort_cuda_order.npz is that dump with the keys / indices narrowed to uint8
and the random M = 16 / 32 cases cut to 4000 rows (see the end of this file)."""
import itertools
import os
import sys

import numpy as np
import torch



def bitonic32(keys, lt_mode):
    """keys: [n, M] float -> argsort indices [n, M] from the emulated network (numpy, vectorised over rows)"""
    n, M = keys.shape
    P = 32
    k = np.zeros((n, P), dtype=np.float64)
    k[:, :M] = keys
    v = np.tile(np.arange(P), (n, 1))
    valid = np.zeros((n, P), dtype=bool)
    valid[:, :M] = True

    def comp(a, b):
        return (a < b) if lt_mode else (a > b)

    def ce(pos, stride, dirflag):
        a, b = pos, pos + stride
        swap = (comp(k[:, a], k[:, b]) & valid[:, a]) | ~valid[:, b]
        ex = swap == dirflag
        for arr in (k, v, valid):
            ta, tb = arr[:, a].copy(), arr[:, b].copy()
            arr[:, a] = np.where(ex, tb, ta)
            arr[:, b] = np.where(ex, ta, tb)

    size = 2
    while size < P:
        stride = size // 2
        while stride > 0:
            for t in range(P // 2):
                flag = (t & (size // 2)) != 0
                pos = 2 * t - (t & (stride - 1))
                ce(pos, stride, flag)
            stride //= 2
        size *= 2
    stride = P // 2
    while stride > 0:
        for t in range(P // 2):
            pos = 2 * t - (t & (stride - 1))
            ce(pos, stride, False)
        stride //= 2
    return v[:, :M]


def main():
    dev = "cuda:0"
    out = {}
    rng = np.random.default_rng(0)
    for M in (2, 3, 4, 6, 8, 16, 32):
        if M <= 4:
            rows = np.array(list(itertools.product(range(3), repeat=M)), dtype=np.float32)
        elif M == 6:
            rows = np.array(list(itertools.product(range(3), repeat=M)), dtype=np.float32)
        elif M == 8:
            rows = np.array(list(itertools.product(range(3), repeat=M)), dtype=np.float32)
        else:
            rows = rng.integers(0, 4, size=(20000, M)).astype(np.float32)
        t = torch.from_numpy(rows).to(dev)
        idx = torch.argsort(t, dim=1).cpu().numpy()
        idx_stable = torch.argsort(t, dim=1, stable=True).cpu().numpy()
        idx_desc = torch.argsort(t, dim=1, descending=True).cpu().numpy()
        # big batch: does the result depend on the batch size / launch shape?
        big = torch.from_numpy(np.tile(rows, (max(1, 200000 // len(rows)), 1))).to(dev)
        idx_big = torch.argsort(big, dim=1).cpu().numpy()[: len(rows)]
        out[f"rows{M}"] = rows
        out[f"idx{M}"] = idx
        out[f"idx_desc{M}"] = idx_desc
        res = {"stable": (idx == idx_stable).all(), "batch_independent": (idx == idx_big).all()}
        for lt in (True, False):
            res["model_lt" if lt else "model_gt"] = (bitonic32(rows, lt) == idx).all()
        # non-contiguous input (the reference sorts a reshape of a transposed view, sparse.py:169-172)
        tt = t.t().contiguous().t()
        res["strided_same"] = (torch.argsort(tt, dim=1).cpu().numpy() == idx).all()
        # bf16 / fp16 keys
        for dt in (torch.bfloat16, torch.float16):
            res[str(dt)] = (torch.argsort(t.to(dt), dim=1).cpu().numpy() == idx).all()
        print(M, len(rows), {k: bool(v) for k, v in res.items()}, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/argsort_probe.npz", **out)
    small = {}
    for M in (2, 3, 4, 6, 8, 16, 32):
        n = 4000 if M >= 16 else None
        small[f"rows{M}"] = out[f"rows{M}"][:n].astype(np.uint8)
        small[f"idx{M}"] = out[f"idx{M}"][:n].astype(np.uint8)
    np.savez_compressed("gpurun_out/argsort_cuda_order.npz", **small)  # -> tests/golden/argsort_cuda_order.npz


if __name__ == "__main__":
    main()

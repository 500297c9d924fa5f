"""Shared helpers for the parity tests: golden loader, bit-exact comparison, case runner."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_G = None


def golden():
    global _G
    if _G is None:
        z = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
        with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
            meta = json.load(f)["cases"]
        _G = (z, meta)
    return _G


def golden_cases(prefix=None, kind=None):
    z, meta = golden()
    out = []
    for name, m in sorted(meta.items(), key=lambda kv: (kv[0].split("/")[0], int(kv[0].split("/")[1]))):
        if prefix and not name.startswith(prefix + "/"):
            continue
        if kind and m["kind"] != kind:
            continue
        out.append(name)
    return out


def case(name):
    z, meta = golden()
    m = meta[name]
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}
    return m, d


def _nan32(u):
    return (u & 0x7FFFFFFF) > 0x7F800000


def _nan16(u):  # bf16
    return (u & 0x7FFF) > 0x7F80


def _nan16h(u):  # fp16
    return (u & 0x7FFF) > 0x7C00


def assert_bits_equal(got, want, what="", dtype="float32"):
    """Bit-exact equality; NaNs compare by class (payload/sign of a NaN differs between x86
    and the GPU for the same IEEE operation, e.g. inf - inf)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} vs {want.dtype}"
    isnan = {"float32": _nan32, "bfloat16": _nan16, "float16": _nan16h}[dtype]
    gn, wn = isnan(got), isnan(want)
    bad = (gn != wn) | (~gn & (got != want))
    if bad.any():
        idx = np.argwhere(bad)[:8]
        lines = [f"{tuple(i)}: got 0x{int(got[tuple(i)]):08x} want 0x{int(want[tuple(i)]):08x}" for i in idx]
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} elements differ\n" + "\n".join(lines))


def f32(u32):
    return np.ascontiguousarray(u32).view(np.float32)


def bits(f):
    return np.ascontiguousarray(f, dtype=np.float32).view(np.uint32)

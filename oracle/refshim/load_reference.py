"""Import the reference's *own* numerics python (unmodified, from /root/reference) in this
container.  TEST SCAFFOLDING ONLY -- used by tests/golden/make_golden.py to generate the
committed fixtures; never imported by the product, never available on the GPU box.

Recipe: stand-ins for the three missing PyPI packages live beside
this file; bare ``dmx`` / ``dmx.compressor`` namespace modules are pre-registered so the
reference's heavyweight package ``__init__`` (transformers.utils.fx, graphviz, evaluate,
skopt ...) is skipped, and ``dmx.compressor.{quant,numerical,sparse}`` import as they are.
The reference JIT-builds its ``quant_cpu`` extension at import
(quant/quant_function.py:6-13); TORCH_EXTENSIONS_DIR is pointed at a scratch dir.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DMX_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src")
_loaded = None


def available():
    return os.path.isdir(os.path.join(_SRC, "dmx", "compressor", "numerical"))


def load():
    """-> (numerical_module, sparse_module, quant_module) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/dmxq_ref_torch_ext")
    for name, sub in (("dmx", "dmx"), ("dmx.compressor", os.path.join("dmx", "compressor"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(_SRC, sub)]
            sys.modules[name] = m
    sys.modules["dmx"].compressor = sys.modules["dmx.compressor"]
    import importlib

    quant = importlib.import_module("dmx.compressor.quant")
    numerical = importlib.import_module("dmx.compressor.numerical")
    sparse = importlib.import_module("dmx.compressor.sparse")
    _loaded = (numerical, sparse, quant)
    return _loaded

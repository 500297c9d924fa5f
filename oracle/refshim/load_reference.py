"""Import the reference's *own* python (unmodified) in this container and on the GPU box.
TEST SCAFFOLDING ONLY -- used by the golden-vector generators, by the plugin parity tests
(tests/test_plugin_reference.py, tests/test_plugin_gpu.py) and by bench.py's reference arms; never
imported by the product.

Where the sources come from (first that exists):
  1. ``$DMX_REFERENCE_ROOT/src``            (default /root/reference: the authoring container)
  2. ``oracle/_ref/pysrc``                  (a verbatim staging of the reference's ``src/dmx`` python
     made by ``oracle/build_ref.py stage_python``: git-ignored like the compiled ``oracle/_ref/*.so``,
     so it never enters history, but it travels to the GPU box where /root/reference does not exist)

Recipe:
  * stand-ins for the missing PyPI packages live beside this file (``parse``, ``bidict``, ``pptree``;
    for the whole package also ``graphviz``, ``evaluate``, ``skopt`` and a ``transformers.utils.fx``
    shim -- the installed transformers 5.5 removed it, the reference pins < 4.50);
  * ``load()``: bare ``dmx`` / ``dmx.compressor`` namespace modules are pre-registered so the reference's
    heavyweight package ``__init__`` is skipped and ``dmx.compressor.{quant,numerical,sparse}`` import as they are;
  * ``load_full()``: the whole package (``dmx.compressor.nn`` modules, ``DmxModel``, ``config_rules``).
  * the reference JIT-builds ``quant_cpu`` / ``quant_cuda`` at import (quant/quant_function.py:6-28).  Those two
    extensions are exactly what ``oracle/build_ref.py`` compiles ahead of time from the same unmodified sources
    (``oracle/_ref/ref_quant_{cpu,cuda}.so``), so during the import ``torch.utils.cpp_extension.load`` is answered
    with the prebuilt binaries (no ninja run, nothing written to ~/.cache); if a binary is missing the real JIT runs.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
REF_ROOT = os.environ.get("DMX_REFERENCE_ROOT", "/root/reference")
_CANDIDATES = [os.path.join(REF_ROOT, "src"), os.path.join(ORACLE, "_ref", "pysrc")]
_loaded = None
_loaded_full = None


def source_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "dmx", "compressor", "numerical")):
            return c
    return None


def available():
    return source_root() is not None


def _prebuilt(name):
    if ORACLE not in sys.path:
        sys.path.insert(0, ORACLE)
    import build_ref

    return build_ref.load("ref_" + name)


class _answer_jit_with_prebuilt:
    """while the reference imports: cpp_extension.load(name='quant_cpu'|'quant_cuda') -> oracle/_ref/ref_<name>.so"""

    def __enter__(self):
        from torch.utils import cpp_extension as ce

        self.ce, self.orig = ce, ce.load

        def load(name, sources, *a, **k):
            if name in ("quant_cpu", "quant_cuda"):
                try:
                    mod = _prebuilt(name)
                except Exception:
                    mod = None
                if mod is not None:
                    return mod
            os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/dmxq_ref_torch_ext")
            return self.orig(name, sources, *a, **k)

        ce.load = load
        return self

    def __exit__(self, *a):
        self.ce.load = self.orig


def _paths():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)


def load():
    """-> (numerical_module, sparse_module, quant_module) of the reference (numerics only)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if _loaded_full is not None:
        import dmx.compressor as dc

        _loaded = (dc.numerical, dc.sparse, importlib.import_module("dmx.compressor.quant"))
        return _loaded
    src = source_root()
    if src is None:
        raise RuntimeError(f"reference sources not found under {_CANDIDATES}")
    _paths()
    for name, sub in (("dmx", "dmx"), ("dmx.compressor", os.path.join("dmx", "compressor"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(src, sub)]
            sys.modules[name] = m
    sys.modules["dmx"].compressor = sys.modules["dmx.compressor"]
    with _answer_jit_with_prebuilt():
        quant = importlib.import_module("dmx.compressor.quant")
        numerical = importlib.import_module("dmx.compressor.numerical")
        sparse = importlib.import_module("dmx.compressor.sparse")
    _loaded = (numerical, sparse, quant)
    return _loaded


def _stub_transformers_fx():
    """transformers.utils.fx was removed from the installed transformers; the reference's fx/tracer.py needs the
    names to exist at import.  HF-model tracing is NOT made to work by this: plain-torch models and
    hand-assembled dmx.compressor.nn stacks are what the tests use."""
    import torch.fx
    import transformers
    import transformers.modeling_utils as mu
    import transformers.utils as tu

    if "transformers.utils.fx" not in sys.modules:
        fx = types.ModuleType("transformers.utils.fx")

        class HFTracer(torch.fx.Tracer):
            def __init__(self, autowrap_modules=(), autowrap_functions=()):
                super().__init__(autowrap_modules=tuple(autowrap_modules), autowrap_functions=tuple(autowrap_functions))

            def trace(self, root, concrete_args=None, dummy_inputs=None, **kw):
                # the real HFTracer propagates meta tensors built from dummy_inputs; plain-torch models trace without
                return super().trace(root, concrete_args=concrete_args)

        def get_concrete_args(model, input_names):
            import inspect

            sig = inspect.signature(model.forward)
            return {p.name: p.default for p in sig.parameters.values() if p.name not in input_names}

        fx.HFTracer = HFTracer
        fx.get_concrete_args = get_concrete_args
        fx._generate_supported_model_class_names = lambda *a, **k: []
        fx.check_if_model_is_supported = lambda *a, **k: None
        sys.modules["transformers.utils.fx"] = fx
        tu.fx = fx
    if not hasattr(mu, "ModelOutput"):
        from transformers.utils.generic import ModelOutput

        mu.ModelOutput = ModelOutput
    return transformers


def load_full():
    """-> the reference's whole ``dmx.compressor`` package (its real ``__init__``: format aliases, config_rules,
    ``nn`` modules, ``DmxModel``)."""
    global _loaded_full
    if _loaded_full is not None:
        return _loaded_full
    if _loaded is not None:
        # load() registered bare namespace stand-ins for the two package levels: drop those (the already imported
        # dmx.compressor.{quant,numerical,sparse} submodules stay in sys.modules and are reused by the real __init__)
        for name in ("dmx.compressor", "dmx"):
            sys.modules.pop(name, None)
    src = source_root()
    if src is None:
        raise RuntimeError(f"reference sources not found under {_CANDIDATES}")
    _paths()
    stubs = os.path.join(HERE, "fullpkg")
    if stubs not in sys.path:
        sys.path.insert(0, stubs)
    _stub_transformers_fx()
    if src not in sys.path:
        sys.path.insert(0, src)
    with _answer_jit_with_prebuilt():
        pkg = importlib.import_module("dmx.compressor")
    for name, mod in list(sys.modules.items()):  # submodules imported before the real package object existed
        parent, _, child = name.rpartition(".")
        if parent.startswith("dmx") and parent in sys.modules and not hasattr(sys.modules[parent], child):
            setattr(sys.modules[parent], child, mod)
    _loaded_full = pkg
    return pkg

"""ctypes binding of libdmxq.so (include/dmxq.h) -- the thin torch shim.

torch is used for what it is good at here: device memory (``torch.empty``), the current CUDA
stream and the device guard.  No torch types cross the C ABI: tensors become ``dmxq_tensor``
views (data pointer, dtype, shape, element strides).

There is NO CPU fallback: if the shared library is missing, importing this module raises;
if a CPU tensor reaches a cast, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdmxq.so")
MAX_DIMS = 8
MAX_STAGES = 4
ABI_VERSION = 3

# enums of include/dmxq.h
F32, BF16, F16 = 0, 1, 2
ROUND = {"nearest": 0, "stochastic": 1, "up": 2, "down": 3}
TIE_AWAY, TIE_EVEN = 0, 1
SCALE_DIV, SCALE_RECIP = 0, 1  # SBFP block scale: max / man_scaling (reference on CPU tensors) | max * fp32(1 / man_scaling) (on CUDA tensors)
NM_STABLE, NM_TORCH_CUDA = 0, 1  # N:M tie order: stable argsort (reference on CPU tensors) | torch's CUDA bitonic order
ST_NONE, ST_NM, ST_BFP, ST_SBFP, ST_FLOAT, ST_FIXED, ST_MXFP, ST_SCALE = 0, 1, 2, 3, 4, 5, 6, 7

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * MAX_DIMS), ("stride", C.c_int64 * MAX_DIMS)]


class Stage(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "block", "precision", "fraction", "man", "exp", "bias", "flush", "is_unsigned", "fp16_flush",
        "symmetric", "clamp", "rounding", "tie", "n_keep", "sc_man", "sc_exp", "sc_bias", "sc_flush", "sc_unsigned",
        "sc_fp16_flush", "sc_rounding")] + [("scale", C.c_float), ("zero_point", C.c_float),
                                            ("scale_mode", C.c_int32), ("nm_order", C.c_int32),
                                            ("vec", C.c_void_p), ("vec_len", C.c_int32), ("vec_op", C.c_int32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python dmx_compressor_b200/build.py` "
            "(nvcc, sm_100a). dmx_compressor_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    TP, SP, VP, I, I64 = C.POINTER(Tensor), C.POINTER(Stage), C.c_void_p, C.c_int, C.c_int64
    sigs = {
        "dmxq_abi_version": ([], I),
        "dmxq_last_error": ([], C.c_char_p),
        "dmxq_status_string": ([I], C.c_char_p),
        "dmxq_launch_count": ([], I64),
        "dmxq_cast_chain": ([TP, TP, I, SP, I, TP, TP, VP, VP], I),
        "dmxq_cast_chain_multi": ([TP, TP, I, I, SP, I, VP, VP], I),
        "dmxq_cast_chain_philox": ([TP, TP, I, SP, I, C.c_uint64, C.c_uint64, VP], I),
        "dmxq_philox_fill": ([VP, I64, I, C.c_uint64, C.c_uint64, VP], I),
        "dmxq_bfp_qdq": ([TP, TP, I, I, I, I, I, VP, VP], I),
        "dmxq_sbfp_qdq": ([TP, TP] + [I] * 14 + [VP], I),
        "dmxq_float_qdq": ([TP, TP] + [I] * 7 + [VP, VP], I),
        "dmxq_fixed_qdq": ([TP, TP] + [I] * 6 + [VP, VP, I64, I, I64, VP, VP], I),
        "dmxq_nm_prune": ([TP, TP, TP, TP, I, I, I, I, VP], I),
        "dmxq_add_cast": ([TP, TP, TP, SP, SP, SP, VP], I),
        "dmxq_softmax_cast": ([TP, TP, TP, SP, SP, SP, SP, I, VP], I),
        "dmxq_bfp_pack": ([TP, VP, VP, I, I, VP], I),
        "dmxq_bfp_unpack": ([VP, VP, TP, I, I, VP], I),
        "dmxq_sbfp_pack": ([TP, VP, VP, SP, VP, VP], I),
        "dmxq_sbfp_unpack": ([VP, VP, TP, SP, VP], I),
        "dmxq_block_quantize": ([TP, TP, I, I, I, I, VP, VP, VP], I),
        "dmxq_minmax": ([TP, I, VP, VP, VP], I),
        "dmxq_amax_multi": ([TP, I, VP, VP], I),
        "dmxq_histc": ([TP, I, C.c_float, C.c_float, VP, VP, VP, VP], I),
        "dmxq_cast_chain_host": ([VP, VP, I, I, I64, I64, SP, I, I], I),
        "dmxq_host_alloc": ([I64], VP),
        "dmxq_host_free": ([VP], None),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.argtypes = args
        fn.restype = res
    if lib.dmxq_abi_version() != ABI_VERSION:
        raise ImportError(f"libdmxq ABI {lib.dmxq_abi_version()} != binding ABI {ABI_VERSION}")
    return lib, list(sigs)


lib, EXPORTS = _load()


def check(rc: int, what: str = "dmxq") -> None:
    """Non-zero status -> RuntimeError, the reference's TORCH_CHECK convention
    (Q/quant_cuda/quant_cuda.cpp:7-11)."""
    if rc != 0:
        msg = lib.dmxq_last_error().decode(errors="replace")
        if rc == -1 and ("not a multiple of block size" in msg):
            raise AssertionError(msg)  # S/sparse.py:166-168 asserts
        raise RuntimeError(f"{what}: {lib.dmxq_status_string(rc).decode()} ({rc}): {msg}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise RuntimeError(f"dmxq: unsupported dtype {dt} (float32, bfloat16, float16 only)") from None


def require_cuda(t: torch.Tensor, name: str = "x") -> None:
    if not isinstance(t, torch.Tensor):
        raise AssertionError(f"{name} is not a torch.Tensor")
    if not t.is_cuda:
        # same message as the reference's CHECK_CUDA (Q/quant_cuda/quant_cuda.cpp:7)
        raise RuntimeError(f"{name} must be a CUDA tensor (dmx_compressor_b200 is CUDA-only: no CPU fallback)")


_VIEW_TEMPLATES: dict = {}


def view(t: torch.Tensor) -> Tensor:
    """dmxq_tensor descriptor of ``t``.  The (dtype, shape, stride) part is memoised -- a model issues the same few
    layouts over and over -- so a call costs one struct copy plus the data pointer."""
    if type(t) is not torch.Tensor and hasattr(t, "materialise"):
        t = t.materialise()  # a deferred cast (elide.Lazy) reaching the ABI directly: run it first
    key = (t.dtype, t.shape, t.stride())
    tmpl = _VIEW_TEMPLATES.get(key)
    if tmpl is None:
        if t.dim() > MAX_DIMS:
            raise RuntimeError(f"dmxq: tensors of more than {MAX_DIMS} dims are not supported")
        tmpl = Tensor()
        tmpl.dtype = dtype_code(t.dtype)
        tmpl.ndim = t.dim()
        for i, (n, s) in enumerate(zip(t.shape, t.stride())):
            tmpl.shape[i] = n
            tmpl.stride[i] = s
        if len(_VIEW_TEMPLATES) > 4096:
            _VIEW_TEMPLATES.clear()
        _VIEW_TEMPLATES[key] = tmpl
    v = Tensor.from_buffer_copy(tmpl)
    v.data = t.data_ptr()
    return v


def views(ts):
    """dmxq_tensor descriptors of many tensors at once -> (pointer usable as ``const dmxq_tensor *``, keep-alive object).
    One numpy fill instead of one ctypes struct per tensor (~2 us each through ``view``): struct dmxq_tensor is 18 int64 words --
    data | dtype + (ndim << 32) | shape[8] | stride[8]."""
    import numpy as np

    n = len(ts)
    a = np.zeros((n, 18), dtype=np.int64)
    a[:, 0] = [t.data_ptr() for t in ts]
    nd = ts[0].dim() if n else 0
    if all(t.dim() == nd for t in ts) and nd <= MAX_DIMS:
        a[:, 1] = [dtype_code(t.dtype) | (nd << 32) for t in ts]
        if nd:
            a[:, 2:2 + nd] = [t.shape for t in ts]
            a[:, 10:10 + nd] = [t.stride() for t in ts]
    else:
        for i, t in enumerate(ts):
            d = t.dim()
            if d > MAX_DIMS:
                raise RuntimeError(f"dmxq: tensors of more than {MAX_DIMS} dims are not supported")
            a[i, 1] = dtype_code(t.dtype) | (d << 32)
            a[i, 2:2 + d] = t.shape
            a[i, 10:10 + d] = t.stride()
    return a.ctypes.data_as(C.POINTER(Tensor)), a


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device) -> int:
    """the caller's current CUDA stream on ``device`` as a raw handle"""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib.dmxq_launch_count())

"""dmx_compressor_b200 -- B200-native (sm_100a) CastTo / Sparsify numerics path of
dmx-compressor, behind the reference's own Format / CastTo / Sparsify interface.

CUDA only.  Importing the package loads dmx_compressor_b200/lib/libdmxq.so and raises
ImportError if it has not been built (python dmx_compressor_b200/build.py).
"""
from . import _lib  # noqa: F401  (fails loudly when libdmxq.so is missing)
from . import ops  # noqa: F401
from . import numerical  # noqa: F401

__version__ = "0.1.0"

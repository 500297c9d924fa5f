"""Numerical formats -- host-side mirror of the reference's ``Format`` plugin surface
(reference src/dmx/compressor/numerical/format.py), with ``cast()`` routed to libdmxq.

Same class names, constructor arguments, shorthand grammar and ``repr`` round trip as the
reference, so a ``CastTo("BFP[8|8]{64}(SN)")`` means the same thing in both.  What differs is
the implementation of ``cast``: one fused CUDA kernel per call instead of the reference's
python split / per-chunk native call / cat loop (format.py:322-341, :453-479).

Every format also exposes ``stage()`` -- its description as a ``dmxq_stage`` -- so chains of
casts (weight hypernet, output-cast -> input-cast pairs) fuse into a single kernel
(``dmx_compressor_b200.ops.cast_chain``).
"""
from __future__ import annotations

import re
from typing import Optional

import torch

from .. import _lib as L
from .. import ops

ROUNDING_MODE = {"U": "up", "D": "down", "N": "nearest", "S": "stochastic"}
ROUNDING_MODE_INV = {v: k for k, v in ROUNDING_MODE.items()}

# FixedPoint "nearest" tie behaviour: the reference's CUDA kernels round half away from zero,
# its CPU extension ties to even.  CUDA tensors get the CUDA behaviour.
DEFAULT_TIE = L.TIE_AWAY


class Format:
    r"""Abstract tensor numerical format (reference format.py:33-77)."""

    blocked: bool
    bfp_id: Optional[int] = None

    def __str__(self) -> str:
        raise NotImplementedError

    def cast(self, *input):
        raise NotImplementedError

    def stage(self) -> Optional[L.Stage]:
        """dmxq_stage describing this format, or None when it is a no-op (SAME)."""
        raise NotImplementedError

    def _memo_stage(self, key, build):
        """the ctypes stage struct is rebuilt only when one of the attributes it encodes changed (a cast is issued per
        module call: building the struct -- and formatting ``repr`` for the FLOAT16 special case -- every time showed
        up in the host-side cost of small casts)"""
        c = self.__dict__.get("_stage_memo")
        if c is None or c[0] != key:
            c = (key, build())
            self.__dict__["_stage_memo"] = c
        return c[1]

    @property
    def bytes_per_elem(self) -> Optional[float]:
        raise NotImplementedError

    @property
    def bit_precision(self) -> Optional[float]:
        raise NotImplementedError

    @staticmethod
    def from_shorthand(sh: str):
        if sh.startswith("SAME"):
            return Same.from_shorthand(sh)
        elif sh.startswith("XP"):
            return FixedPoint.from_shorthand(sh)
        elif sh.startswith("FP"):
            return FloatingPoint.from_shorthand(sh)
        elif sh.startswith("BFP"):
            return BlockFloatingPoint.from_shorthand(sh)
        elif sh.startswith("SBFP"):
            return ScaledBlockFloatingPoint.from_shorthand(sh)
        elif sh.startswith("MXINT"):
            return MXINT.from_shorthand(sh)
        elif sh.startswith("MXFP"):
            return MXFP.from_shorthand(sh)
        else:
            raise ValueError(f"unrecognized format shorthand: {sh}")

    def __eq__(self, other):
        return isinstance(other, Format) and repr(self) == repr(other)

    def __hash__(self):
        return hash(repr(self))


def _bad(sh):
    return ValueError(f"unrecognized format shorthand: {sh}")


class Same(Format):
    r"""Dummy format: ``cast`` returns a copy (reference format.py:80-107: ``x.clone()``)."""

    blocked = False

    def cast(self, x, *args):
        return x.clone()

    def stage(self):
        return None

    @property
    def bytes_per_elem(self) -> None:
        return None

    @property
    def bit_precision(self) -> None:
        return None

    @classmethod
    def from_shorthand(cls, sh: str):
        return cls()

    def __str__(self) -> str:
        return "Dummy numerical format: no casting"

    def __repr__(self) -> str:
        return "SAME"


class FixedPoint(Format):
    r"""Fixed point simulated in fp32 (reference format.py:110-170)."""

    blocked = False
    _RX = re.compile(r"^XP\[(?P<precision>[-+]?\d+),(?P<fraction>[-+]?\d+)\]\((?P<clamp>\w)(?P<symmetric>\w)(?P<rounding>\w)\)$")

    def __init__(self, precision, fraction, clamp=True, symmetric=True, rounding="nearest", tie=None):
        assert 1 <= precision <= 24, f"highest integer precision simulated by FP32 is 25, got {precision}"
        self.precision = precision
        self.fraction = fraction
        self.clamp = clamp
        self.symmetric = symmetric
        self.rounding = rounding
        self.tie = DEFAULT_TIE if tie is None else tie

    def cast(self, x, *args):
        return ops.fixed_qdq(x, self.precision, self.fraction, self.clamp, self.symmetric, self.rounding, self.tie,
                             out_dtype=torch.float32)

    def stage(self, scale: float = 1.0, zero_point: float = 0.0):
        key = (self.precision, self.fraction, self.clamp, self.symmetric, self.rounding, self.tie, scale, zero_point)
        return self._memo_stage(key, lambda: ops.fixed_stage(*key))

    @property
    def bytes_per_elem(self) -> float:
        return self.precision / 8.0

    @property
    def bit_precision(self) -> float:
        return float(self.precision)

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise _bad(sh)
        return cls(precision=int(m["precision"]), fraction=int(m["fraction"]), clamp=m["clamp"] == "C",
                   symmetric=m["symmetric"] == "S", rounding=ROUNDING_MODE[m["rounding"]])

    def __str__(self) -> str:
        return (f"Simulated fixed point format: precision bits = {self.precision}, fraction bits = {self.fraction}, \n"
                f"casting behavior: symmetric = {self.symmetric}, clamp = {self.clamp}, rounding = {self.rounding}")

    def __repr__(self) -> str:
        return (f"XP[{self.precision},{'0' if self.fraction == 0 else f'{self.fraction:+d}'}]"
                f"({'C' if self.clamp else '_'}{'S' if self.symmetric else '_'}{ROUNDING_MODE_INV[self.rounding]})")


class FloatingPoint(Format):
    r"""Low-bit floating point simulated in fp32 (reference format.py:173-270)."""

    blocked = False
    _RX = re.compile(r"^FP\[(?P<sign>\d+)\|(?P<exponent>\d+)\|(?P<mantissa>\d+),(?P<bias>[-+]?\d+)\]\((?P<flush>\w)(?P<rounding>[A-Za-z])\)$")

    def __init__(self, mantissa=23, exponent=8, bias=None, flush_subnormal=True, unsigned=False, rounding="nearest"):
        assert 0 <= mantissa <= 23, f"number of mantisa bits simulatable by FP32 is between 0 and 23, got{mantissa}"
        assert 0 < exponent <= 8, f"number of exponent bits simulatable by FP32 is between 1 and 8, got {exponent}"
        if bias is None:
            bias = 2 ** (exponent - 1) - 1
        _bias_min = 127 if exponent == 8 else -128 + 2**exponent
        assert _bias_min <= bias <= 127, (f"exponent bias simulatable by FP32 for {exponent}-bit exponent is constrained "
                                          f"between {_bias_min} and 127, got {bias}")
        self.mantissa = mantissa
        self.exponent = exponent
        self.bias = bias
        self.flush_subnormal = flush_subnormal
        self.unsigned = unsigned
        self.rounding = rounding

    def _identity_for(self, dtype) -> bool:  # reference format.py:209-212
        r = repr(self)
        return (dtype == torch.float32 and r == "FP[1|8|23,127](_N)") or (dtype == torch.float16 and r == "FP[1|5|10,15](_N)")

    def cast(self, x, *args):
        if self._identity_for(x.dtype):
            return x.abs() if self.unsigned else x
        return ops.float_qdq(x, self.mantissa, self.exponent, self.bias, self.flush_subnormal, self.unsigned,
                             repr(self) == "FP[1|5|10,15](FN)", self.rounding, out_dtype=torch.float32)

    def stage(self):
        key = (self.mantissa, self.exponent, self.bias, self.flush_subnormal, self.unsigned, self.rounding)
        return self._memo_stage(key, lambda: ops.float_stage(self.mantissa, self.exponent, self.bias, self.flush_subnormal, self.unsigned,
                                                             repr(self) == "FP[1|5|10,15](FN)", self.rounding))

    @property
    def largest_representable_power_of_two(self):
        return 2 ** (2 ** (self.exponent - 1))

    @property
    def bytes_per_elem(self) -> float:
        return (self.mantissa + self.exponent + 1) / 8.0

    @property
    def bit_precision(self) -> float:
        return float(self.mantissa + self.exponent if self.unsigned else 1 + self.mantissa + self.exponent)

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise _bad(sh)
        return cls(mantissa=int(m["mantissa"]), exponent=int(m["exponent"]), bias=int(m["bias"]),
                   flush_subnormal=m["flush"] == "F", unsigned=int(m["sign"]) == 0, rounding=ROUNDING_MODE[m["rounding"]])

    def __str__(self) -> str:
        return (f"Simulated floating point format: mantissa bits = {self.mantissa}, exponent bits = {self.exponent}, "
                f"exponent bias = {self.bias}, unsigned = {self.unsigned}, \ncasting behavior: flush subnormal = "
                f"{self.flush_subnormal}, rounding = {self.rounding}")

    def __repr__(self) -> str:
        return (f"FP[{'0' if self.unsigned else '1'}|{self.exponent}|{self.mantissa},{self.bias}]"
                f"({'F' if self.flush_subnormal else '_'}{ROUNDING_MODE_INV[self.rounding]})")


class BlockFloatingPoint(Format):
    r"""Block floating point simulated in fp32 (reference format.py:273-397): one shared
    exponent per block of ``block_size`` elements along ``block_dim``."""

    blocked = True
    _RX = re.compile(r"^BFP\[(?P<precision>\d+)\|8\]\{(?P<block_size>\d+)\}\((?P<symmetric>\w)(?P<rounding>[A-Za-z])\)$")

    def __init__(self, precision=8, block_size=64, symmetric=True, rounding="nearest"):
        assert 2 <= precision <= 25, f"highest integer precision simulated by FP32 is 25, got {precision}"
        assert block_size > 0, f"block size has to be positive, got {block_size}"
        self.precision = precision
        self.block_size = block_size
        self.symmetric = symmetric
        self.rounding = rounding

    def cast(self, x: torch.Tensor, block_dim: int = -1):
        return ops.bfp_qdq(x, block_dim, self.block_size, self.precision, self.symmetric, self.rounding,
                           out_dtype=torch.float32)

    def stage(self):
        key = (self.block_size, self.precision, self.symmetric, self.rounding)
        return self._memo_stage(key, lambda: ops.bfp_stage(*key))

    # packed storage: the real format whose size `bytes_per_elem` reports (reference format.py:345-347)
    def pack(self, x: torch.Tensor):
        """-> (mantissas, exponents); blocks along the last dim; symmetric nearest formats, precision <= 8"""
        assert self.symmetric and self.rounding == "nearest" and self.precision <= 8, "packed storage: symmetric nearest, <= 8 bits"
        return ops.bfp_pack(x, self.block_size, self.precision)

    def unpack(self, mantissas: torch.Tensor, exponents: torch.Tensor, dtype=torch.float32):
        """dequantise packed storage; equals ``cast`` of the original tensor bit for bit"""
        return ops.bfp_unpack(mantissas, exponents, self.block_size, self.precision, dtype)

    @property
    def bytes_per_elem(self) -> float:
        return (self.precision + 8.0 / self.block_size) / 8.0

    @property
    def bit_precision(self) -> float:
        return self.precision + 8.0 / self.block_size

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise _bad(sh)
        return cls(precision=int(m["precision"]), block_size=int(m["block_size"]), symmetric=m["symmetric"] == "S",
                   rounding=ROUNDING_MODE[m["rounding"]])

    def __str__(self) -> str:
        return (f"Simulated block floating point format: precision bits = {self.precision}, block size = {self.block_size}\n"
                f"casting behavior: symmetric = {self.symmetric}, rounding = {self.rounding}")

    def __repr__(self) -> str:
        return f"BFP[{self.precision}|8]{{{self.block_size}}}({'S' if self.symmetric else '_'}{ROUNDING_MODE_INV[self.rounding]})"


class ScaledBlockFloatingPoint(Format):
    r"""Scaled block floating point (reference format.py:400-511): integer block format times a
    low-bit floating-point scaler per block."""

    blocked = True
    _RX = re.compile(r"^SBFP<(?P<block_format_sh>.+?)><(?P<scaler_format_sh>.+?)>\{(?P<block_size>\d+)\}$")

    def __init__(self, block_format: FixedPoint, scaler_format: FloatingPoint, block_size=64):
        assert isinstance(block_format, FixedPoint), "block format needs to be fixed point"
        assert isinstance(scaler_format, FloatingPoint), "scaler format needs to be floating point"
        assert block_format.fraction == 0, "block format needs to have zero fraction"
        assert block_format.symmetric, "block format needs to have symmetric range"
        assert block_size > 0, f"block size has to be positive, got {block_size}"
        self.block_format = block_format
        self.scaler_format = scaler_format
        self.block_size = block_size
        self.man_scaling = 2 ** (self.block_format.precision - 1) - 1
        # the reference picks the scaler bias from tensor values only when d-Matrix's private
        # `numerics` module is installed (format.py:13-20, 438-446); without it -- as here -- the
        # shorthand's constant bias is used.
        self.scaler_format_exponent_bias_determined = False

    def cast(self, x: torch.Tensor, block_dim: int = -1) -> torch.Tensor:
        self.scaler_format_exponent_bias_determined = True
        return ops.cast_chain(x, [self.stage()], block_dim, out_dtype=torch.float32)

    def stage(self):
        b, s = self.block_format, self.scaler_format
        # the block-scale rule goes with the tie rule (ops._scale_mode): tie "away" = the reference on CUDA tensors
        # (half-away XP rounding, scale = max * fp32(1/man_scaling)), tie "even" = on CPU tensors (ties-even, max / man_scaling);
        # `scale_mode` on the format overrides it
        key = (self.block_size, b.precision, b.clamp, b.rounding, b.tie, s.mantissa, s.exponent, s.bias, s.flush_subnormal, s.unsigned,
               s.rounding, getattr(self, "scale_mode", None))
        return self._memo_stage(key, lambda: ops.sbfp_stage(self.block_size, b.precision, b.clamp, b.rounding, b.tie, s.mantissa,
                                                            s.exponent, s.bias, s.flush_subnormal, s.unsigned,
                                                            repr(s) == "FP[1|5|10,15](FN)", s.rounding, key[-1]))

    # packed storage: the real format whose size `bytes_per_elem` reports (reference format.py:481-486)
    def pack(self, x: torch.Tensor, return_inexact=False):
        """-> (mantissas, scalers[, n_inexact]); blocks along the last dim.  Sign-magnitude mantissas (nibbles for
        precision <= 4) and one scaler byte per block; ``n_inexact`` counts blocks the bytes cannot hold."""
        return ops.sbfp_pack(x, self.stage(), return_inexact)

    def unpack(self, mantissas: torch.Tensor, scalers: torch.Tensor, dtype=torch.float32):
        """dequantise packed storage; equals ``cast`` of the original tensor bit for bit"""
        return ops.sbfp_unpack(mantissas, scalers, self.stage(), dtype)

    @property
    def bytes_per_elem(self) -> float:
        return self.block_format.bytes_per_elem + self.scaler_format.bytes_per_elem / self.block_size

    @property
    def bit_precision(self) -> float:
        return self.block_format.bit_precision + self.scaler_format.bit_precision / self.block_size

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise _bad(sh)
        return cls(block_format=FixedPoint.from_shorthand(m["block_format_sh"]),
                   scaler_format=FloatingPoint.from_shorthand(m["scaler_format_sh"]), block_size=int(m["block_size"]))

    def __str__(self) -> str:
        return (f"Simulated scaled block floating point format: block format = {self.block_format}, scaler format = "
                f"{self.scaler_format},\n block size = {self.block_size}")

    def __repr__(self) -> str:
        return f"SBFP<{repr(self.block_format)}><{repr(self.scaler_format)}>{{{self.block_size}}}"


class MXFP(Format):
    r"""MXFP (reference format.py:514-602): low-bit float elements times a power-of-two block scale."""

    blocked = True
    _RX = re.compile(r"^MXFP(?P<precision>\d+)\[E(?P<exponent>\d+)M(?P<mantissa>\d+)\]\{(?P<block_size>\d+)\}$")

    def __init__(self, element_format: FloatingPoint, block_size=32):
        assert isinstance(element_format, FloatingPoint), "block format needs to be floating point"
        assert block_size > 0, f"block size has to be positive, got {block_size}"
        self.element_format = element_format
        self.scaler_format = FloatingPoint(mantissa=0, exponent=8, bias=127, unsigned=True)
        self.block_size = block_size

    def cast(self, x: torch.Tensor, block_dim: int = -1) -> torch.Tensor:
        return ops.cast_chain(x, [self.stage()], block_dim, out_dtype=torch.float32)

    def stage(self):
        e = self.element_format
        assert e.bias == 2 ** (e.exponent - 1) - 1 and not e.flush_subnormal and not e.unsigned and e.rounding == "nearest", \
            "MXFP element formats are E<e>M<m> with the default bias, subnormals kept, nearest rounding"
        key = (self.block_size, e.mantissa, e.exponent)
        return self._memo_stage(key, lambda: ops.mxfp_stage(*key))

    @property
    def bytes_per_elem(self) -> float:
        return self.element_format.bytes_per_elem + self.scaler_format.bytes_per_elem / self.block_size

    @property
    def bit_precision(self) -> float:
        return self.element_format.bit_precision + 8.0 / self.block_size

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RX.match(sh)
        if m is None:
            raise _bad(sh)
        assert int(m["precision"]) == int(m["exponent"]) + int(m["mantissa"]) + 1
        e = int(m["exponent"])
        return cls(element_format=FloatingPoint(mantissa=int(m["mantissa"]), exponent=e, bias=2 ** (e - 1) - 1, flush_subnormal=False,
                                                unsigned=False, rounding="nearest"), block_size=int(m["block_size"]))

    def __str__(self) -> str:
        return (f"Simulated MXFP format: element format = {self.element_format}, scaler format = {self.scaler_format},\n"
                f" block size = {self.block_size}")

    def __repr__(self) -> str:
        e = self.element_format
        return f"MXFP{e.exponent + e.mantissa + 1}[E{e.exponent}M{e.mantissa}]{{{self.block_size}}}"

    def __reduce__(self):
        return (self.__class__, (self.element_format, self.block_size))


class MXINT(BlockFloatingPoint):
    r"""MXINT (reference format.py:603-653): BlockFloatingPoint, symmetric, nearest."""

    _RXM = re.compile(r"^MXINT(?P<precision>\d+)\{(?P<block_size>\d+)\}$")

    def __init__(self, precision=8, block_size=32):
        super().__init__(precision=precision, block_size=block_size, symmetric=True, rounding="nearest")

    @classmethod
    def from_shorthand(cls, sh: str):
        m = cls._RXM.match(sh)
        if m is None:
            raise _bad(sh)
        return cls(precision=int(m["precision"]), block_size=int(m["block_size"]))

    def __str__(self) -> str:
        return (f"Simulated MXINT format: precision bits = {self.precision}, block size = {self.block_size}\n"
                f"casting behavior: symmetric = {self.symmetric}, rounding = {self.rounding}")

    def __repr__(self) -> str:
        return f"MXINT{self.precision}{{{self.block_size}}}"

    def __reduce__(self):
        return (self.__class__, (self.precision, self.block_size))

"""CPU-side checks of the drop-in boundary: libdmxq.so loads, exports every symbol that
include/dmxq.h declares, validates arguments like the reference does, and the host-side mirror
of the reference interface parses / prints the same shorthands.  No compute calls (no GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from util import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dmxq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dmxq_[a-z0-9_]+)\s*\(", hdr))
    assert {"dmxq_cast_chain", "dmxq_bfp_qdq", "dmxq_sbfp_qdq", "dmxq_float_qdq", "dmxq_fixed_qdq", "dmxq_nm_prune",
            "dmxq_block_quantize", "dmxq_minmax", "dmxq_cast_chain_host", "dmxq_last_error", "dmxq_abi_version"} <= declared
    lib = C.CDLL(os.path.join(ROOT, "dmx_compressor_b200", "lib", "libdmxq.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"libdmxq.so does not export {name}"
    assert lib.dmxq_abi_version() == 3
    from dmx_compressor_b200 import _lib

    assert set(_lib.EXPORTS) == declared, "python binding and header disagree"


def test_struct_layout_matches_header():
    from dmx_compressor_b200 import _lib as L

    assert C.sizeof(L.Tensor) == 8 + 4 + 4 + 8 * 8 * 2
    assert C.sizeof(L.Stage) == 22 * 4 + 2 * 4 + 2 * 4 + 8 + 2 * 4
    assert L.Stage.vec.offset == 104 and L.Stage.vec_len.offset == 112


def test_argument_validation_without_a_device():
    """bad arguments are rejected before any CUDA call, with the reference's wording"""
    from dmx_compressor_b200 import _lib as L

    x = torch.zeros(4, 64)
    vx, vy = L.view(x), L.view(torch.zeros(4, 64))
    rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vy), -1, 64, 1, 1, 0, None, None)
    assert rc == -1 and b"highest integer precision" in L.lib.dmxq_last_error()
    rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vy), -1, 0, 8, 1, 0, None, None)
    assert rc == -1 and b"block size has to be positive" in L.lib.dmxq_last_error()
    rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vy), 5, 64, 8, 1, 0, None, None)
    assert rc == -1 and b"block_dim" in L.lib.dmxq_last_error()
    rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vy), -1, 64, 8, 1, 1, None, None)
    assert rc == -1 and b"random tensor" in L.lib.dmxq_last_error()
    rc = L.lib.dmxq_nm_prune(C.byref(vx), None, C.byref(vy), None, -1, 2, 5, 0, None)
    assert rc == -1 and b"not a multiple of block size" in L.lib.dmxq_last_error()
    with pytest.raises(AssertionError):
        L.check(rc)
    rc = L.lib.dmxq_nm_prune(C.byref(vx), None, C.byref(vy), None, -1, 2, 4, 7, None)
    assert rc == -1 and b"nm_order" in L.lib.dmxq_last_error()
    rc = L.lib.dmxq_float_qdq(C.byref(vx), C.byref(vy), 24, 8, 127, 0, 0, 0, 0, None, None)
    assert rc == -1
    vz = L.view(torch.zeros(4, 32))
    rc = L.lib.dmxq_bfp_qdq(C.byref(vx), C.byref(vz), -1, 64, 8, 1, 0, None, None)
    assert rc == -1 and b"same shape" in L.lib.dmxq_last_error()
    # empty tensors are a no-op
    ve = L.view(torch.zeros(0, 64))
    assert L.lib.dmxq_bfp_qdq(C.byref(ve), C.byref(ve), -1, 64, 8, 1, 0, None, None) == 0


def test_no_cpu_fallback():
    from dmx_compressor_b200.numerical import CastTo, Format
    from dmx_compressor_b200.sparse import Sparsify
    from dmx_compressor_b200 import quant

    x = torch.randn(4, 64)
    for call in (lambda: Format.from_shorthand("BFP[8|8]{64}(SN)").cast(x, -1), lambda: CastTo("FP[1|5|10,15](FN)")(x),
                 lambda: CastTo("XP[8,0](CSN)")(x), lambda: Sparsify(x.shape, "BTOPK{2:4,-1}(U)")(x),
                 lambda: quant.float_quantize(x, 5, 10, rounding="nearest"), lambda: quant.block_quantize(x, 8, rounding="nearest"),
                 lambda: quant.fixed_point_quantize(x, 8, 0, rounding="nearest")):
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            call()
    # packed storage too: no host encoder hides behind the CUDA one
    bfp, sbfp = Format.from_shorthand("BFP[8|8]{64}(SN)"), Format.from_shorthand("SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}")
    for call in (lambda: bfp.pack(x), lambda: sbfp.pack(x), lambda: sbfp.unpack(torch.zeros(4, 32, dtype=torch.uint8), torch.zeros(4, 4, dtype=torch.uint8))):
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            call()


def test_round2_entry_points_validate_without_a_gpu():
    """dmxq_softmax_cast / dmxq_cast_chain_philox / the SCALE stage reject what they do not take before any CUDA call, and have
    no host stand-in"""
    from dmx_compressor_b200 import _lib as L
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format

    x = torch.zeros(4, 64)
    assert not ops.softmax_supported(x)
    for call in (lambda: ops.softmax_cast(x), lambda: ops.cast_chain(x, [Format.from_shorthand("BFP[8|8]{64}(SS)").stage()], -1, philox=(1, 2)),
                 lambda: ops.scale_stage(torch.ones(64)), lambda: ops.philox_fill((4,), 1, device="cpu") and None):
        with pytest.raises((RuntimeError, AssertionError)):
            call()
    vx, vy = L.view(x), L.view(torch.zeros(4, 64))
    for n, msg in ((32, b"33..2048"), (4096, b"33..2048")):
        t = torch.zeros(4, n)
        v = L.view(t)
        assert L.lib.dmxq_softmax_cast(C.byref(v), None, C.byref(v), None, None, None, None, 0, None) == -2 and msg in L.lib.dmxq_last_error()
    st = Format.from_shorthand("BFP[8|8]{64}(SN)").stage()
    assert L.lib.dmxq_softmax_cast(C.byref(vx), None, C.byref(vy), None, None, None, C.byref(st), 9, None) == -1
    assert L.lib.dmxq_softmax_cast(C.byref(vx), None, C.byref(vy), C.byref(st), None, None, None, 0, None) == -1  # add stages without an addend
    vt = L.view(torch.zeros(64, 4).t())
    assert L.lib.dmxq_softmax_cast(C.byref(vt), None, C.byref(vt), None, None, None, None, 0, None) == -2  # softmax dim not contiguous
    sc = ops.make_stage(kind=L.ST_SCALE, vec=None, vec_len=64, vec_op=0)
    assert L.lib.dmxq_cast_chain(C.byref(vx), C.byref(vy), -1, C.byref(sc), 1, None, None, None, None) == -1 and b"device vector" in L.lib.dmxq_last_error()
    assert L.lib.dmxq_philox_fill(None, 16, 0, 1, 2, None) == -1


def test_packed_sbfp_argument_checks_without_a_gpu():
    """dmxq_sbfp_pack validates the format description before anything touches the device"""
    from dmx_compressor_b200 import _lib as L
    from dmx_compressor_b200 import ops
    import ctypes as C

    v = L.view(torch.zeros(4, 64))
    buf = (C.c_uint8 * 512)()
    ok = ops.sbfp_stage(16, 4, True, "nearest", L.TIE_AWAY, 4, 4, 7, True, True, False, "nearest")
    assert L.lib.dmxq_sbfp_pack(C.byref(v), buf, buf, None, None, None) == -1                      # null format
    bad = [ops.sbfp_stage(16, 4, True, "nearest", L.TIE_EVEN, 4, 4, 7, True, True, False, "nearest"),  # CPU tie rule
           ops.sbfp_stage(16, 4, True, "nearest", L.TIE_AWAY, 4, 4, 7, False, True, False, "nearest"),  # subnormal-keeping scaler
           ops.sbfp_stage(16, 4, True, "nearest", L.TIE_AWAY, 7, 4, 7, True, True, False, "nearest"),   # 11-bit scaler
           ops.sbfp_stage(12, 4, True, "nearest", L.TIE_AWAY, 4, 4, 7, True, True, False, "nearest"),   # block not a power of two
           ops.sbfp_stage(16, 4, False, "nearest", L.TIE_AWAY, 4, 4, 7, True, True, False, "nearest")]  # unclamped block format
    for st in bad:
        assert L.lib.dmxq_sbfp_pack(C.byref(v), buf, buf, C.byref(st), None, None) == -2, L.lib.dmxq_last_error()
    vr = L.view(torch.zeros(4, 72))
    assert L.lib.dmxq_sbfp_pack(C.byref(vr), buf, buf, C.byref(ok), None, None) == -1              # K % block != 0
    ve = L.view(torch.zeros(0, 64))
    assert L.lib.dmxq_sbfp_pack(C.byref(ve), buf, buf, C.byref(ok), None, None) == 0               # empty: nothing to do


SHORTHANDS = ["SAME", "XP[8,0](CSN)", "XP[4,0](CSN)", "XP[8,+4](C_U)", "XP[8,-2](_SD)", "FP[1|5|10,15](FN)", "FP[1|8|7,127](FN)",
              "FP[1|4|3,7](_N)", "FP[0|4|4,7](FN)", "BFP[8|8]{64}(SN)", "BFP[4|8]{128}(_N)", "BFP[24|8]{1}(SN)", "BFP[8|8]{64}(SS)",
              "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}", "MXINT8{32}"]


@pytest.mark.parametrize("sh", SHORTHANDS)
def test_format_shorthand_round_trip(sh):  # reference grammar, SURVEY.md appendix A
    from dmx_compressor_b200.numerical import Format

    assert repr(Format.from_shorthand(sh)) == sh


def test_format_errors_match_reference():
    from dmx_compressor_b200.numerical import BlockFloatingPoint, FixedPoint, FloatingPoint, Format
    from dmx_compressor_b200.sparse import BlockTopK, Sparseness

    with pytest.raises(ValueError, match="unrecognized format shorthand"):
        Format.from_shorthand("INT8")
    with pytest.raises(ValueError, match="unrecognized sparseness shorthand"):
        Sparseness.from_shorthand("NM{2:4}")
    with pytest.raises(AssertionError):
        BlockFloatingPoint(precision=1)
    with pytest.raises(AssertionError):
        BlockFloatingPoint(block_size=0)
    with pytest.raises(AssertionError):
        FixedPoint(25, 0)
    with pytest.raises(AssertionError):
        FloatingPoint(mantissa=24)
    with pytest.raises(AssertionError):
        FloatingPoint(mantissa=3, exponent=4, bias=-200)
    with pytest.raises(AssertionError):
        BlockTopK(K=5, block_size=4)
    for sh in ["DENSE", "TOPK{0.5}(U)", "BTOPK{2:4,-1}(U)", "BTOPK{4:8,1}(M)", "BERN"]:
        assert repr(Sparseness.from_shorthand(sh)) == sh


def test_castto_state_and_config_surface():
    from dmx_compressor_b200.numerical import CastTo, CastToDict, Same

    c = CastTo("XP[8,0](CSN)", qscheme=torch.per_channel_affine, ch_axis=1)
    assert set(c.state_dict()) >= {"scale", "zero_point", "fake_quant_enabled", "observer_enabled"}
    assert c._fq_on and not c._obs_on
    c.disable_fake_quant()
    sd = c.state_dict()
    d = CastTo("XP[8,0](CSN)", qscheme=torch.per_channel_affine, ch_axis=1)
    d.load_state_dict(sd)
    assert not d._fq_on
    assert torch.equal(d(torch.ones(2, 3)), torch.ones(2, 3))  # fake-quant off: identity, no kernel
    c.set_format("BFP[8|8]{64}(SN)")
    assert repr(c.format) == "BFP[8|8]{64}(SN)" and c.activation_post_process.dtype is c.format
    dd = CastToDict({"input_cast": CastTo(), "multiplier_cast": CastTo()})
    dd.set_format(["BFP[8|8]{64}(SN)", None])
    assert repr(dd["input_cast"].format) == "BFP[8|8]{64}(SN)" and isinstance(dd["multiplier_cast"].format, Same)
    x = torch.randn(3)
    y = CastTo("SAME")(x)
    assert torch.equal(x, y) and y.data_ptr() != x.data_ptr()  # SAME clones (reference format.py:89-90)


def test_packed_pair_contract_is_checked_on_the_host():
    """the unpack entry points take raw pointers: the python binding refuses (mantissas, per-block bytes) pairs whose shapes
    do not belong together before anything reaches the device"""
    from dmx_compressor_b200 import ops

    u8 = lambda *shape: torch.zeros(*shape, dtype=torch.uint8)
    assert ops._check_packed(u8(4, 32), u8(4, 4), 16, 4, "scalers") == 64            # nibbles: K = 2 * 32
    assert ops._check_packed(torch.zeros(4, 64, dtype=torch.int8), u8(4, 1), 64, 8, "exponents") == 64
    assert ops._check_packed(u8(2, 3, 128), u8(2, 3, 8), 16, 8, "scalers") == 128
    for mant, side, bs, prec in ((u8(4, 32), u8(4, 8), 16, 4),                        # too many scaler bytes
                                 (u8(4, 32), u8(3, 4), 16, 4),                        # other leading shape
                                 (u8(4, 32), torch.zeros(4, 4), 16, 4),               # not bytes
                                 (u8(4, 36), u8(4, 4), 16, 4),                        # K = 72 is not a whole number of blocks
                                 (u8(4, 64)[:, ::2], u8(4, 4), 16, 4)):               # strided mantissas
        with pytest.raises(RuntimeError, match="packed storage"):
            ops._check_packed(mant, side, bs, prec, "scalers")


def test_binding_signatures_match_header_prototypes():
    """every prototype in include/dmxq.h, parameter by parameter, against the ctypes argtypes / restype the
    binding installs: a width or arity mismatch here is silent stack corruption on the device path"""
    from dmx_compressor_b200 import _lib as L

    hdr = open(os.path.join(ROOT, "include", "dmxq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)

    def classify(decl):
        decl = " ".join(decl.replace("const", " ").split())
        if decl in ("void", ""):
            return None
        if "*" in decl:
            base = decl.split("*")[0].strip()
            return {"dmxq_tensor": C.POINTER(L.Tensor), "dmxq_stage": C.POINTER(L.Stage),
                    "char": C.c_char_p}.get(base, C.c_void_p)
        base = decl.rsplit(" ", 1)[0] if " " in decl else decl
        return {"int": C.c_int, "int32_t": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "float": C.c_float,
                "size_t": C.c_size_t}[base]

    protos = re.findall(r"([A-Za-z_][\w\s\*]*?)\b(dmxq_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) == len(L.EXPORTS)
    for ret, name, params in protos:
        fn = getattr(L.lib, name)
        want = [classify(p) for p in params.split(",")]
        want = [w for w in want if w is not None]
        assert list(fn.argtypes) == want, f"{name}: binding {fn.argtypes} vs header {want}"
        ret = " ".join(ret.replace("const", " ").replace("DMXQ_API", " ").replace("extern", " ").split())
        want_ret = None if ret == "void" else classify(ret + " r")
        assert fn.restype == want_ret, f"{name}: restype {fn.restype} vs header {ret!r}"

"""dmxq_softmax_cast (the attention block's last full-size passes): torch's CUDA softmax bit for bit,
with the mask add + its casts in front and the casts of the probabilities behind, all against the unfused sequence of the
same library calls / torch ops on the same GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format

DEV = "cuda:0"
F = lambda sh: Format.from_shorthand(sh).stage()
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def _bits(t):
    return t.contiguous().view(torch.int32 if t.element_size() == 4 else torch.int16)


def _scores(rows, n, dt, seed, scale=4.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(rows, n, device=DEV, generator=g) * scale
    x[::7] *= 8.0                                   # peaked rows: most probabilities underflow to 0 / denormals
    x[1::5, n // 3:] = float("-inf")                # causal-mask-like tails
    if rows > 3:
        x[3] = float("-inf")                        # a fully masked row: NaN in torch, NaN here
        x[2, 5] = float("inf")
    return x.to(dt)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("n", [40, 64, 72, 128, 200, 256, 504, 512, 1000, 1024, 1536, 2040, 2048])
def test_softmax_equals_torch_bitwise(dt, n):
    x = _scores(300, n, DT[dt], n)
    want = torch.softmax(x, dim=-1)
    got = ops.softmax_cast(x)
    wn, gn = torch.isnan(want), torch.isnan(got)
    assert torch.equal(wn, gn)
    assert torch.equal(_bits(torch.where(wn, torch.zeros_like(want), want)), _bits(torch.where(gn, torch.zeros_like(got), got)))


@pytest.mark.parametrize("dt", list(DT))
def test_softmax_nan_and_inf_rows_match_torch(dt):
    """a NaN (or +inf) anywhere in a row makes the whole row NaN in torch's kernel (the sum is NaN); the maximum is taken with
    fmaxf here instead of torch's comparison chain -- same rows, same NaN pattern after rounding to the dtype; the other rows stay
    bit-identical.  Rows whose spread exceeds 110 exercise the skipped (exactly zero) exponentials and the literal division."""
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(64, 2048, device=DEV, generator=g) * 5
    x[3, 17] = float("nan")
    x[5, 2047] = float("nan")          # last element of its lane: torch's chain keeps the NaN as that lane's maximum
    x[7, 0] = float("nan")             # first element: torch's chain drops it from the maximum
    x[9, 100] = float("inf")
    x[11, :] = float("-inf")
    x[13, 64:] -= 400.0                # tails far below the maximum: exp underflows to exactly 0
    x[15, 1:] -= 95.0                  # quotients in the denormal range: the literal division path
    x = x.to(DT[dt])
    want, got = torch.softmax(x, -1), ops.softmax_cast(x)
    assert torch.equal(torch.isnan(want), torch.isnan(got))
    assert torch.equal(_bits(want), _bits(got))  # NaN rows included: both produce the canonical NaN of the dtype
    post = [F("FP[1|5|10,15](FN)"), F("BFP[8|8]{64}(SN)")]
    assert torch.equal(_bits(ops.cast_chain(want, post, -1)), _bits(ops.softmax_cast(x, post)))


@pytest.mark.parametrize("dt", ["float32", "bfloat16"])
def test_softmax_equals_torch_at_scale(dt):
    """the OPT-125m attention shape: [96, 2048, 2048] probabilities"""
    g = torch.Generator(device=DEV).manual_seed(3)
    x = (torch.randn(96, 2048, 2048, device=DEV, generator=g) * 3).to(DT[dt])
    mask = torch.full((2048, 2048), float("-inf"), device=DEV).triu(1).to(DT[dt])
    x += mask
    want = torch.softmax(x, dim=-1)
    got = ops.softmax_cast(x)
    assert torch.equal(_bits(want), _bits(got))


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("post", [["FP[1|5|10,15](FN)"], ["FP[1|5|10,15](FN)", "BFP[8|8]{64}(SN)"], ["BFP[4|8]{64}(SN)"], ["BFP[8|8]{16}(SN)"],
                                  ["FP[1|4|3,7](_N)"], ["FP[1|5|10,15](FN)", "MXFP8[E4M3]{32}"], ["XP[8,+7](CSN)"],
                                  ["FP[1|5|10,15](FN)", "SBFP<XP[4,0](CSN)><FP[0|4|4,7](FN)>{16}"]])
def test_softmax_post_chain_equals_unfused(dt, post):
    for n in (128, 1024, 2048):
        x = _scores(257, n, DT[dt], 11 + n, scale=2.0)
        x[3] = 0.0  # (no NaN rows here: NaN payloads through the casts are compared in the parity suite)
        st = [F(s) for s in post]
        want = ops.cast_chain(torch.softmax(x, dim=-1), st, -1)
        got = ops.softmax_cast(x, st)
        wn, gn = torch.isnan(want), torch.isnan(got)  # (MX: all-zero blocks are 0 / 0)
        assert torch.equal(wn, gn)
        assert torch.equal(_bits(torch.where(wn, torch.zeros_like(want), want)), _bits(torch.where(gn, torch.zeros_like(got), got))), (dt, post, n)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("with_casts", [True, False])
def test_softmax_with_mask_add_equals_unfused(dt, with_casts):
    """ResAdd (attention-mask add, FLOAT16 casts on both inputs and the output) -> Softmax -> FLOAT16 -> BFP16, one launch"""
    B, H, S = 2, 3, 256
    g = torch.Generator(device=DEV).manual_seed(5)
    x = (torch.randn(B, H, S, S, device=DEV, generator=g) * 3).to(DT[dt])
    x.view(-1)[::97] *= 1e-6  # values the FLOAT16 cast flushes
    mask = torch.full((S, S), float("-inf"), device=DEV).triu(1).to(DT[dt])[None, None].expand(B, 1, S, S)
    mask = (mask + torch.zeros(B, 1, S, S, device=DEV, dtype=DT[dt])).contiguous()  # a real [B,1,S,S] tensor, H broadcast
    f16 = F("FP[1|5|10,15](FN)") if with_casts else None
    post = [F("FP[1|5|10,15](FN)"), F("BFP[8|8]{64}(SN)")]
    z = ops.add_cast(x, mask.expand(B, H, S, S), f16, f16, f16)
    want = ops.cast_chain(torch.softmax(z, dim=-1), post, -1)
    got = ops.softmax_cast(x, post, addend=mask.expand(B, H, S, S), stage_x=f16, stage_addend=f16, stage_sum=f16)
    assert torch.equal(_bits(want), _bits(got))
    # a 2-d mask broadcast over batch and heads, and a full-shape addend
    m2 = mask[0, 0].contiguous()
    z = ops.add_cast(x, m2, f16, f16, f16)
    assert torch.equal(_bits(ops.cast_chain(torch.softmax(z, dim=-1), post, -1)), _bits(ops.softmax_cast(x, post, addend=m2, stage_x=f16, stage_addend=f16, stage_sum=f16)))
    full = (torch.randn(B, H, S, S, device=DEV, generator=g)).to(DT[dt])
    z = ops.add_cast(x, full, f16, f16, f16)
    assert torch.equal(_bits(ops.cast_chain(torch.softmax(z, dim=-1), post, -1)), _bits(ops.softmax_cast(x, post, addend=full, stage_x=f16, stage_addend=f16, stage_sum=f16)))


def test_softmax_refuses_what_torch_runs_differently():
    for shape in ((4, 32), (4, 4096), (4, 100)):  # <= 32: sub-warp rows; > 2048: ATen's block softmax; 100 bf16: not whole vectors
        x = torch.randn(*shape, device=DEV, dtype=torch.bfloat16)
        assert not ops.softmax_supported(x)
        with pytest.raises(RuntimeError, match="unsupported"):
            ops.softmax_cast(x)
    x = torch.randn(8, 256, device=DEV)
    assert ops.softmax_supported(x) and not ops.softmax_supported(x.t()) and not ops.softmax_supported(x, 0)
    with pytest.raises(RuntimeError):
        ops.softmax_cast(x, [F("BFP[8|8]{64}(SN)")] * 5)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.softmax_cast(torch.randn(8, 96, device=DEV), [F("BFP[8|8]{64}(SN)")])  # 96 is not a whole number of blocks

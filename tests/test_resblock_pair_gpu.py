"""Parity of the fused ResBlock-pair kernel (conv1 -> leaky-ReLU -> conv2 -> fused epilogue, the intermediate tile
kept in shared memory) through the C ABI (m2s_resblock_pair_fwd) against an fp64 restatement of models.py:36-48 on
the same fp16-rounded operands.  The intermediate is rounded to fp16 exactly where the kernel rounds it; what remains
is fp32 accumulation order plus rare 1-ulp flips of that rounding, hence 2e-3 of the output scale."""
import pytest
import torch

from tests.util import multi_tap_reference

pytestmark = pytest.mark.gpu

CASES = [
    # B, L, C, k, dilation
    (2, 1000, 64, 3, 1),
    (2, 1000, 64, 3, 3),
    (1, 5000, 64, 11, 5),
    (3, 777, 32, 7, 3),          # 64-byte rows (SWIZZLE_64B), ragged tile edge
    (1, 40000, 32, 11, 5),       # many tiles per CTA (double-buffered accumulators / T tile)
    (2, 3000, 128, 3, 1),        # N = 128: single-buffered
    (1, 2000, 128, 11, 5),
    (1, 100, 64, 7, 5),          # shorter than one tile
]


def _lrelu(v, s):
    return torch.where(v >= 0, v, v * s)


def _reference(x16, w1, b1, d, w2, b2, L):
    k1, k2 = w1.shape[0], w2.shape[0]
    t = multi_tap_reference(x16.float(), w1, [-(k1 - 1 - j) * d for j in range(k1)], L) + b1.double().cpu()
    t = _lrelu(t, 0.1).float().half().float()            # the fp16 T tile
    return multi_tap_reference(t, w2, [-(k2 - 1 - j) for j in range(k2)], L) + b2.double().cpu()


def _inputs(B, L, C, k, seed=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, C, generator=g).half()
    w1 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).half().float()
    w2 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).half().float()
    b1 = torch.randn(C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    return [t.cuda() for t in (x, w1, b1, w2, b2)]


@pytest.mark.parametrize("case", CASES)
def test_fused_pair_matches_two_convs(case):
    from mri2speech_b200 import _lib
    B, L, C, k, d = case
    _lib.set_knob("fuse_max_n", 128)  # N = 128 is supported (single-buffered) though the Generator does not use it
    x, w1, b1, w2, b2 = _inputs(B, L, C, k)
    ref = _reference(x, w1, b1, d, w2, b2, L)
    d16 = torch.zeros(B, L, C, device="cuda", dtype=torch.float16)
    out = _lib.resblock_pair_fwd(x, w1, b1, d, w2, b2, out16=d16)
    scale = max(1.0, ref.abs().max().item())
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < 2e-3 * scale, (case, err)
    assert (d16.double().cpu() - ref).abs().max().item() < 3e-3 * scale


def test_fused_pair_resblock_epilogue_and_mask():
    """The vocoder's conv2 programs: residual through the inverse leaky-ReLU, MRF accumulate, 1/3 scale, leaky-ReLU,
    ragged length mask; fp16-only output."""
    from mri2speech_b200 import _lib
    B, L, C, k, d = 3, 900, 64, 7, 3
    x, w1, b1, w2, b2 = _inputs(B, L, C, k, seed=9)
    g = torch.Generator().manual_seed(10)
    res = torch.randn(B, L, C, generator=g).cuda()
    acc = torch.randn(B, L, C, generator=g).cuda()
    lens = torch.tensor([900, 411, 5], dtype=torch.int32).cuda()
    d16 = torch.full((B, L, C), 3.0, device="cuda", dtype=torch.float16)
    _lib.resblock_pair_fwd(x, w1, b1, d, w2, b2, res=res, res_inv_slope=10.0, accum=acc, out_scale=1.0 / 3.0,
                           act=_lib.ACT_LRELU, act_slope=0.01, lens=lens, out16=d16, want_d32=False)
    conv = _reference(x, w1, b1, d, w2, b2, L)
    r = res.double().cpu()
    v = (conv + torch.where(r >= 0, r, r * 10.0) + acc.double().cpu()) / 3.0
    v = _lrelu(v, 0.01)
    t = torch.arange(L).view(1, L, 1)
    v = v * (t < lens.cpu().view(B, 1, 1)).double()
    assert (d16.double().cpu() - v).abs().max().item() < 3e-3 * max(1.0, v.abs().max().item())


def test_fused_pair_plain_residual_fp32_out():
    from mri2speech_b200 import _lib
    B, L, C, k, d = 2, 2000, 32, 3, 1
    x, w1, b1, w2, b2 = _inputs(B, L, C, k, seed=4)
    res = torch.randn(B, L, C, generator=torch.Generator().manual_seed(2)).cuda()
    out = _lib.resblock_pair_fwd(x, w1, b1, d, w2, b2, res=res, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1)
    r = res.double().cpu()
    v = _lrelu(_reference(x, w1, b1, d, w2, b2, L) + torch.where(r >= 0, r, r * 10.0), 0.1)
    assert (out.double().cpu() - v).abs().max().item() < 2e-3 * max(1.0, v.abs().max().item())


def test_unsupported_pair_is_refused():
    from mri2speech_b200 import _lib
    _lib.set_knob("fuse_max_n", 128)
    x, w1, b1, w2, b2 = _inputs(1, 300, 256, 3)          # N = 256 > 128
    with pytest.raises(_lib.M2SError):
        _lib.resblock_pair_fwd(x, w1, b1, 1, w2, b2)


@pytest.mark.parametrize("case", [(2, 3000, 64, 3, 1), (2, 5000, 32, 11, 5), (1, 2500, 64, 7, 3)])
def test_fused_pair_split_fp16_stream(case):
    """The fused pair on the split-fp16 residual stream: x arrives as (hi, lo) planes (hi is conv1's operand, hi + lo
    the residual), the result leaves as (hi, lo) planes in DIFFERENT buffers (the kernel re-reads a halo of x)."""
    from mri2speech_b200 import _lib
    B, L, C, k, d = case
    x, w1, b1, w2, b2 = _inputs(B, L, C, k, seed=13)
    g = torch.Generator().manual_seed(14)
    x_lo = (torch.randn(B, L, C, generator=g) * 0.4).half().cuda()      # any lo plane: the residual is hi + lo
    res = x.float() + x_lo.float() / 2048.0
    out_hi = torch.full((B, L, C), 9.0, device="cuda", dtype=torch.float16)
    out_lo = torch.full((B, L, C), 9.0, device="cuda", dtype=torch.float16)
    _lib.resblock_pair_fwd(x, w1, b1, d, w2, b2, res_hi=x, res_lo=x_lo, res_inv_slope=10.0, act=_lib.ACT_LRELU,
                           act_slope=0.1, out16=out_hi, out16_lo=out_lo, want_d32=False)
    ref = _lib.resblock_pair_fwd(x, w1, b1, d, w2, b2, res=res, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1)
    rec = out_hi.float() + out_lo.float() / 2048.0
    assert (rec - ref).abs().max().item() < 2e-6 * max(1.0, ref.abs().max().item())
    r = res.double().cpu()
    v = _lrelu(_reference(x, w1, b1, d, w2, b2, L) + torch.where(r >= 0, r, r * 10.0), 0.1)
    assert (rec.double().cpu() - v).abs().max().item() < 2e-3 * max(1.0, v.abs().max().item())

"""Parity of the conv engine with fp16 operands (tcgen05 kind::f16, fp32 accumulate) through the C ABI against an
fp64 restatement of its contract evaluated on the SAME fp16-rounded operands: what remains is fp32 accumulation
order, so the tolerance is tight (1e-4 relative to the output scale).  Covers both SMEM layouts (64 halves per
128-byte row, SWIZZLE_128B; 32 halves per 64-byte row, SWIZZLE_64B for c_in <= 32), the single-CTA and the CTA-pair
kernel, row-shifted taps, ragged M, and the fp16 second output."""
import pytest
import torch

from tests.util import multi_tap_reference

pytestmark = pytest.mark.gpu

CASES = [
    # B, L, Cin, N, shifts
    (2, 300, 64, 64, [0]),
    (2, 300, 32, 32, [0]),                       # 64-byte rows (SWIZZLE_64B)
    (2, 300, 32, 32, [-2, -1, 0]),               # 64-byte rows, row-shifted descriptors
    (1, 2000, 32, 32, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0]),
    (2, 300, 16, 16, [-1, 0]),                   # c_in < one row
    (3, 700, 64, 64, [-2, -1, 0]),
    (2, 1000, 128, 128, [-18, -15, -12, -9, -6, -3, 0]),
    (2, 1500, 256, 256, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0]),
    (2, 100, 64, 512, [0, 1, 2, 3, 4, 5, 6]),
    (4, 4000, 256, 256, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0]),   # CTA-pair kernel, N = 256
    (1, 20000, 128, 128, [-6, -3, 0]),                                           # CTA-pair kernel, N = 128
    (1, 40000, 64, 64, [-10, -5, 0]),                                            # CTA-pair (operand-bound narrow layer)
    (1, 60000, 32, 32, [-10, -9, -8, -7, -6, -5, -4, -3, -2, -1, 0]),            # CTA-pair, 64-byte rows
    (2, 5003, 64, 320, [1, 0, -1]),                                              # CTA-pair, 2 N tiles, ragged M
    (1, 20001, 208, 208, [0]),                                                   # K tail: 208 = 3 x 64 + 16
    (2, 129, 56, 104, [0]),                                                      # K tail 56 (not a multiple of 16)
    (1, 5, 32, 32, [-2, -1, 0]),
]


def _inputs(B, L, C, N, taps, seed=11):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, L, C, generator=g).half()
    w = (torch.randn(taps, N, C, generator=g) / (C * taps) ** 0.5).half().float()
    bias = torch.randn(N, generator=g)
    return a.cuda(), w.cuda(), bias.cuda()


@pytest.mark.parametrize("case", CASES)
def test_fp16_conv_matches_contract(case):
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = case
    a, w, bias = _inputs(B, L, C, N, len(shifts))
    ref = multi_tap_reference(a.float(), w, shifts, L) + bias.double().cpu()
    d16 = torch.zeros(B, L, N, device="cuda", dtype=torch.float16)
    d = _lib.conv_fwd(a, w, shifts, L, bias=bias, out16=d16)
    scale = max(1.0, ref.abs().max().item())
    err = (d.double().cpu() - ref).abs().max().item()
    assert err < 1e-4 * scale, (case, err)
    err16 = (d16.double().cpu() - ref).abs().max().item()
    assert err16 < 1.5e-3 * scale, (case, err16)  # fp16 output rounding: 2^-11 relative


def test_fp16_only_output_and_epilogue():
    """fp16-only output (no fp32 store) with the full epilogue: residual through the inverse leaky-ReLU, accumulate,
    scale, leaky-ReLU, length mask."""
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = 3, 400, 64, 64, [-6, -3, 0]
    a, w, bias = _inputs(B, L, C, N, 3, seed=5)
    g = torch.Generator().manual_seed(6)
    res = torch.randn(B, L, N, generator=g).cuda()
    acc = torch.randn(B, L, N, generator=g).cuda()
    lens = torch.tensor([400, 123, 7], dtype=torch.int32).cuda()
    d16 = torch.full((B, L, N), 7.0, device="cuda", dtype=torch.float16)
    _lib.conv_fwd(a, w, shifts, L, bias=bias, res=res, res_inv_slope=10.0, accum=acc, out_scale=1.0 / 3.0,
                  act=_lib.ACT_LRELU, act_slope=0.1, lens=lens, out16=d16, want_d32=False)
    conv = multi_tap_reference(a.float(), w, shifts, L) + bias.double().cpu()
    r = res.double().cpu()
    r = torch.where(r >= 0, r, r * 10.0)
    v = (conv + r + acc.double().cpu()) / 3.0
    v = torch.where(v >= 0, v, v * 0.1)
    t = torch.arange(L).view(1, L, 1)
    v = v * (t < lens.cpu().view(B, 1, 1)).double()
    assert (d16.double().cpu() - v).abs().max().item() < 2e-3 * max(1.0, v.abs().max().item())


def test_fp16_saturates_instead_of_inf():
    from mri2speech_b200 import _lib
    a = torch.full((1, 128, 64), 100.0, device="cuda", dtype=torch.float16)
    w = torch.full((1, 64, 64), 100.0, device="cuda")
    d16 = torch.zeros(1, 128, 64, device="cuda", dtype=torch.float16)
    d = _lib.conv_fwd(a, w, [0], 128, out16=d16)
    assert torch.isfinite(d16).all() and d16.max().item() == 65504.0
    assert abs(d.max().item() - 640000.0) < 1.0


def test_operand_format_mismatch_is_refused():
    from mri2speech_b200 import _lib
    a = torch.zeros(1, 128, 12, device="cuda", dtype=torch.float16)  # c_in % 8 != 0
    w = torch.zeros(1, 16, 12, device="cuda")
    with pytest.raises(_lib.M2SError):
        _lib.conv_fwd(a, w, [0], 128)


def _split(v):
    """(hi, lo) fp16 planes of an fp32 tensor: hi = fp16(v), lo = fp16((v - hi) * 2048)."""
    hi = v.half()
    lo = ((v - hi.float()) * 2048.0).half()
    return hi, lo


@pytest.mark.parametrize("case", [(3, 400, 64, 64, [-6, -3, 0], True), (2, 700, 32, 32, [-2, -1, 0], False),
                                  (1, 30000, 128, 128, [-10, -5, 0], True), (2, 333, 256, 256, [-1, 0], False)])
def test_split_fp16_residual_stream(case):
    """Split-fp16 residual stream (include/m2s.h): residual read as float(hi) + float(lo), output written as the
    (hi, lo) planes.  hi must be bit-identical to the plain fp16 output, hi + lo must carry the fp32 result to
    ~2^-21 relative, and the residual read must agree with the fp32-residual program on the same values."""
    from mri2speech_b200 import _lib
    B, L, C, N, shifts, with_acc = case
    a, w, bias = _inputs(B, L, C, N, len(shifts), seed=21)
    g = torch.Generator().manual_seed(22)
    res = torch.randn(B, L, N, generator=g).cuda()
    res_hi, res_lo = _split(res)
    res_q = res_hi.float() + res_lo.float() / 2048.0          # what the split stream carries
    assert (res_q - res).abs().max().item() < 4e-6
    acc = torch.randn(B, L, N, generator=g).cuda() if with_acc else None
    lens = torch.tensor(([L, L // 3, 7] * B)[:B], dtype=torch.int32).cuda()
    kw = dict(bias=bias, res_inv_slope=10.0, accum=acc, out_scale=1.0 / 3.0 if with_acc else 1.0,
              act=_lib.ACT_LRELU, act_slope=0.1, lens=lens)
    # fp32-residual program on the values the split stream carries: the reference for this test
    ref16 = torch.zeros(B, L, N, device="cuda", dtype=torch.float16)
    ref = _lib.conv_fwd(a, w, shifts, L, res=res_q, out16=ref16, **kw)
    hi = torch.full((B, L, N), 5.0, device="cuda", dtype=torch.float16)
    lo = torch.full((B, L, N), 5.0, device="cuda", dtype=torch.float16)
    _lib.conv_fwd(a, w, shifts, L, res_hi=res_hi, res_lo=res_lo, out16=hi, out16_lo=lo, want_d32=False, **kw)
    assert torch.equal(hi, ref16)
    rec = hi.float() + lo.float() / 2048.0
    scale = max(1.0, ref.abs().max().item())
    assert (rec - ref).abs().max().item() < 2e-6 * scale
    # rows past the length mask are zero in both planes
    t = torch.arange(L, device="cuda").view(1, L, 1)
    dead = (t >= lens.view(B, 1, 1)).expand(B, L, N)
    if dead.any():
        assert hi[dead].abs().max().item() == 0 and lo[dead].abs().max().item() == 0


def test_split_fp16_residual_refuses_unsupported_epilogues():
    from mri2speech_b200 import _lib
    a, w, bias = _inputs(1, 200, 64, 64, 1, seed=2)
    res_hi, res_lo = _split(torch.randn(1, 200, 64).cuda())
    d16 = torch.zeros(1, 200, 64, device="cuda", dtype=torch.float16)
    with pytest.raises(_lib.M2SError):   # SiLU is not one of the two ResBlock programs
        _lib.conv_fwd(a, w, [0], 200, bias=bias, res_hi=res_hi, res_lo=res_lo, act=_lib.ACT_SILU, out16=d16)
    with pytest.raises(_lib.M2SError):   # lo plane missing
        _lib.conv_fwd(a, w, [0], 200, bias=bias, res_hi=res_hi, out16=d16)
    with pytest.raises(_lib.M2SError):   # lo output without the hi output
        _lib.conv_fwd(a, w, [0], 200, bias=bias, out16_lo=d16)

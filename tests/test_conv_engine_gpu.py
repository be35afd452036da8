"""Parity of the multi-tap implicit-GEMM conv engine (tcgen05 and CUDA-core builds) through the C ABI
(m2s_conv_fwd) against an fp64 restatement of its contract.  TF32 tolerance: operands carry 10 mantissa
bits -> relative error <= ~2^-10 per product, stated per test."""
import pytest
import torch

from tests.util import multi_tap_reference

pytestmark = pytest.mark.gpu

CASES = [
    # B, L, Cin, N, shifts
    (2, 300, 32, 32, [0]),
    (1, 1000, 128, 256, [0]),
    (2, 300, 16, 16, [0]),
    (1, 257, 208, 512, [0]),
    (3, 700, 64, 64, [-2, -1, 0]),
    (2, 1000, 128, 128, [-18, -15, -12, -9, -6, -3, 0]),
    (2, 1500, 256, 256, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0]),
    (2, 100, 64, 512, [0, 1, 2, 3, 4, 5, 6]),
    (2, 500, 64, 320, [1, 0, -1]),
    (4, 4000, 256, 256, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0]),   # CTA-pair kernel, N = 256
    (1, 20000, 128, 128, [-6, -3, 0]),                                           # CTA-pair kernel, N = 128
    (2, 3000, 64, 512, [0, 1, 2, 3, 4, 5, 6]),                                    # CTA-pair kernel, 2 N-tiles
    (2, 5003, 64, 320, [1, 0, -1]),                                               # CTA-pair, N = 320 -> 2 x 160, ragged M
    (1, 20001, 208, 208, [0]),                                                    # CTA-pair GEMM, odd channel counts
    (1, 5, 32, 32, [-2, -1, 0]),          # shorter than one tile
    (2, 129, 56, 104, [0]),               # odd encoder channel counts (N padded to 16 inside)
]


def _inputs(B, L, C, N, taps, seed=11):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, L, C, generator=g)
    w = torch.randn(taps, N, C, generator=g) / (C * taps) ** 0.5
    bias = torch.randn(N, generator=g)
    return a.cuda(), w.cuda(), bias.cuda()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_conv_matches_contract(case, impl):
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = case
    a, w, bias = _inputs(B, L, C, N, len(shifts))
    ref = multi_tap_reference(a, w, shifts, L) + bias.double().cpu()
    d = _lib.conv_fwd(a, w, shifts, L, impl=_lib.IMPL_TCGEN05 if impl == "tcgen05" else _lib.IMPL_SIMT, bias=bias)
    err = (d.double().cpu() - ref).abs().max().item()
    tol = 2e-5 if impl == "simt" else 6e-3 * max(1.0, ref.abs().max().item())
    assert err < tol, (case, impl, err)


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_fused_epilogue(impl):
    """bias + inverse-leaky-ReLU residual + accumulate + scale + leaky-ReLU + length mask."""
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = 3, 400, 64, 64, [-6, -3, 0]
    a, w, bias = _inputs(B, L, C, N, 3, seed=5)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, L, N, generator=g)
    res = torch.nn.functional.leaky_relu(x, 0.1).cuda()
    acc = torch.randn(B, L, N, generator=g).cuda()
    lens = torch.tensor([100, 3, 57], dtype=torch.int32).cuda()
    ref = multi_tap_reference(a, w, shifts, L) + bias.double().cpu() + x.double() + acc.double().cpu()
    ref = torch.nn.functional.leaky_relu(ref / 3.0, 0.01)
    t = torch.arange(L).view(1, L, 1)
    ref = ref * (t < (lens.cpu().view(-1, 1, 1) * 4)).double()
    d = _lib.conv_fwd(a, w, shifts, L, impl=_lib.IMPL_TCGEN05 if impl == "tcgen05" else _lib.IMPL_SIMT,
                      bias=bias, res=res, res_inv_slope=10.0, accum=acc, out_scale=1.0 / 3.0,
                      act=_lib.ACT_LRELU, act_slope=0.01, lens=lens, len_scale=4)
    err = (d.double().cpu() - ref).abs().max().item()
    assert err < (5e-5 if impl == "simt" else 6e-3), err


def test_pair_kernel_fused_epilogue_and_matches_single_cta():
    """cta_group::2 kernel with residual + accumulate + scale + activation + length mask; must agree with the
    single-CTA kernel to fp32 accumulation-order noise."""
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = 3, 6000, 128, 128, [-18, -15, -12, -9, -6, -3, 0]
    a, w, bias = _inputs(B, L, C, N, len(shifts), seed=9)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, L, N, generator=g)
    res = torch.nn.functional.leaky_relu(x, 0.1).cuda()
    acc = torch.randn(B, L, N, generator=g).cuda()
    lens = torch.tensor([1500, 40, 777], dtype=torch.int32).cuda()
    kw = dict(bias=bias, res=res, res_inv_slope=10.0, accum=acc, out_scale=1.0 / 3.0, act=_lib.ACT_LRELU,
              act_slope=0.01, lens=lens, len_scale=4)
    ref = multi_tap_reference(a, w, shifts, L) + bias.double().cpu() + x.double() + acc.double().cpu()
    ref = torch.nn.functional.leaky_relu(ref / 3.0, 0.01)
    t = torch.arange(L).view(1, L, 1)
    ref = ref * (t < (lens.cpu().view(-1, 1, 1) * 4)).double()
    _lib.set_knob("pair", 1)
    d_pair = _lib.conv_fwd(a, w, shifts, L, **kw).cpu()
    _lib.set_knob("pair", 0)
    try:
        d_single = _lib.conv_fwd(a, w, shifts, L, **kw).cpu()
    finally:
        _lib.set_knob("pair", 1)
    assert (d_pair.double() - ref).abs().max().item() < 6e-3
    assert (d_pair - d_single).abs().max().item() < 1e-4


def test_pitch_mask_and_row_offset():
    """3x3 stride-1 conv over a zero-bordered image flattened to rows (the encoder's use of the engine)."""
    from mri2speech_b200 import _lib
    N_img, H, W, C, Co = 2, 12, 20, 32, 48
    pitch = W + 2
    g = torch.Generator().manual_seed(3)
    img = torch.randn(N_img, C, H, W, generator=g)
    wt = torch.randn(Co, C, 3, 3, generator=g) / (9 * C) ** 0.5
    ref = torch.nn.functional.conv2d(img.double(), wt.double(), padding=1)          # (N, Co, H, W)
    padded = torch.nn.functional.pad(img, (1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(N_img, (H + 2) * pitch, C)
    w_eng = wt.permute(2, 3, 0, 1).reshape(9, Co, C).contiguous()
    shifts = [dy * pitch + dx for dy in range(3) for dx in range(3)]
    for impl in (_lib.IMPL_SIMT, _lib.IMPL_TCGEN05):
        out = torch.full((N_img, (H + 2) * pitch, Co), 7.0).cuda()
        out[:, : pitch + 1] = 0
        out[:, H * pitch + pitch + 1:] = 0
        _lib.conv_fwd(padded.cuda().contiguous(), w_eng.cuda(), shifts, H * pitch, impl=impl,
                      pitch_mask=(pitch, 1, H + 1, 1, W + 1), d_row_offset=pitch + 1, out=out)
        got = out.cpu().view(N_img, H + 2, pitch, Co)
        assert got[:, 0].abs().max() == 0 and got[:, -1].abs().max() == 0
        assert got[:, :, 0].abs().max() == 0 and got[:, :, -1].abs().max() == 0
        inner = got[:, 1:H + 1, 1:W + 1].permute(0, 3, 1, 2).double()
        err = (inner - ref).abs().max().item()
        assert err < (2e-5 if impl == _lib.IMPL_SIMT else 6e-3), (impl, err)

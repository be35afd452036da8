"""GPU: no kernel writes outside its output.  Outputs are allocated with sentinel-filled guard rows after every batch
item (compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are hunted this way): ragged M (rows
that do not fill a tile), ragged N, the fp16 second output, the fused pair kernel's partially kept tiles."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SENTINEL = 12345.0


def _guarded(B, rows, guard, N, dtype):
    return torch.full((B, rows + guard, N), SENTINEL, device="cuda", dtype=dtype)


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("case", [(2, 131, 64, 64, [-2, -1, 0]), (1, 5003, 64, 320, [1, 0, -1]), (3, 300, 32, 32, [0]),
                                  (2, 777, 128, 104, [0]), (1, 20001, 128, 128, [-6, -3, 0])])
def test_engine_respects_output_bounds(case, half):
    from mri2speech_b200 import _lib
    B, L, C, N, shifts = case
    g = torch.Generator().manual_seed(1)
    a = torch.randn(B, L, C, generator=g).cuda()
    if half:
        a = a.half()
    w = (torch.randn(len(shifts), N, C, generator=g) / (C * len(shifts)) ** 0.5).cuda()
    guard = 9
    d32 = _guarded(B, L, guard, N, torch.float32)
    d16 = _guarded(B, L, guard, N, torch.float16)
    _lib.conv_fwd(a, w, shifts, L, out=d32, out16=d16 if half else None, d_rows=L + guard)
    assert torch.all(d32[:, L:] == SENTINEL), "fp32 output: guard rows overwritten"
    assert not torch.any(d32[:, :L] == SENTINEL)
    if half:
        assert torch.all(d16[:, L:] == SENTINEL), "fp16 output: guard rows overwritten"


@pytest.mark.parametrize("case", [(2, 1000, 64, 3, 1), (3, 777, 32, 7, 3), (1, 247, 64, 11, 5), (2, 503, 32, 11, 1)])
def test_fused_pair_respects_output_bounds(case):
    from mri2speech_b200 import _lib
    B, L, C, k, d = case
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, L, C, generator=g).half().cuda()
    w1 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    w2 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    b1 = torch.zeros(C, device="cuda")
    guard = 7
    d32 = _guarded(B, L, guard, C, torch.float32)
    d16 = _guarded(B, L, guard, C, torch.float16)
    res = torch.randn(B, L + guard, C, generator=g).cuda()
    _lib.resblock_pair_fwd(x, w1, b1, d, w2, b1, res=res, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1,
                           out=d32, out16=d16, d_rows=L + guard)
    assert torch.all(d32[:, L:] == SENTINEL) and torch.all(d16[:, L:] == SENTINEL)
    assert not torch.any(d32[:, :L] == SENTINEL)


def test_generator_and_encoder_leave_neighbours_alone():
    """Outputs of the full forwards live in the middle of a larger sentinel-filled allocation."""
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from mri2speech_b200.vocoder import Generator
    from tests.util import load_config
    torch.manual_seed(1234)
    gen = Generator(load_config(), precision="fp16").cuda().eval()
    mel = synth.synthetic_mels(2, 13).cuda()
    with torch.no_grad():
        wav = gen(mel)
    assert wav.shape == (2, 1, 13 * 420) and torch.isfinite(wav).all()
    ac = build_acoustic_model(precision="fp16").cuda().eval()
    clip = synth.synthetic_clip_u8(5, 7).unsqueeze(0).cuda()
    with torch.no_grad():
        out = ac(clip)
    assert out.shape == (1, 7, 64) and torch.isfinite(out).all()

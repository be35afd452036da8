"""Fused MBConv kernels (csrc/mbconv_sm100.cu) against the five-launch path they replace.

The reference block is timm's InvertedResidual + SqueezeExcite as called by EffNetV2B2Backbone.forward
(mri2speech_code/mri_acoustic_model.py:28-48).  M2S_MBCONV selects the path when the handle is created: 0 = expand GEMM,
depthwise kernel, SE MLP, SE scale pass, project GEMM; bit0 = expand GEMM with the depthwise conv + squeeze as its
epilogue; bit1 = SE scale applied to the project GEMM's A operand in SMEM; bit2 = the stride-1 EdgeResidual blocks of
stages 1-2 (3x3 expand -> SiLU -> 1x1 project) in one kernel with the expanded tile in SMEM (csrc/fused_er_sm100.cu);
bit4 = stage 0's 3x3 convs with two pixels per GEMM row (N = 32 instead of 16: interleaved A planes of the conv engine);
bit5 = the residual stream in fp16 (shortcuts read from the fp16 operand copies, no fp32 copies of the block outputs);
bit6 = the stride-2 EdgeResidual blocks read a space-to-depth copy of their input (9 K-window taps over the 4 parity
planes, with bit2) instead of an im2col matrix; bit7 = no copy either: the fused kernel's TMA loads gather the
space-to-depth tile from the NHWC input (5-D box, out-of-bounds zero fill as the conv's padding).  Both kernels keep the arithmetic of the
launches they replace (fp32 depthwise accumulation in tap order, fp32 scale * fp16 activation rounded once), so the
encoder features must agree far below the fp16 operand noise (2^-11); the oracle comparison of the fused default runs
in tests/test_acoustic_gpu.py / test_bench_paths_gpu.py / test_scaled_init_gpu.py."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _encode(mode, frames, randomize_bn=True, dw32=None):
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    m = build_acoustic_model(precision="fp16")
    if randomize_bn:
        synth.randomize_batchnorm(m)
    m = m.cuda().eval()
    old = os.environ.get("M2S_MBCONV")
    os.environ["M2S_MBCONV"] = str(mode)
    try:
        m.refresh()                       # the knob is read when the handle is created
        out = m.encode_frames(frames)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("M2S_MBCONV", None)
        else:
            os.environ["M2S_MBCONV"] = old
    return out.cpu(), m.launches_per_forward()


def _frames(n, seed=0):
    from mri2speech_b200 import synth
    return synth.synthetic_clip(seed, n).cuda()


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 12, 16, 23, 32, 39, 55, 68, 119, 196, 247])
@pytest.mark.parametrize("n", [5, 301])
def test_fused_equals_unfused(mode, n):
    """5 frames: partial tiles (two 8x8 frames per tile, odd count); 301 frames: several tiles per CTA (ring phases wrap)."""
    frames = _frames(n)
    ref, l0 = _encode(0, frames)
    got, l1 = _encode(mode, frames)
    assert torch.isfinite(got).all()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    # bit 5 (32) = the residual stream in fp16: one more fp16 rounding per block with a shortcut (tools/
    # emulate_residual_rounding.py: ~3e-4 of the features' abs-max on these weights); everything else keeps the arithmetic
    # bit 1 (2): on 16 x 16 frames the excite scale multiplies the project WEIGHTS (one frame per tile) instead of the
    # activations: one fp16 rounding per product term either way, but not the same one
    tol = 2e-3 if mode & 32 else (5e-4 if mode & 2 else 2e-4)
    assert err <= tol * scale, (mode, n, err, scale)
    assert l1 < l0 or mode in (16, 32)               # fewer launches per forward (16 adds two border passes, 32 none)
    if mode == 3:
        assert l0 - l1 == 18 + 20                    # 18 stride-1 blocks lose the depthwise launch, all 20 the scale pass
    if mode == 4:
        assert l0 - l1 == 4                          # stage 1's three EdgeResidual blocks and stage 2's first (weights resident in SMEM)
    if mode == 12:
        assert l0 - l1 == 6                          # ... and stage 2's other two (258 KB of weights streamed per tile; slower, opt-in)


def test_fused_default_is_on():
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    m = build_acoustic_model(precision="fp16").cuda().eval()
    m.refresh()
    fused = m.launches_per_forward()
    os.environ["M2S_MBCONV"] = "0"
    try:
        m.refresh()
        plain = m.launches_per_forward()
    finally:
        os.environ.pop("M2S_MBCONV", None)
    assert plain - fused == 42                       # 42 launches fewer, 2 border-column passes more, the 2 im2col passes gone


def test_stride2_depthwise_tma_equals_slab_kernel():
    """dwconv_s2_tma_kernel (TMA tiles, channel-pair lanes) against dwconv_kernel<2>: same fp32 arithmetic in the same
    order, so the two stride-2 blocks (stage 3 / stage 5, the second on an unpadded 16 x 16 input with an odd frame
    count in its two-frame loads) must reproduce the features to fp32 summation-order noise."""
    frames = _frames(7, seed=3)
    os.environ["M2S_DWCONV_S2_TMA"] = "0"
    try:
        ref, _ = _encode(0, frames)
    finally:
        os.environ.pop("M2S_DWCONV_S2_TMA", None)
    got, _ = _encode(0, frames)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 2e-4 * scale

"""Edge cases of the drop-in modules on the GPU: single frame, zero-length utterances, odd lengths, big-ish batch."""
import pytest
import torch

from tests.util import load_config

pytestmark = pytest.mark.gpu


def _gen():
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    return Generator(load_config()).cuda().eval()


def _ac():
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    return build_acoustic_model().cuda().eval()


def _cpu(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def test_generator_single_frame_and_odd_lengths():
    from oracle.vocoder import generator_forward, snr_db
    g = _gen()
    for T in (1, 2, 7, 129):
        mel = torch.randn(1, 64, T, generator=torch.Generator().manual_seed(T)) * 2 - 5
        ref = generator_forward(_cpu(g), load_config(), mel)
        with torch.no_grad():
            wav = g(mel.cuda()).cpu()
        assert wav.shape == (1, 1, T * 420)
        assert snr_db(ref, wav, True) >= 40.0, T


def test_generator_zero_length_utterance_in_batch():
    from oracle.vocoder import generator_forward, snr_db
    g = _gen()
    mel = torch.randn(3, 64, 12, generator=torch.Generator().manual_seed(5)) * 2 - 5
    lens = torch.tensor([12, 0, 5], dtype=torch.int32)
    with torch.no_grad():
        wav = g(mel.cuda(), lengths=lens.cuda()).cpu()
    assert torch.isfinite(wav).all()
    for b in (0, 2):
        n = int(lens[b])
        ref = generator_forward(_cpu(g), load_config(), mel[b:b + 1, :, :n])
        assert snr_db(ref[0, 0], wav[b, 0, : n * 420], True) >= 40.0


def test_generator_identical_rows_give_identical_waveforms():
    """Utterance independence inside a batch (the property that makes utterance sharding exact)."""
    g = _gen()
    one = torch.randn(1, 64, 33, generator=torch.Generator().manual_seed(6)) * 2 - 5
    batch = one.repeat(5, 1, 1).cuda()
    with torch.no_grad():
        wav = g(batch)
    assert all(torch.equal(wav[0], wav[i]) for i in range(1, 5))


def test_acoustic_single_frame_and_ragged_with_empty():
    from mri2speech_b200 import synth
    from oracle.acoustic import acoustic_forward
    m = _ac()
    clip = synth.synthetic_clip(2, 3)
    with torch.no_grad():
        one = m(clip[:1].unsqueeze(0).unsqueeze(2).cuda()).cpu()       # T = 1
        ref1 = acoustic_forward(_cpu(m), clip[:1].unsqueeze(0).unsqueeze(2))
    assert one.shape == (1, 1, 64) and (one - ref1).abs().max().item() < 1e-3
    batch = torch.stack([clip, clip.flip(0), clip])
    lens = torch.tensor([3, 0, 2], dtype=torch.int32)
    with torch.no_grad():
        got = m(batch.cuda(), lengths=lens).cpu()                       # (B,T,H,W) form, one empty utterance
        ref = acoustic_forward(_cpu(m), batch.unsqueeze(2), lengths=lens.tolist())
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 1e-3
    assert got[1].abs().max().item() == 0.0


def test_acoustic_frame_independence_and_shape_errors():
    from mri2speech_b200 import synth
    m = _ac()
    clip = synth.synthetic_clip(4, 4).cuda()
    f_all = m.encode_frames(clip)
    f_rev = m.encode_frames(clip.flip(0))
    assert torch.allclose(f_all, f_rev.flip(0), atol=1e-6)             # eval-mode BN: frames are independent
    with pytest.raises(ValueError):
        m(torch.zeros(2, 3, 256, device="cuda"))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 3, 256, 256, device="cuda"))


@pytest.mark.parametrize("precision,rel_tol", [("fp16", 5e-3), ("fp32", 2e-4)])
@pytest.mark.parametrize("hw", [(128, 160), (288, 256)])
def test_encoder_other_frame_sizes(precision, rel_tol, hw):
    """The module API takes any frame size that is a multiple of 32 (the reference's CLIs always resize to 256 x 256, its
    nn.Module does not care).  Non-square / non-256 frames walk the general paths: other pitches for the two-pixel rows of
    stage 0 and for the space-to-depth TMA boxes of the stride-2 blocks, and the unfused InvertedResidual launches
    wherever the fused kernels' 16 x 16 / 8 x 8 geometry does not apply."""
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from oracle.acoustic import encoder_forward
    torch.manual_seed(1234)
    m = build_acoustic_model(precision=precision)
    synth.randomize_batchnorm(m)
    m = m.cuda().eval()
    frames = torch.rand(3, hw[0], hw[1], generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ref = encoder_forward(_cpu(m), frames.unsqueeze(1))
    got = m.encode_frames(frames.cuda()).cpu()
    assert got.shape == (3, 208)
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    assert rel < rel_tol, (hw, rel)

"""Parity on the code paths the benches actually run ("Parity first").

The small-clip tests elsewhere encode <= 9 frames: the engine then picks msub <= 2 tiles, the encoder's chunk loop
runs once and a ragged batch's frame map never spans a chunk.  Here:
  * >= 1100 frames through the encoder with M2S_ENCODER_CHUNK in {64, 1024, default}: the 512 / 1024-row tiles, the
    CTA-pair dispatch and multi-chunk passes, 32 sampled frames against the CPU oracle;
  * a ragged batch whose compact frame list spans a chunk boundary, every valid frame against the oracle;
  * BASELINE.json configs[0]: one 150-frame clip end to end (run_mri_video_inference's chain) against the oracle at
    the north-star gates (mel 1e-3 max-abs, waveform SNR >= 40 dB raw AND mean-removed), tf32 and fp16 builds.
"""
import os

import pytest
import torch

from tests.util import load_config

pytestmark = pytest.mark.gpu


def _acoustic(precision, chunk=None, randomize_bn=True):
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    if chunk is None:
        os.environ.pop("M2S_ENCODER_CHUNK", None)
    else:
        os.environ["M2S_ENCODER_CHUNK"] = str(chunk)   # read by m2s_acoustic_create (first forward)
    torch.manual_seed(1234)
    m = build_acoustic_model(precision=precision)
    if randomize_bn:
        synth.randomize_batchnorm(m)
    return m.cuda().eval()


def _cpu_sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


@pytest.fixture(autouse=True)
def _restore_chunk_env():
    yield
    os.environ.pop("M2S_ENCODER_CHUNK", None)


@pytest.mark.parametrize("precision,rel_tol", [("tf32", 5e-3), ("fp16", 5e-3)])
@pytest.mark.parametrize("chunk", [64, 1024, None])
def test_encoder_1100_frames_sampled_vs_oracle(precision, rel_tol, chunk):
    from mri2speech_b200 import synth
    from oracle.acoustic import encoder_forward
    m = _acoustic(precision, chunk)
    n = 1100
    clips = [synth.synthetic_clip(20 + i, 110) for i in range(10)]
    frames = torch.cat(clips)                                   # (1100, 256, 256)
    got = m.encode_frames(frames.cuda()).cpu()
    assert got.shape == (n, 208) and torch.isfinite(got).all()
    pick = torch.randperm(n, generator=torch.Generator().manual_seed(6))[:32].sort().values
    pick[0], pick[-1] = 0, n - 1                                 # first / last frame, a chunk's first / last rows
    pick[1], pick[2] = 1023, 1024                                # either side of a 1024-frame chunk boundary
    with torch.no_grad():
        ref = encoder_forward(_cpu_sd(m), frames[pick].unsqueeze(1))
    rel = (got[pick] - ref).abs().max().item() / ref.abs().max().item()
    print(f"[{precision} chunk={chunk or 2048}] 1100 frames, 32 sampled: feature rel err {rel:.2e}")
    assert rel < rel_tol
    # frame independence at size: the same frames encoded on their own (small tiles, one chunk) agree closely
    solo = m.encode_frames(frames[pick].cuda()).cpu()
    assert (solo - got[pick]).abs().max().item() / ref.abs().max().item() < 2e-3


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_ragged_batch_spanning_chunk_boundary_vs_oracle(precision):
    from mri2speech_b200 import synth
    from oracle.acoustic import acoustic_forward
    m = _acoustic(precision, chunk=64)
    lens = [50, 20, 30]                                          # compact list: clip 1 occupies frames 50..69
    T = max(lens)
    x = torch.zeros(3, T, 256, 256)
    clips = [synth.synthetic_clip(40 + i, ln) for i, ln in enumerate(lens)]
    for b, c in enumerate(clips):
        x[b, : lens[b]] = c
    with torch.no_grad():
        got = m(x.cuda(), lengths=torch.tensor(lens, dtype=torch.int32)).cpu()
    sd = _cpu_sd(m)
    for b, c in enumerate(clips):
        ref = acoustic_forward(sd, c[None, :, None])[0]
        err = (got[b, : lens[b]] - ref).abs().max().item()
        print(f"[{precision}] ragged clip {b} ({lens[b]} frames): mel max-abs err {err:.2e} (abs-max {ref.abs().max():.2e})")
        assert err < 1e-3
        assert err / ref.abs().max().item() < 2e-2
        assert got[b, lens[b]:].abs().max().item() == 0.0 if lens[b] < T else True


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_config0_one_150_frame_clip_end_to_end(precision):
    """BASELINE.json configs[0]: what scripts/run_mri_video_inference.py does for one clip (reference :218-243)."""
    from mri2speech_b200 import synth
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.vocoder import Generator
    from oracle.acoustic import acoustic_forward
    from oracle.glue import mel_glue
    from oracle.vocoder import generator_forward, snr_db
    h = load_config()
    ac = _acoustic(precision, randomize_bn=False)               # north star: random-init weights
    torch.manual_seed(1234)
    gen = Generator(h, precision=precision)
    mean, std = synth.synthetic_scaler()
    clip = synth.synthetic_clip(0, 150)
    pipe = MriToSpeech(ac, gen, mean, std)
    out = pipe.infer([clip.cuda()])[0]
    mel_ref = acoustic_forward(_cpu_sd(ac), clip[None, :, None])[0]
    _, mel_log, voc_in = mel_glue(mel_ref, mean, std)
    wav_ref = generator_forward({k: v.detach().cpu() for k, v in gen.state_dict().items()}, h, voc_in.unsqueeze(0))[0, 0]
    mel_err = (out["mel_norm"].cpu() - mel_ref).abs().max().item()
    wav = out["audio"].cpu()
    raw, mr = snr_db(wav_ref, wav, False), snr_db(wav_ref, wav, True)
    print(f"[{precision}] configs[0] 150 frames: mel max-abs err {mel_err:.2e} (gate 1e-3, abs-max {mel_ref.abs().max():.2e}), "
          f"waveform SNR raw {raw:.1f} dB / mean-removed {mr:.1f} dB (gate 40)")
    assert wav.shape == (150 * 420,)
    assert mel_err < 1e-3
    assert raw >= 40.0 and mr >= 40.0


@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_packed_forward_equals_padded_forward(dtype):
    """m2s_acoustic_forward_packed (no input padding; how MriToSpeech.infer feeds micro-batches) is bit-identical to the
    padded call with lengths, for float32 and uint8 frames, and rejects inconsistent lengths."""
    from mri2speech_b200 import synth
    m = _acoustic("fp16", chunk=8)                                # 3 + 7 + 5 frames: two encoder chunks
    lens = [3, 7, 5]
    mk = synth.synthetic_clip_u8 if dtype == "uint8" else synth.synthetic_clip
    clips = [mk(60 + i, ln) for i, ln in enumerate(lens)]
    T = max(lens)
    padded = torch.zeros(3, T, 256, 256, dtype=clips[0].dtype)
    for b, c in enumerate(clips):
        padded[b, : lens[b]] = c
    lt = torch.tensor(lens, dtype=torch.int32)
    with torch.no_grad():
        a = m(padded.cuda(), lengths=lt).clone()
        b = m.forward_packed(torch.cat(clips).cuda(), lt).clone()
    assert a.shape == b.shape == (3, T, 64)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        m.forward_packed(torch.cat(clips).cuda(), torch.tensor([3, 7, 6], dtype=torch.int32))


def test_infer_overlapped_micro_batches_and_host_audio():
    """MriToSpeech.infer with several micro-batches (double-buffered staging on a copy stream) and pinned-host audio:
    every clip equals its own single-clip run, whatever the micro-batch it landed in."""
    from mri2speech_b200 import synth
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.vocoder import Generator
    ac = _acoustic("fp16")
    torch.manual_seed(1234)
    gen = Generator(load_config(), precision="fp16")
    mean, std = synth.synthetic_scaler()
    pipe = MriToSpeech(ac, gen, mean, std)
    lens = [9, 4, 7, 2, 5, 8]
    clips = [synth.synthetic_clip_u8(70 + i, ln).pin_memory() for i, ln in enumerate(lens)]
    outs = pipe.infer(clips, max_batch_frames=16, audio_to_host=True)     # 4 micro-batches
    assert len(MriToSpeech.plan_micro_batches(lens, 16)) >= 3
    for i, (c, o) in enumerate(zip(clips, outs)):
        solo = pipe.infer([c])[0]
        assert not o["audio"].is_cuda and o["audio"].shape == (lens[i] * 420,)
        assert (o["mel_norm"] - solo["mel_norm"]).abs().max().item() < 1e-5, i
        # the vocoder's tiling depends on the batch shape: fp32 summation order may differ
        assert (o["audio"] - solo["audio"].cpu()).abs().max().item() < 1e-4, i

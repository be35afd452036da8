"""Parity at BASELINE.json's full sizes, where the CPU oracle cannot run the whole batch in test time.

configs[1] (HiFi-GAN Generator only, batch 32 x 256 mel frames -> 32 x 107 520 samples): the fp16 build `bench.py`
measures and the tf32 build are checked through properties that do not depend on the size --
  * determinism: two runs are bit-identical;
  * batch-permutation equivariance, bit-exact: an utterance's waveform does not depend on its slot in the batch
    (the property that makes utterance sharding exact, here on the CTA-pair / fused-pair kernels that only run at size);
  * every utterance of the batch equals its own B = 1 run (other tile shapes and kernels: fp32 summation order may
    differ, so a tight SNR instead of bit-exactness);
  * ragged lengths at full size: the valid part of each utterance equals the B = 1 run of its truncated mel;
and the oracle itself is run on two utterances of the batch (T = 256, ~0.5 s of CPU each) at the north-star gate.
configs[2]'s shape (ragged clips of 150-600 frames) goes through the same properties on the acoustic model's
BiLSTM + head, whose oracle is cheap at any length; the encoder is frame-independent (tests/test_edge_cases_gpu.py).
"""
import pytest
import torch

from tests.util import load_config

pytestmark = pytest.mark.gpu

HOP = 420


def _gen(precision):
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    return Generator(load_config(), precision=precision).cuda().eval()


def _snr(a, b):
    from oracle.vocoder import snr_db
    return snr_db(a, b, True)   # mean-removed: random-init weights give an almost-DC waveform


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_config2_full_size_properties(precision):
    from mri2speech_b200 import synth
    from oracle.vocoder import generator_forward
    B, T = 32, 256
    g = _gen(precision)
    mel = synth.synthetic_mels(B, T, seed=77)
    x = mel.cuda()
    with torch.no_grad():
        wav = g(x).clone()
        again = g(x).clone()
    assert wav.shape == (B, 1, T * HOP) and torch.isfinite(wav).all()
    assert torch.equal(wav, again)                                       # determinism
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        wav_p = g(x[perm.cuda()].contiguous()).clone()
    assert torch.equal(wav_p, wav[perm.cuda()])                          # slot independence, bit-exact
    for b in (0, 13, 31):                                                # each utterance == its own B = 1 run
        with torch.no_grad():
            solo = g(x[b:b + 1].contiguous())
        s = _snr(solo[0, 0].cpu(), wav[b, 0].cpu())
        assert s >= 60.0, (b, s)
    sd = {k: v.detach().cpu() for k, v in g.state_dict().items()}
    for b in (5, 22):                                                    # the oracle on two utterances of the batch
        ref = generator_forward(sd, load_config(), mel[b:b + 1])
        s = _snr(ref[0, 0], wav[b, 0].cpu())
        print(f"[{precision}] full-size batch, utterance {b}: mean-removed SNR vs the CPU oracle {s:.1f} dB")
        assert s >= 40.0


def test_config2_full_size_ragged():
    from mri2speech_b200 import synth
    B, T = 32, 256
    g = _gen("fp16")
    mel = synth.synthetic_mels(B, T, seed=78)
    lens = torch.randint(1, T + 1, (B,), generator=torch.Generator().manual_seed(9), dtype=torch.int32)
    lens[0], lens[1], lens[2] = T, 1, 255
    x = mel.cuda()
    with torch.no_grad():
        wav = g(x, lengths=lens.cuda()).clone()
    assert torch.isfinite(wav).all()
    for b in (0, 1, 2, 7, 19, 31):
        n = int(lens[b])
        with torch.no_grad():
            solo = g(x[b:b + 1, :, :n].contiguous())
        s = _snr(solo[0, 0].cpu(), wav[b, 0, : n * HOP].cpu())
        assert s >= 50.0, (b, n, s)
        if n < T:   # past the clip every utterance carries the same constant (conv_post of a zeroed tensor)
            tail = wav[b, 0, n * HOP:]
            assert (tail - tail[0]).abs().max().item() == 0.0


def test_config3_shape_bilstm_head_ragged_600():
    """configs[2]: ragged clips of 150-600 frames.  The recurrence + head against torch.nn.LSTM on the CPU at the
    longest length, batch 8, fp16 build (the input projection and the head run on the tensor cores)."""
    from mri2speech_b200.acoustic import build_acoustic_model
    from oracle.acoustic import bilstm_head_forward
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision="fp16").cuda().eval()
    B, T = 8, 600
    g = torch.Generator().manual_seed(4)
    feats = torch.randn(B, T, 208, generator=g) * 0.5
    lens = torch.tensor([600, 150, 599, 333, 151, 600, 420, 287], dtype=torch.int32)
    with torch.no_grad():
        mel = ac.rnn_head(feats.cuda(), lens).cpu()
    sd = {k: v.detach().cpu() for k, v in ac.state_dict().items()}
    for b in range(B):
        n = int(lens[b])
        ref = bilstm_head_forward(sd, feats[b:b + 1, :n])
        err = (mel[b, :n] - ref[0]).abs().max().item()
        assert err < 1e-3, (b, n, err)
        if n < T:
            assert mel[b, n:].abs().max().item() == 0.0

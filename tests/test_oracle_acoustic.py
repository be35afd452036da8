"""CPU: structural pins of the acoustic-model oracle (encoder parity is UNPINNED: timm is absent)."""
import os

import numpy as np
import torch

from tests.util import GOLDEN


def _model():
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    return build_acoustic_model()


def test_backbone_parameter_count_and_names():
    m = _model()
    assert sum(p.numel() for p in m.cnn.parameters()) == 8391406
    sd = m.state_dict()
    for key in ("cnn.backbone.conv_stem.weight", "cnn.backbone.bn1.running_var",
                "cnn.backbone.blocks.0.1.conv.weight", "cnn.backbone.blocks.1.0.conv_exp.weight",
                "cnn.backbone.blocks.2.2.conv_pwl.weight", "cnn.backbone.blocks.3.0.se.conv_reduce.bias",
                "cnn.backbone.blocks.5.9.bn3.weight", "rnn.lstm.weight_hh_l0_reverse", "head.bias"):
        assert key in sd, key
    assert sd["cnn.backbone.blocks.3.0.se.conv_reduce.weight"].shape == (14, 224, 1, 1)
    assert sd["cnn.backbone.blocks.5.1.se.conv_reduce.weight"].shape == (52, 1248, 1, 1)
    assert sd["rnn.lstm.weight_ih_l0"].shape == (2560, 208) and sd["head.weight"].shape == (64, 640)
    assert m.cnn.out_channels == 208 and m.n_mels == 64


def test_encoder_output_shape_and_golden():
    from mri2speech_b200 import synth
    from oracle.acoustic import encoder_forward
    sd = _model().state_dict()
    clip = synth.synthetic_clip(0, 6)[:2]
    with torch.no_grad():
        f = encoder_forward(sd, clip.unsqueeze(1))
    assert f.shape == (2, 208)
    z = np.load(os.path.join(GOLDEN, "acoustic_oracle_seed1234_clip0_t6.npz"))
    # per-frame independence (eval-mode BN): first two frames of the 6-frame golden
    assert np.abs(f.numpy() - z["feats"][:2]).max() < 1e-6


def test_explicit_lstm_equals_torch_lstm():
    from oracle.acoustic import bilstm_head_forward
    sd = _model().state_dict()
    feats = torch.randn(2, 9, 208, generator=torch.Generator().manual_seed(2))
    a = bilstm_head_forward(sd, feats, explicit=False)
    b = bilstm_head_forward(sd, feats, explicit=True)
    assert (a - b).abs().max().item() < 1e-5


def test_mel_glue_closed_form():
    from mri2speech_b200 import synth
    from oracle.glue import mel_glue
    mean, std = synth.synthetic_scaler()
    pred = torch.randn(7, 64, generator=torch.Generator().manual_seed(1)) * 3
    mel_db, mel_log, voc = mel_glue(pred, mean, std)
    closed = torch.clamp(mel_db * (np.log(10.0) / 10.0), min=float(np.log(1e-5)))
    assert (mel_log - closed).abs().max().item() < 1e-4
    assert voc.shape == (64, 7) and mel_db.shape == (7, 64)
    assert mel_log.min().item() >= np.log(1e-5) - 1e-5


def test_second_encoder_oracle_from_torchvision_blocks():
    """oracle/acoustic.py (restatement of timm's topology) == oracle/acoustic_tv.py (torchvision's own FusedMBConv /
    MBConv / SqueezeExcitation blocks + TF-same padding) on the same state_dict: default init and scaled init."""
    from mri2speech_b200 import synth
    from oracle.acoustic import encoder_forward
    from oracle.acoustic_tv import encoder_forward_tv
    from oracle.scaled_init import calibration_frames, scale_acoustic
    clip = synth.synthetic_clip(2, 2)
    m = _model()
    for scaled in (False, True):
        if scaled:
            scale_acoustic(m, calibration_frames(4))
        sd = m.state_dict()
        with torch.no_grad():
            a = encoder_forward(sd, clip.unsqueeze(1))
        b = encoder_forward_tv(sd, clip.unsqueeze(1))
        assert a.shape == b.shape == (2, 208)
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, a.abs().max().item())
        if scaled:
            assert 0.2 < a.std().item() < 20.0          # the scaled init does put the features at O(1)


def test_scaled_init_ranges():
    """The scaled-init helpers do what the parity tests rely on: O(1) mel, a waveform that uses tanh's range."""
    from mri2speech_b200 import synth
    from mri2speech_b200.vocoder import Generator
    from oracle.acoustic import acoustic_forward
    from oracle.glue import mel_glue
    from oracle.scaled_init import calibration_frames, scale_acoustic, scale_generator
    from oracle.vocoder import generator_forward
    from tests.util import load_config
    m = _model()
    scale_acoustic(m, calibration_frames(6))
    clip = synth.synthetic_clip(1, 4)
    mel = acoustic_forward(m.state_dict(), clip[None, :, None])
    assert 0.3 < mel.std().item() < 5.0
    h = load_config()
    torch.manual_seed(1234)
    g = Generator(h)
    scale_generator(g, h, synth.synthetic_mels(1, 32, seed=5))
    _, _, voc = mel_glue(mel[0], *synth.synthetic_scaler())
    wav = generator_forward(g.state_dict(), h, voc.unsqueeze(0))
    assert wav.std().item() > 0.1 and wav.abs().max().item() <= 1.0

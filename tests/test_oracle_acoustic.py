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

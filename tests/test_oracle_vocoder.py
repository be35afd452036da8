"""CPU: the vocoder oracle is pinned to the reference's own Generator.

* against the committed golden vectors (generated from /root/reference/models.py by oracle/gen_golden.py);
* against the live reference when /root/reference is present (build container only);
* structural facts of SURVEY.md 8a-5: 233 tensors / 14 394 370 parameters, causal ResBlocks, look-ahead 6."""
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, load_config


def _state_dict():
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    return Generator(load_config()).state_dict()


def test_oracle_matches_reference_golden():
    from oracle.vocoder import generator_forward, snr_db
    z = np.load(os.path.join(GOLDEN, "vocoder_ref_seed1234_b2_t24.npz"))
    wav = generator_forward(_state_dict(), load_config(), torch.from_numpy(z["mel"]))
    assert np.abs(wav.numpy() - z["wav"]).max() < 1e-6
    assert snr_db(torch.from_numpy(z["wav"]), wav, True) > 100.0


def test_oracle_ragged_equals_b1_golden():
    from oracle.vocoder import generator_forward
    z = np.load(os.path.join(GOLDEN, "vocoder_ref_seed1234_ragged.npz"))
    wav = generator_forward(_state_dict(), load_config(), torch.from_numpy(z["mel"]), lengths=z["lens"].tolist())
    for b, key in enumerate(("wav0", "wav1")):
        n = z[key].shape[0]
        assert np.abs(wav[b, 0, :n].numpy() - z[key]).max() < 1e-6


def test_state_dict_layout():
    sd = _state_dict()
    assert len(sd) == 233
    assert sum(v.numel() for v in sd.values()) == 14394370
    assert sd["ups.0.weight_g"].shape == (512, 1, 1) and sd["ups.0.weight_v"].shape == (512, 256, 20)
    assert "conv_pre.weight" in sd and "conv_pre.weight_g" not in sd
    assert sd["resblocks.11.convs2.2.weight_v"].shape == (32, 32, 11)


def test_receptive_field_is_causal_with_lookahead_6():
    """Perturbing mel frame 10 of 40 changes samples 1435..10966 only (SURVEY.md 8a-5 probe)."""
    from oracle.vocoder import generator_forward
    sd, h = _state_dict(), load_config()
    mel = torch.randn(1, 64, 40, generator=torch.Generator().manual_seed(0))
    mel2 = mel.clone()
    mel2[:, :, 10] += 1.0
    d = (generator_forward(sd, h, mel) - generator_forward(sd, h, mel2)).abs()[0, 0]
    nz = torch.nonzero(d > 0).flatten()
    assert nz.min().item() >= (10 - 6) * 420 - 260 and nz.min().item() <= (10 - 6) * 420
    assert nz.max().item() < 40 * 420


def test_against_live_reference_when_present():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle.vocoder import generator_forward
    g_ref, h = ref_import.reference_generator(1234)
    sd = _state_dict()
    ref_sd = g_ref.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    assert all(torch.equal(sd[k], ref_sd[k]) for k in sd), "seeded init diverged from the reference"
    mel = torch.randn(1, 64, 12, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y_ref = g_ref(mel)
    assert (generator_forward(sd, h, mel) - y_ref).abs().max().item() < 1e-6
    # float64 restatement == reference to fp32 rounding
    assert (generator_forward(sd, h, mel, dtype=torch.float64).float() - y_ref).abs().max().item() < 1e-6


def _config_rb2():
    h = load_config()
    h["resblock"] = "2"
    h["resblock_dilation_sizes"] = [[1, 3], [1, 3], [1, 3]]
    return h


def test_resblock2_oracle_matches_reference_golden():
    """Configs with "resblock": "2" (reference models.py:58-85, :95).  The golden comes from the reference's own
    Generator under seed 1234; the product's Generator must re-create the same tensors from the same seed (names
    resblocks.N.convs.M.*), and the oracle's ResBlock2 branch must reproduce the waveform, ragged clip included."""
    from mri2speech_b200.vocoder import Generator
    from oracle.vocoder import generator_forward
    h = _config_rb2()
    torch.manual_seed(1234)
    sd = Generator(h).state_dict()
    assert "resblocks.0.convs.1.weight_v" in sd and "resblocks.0.convs1.0.weight_v" not in sd
    assert sd["resblocks.11.convs.1.weight_v"].shape == (32, 32, 11)
    z = np.load(os.path.join(GOLDEN, "vocoder_ref_seed1234_resblock2.npz"))
    mel = torch.from_numpy(z["mel"])
    wav = generator_forward(sd, h, mel)
    assert np.abs(wav.numpy() - z["wav"]).max() < 1e-6
    rag = generator_forward(sd, h, mel, lengths=z["lens"].tolist())
    n = z["wav1_ragged"].shape[0]
    assert np.abs(rag[1, 0, :n].numpy() - z["wav1_ragged"]).max() < 1e-6

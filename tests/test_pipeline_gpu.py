"""End-to-end rtMRI -> mel -> wav through the public pipeline and the CLI entry point, vs the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import ROOT, load_config

pytestmark = pytest.mark.gpu


def _models():
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    gen = Generator(load_config())
    ac = build_acoustic_model()
    synth.randomize_batchnorm(ac)
    return ac, gen


def test_ragged_pipeline_matches_oracle_per_clip():
    from mri2speech_b200 import synth
    from mri2speech_b200.pipeline import MriToSpeech
    from oracle.acoustic import acoustic_forward
    from oracle.glue import mel_glue
    from oracle.vocoder import generator_forward, snr_db
    ac, gen = _models()
    mean, std = synth.synthetic_scaler()
    clips = [synth.synthetic_clip(i, t) for i, t in enumerate((5, 9, 3))]
    sd_ac = {k: v.clone() for k, v in ac.state_dict().items()}
    sd_gen = {k: v.clone() for k, v in gen.state_dict().items()}
    pipe = MriToSpeech(ac, gen, mean, std)
    # max_batch_frames small enough to force two micro-batches
    out = pipe.infer([c.cuda() for c in clips], max_batch_frames=20)
    for i, clip in enumerate(clips):
        m = acoustic_forward(sd_ac, clip.unsqueeze(0).unsqueeze(2))[0]
        mel_db, mel_log, voc_in = mel_glue(m, mean, std)
        wav = generator_forward(sd_gen, load_config(), voc_in.unsqueeze(0))[0, 0]
        assert out[i]["audio"].shape == (clip.shape[0] * 420,)
        assert (out[i]["mel_norm"].cpu() - m).abs().max().item() < 1e-3
        assert (out[i]["mel_db"].cpu() - mel_db).abs().max().item() < 2e-2      # dB scale (std up to 15)
        assert (out[i]["mel_log"].cpu() - mel_log).abs().max().item() < 5e-3
        assert snr_db(wav, out[i]["audio"].cpu(), True) >= 40.0


def test_mel_glue_matches_reference_lines():
    from mri2speech_b200 import synth
    from mri2speech_b200.pipeline import mel_glue
    from oracle.glue import mel_glue as ref_glue
    mean, std = synth.synthetic_scaler()
    pred = torch.randn(2, 11, 64, generator=torch.Generator().manual_seed(3)) * 3.0   # drives some bins into the clamp
    lens = torch.tensor([11, 6], dtype=torch.int32)
    db, lg, voc = mel_glue(pred.cuda(), torch.from_numpy(mean), torch.from_numpy(std), lens)
    rdb, rlg, rvoc = ref_glue(pred, mean, std)
    assert (db.cpu()[0] - rdb[0]).abs().max().item() < 1e-4
    assert (lg.cpu()[0] - rlg[0]).abs().max().item() < 1e-4
    assert (voc.cpu()[0] - rvoc[0]).abs().max().item() < 1e-4
    assert (lg.cpu()[1, :6] - rlg[1, :6]).abs().max().item() < 1e-4
    assert db.cpu()[1, 6:].abs().max().item() == 0 and voc.cpu()[1, :, 6:].abs().max().item() == 0
    with pytest.raises(ValueError):
        mel_glue(pred.cuda(), torch.zeros(3), torch.ones(3))


def test_cli_run_with_synthetic_checkpoints(tmp_path):
    """scripts/run_mri_video_inference.run(): checkpoint / config / scaler formats of the reference."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import run_mri_video_inference as cli
    from mri2speech_b200 import synth
    ac, gen = _models()
    torch.save({"epoch": 1, "model_state_dict": ac.state_dict(), "val_loss": 0.0}, tmp_path / "mri.pt")
    torch.save({"generator": gen.state_dict()}, tmp_path / "g_00000001")
    synth.write_scaler_json(tmp_path / "scaler.json")
    args = cli.parse_args(["--video", str(tmp_path / "clip7.mp4"), "--mri-checkpoint", str(tmp_path / "mri.pt"),
                           "--scaler-json", str(tmp_path / "scaler.json"),
                           "--hifigan-config", os.path.join(ROOT, "config_custom.json"),
                           "--hifigan-checkpoint", str(tmp_path / "g_00000001"), "--output-dir", str(tmp_path / "out"),
                           "--mri-code-dir", os.path.join(ROOT, "mri2speech_code")])
    frames = synth.synthetic_clip(7, 6)
    audio, mel_db, mel_log = cli.run(args, frames=frames)
    assert audio.shape == (6 * 420,) and mel_db.shape == (6, 64) and mel_log.shape == (6, 64)
    out = tmp_path / "out"
    assert (out / "clip7_generated.wav").exists() and (out / "clip7_mel.npy").exists()
    assert (out / "clip7_mel_log.npy").exists()
    assert np.load(out / "clip7_mel.npy").shape == (6, 64)
    with pytest.raises(FileNotFoundError):
        cli.run(args)                                        # the video file does not exist
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"mean": [0.0]}))
    with pytest.raises(KeyError):
        cli.load_scaler(bad)
    with pytest.raises(ValueError):
        cli.frames_to_tensor(torch.zeros(2, 2))

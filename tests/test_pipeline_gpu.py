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


def test_cli_uint8_ingest_and_mask(tmp_path):
    """The CLI on raw uint8 frames (device-side normalisation) equals the float path on frames normalised by the
    oracle's restatement of _preprocess_frame; --mask-type applies the articulator mask in memory."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import run_mri_video_inference as cli
    from mri2speech_b200 import synth
    from oracle import ingest
    ac, gen = _models()
    torch.save({"model_state_dict": ac.state_dict()}, tmp_path / "mri.pt")
    torch.save({"generator": gen.state_dict()}, tmp_path / "g_00000001")
    synth.write_scaler_json(tmp_path / "scaler.json")
    base = ["--video", str(tmp_path / "clip9.mp4"), "--mri-checkpoint", str(tmp_path / "mri.pt"),
            "--scaler-json", str(tmp_path / "scaler.json"), "--hifigan-config", os.path.join(ROOT, "config_custom.json"),
            "--hifigan-checkpoint", str(tmp_path / "g_00000001"), "--mri-code-dir", os.path.join(ROOT, "mri2speech_code")]
    u8 = synth.synthetic_clip_u8(9, 5)
    a_u8, mel_u8, _ = cli.run(cli.parse_args(base + ["--output-dir", str(tmp_path / "o1")]), frames=u8)
    f32 = torch.from_numpy(ingest.preprocess_clip(u8.numpy()))
    a_f, mel_f, _ = cli.run(cli.parse_args(base + ["--output-dir", str(tmp_path / "o2")]), frames=f32)
    assert np.abs(mel_u8 - mel_f).max() < 5e-3                     # dB scale (std 8-15 per normalised unit)
    assert a_u8.shape == a_f.shape == (5 * 420,)
    a_m, mel_m, _ = cli.run(cli.parse_args(base + ["--output-dir", str(tmp_path / "o3"), "--mask-type", "tongue",
                                                   "--mask-alpha", "0.0"]), frames=u8)
    assert np.abs(mel_m - mel_u8).max() > 1e-3                     # the mask changes the prediction
    with pytest.raises(ValueError):
        cli.run(cli.parse_args(base + ["--output-dir", str(tmp_path / "o4"), "--mask-type", "lip"]), frames=f32)


def test_alternate_vocoder_clis(tmp_path):
    """mel_to_audio_synthesis.py (ragged batch over .npy mels, bin padding) and inference_e2e.py (int16 wavs) against
    direct Generator calls."""
    import sys
    sys.path.insert(0, ROOT)
    import inference_e2e
    import mel_to_audio_synthesis as m2a
    from scipy.io import wavfile
    _, gen = _models()
    torch.save({"generator": gen.state_dict()}, tmp_path / "g_00000002")
    import shutil
    shutil.copy(os.path.join(ROOT, "config_custom.json"), tmp_path / "config.json")
    mels = tmp_path / "mels"
    mels.mkdir()
    g = torch.Generator().manual_seed(3)
    m_a = torch.randn(64, 9, generator=g) * 2 - 5
    m_b = torch.randn(60, 5, generator=g) * 2 - 5                 # 60 bins: zero-padded to 64 (:76-87)
    np.save(mels / "a_mel.npy", m_a.numpy())
    np.save(mels / "b.npy", m_b.numpy())
    res = m2a.main(["--input", str(mels), "--checkpoint_file", str(tmp_path / "g_00000002"),
                    "--config", os.path.join(ROOT, "config_custom.json"), "--output_dir", str(tmp_path / "syn")])
    assert sorted(n for n, _ in res) == ["a", "b"]
    gen = gen.cuda().eval()
    with torch.no_grad():
        ref_a = gen(m_a.cuda())[0, 0].cpu().numpy()
        ref_b = gen(torch.nn.functional.pad(m_b, (0, 0, 0, 4)).cuda())[0, 0].cpu().numpy()
    sr, wa = wavfile.read(tmp_path / "syn" / "a_from_mel.wav")
    _, wb = wavfile.read(tmp_path / "syn" / "b_from_mel.wav")
    assert sr == 11413 and wa.shape == (9 * 420,) and wb.shape == (5 * 420,)
    # PCM_16 like the reference's sf.write default (io_formats.write_wav_pcm16): one LSB = 1/32767
    assert wa.dtype == np.int16 and wb.dtype == np.int16
    assert np.abs(wa / 32767.0 - ref_a).max() < 2e-4 and np.abs(wb / 32767.0 - ref_b).max() < 2e-4   # ragged batch == own B=1 run
    assert (tmp_path / "syn" / "overall_synthesis_stats.json").exists()
    outs = inference_e2e.main(["--input_mels_dir", str(mels), "--output_dir", str(tmp_path / "e2e"),
                               "--checkpoint_file", str(tmp_path / "g_00000002")])
    assert len(outs) == 2
    sr, ia = wavfile.read(tmp_path / "e2e" / "a_mel_generated_e2e.wav")
    assert sr == 11413 and ia.dtype == np.int16 and np.abs(ia.astype(np.float32) / 32768.0 - ref_a).max() < 1e-3


def test_export_predicted_mels_cli(tmp_path):
    """scripts/export_predicted_mels.py: samples/<ID>/mri.npy -> <ID>.npy of shape (64, T) log-mel, ragged batches equal
    each clip's own B=1 result; --cpu and a missing samples directory are refused (SystemExit, like the reference)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import export_predicted_mels as cli
    from mri2speech_b200 import pipeline, synth
    ac, _ = _models()
    torch.save({"model_state_dict": ac.state_dict()}, tmp_path / "mri.pt")
    synth.write_scaler_json(tmp_path / "scaler.json")
    proc = tmp_path / "processed"
    clips = {"s01": synth.synthetic_clip(1, 5), "s02": synth.synthetic_clip(2, 3)}
    for name, clip in clips.items():
        (proc / "samples" / name).mkdir(parents=True)
        arr = clip.numpy() if name == "s01" else clip.numpy()[:, None]           # (T,H,W) and (T,1,H,W) both occur
        np.save(proc / "samples" / name / "mri.npy", arr)
    base = ["--processed_dir", str(proc), "--mri_checkpoint", str(tmp_path / "mri.pt"), "--scaler_json",
            str(tmp_path / "scaler.json"), "--output_dir", str(tmp_path / "mels"),
            "--mri_code_dir", os.path.join(ROOT, "mri2speech_code")]
    cli.export_mels(cli.parse_args(base))
    mean, std = pipeline.load_scaler(tmp_path / "scaler.json")
    ac = ac.cuda().eval()
    for name, clip in clips.items():
        got = np.load(tmp_path / "mels" / f"{name}.npy")
        assert got.shape == (64, clip.shape[0]) and got.dtype == np.float32
        with torch.no_grad():
            pred = ac(clip.unsqueeze(0).cuda())
            _, mel_log, _ = pipeline.mel_glue(pred, torch.from_numpy(mean), torch.from_numpy(std))
        assert np.abs(got - mel_log[0].t().cpu().numpy()).max() < 2e-3
    with pytest.raises(SystemExit):
        cli.export_mels(cli.parse_args(base + ["--cpu", "--overwrite"]))
    with pytest.raises(SystemExit):
        cli.export_mels(cli.parse_args(["--processed_dir", str(tmp_path / "nope")] + base[2:]))

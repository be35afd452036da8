"""CPU: output / interchange formats of the path -- wav flavours, mel bin fitting, the .npy mel
conventions, the asynchronous writer -- against the reference's file conventions (cited in io_formats.py)."""
import numpy as np
import pytest
import torch
from scipy.io import wavfile

from mri2speech_b200 import io_formats


def test_wav_flavours(tmp_path):
    audio = np.linspace(-0.999, 0.999, 2048).astype(np.float32)
    # the reference CLIs' sf.write(path, float32, sr): PCM_16, round(x * 32767) -- whichever library writes it
    io_formats.write_wav_pcm16(tmp_path / "p.wav", audio, 11413)
    sr, p = wavfile.read(tmp_path / "p.wav")
    assert sr == 11413 and p.dtype == np.int16
    assert np.array_equal(p, np.rint(audio.astype(np.float64) * 32767.0).astype(np.int16))
    io_formats.write_wav_float32(tmp_path / "f.wav", audio, 11413)
    sr, a = wavfile.read(tmp_path / "f.wav")
    assert sr == 11413 and a.dtype == np.float32 and np.array_equal(a, audio)
    io_formats.write_wav_int16(tmp_path / "i.wav", audio, 11413)
    sr, b = wavfile.read(tmp_path / "i.wav")
    # inference_e2e.py:52-57: audio * 32768 -> astype(int16) (truncation towards zero)
    assert sr == 11413 and b.dtype == np.int16 and np.array_equal(b, (audio * 32768.0).astype("int16"))


def test_fit_mel_bins_pad_and_truncate():
    mel = torch.arange(2 * 5 * 3, dtype=torch.float32).view(2, 5, 3)
    up = io_formats.fit_mel_bins(mel, 8)
    assert up.shape == (2, 8, 3) and torch.equal(up[:, :5], mel) and up[:, 5:].abs().sum() == 0
    down = io_formats.fit_mel_bins(mel, 4)
    assert down.shape == (2, 4, 3) and torch.equal(down, mel[:, :4])
    assert io_formats.fit_mel_bins(mel, 5) is mel


def test_mel_file_to_tensor_shapes():
    assert io_formats.mel_file_to_tensor(np.zeros((64, 7))).shape == (1, 64, 7)
    assert io_formats.mel_file_to_tensor(np.zeros((3, 64, 7))).shape == (1, 64, 7)   # batch > 1: first item
    with pytest.raises(ValueError):
        io_formats.mel_file_to_tensor(np.zeros(5))


def test_load_processed_clip(tmp_path):
    d = tmp_path / "samples" / "A01"
    d.mkdir(parents=True)
    clip = np.random.rand(4, 1, 8, 8).astype(np.float32)
    np.save(d / "mri.npy", clip)
    out = io_formats.load_processed_clip(d)
    assert out.shape == (4, 8, 8) and np.array_equal(out, clip[:, 0])
    np.save(d / "mri.npy", np.zeros((4, 8)))
    with pytest.raises(ValueError):
        io_formats.load_processed_clip(d)


def test_async_writer_writes_and_reports_errors(tmp_path):
    res = {"audio": torch.linspace(-1, 1, 840), "mel_db": torch.zeros(2, 64), "mel_log": torch.ones(2, 64)}
    with io_formats.AsyncWriter() as w:
        paths = io_formats.save_clip_outputs(w, res, tmp_path / "out", "clip3", 11413)
    assert [p.name for p in paths] == ["clip3_generated.wav", "clip3_mel.npy", "clip3_mel_log.npy"]
    sr, a = wavfile.read(paths[0])
    assert sr == 11413 and a.shape == (840,) and a.dtype == np.int16     # the reference's <stem>_generated.wav format
    assert np.load(paths[1]).shape == (2, 64) and np.load(paths[2]).mean() == 1.0

    def boom(_):
        raise OSError("disk full")
    w = io_formats.AsyncWriter()
    w.submit(torch.zeros(3), boom)
    with pytest.raises(OSError, match="disk full"):
        w.close()

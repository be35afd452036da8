"""Output / interchange formats of the path and an asynchronous writer.

Files the reference CLIs produce and consume, kept byte-compatible:
  * ``<stem>_generated.wav`` / ``<stem>_from_mel.wav``: 16-bit PCM at h.sampling_rate
    (scripts/run_mri_video_inference.py:166-170, mel_to_audio_synthesis.py:101-103: ``sf.write(path, float32_array, sr)``
    -- soundfile's default subtype for WAV is PCM_16, libsndfile converts with round(x * 32767); the same file is
    written here whether or not soundfile is installed: see ``write_wav_pcm16``);
  * ``<name>_generated_e2e.wav``: int16 PCM, ``(audio * 32768).astype(int16)`` (inference_e2e.py:52-57);
  * ``<stem>_mel.npy`` (T, n_mels) dB, ``<stem>_mel_log.npy`` (T, n_mels) log-power (run_mri_video_inference.py:171,248);
  * ``<stem>.npy`` (n_mels, T) log-power for HiFi-GAN fine-tuning (scripts/export_predicted_mels.py:98-99);
  * ``samples/<ID>/mri.npy`` (T, H, W) input clips (export_predicted_mels.py:84-90).
Disk I/O never sits on the GPU's critical path: results are copied to pinned host memory on a side stream and a
writer thread does the encoding, so the next micro-batch computes while the previous one is written.
"""
from __future__ import annotations

import os
import queue
import threading
from pathlib import Path
from typing import Callable, List, Optional

import numpy as np
import torch

MAX_WAV_VALUE = 32768.0  # meldataset.py:13 of the reference


def write_wav_pcm16(path, audio: np.ndarray, sampling_rate: int) -> None:
    """What the reference's ``sf.write(path, float32_audio, sr)`` leaves on disk: a PCM_16 WAV (soundfile's default
    subtype for the WAV container), samples = round-to-nearest(x * 32767) as libsndfile converts normalised floats.
    The subtype is explicit in BOTH branches, so the file format does not depend on which library is installed."""
    audio = np.asarray(audio, dtype=np.float32).reshape(-1)
    try:
        import soundfile as sf
        sf.write(str(path), audio, int(sampling_rate), subtype="PCM_16")
    except ImportError:
        from scipy.io import wavfile
        pcm = np.clip(np.rint(audio.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)
        wavfile.write(str(path), int(sampling_rate), pcm)


def write_wav_float32(path, audio: np.ndarray, sampling_rate: int) -> None:
    """IEEE-float WAV (format 3): NOT what the reference CLIs write -- a lossless option for callers that want the
    generator's float32 samples on disk."""
    audio = np.asarray(audio, dtype=np.float32).reshape(-1)
    try:
        import soundfile as sf
        sf.write(str(path), audio, int(sampling_rate), subtype="FLOAT")
    except ImportError:
        from scipy.io import wavfile
        wavfile.write(str(path), int(sampling_rate), audio)


def write_wav_int16(path, audio: np.ndarray, sampling_rate: int) -> None:
    """inference_e2e.py:52-57: scale by MAX_WAV_VALUE, truncating cast to int16."""
    from scipy.io import wavfile
    pcm = (np.asarray(audio, dtype=np.float32).reshape(-1) * MAX_WAV_VALUE).astype("int16")
    wavfile.write(str(path), int(sampling_rate), pcm)


def fit_mel_bins(mel: torch.Tensor, num_mels: int) -> torch.Tensor:
    """(B, M, T) -> (B, num_mels, T): truncate extra bins / zero-pad missing ones (mel_to_audio_synthesis.py:76-87)."""
    have = mel.size(1)
    if have > num_mels:
        return mel[:, :num_mels, :]
    if have < num_mels:
        return torch.nn.functional.pad(mel, (0, 0, 0, num_mels - have), "constant", 0)
    return mel


def mel_file_to_tensor(mel_np: np.ndarray) -> torch.Tensor:
    """.npy mel -> (1, M, T) float32 (mel_to_audio_synthesis.py:60-71: 2-D gets a batch axis, 3-D keeps item 0)."""
    t = torch.as_tensor(np.asarray(mel_np), dtype=torch.float32)
    if t.dim() == 2:
        return t.unsqueeze(0)
    if t.dim() == 3:
        return t[0:1]
    raise ValueError(f"Invalid mel spectrogram dimensions: {tuple(t.shape)}")


def load_processed_clip(sample_dir) -> np.ndarray:
    """samples/<ID>/mri.npy -> (T, H, W) float32 (export_predicted_mels.py:84-90 squeezes a channel axis)."""
    arr = np.load(Path(sample_dir) / "mri.npy")
    if arr.ndim == 4:
        arr = arr[:, 0] if arr.shape[1] == 1 else arr[..., 0]
    if arr.ndim != 3:
        raise ValueError(f"mri.npy must be (T,H,W) or (T,1,H,W), got {arr.shape}")
    return arr


class AsyncWriter:
    """Pinned D2H on a side stream + one writer thread.  ``submit(tensor, fn)`` returns immediately; ``fn(ndarray)``
    runs on the writer thread once the copy has landed.  ``close()`` drains the queue and re-raises the first
    writer error."""

    def __init__(self, max_pending: int = 64):
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self._err: Optional[BaseException] = None
        self._copy_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._t = threading.Thread(target=self._run, name="m2s-writer", daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            host, event, fn = item
            try:
                if event is not None:
                    event.synchronize()
                fn(host.numpy())
            except BaseException as exc:  # surfaced by close()
                if self._err is None:
                    self._err = exc

    def submit(self, t: torch.Tensor, fn: Callable[[np.ndarray], None]) -> None:
        if t.is_cuda:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._copy_stream.wait_stream(torch.cuda.current_stream(t.device))
            with torch.cuda.stream(self._copy_stream):
                host.copy_(t, non_blocking=True)
                t.record_stream(self._copy_stream)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._q.put((host, ev, fn))
        else:
            self._q.put((t.detach().contiguous(), None, fn))

    def close(self) -> None:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def save_clip_outputs(writer: AsyncWriter, result: dict, output_dir, stem: str, sampling_rate: int) -> List[Path]:
    """The four files of run_mri_video_inference.py for one clip (the PNG is the CLI's business), asynchronously."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    wav, mel, mel_log = out / f"{stem}_generated.wav", out / f"{stem}_mel.npy", out / f"{stem}_mel_log.npy"
    writer.submit(result["audio"], lambda a, p=wav: write_wav_pcm16(p, a, sampling_rate))
    writer.submit(result["mel_db"], lambda a, p=mel: np.save(p, a.astype(np.float32)))
    writer.submit(result["mel_log"], lambda a, p=mel_log: np.save(p, a.astype(np.float32)))
    return [wav, mel, mel_log]

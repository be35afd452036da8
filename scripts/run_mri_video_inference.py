#!/usr/bin/env python
"""rtMRI video -> speech on the B200 path.  Same flags, outputs and helper names as the reference CLI
(scripts/run_mri_video_inference.py:187-255): <stem>_generated.wav, <stem>_mel.npy (dB, (T,64)),
<stem>_mel.png, <stem>_mel_log.npy.  Helpers keep their names because other tools import them
(scripts/mri_gradcam_formant.py:25-30).  Video decoding / file writing stay on the CPU; the three model
stages and the mel glue run through libm2s (no CPU fallback: a CUDA sm_100 device is required)."""
import argparse
import json
import sys
from pathlib import Path

PROJECT_ROOT = Path(__file__).resolve().parents[1]
if str(PROJECT_ROOT) not in sys.path:
    sys.path.insert(0, str(PROJECT_ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models import Generator  # noqa: E402
from mri2speech_b200 import pipeline  # noqa: E402


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def _ensure_sys_path(path):
    if path and Path(path).exists():
        sys.path.insert(0, str(path))


def _preprocess_frame(frame: np.ndarray, target_size=(256, 256)) -> np.ndarray:
    """BGR/gray frame -> float32 (H,W) in [0,1]: gray, bilinear resize, z-score, then min-max
    (net effect: per-frame min-max; a constant frame maps to zeros)."""
    import cv2
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
    if gray.shape[::-1] != tuple(target_size):
        gray = cv2.resize(gray, tuple(target_size), interpolation=cv2.INTER_LINEAR)
    gray = gray.astype(np.float32)
    sd = gray.std()
    gray = (gray - gray.mean()) / sd if sd > 0 else gray - gray.mean()
    lo, hi = gray.min(), gray.max()
    return (gray - lo) / (hi - lo) if hi > lo else np.zeros_like(gray)


def load_video_frames(video_path, target_size=(256, 256), max_frames=None) -> torch.Tensor:
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise ValueError(f"Unable to open video: {video_path}")
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if max_frames is not None:
        total = min(total, max_frames)
    frames = []
    while len(frames) < total:
        ok, frame = cap.read()
        if not ok:
            break
        frames.append(_preprocess_frame(frame, target_size))
    cap.release()
    if not frames:
        raise ValueError("No frames could be read from video")
    return torch.from_numpy(np.asarray(frames, dtype=np.float32))


def load_video_frames_u8(video_path, target_size=(256, 256), max_frames=None) -> torch.Tensor:
    """Decoded frames as raw uint8 gray (T,H,W): the gray conversion and the bilinear resize of _preprocess_frame
    (:35-40 of the reference, both uint8 -> uint8 in OpenCV) stay on the host; the float cast, z-score and min-max
    (:41-53) run on the device, fused into the encoder's stem load (m2s_acoustic_forward_u8) -- 4x fewer bytes over
    PCIe / HBM than float32 frames."""
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise ValueError(f"Unable to open video: {video_path}")
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if max_frames is not None:
        total = min(total, max_frames)
    frames = []
    while len(frames) < total:
        ok, frame = cap.read()
        if not ok:
            break
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        if gray.shape[::-1] != tuple(target_size):
            gray = cv2.resize(gray, tuple(target_size), interpolation=cv2.INTER_LINEAR)
        frames.append(gray)
    cap.release()
    if not frames:
        raise ValueError("No frames could be read from video")
    return torch.from_numpy(np.asarray(frames, dtype=np.uint8))


load_scaler = pipeline.load_scaler


def load_hifigan(config_path, checkpoint_path, device: torch.device):
    with open(config_path, "r", encoding="utf-8") as f:
        h = AttrDict(json.load(f))
    generator = Generator(h).to(device)
    ckpt = torch.load(checkpoint_path, map_location=device)
    if "generator" not in ckpt:
        raise KeyError("HiFi-GAN checkpoint missing 'generator' state")
    generator.load_state_dict(ckpt["generator"])
    generator.eval()
    # weight-norm is folded inside libm2s; the reference's best-effort removal is harmless and kept
    from torch.nn.utils import remove_weight_norm
    for module in list(generator.ups) + [generator.conv_post]:
        try:
            remove_weight_norm(module)
        except (ValueError, AttributeError):
            pass
    for res in generator.resblocks:
        try:
            res.remove_weight_norm()
        except (ValueError, AttributeError):
            pass
    return generator, h


def build_mri_model(args, device: torch.device):
    code_dir = Path(args.mri_code_dir) if args.mri_code_dir else None
    if code_dir is None:
        code_dir = Path(args.mri_checkpoint).resolve().parent.parent / "mri2speech_code"
        if not code_dir.exists():
            code_dir = PROJECT_ROOT / "mri2speech_code"
    _ensure_sys_path(code_dir)
    try:
        from mri_acoustic_model import build_acoustic_model
    except ImportError as exc:
        raise ImportError("Failed to import mri_acoustic_model. Use --mri-code-dir to point to the "
                          "mri2speech_code directory.") from exc
    model = build_acoustic_model(n_mels=args.n_mels, cnn_pretrained=False, rnn_hidden=args.rnn_hidden,
                                 dropout=args.dropout, use_checkpoint=False, ckpt_segments=2,
                                 use_reentrant=False).to(device)
    checkpoint = torch.load(args.mri_checkpoint, map_location=device)
    state_dict = checkpoint.get("model_state_dict", checkpoint)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    if missing:
        print(f"[WARN] Missing keys when loading MRI model: {missing}")
    if unexpected:
        print(f"[WARN] Unexpected keys when loading MRI model: {unexpected}")
    model.eval()
    return model


def frames_to_tensor(frames: torch.Tensor, use_channel: bool = True) -> torch.Tensor:
    if frames.dim() != 3:
        raise ValueError(f"Expected frames tensor of shape (T,H,W), got {tuple(frames.shape)}")
    frames = frames.unsqueeze(0)
    return frames.unsqueeze(2) if use_channel else frames


def denormalize_mel(mel_normalized: torch.Tensor, mean: np.ndarray, std: np.ndarray) -> torch.Tensor:
    mel_db, _, _ = pipeline.mel_glue(mel_normalized, torch.from_numpy(np.asarray(mean, np.float32)),
                                     torch.from_numpy(np.asarray(std, np.float32)), want_log=False)
    return mel_db


def _write_wav(path, audio: np.ndarray, sampling_rate: int):
    try:
        import soundfile as sf
        sf.write(path, audio, sampling_rate)
    except ImportError:
        from scipy.io import wavfile
        wavfile.write(str(path), int(sampling_rate), np.asarray(audio, dtype=np.float32))


def save_outputs(audio: np.ndarray, mel: np.ndarray, output_dir, sampling_rate: int, stem: str):
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    audio_path = output_dir / f"{stem}_generated.wav"
    _write_wav(audio_path, audio, sampling_rate)
    mel_path = output_dir / f"{stem}_mel.npy"
    np.save(mel_path, mel)
    fig_path = output_dir / f"{stem}_mel.png"
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure(figsize=(12, 4))
        plt.imshow(mel.T, aspect="auto", origin="lower", cmap="viridis")
        plt.colorbar()
        plt.title(f"Generated Mel Spectrogram - {stem}")
        plt.xlabel("Time")
        plt.ylabel("Mel bins")
        plt.tight_layout()
        plt.savefig(fig_path, dpi=150)
        plt.close()
    except ImportError:
        print("[WARN] matplotlib is not installed; skipping the mel figure")
        fig_path = None
    return audio_path, mel_path, fig_path


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="rtMRI -> Speech inference (OTN-like MRI model + HiFi-GAN) on B200")
    p.add_argument("--video", required=True, help="Input rtMRI video (.mp4)")
    p.add_argument("--mri-checkpoint", required=True, help="Path to OTN-like MRI checkpoint (.pt)")
    p.add_argument("--scaler-json", required=True, help="Path to scaler.json (contains per-mel mean/std)")
    p.add_argument("--hifigan-config", required=True, help="HiFi-GAN config JSON")
    p.add_argument("--hifigan-checkpoint", required=True, help="HiFi-GAN generator checkpoint")
    p.add_argument("--output-dir", required=True, help="Directory to save generated artifacts")
    p.add_argument("--mri-code-dir", help="Directory containing mri_acoustic_model.py")
    p.add_argument("--max-frames", type=int, default=None, help="Optional max number of frames to process")
    p.add_argument("--n-mels", type=int, default=64)
    p.add_argument("--rnn-hidden", type=int, default=640)
    p.add_argument("--dropout", type=float, default=0.5)
    # in-memory articulator masking (replaces the mask_rtmri_video.py re-encode + second run of this CLI)
    p.add_argument("--mask-type", choices=["lip", "tongue"], default=None, help="Apply a preset articulator mask")
    p.add_argument("--mask-alpha", type=float, default=0.1, help="Residual intensity inside the mask (0-1)")
    p.add_argument("--mask-blur-kernel", type=int, default=11, help="Gaussian blur kernel size for soft edges")
    return p.parse_args(argv)


def run(args, frames: torch.Tensor = None):
    """The body of main(); ``frames`` (T,H,W) may be injected by tests instead of decoding a video."""
    video_path = Path(args.video)
    if frames is None and not video_path.exists():
        raise FileNotFoundError(f"Video file not found: {video_path}")
    mean, std = load_scaler(Path(args.scaler_json))
    if len(mean) != args.n_mels or len(std) != args.n_mels:
        raise ValueError("Scaler mean/std length does not match n_mels")
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on sm_100 CUDA devices only (there is no CPU fallback)")
    device = torch.device("cuda")
    print(f"[INFO] Using device: {device}")
    if frames is None:  # raw uint8 frames: normalisation happens on the device (fused ingest)
        frames = load_video_frames_u8(video_path, target_size=(256, 256), max_frames=args.max_frames)
    if args.mask_type:
        if frames.dtype != torch.uint8:
            raise ValueError("--mask-type applies to raw uint8 frames (the mask is applied before normalisation)")
        from mri2speech_b200 import masking
        mask = torch.from_numpy(masking.preset_mask(args.mask_type, args.mask_alpha, tuple(frames.shape[-2:]),
                                                    args.mask_blur_kernel))
    else:
        mask = None
    frames_tensor = frames_to_tensor(frames, use_channel=True).to(device)

    mri_model = build_mri_model(args, device)
    with torch.no_grad():
        pred_norm = (mri_model(frames_tensor, mask=mask) if mask is not None else mri_model(frames_tensor)).squeeze(0)
    print(f"[INFO] Predicted normalized mel shape: {tuple(pred_norm.shape)}")
    mel_db, mel_log, voc_in = pipeline.mel_glue(pred_norm, torch.from_numpy(mean), torch.from_numpy(std))
    mel_denorm_np = mel_db.cpu().numpy().astype(np.float32)
    print(f"[INFO] Mel (denormalized dB) range: {mel_denorm_np.min():.3f} .. {mel_denorm_np.max():.3f}")
    mel_log_np = mel_log.cpu().numpy().astype(np.float32)
    print(f"[INFO] Mel (log-power) range: {mel_log_np.min():.3f} .. {mel_log_np.max():.3f}")

    generator, hifigan_config = load_hifigan(Path(args.hifigan_config), Path(args.hifigan_checkpoint), device)
    with torch.no_grad():
        audio = generator(voc_in).squeeze().cpu().numpy()
    print(f"[INFO] Generated audio length: {audio.shape[0]} samples")

    stem = video_path.stem
    output_dir = Path(args.output_dir)
    audio_path, mel_path, fig_path = save_outputs(audio, mel_denorm_np, output_dir, hifigan_config.sampling_rate, stem)
    log_mel_path = output_dir / f"{stem}_mel_log.npy"
    np.save(log_mel_path, mel_log_np)
    print("[DONE] Inference complete.")
    print(f"  Audio : {audio_path}")
    print(f"  Mel   : {mel_path}")
    print(f"  LogMel: {log_mel_path}")
    print(f"  Figure: {fig_path}")
    return audio, mel_denorm_np, mel_log_np


def main():
    run(parse_args())


if __name__ == "__main__":
    main()

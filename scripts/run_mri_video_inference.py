#!/usr/bin/env python
"""rtMRI video -> speech on the B200 path.

Drop-in for the reference CLI of the same name (its flags: scripts/run_mri_video_inference.py:187-200; its outputs:
<stem>_generated.wav, <stem>_mel.npy (dB, (T, 64)), <stem>_mel.png, <stem>_mel_log.npy, :166-184,248-249).  The helper
names other tools import from that module (scripts/mri_gradcam_formant.py:25-30) exist here with the same call
signatures; what they do is written for this code base: video decoding and file writing stay on the CPU, the frames
travel as uint8, and the three model stages plus the mel glue run through libm2s (no CPU fallback: a CUDA sm_100
device is required).  `--mask-type / --mask-alpha / --mask-blur-kernel` apply an articulator mask in memory."""
import argparse
import contextlib
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [p for p in (str(REPO),) if p not in sys.path]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from env import AttrDict  # noqa: E402,F401  (re-exported: the reference module defines it too)
from models import Generator  # noqa: E402
from mri2speech_b200 import io_formats, pipeline  # noqa: E402

FRAME_SIZE = (256, 256)


# ---------------------------------------------------------------------------------------------------------------
# video -> frames
# ---------------------------------------------------------------------------------------------------------------
def _gray_frames(video_path, target_size=FRAME_SIZE, max_frames=None):
    """Yield the clip's frames as uint8 gray images of ``target_size`` (gray conversion + bilinear resize, both
    uint8 -> uint8 in OpenCV, i.e. the first half of the reference's _preprocess_frame)."""
    import cv2
    reader = cv2.VideoCapture(str(video_path))
    try:
        if not reader.isOpened():
            raise ValueError(f"Unable to open video: {video_path}")
        budget = int(reader.get(cv2.CAP_PROP_FRAME_COUNT))
        if max_frames is not None:
            budget = min(budget, int(max_frames))
        size = tuple(target_size)
        for _ in range(budget):
            ok, image = reader.read()
            if not ok:
                return
            if image.ndim == 3:
                image = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
            if image.shape[::-1] != size:
                image = cv2.resize(image, size, interpolation=cv2.INTER_LINEAR)
            yield image
    finally:
        reader.release()


def _normalise(gray: np.ndarray) -> np.ndarray:
    """float32 z-score followed by min-max to [0, 1] (a constant frame maps to zeros): the second half of the
    reference's _preprocess_frame; on the B200 path this runs fused into the encoder's stem load instead."""
    g = np.asarray(gray, dtype=np.float32)
    g = g - g.mean()
    spread = g.std()
    if spread > 0:
        g = g / spread
    lo, hi = float(g.min()), float(g.max())
    return (g - lo) / (hi - lo) if hi > lo else np.zeros_like(g)


def _preprocess_frame(frame: np.ndarray, target_size=FRAME_SIZE) -> np.ndarray:
    """One decoded frame (BGR or gray) -> float32 (H, W) in [0, 1], as the reference helper of this name."""
    import cv2
    if frame.ndim == 3:
        frame = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
    if frame.shape[::-1] != tuple(target_size):
        frame = cv2.resize(frame, tuple(target_size), interpolation=cv2.INTER_LINEAR)
    return _normalise(frame)


def _stack(frames, dtype) -> torch.Tensor:
    if not frames:
        raise ValueError("No frames could be read from video")
    return torch.from_numpy(np.stack(frames).astype(dtype, copy=False))


def load_video_frames(video_path, target_size=FRAME_SIZE, max_frames=None) -> torch.Tensor:
    """(T, H, W) float32 frames normalised on the host (the reference's contract for this helper)."""
    return _stack([_normalise(g) for g in _gray_frames(video_path, target_size, max_frames)], np.float32)


def load_video_frames_u8(video_path, target_size=FRAME_SIZE, max_frames=None) -> torch.Tensor:
    """(T, H, W) raw uint8 gray frames: normalisation is left to the device (m2s_acoustic_forward_u8), which moves
    4x fewer bytes over PCIe / HBM than float32 frames."""
    return _stack(list(_gray_frames(video_path, target_size, max_frames)), np.uint8)


def frames_to_tensor(frames: torch.Tensor, use_channel: bool = True) -> torch.Tensor:
    """(T, H, W) -> (1, T, 1, H, W) (or (1, T, H, W) without the channel axis)."""
    if frames.dim() != 3:
        raise ValueError(f"Expected frames tensor of shape (T,H,W), got {tuple(frames.shape)}")
    batch = frames[None]
    return batch[:, :, None] if use_channel else batch


# ---------------------------------------------------------------------------------------------------------------
# checkpoints
# ---------------------------------------------------------------------------------------------------------------
load_scaler = pipeline.load_scaler


def _ensure_sys_path(path):
    if path and Path(path).is_dir() and str(path) not in sys.path:
        sys.path.insert(0, str(path))


def load_hifigan(config_path, checkpoint_path, device: torch.device):
    """(Generator in eval mode, config) from config_custom.json + a ``{"generator": state_dict}`` checkpoint."""
    h = AttrDict(json.loads(Path(config_path).read_text(encoding="utf-8")))
    payload = torch.load(checkpoint_path, map_location=device)
    if "generator" not in payload:
        raise KeyError("HiFi-GAN checkpoint missing 'generator' state")
    generator = Generator(h).to(device).eval()
    generator.load_state_dict(payload["generator"])
    # libm2s folds weight-norm itself when it plans the device weights, so stripping it is optional; the reference's
    # best-effort removal (conv_pre carries none, hence per module and forgiving) is kept so that both code paths
    # see the same parameter names afterwards.
    from torch.nn.utils import remove_weight_norm
    for module in (*generator.ups, generator.conv_post, *generator.resblocks):
        strip = getattr(module, "remove_weight_norm", None) or (lambda m=module: remove_weight_norm(m))
        with contextlib.suppress(ValueError, AttributeError):
            strip()
    return generator, h


def _acoustic_code_dir(args) -> Path:
    if args.mri_code_dir:
        return Path(args.mri_code_dir)
    beside_checkpoint = Path(args.mri_checkpoint).resolve().parents[1] / "mri2speech_code"
    return beside_checkpoint if beside_checkpoint.exists() else REPO / "mri2speech_code"


def build_mri_model(args, device: torch.device):
    """The acoustic model (frame CNN + BiLSTM + mel head) with the checkpoint's weights, in eval mode."""
    _ensure_sys_path(_acoustic_code_dir(args))
    try:
        from mri_acoustic_model import build_acoustic_model
    except ImportError as exc:
        raise ImportError("Failed to import mri_acoustic_model. Use --mri-code-dir to point to the "
                          "mri2speech_code directory.") from exc
    model = build_acoustic_model(n_mels=args.n_mels, cnn_pretrained=False, rnn_hidden=args.rnn_hidden,
                                 dropout=args.dropout, use_checkpoint=False, ckpt_segments=2, use_reentrant=False)
    payload = torch.load(args.mri_checkpoint, map_location=device)
    report = model.to(device).load_state_dict(payload.get("model_state_dict", payload), strict=False)
    for kind, keys in (("Missing", report.missing_keys), ("Unexpected", report.unexpected_keys)):
        if keys:
            print(f"[WARN] {kind} keys when loading MRI model: {keys}")
    return model.eval()


# ---------------------------------------------------------------------------------------------------------------
# mel glue and outputs
# ---------------------------------------------------------------------------------------------------------------
def denormalize_mel(mel_normalized: torch.Tensor, mean: np.ndarray, std: np.ndarray) -> torch.Tensor:
    as_tensor = lambda v: torch.from_numpy(np.asarray(v, np.float32))   # noqa: E731
    return pipeline.mel_glue(mel_normalized, as_tensor(mean), as_tensor(std), want_log=False)[0]


def _mel_figure(mel: np.ndarray, path: Path, title: str):
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot
    except ImportError:
        print("[WARN] matplotlib is not installed; skipping the mel figure")
        return None
    figure, axes = pyplot.subplots(figsize=(12, 4))
    image = axes.imshow(mel.T, aspect="auto", origin="lower", cmap="viridis")
    figure.colorbar(image, ax=axes)
    axes.set(title=title, xlabel="Time", ylabel="Mel bins")
    figure.tight_layout()
    figure.savefig(path, dpi=150)
    pyplot.close(figure)
    return path


def save_outputs(audio: np.ndarray, mel: np.ndarray, output_dir, sampling_rate: int, stem: str):
    """<stem>_generated.wav (float32), <stem>_mel.npy and <stem>_mel.png under ``output_dir``."""
    target = Path(output_dir)
    target.mkdir(parents=True, exist_ok=True)
    audio_path, mel_path = target / f"{stem}_generated.wav", target / f"{stem}_mel.npy"
    io_formats.write_wav_pcm16(audio_path, audio, sampling_rate)
    np.save(mel_path, mel)
    fig_path = _mel_figure(mel, target / f"{stem}_mel.png", f"Generated Mel Spectrogram - {stem}")
    return audio_path, mel_path, fig_path


# ---------------------------------------------------------------------------------------------------------------
# CLI
# ---------------------------------------------------------------------------------------------------------------
def parse_args(argv=None):
    p = argparse.ArgumentParser(description="rtMRI -> Speech inference (OTN-like MRI model + HiFi-GAN) on B200")
    for flag, text in (("--video", "Input rtMRI video (.mp4)"),
                       ("--mri-checkpoint", "Path to OTN-like MRI checkpoint (.pt)"),
                       ("--scaler-json", "Path to scaler.json (contains per-mel mean/std)"),
                       ("--hifigan-config", "HiFi-GAN config JSON"),
                       ("--hifigan-checkpoint", "HiFi-GAN generator checkpoint"),
                       ("--output-dir", "Directory to save generated artifacts")):
        p.add_argument(flag, required=True, help=text)
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


def _articulator_mask(args, frames: torch.Tensor):
    if not args.mask_type:
        return None
    if frames.dtype != torch.uint8:
        raise ValueError("--mask-type applies to raw uint8 frames (the mask is applied before normalisation)")
    from mri2speech_b200 import masking
    return torch.from_numpy(masking.preset_mask(args.mask_type, args.mask_alpha, tuple(frames.shape[-2:]),
                                                args.mask_blur_kernel))


def _span(name: str, values: np.ndarray):
    print(f"[INFO] Mel ({name}) range: {values.min():.3f} .. {values.max():.3f}")


def run(args, frames: torch.Tensor = None):
    """The body of main(); ``frames`` (T,H,W) may be injected by tests instead of decoding a video."""
    video_path = Path(args.video)
    if frames is None and not video_path.exists():
        raise FileNotFoundError(f"Video file not found: {video_path}")
    mean, std = load_scaler(Path(args.scaler_json))
    if not (len(mean) == len(std) == args.n_mels):
        raise ValueError("Scaler mean/std length does not match n_mels")
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on sm_100 CUDA devices only (there is no CPU fallback)")
    device = torch.device("cuda")
    print(f"[INFO] Using device: {device}")

    if frames is None:  # raw uint8 frames: normalisation happens on the device (fused ingest)
        frames = load_video_frames_u8(video_path, target_size=FRAME_SIZE, max_frames=args.max_frames)
    mask = _articulator_mask(args, frames)
    clip = frames_to_tensor(frames, use_channel=True).to(device)

    acoustic = build_mri_model(args, device)
    with torch.no_grad():
        pred_norm = (acoustic(clip) if mask is None else acoustic(clip, mask=mask))[0]
    print(f"[INFO] Predicted normalized mel shape: {tuple(pred_norm.shape)}")
    mel_db, mel_log, voc_in = pipeline.mel_glue(pred_norm, torch.from_numpy(mean), torch.from_numpy(std))
    mel_db_np, mel_log_np = (t.cpu().numpy().astype(np.float32) for t in (mel_db, mel_log))
    _span("denormalized dB", mel_db_np)
    _span("log-power", mel_log_np)

    generator, hifigan_config = load_hifigan(Path(args.hifigan_config), Path(args.hifigan_checkpoint), device)
    with torch.no_grad():
        audio = generator(voc_in).reshape(-1).cpu().numpy()
    print(f"[INFO] Generated audio length: {audio.shape[0]} samples")

    stem, out_dir = video_path.stem, Path(args.output_dir)
    written = dict(zip(("Audio", "Mel", "Figure"),
                       save_outputs(audio, mel_db_np, out_dir, hifigan_config.sampling_rate, stem)))
    written["LogMel"] = out_dir / f"{stem}_mel_log.npy"
    np.save(written["LogMel"], mel_log_np)
    print("[DONE] Inference complete.")
    for label in ("Audio", "Mel", "LogMel", "Figure"):
        print(f"  {label:<6}: {written[label]}")
    return audio, mel_db_np, mel_log_np


def main():
    run(parse_args())


if __name__ == "__main__":
    main()

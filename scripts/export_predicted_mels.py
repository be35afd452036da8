#!/usr/bin/env python
"""Batch export of predicted log-mels on the B200 path.  Same flags and outputs as the reference
(scripts/export_predicted_mels.py:43-118): one <stem>.npy of shape (64, T) per samples/<stem>/mri.npy.
Unlike the reference's serial B=1 loop, clips are packed into ragged batches (--batch-frames);
every clip still equals its own B=1 result.  --cpu is accepted and refused (no CPU fallback)."""
import argparse
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [p for p in (str(REPO),) if p not in sys.path]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mri2speech_b200 import io_formats, pipeline  # noqa: E402


def load_scaler(scaler_path: Path):
    """(mean, std) of scaler.json as float32 tensors."""
    return tuple(torch.from_numpy(v) for v in pipeline.load_scaler(scaler_path))


def build_model(checkpoint_path: Path, n_mels: int, device: torch.device):
    """The acoustic model with the checkpoint's weights (plain state_dict or {"model_state_dict": ...}), eval mode."""
    from mri_acoustic_model import build_acoustic_model
    model = build_acoustic_model(n_mels=n_mels, cnn_pretrained=False, rnn_hidden=640, dropout=0.5,
                                 use_checkpoint=False, ckpt_segments=2, use_reentrant=False)
    payload = torch.load(checkpoint_path, map_location=device)
    report = model.to(device).load_state_dict(payload.get("model_state_dict", payload), strict=False)
    for kind, keys in (("missing", report.missing_keys), ("unexpected", report.unexpected_keys)):
        if keys:
            print(f"[WARN] {kind} keys when loading MRI model: {keys}")
    return model.eval()


def _work_list(samples_dir: Path, output_dir: Path, overwrite: bool):
    """(samples/<stem>/mri.npy, <output_dir>/<stem>.npy) pairs still to do, in name order."""
    clips = sorted((d for d in samples_dir.iterdir() if d.is_dir()), key=lambda d: d.name)
    if not clips:
        raise SystemExit(f"No sample folders found under {samples_dir}")
    todo = []
    for clip_dir in clips:
        target = output_dir / f"{clip_dir.name}.npy"
        if target.exists() and not overwrite:
            continue
        source = clip_dir / "mri.npy"
        if source.is_file():
            todo.append((source, target))
        else:
            print(f"[WARN] MRI file missing for {clip_dir.name}, skipping")
    return todo


def export_mels(args: argparse.Namespace) -> None:
    samples_dir = Path(args.processed_dir).resolve() / "samples"
    if not samples_dir.is_dir():
        raise SystemExit(f"samples directory not found: {samples_dir}")
    output_dir = Path(args.output_dir).resolve()
    output_dir.mkdir(parents=True, exist_ok=True)
    mean, std = load_scaler(Path(args.scaler_json).resolve())
    if mean.numel() != std.numel():
        raise SystemExit("Scaler mean/std length mismatch")
    if args.cpu:
        raise SystemExit("--cpu: this build has no CPU fallback (sm_100 CUDA device required)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this build has no CPU fallback")
    device = torch.device("cuda")
    print(f"[INFO] Using device: {device}")

    code_dir = Path(args.mri_code_dir).resolve() if args.mri_code_dir else REPO / "mri2speech_code"
    if code_dir.is_dir() and str(code_dir) not in sys.path:
        sys.path.insert(0, str(code_dir))
    model = build_model(Path(args.mri_checkpoint).resolve(), mean.numel(), device)
    todo = _work_list(samples_dir, output_dir, args.overwrite)

    with torch.no_grad(), io_formats.AsyncWriter() as writer:
        pending = []

        def flush():
            if not pending:
                return
            lens = [c.shape[0] for c, _ in pending]
            tmax = max(lens)
            H, W = pending[0][0].shape[-2:]
            batch = torch.zeros(len(pending), tmax, H, W, device=device)
            for b, (clip, _) in enumerate(pending):
                batch[b, : clip.shape[0]] = torch.from_numpy(clip).to(device)
            lt = torch.tensor(lens, dtype=torch.int32)
            pred = model(batch, lengths=lt)
            _, mel_log, _ = pipeline.mel_glue(pred, mean, std, lt, want_db=False)
            for b, (_, out_path) in enumerate(pending):  # (64, T) log-mel; D2H + np.save overlap the next batch
                writer.submit(mel_log[b, : lens[b]].transpose(0, 1).contiguous(),
                              lambda a, p=out_path: np.save(p, a.astype(np.float32)))
            pending.clear()

        frames_in_batch = 0
        for mri_path, out_path in todo:
            clip = io_formats.load_processed_clip(mri_path.parent).astype(np.float32)
            if pending and (frames_in_batch + clip.shape[0] > args.batch_frames
                            or clip.shape[-2:] != pending[0][0].shape[-2:]):
                flush()
                frames_in_batch = 0
            pending.append((clip, out_path))
            frames_in_batch += clip.shape[0]
        flush()
    print(f"[DONE] exported {len(todo)} mel files to {output_dir}")


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Export predicted log-mel features for HiFi-GAN fine-tuning (B200).")
    p.add_argument("--processed_dir", required=True, help="rtMRI processed dataset root (contains samples/).")
    p.add_argument("--mri_checkpoint", required=True, help="Path to trained MRI->mel checkpoint (.pt).")
    p.add_argument("--scaler_json", required=True, help="Path to scaler.json (mean/std for denormalization).")
    p.add_argument("--output_dir", required=True, help="Directory for log-mel numpy files (one per sample, [64, T]).")
    p.add_argument("--mri_code_dir", help="Directory containing mri_acoustic_model.py.")
    p.add_argument("--cpu", action="store_true", help="Accepted for compatibility; refused (no CPU fallback).")
    p.add_argument("--overwrite", action="store_true", help="Regenerate files even if they already exist.")
    p.add_argument("--batch_frames", type=int, default=2048, help="Frames per ragged batch (extension).")
    return p.parse_args(argv)


def main() -> None:
    export_mels(parse_args())


if __name__ == "__main__":
    main()

"""Seeded synthetic inputs and weights for tests and benchmarks.

No real rtMRI data or checkpoints are available offline, so every measurement uses these generators:
  * clips: smooth random fields, per-frame min-max normalised to [0,1] like
    scripts/run_mri_video_inference.py:34-54 does with real frames;
  * mels for the vocoder-only configuration: randn * 2 - 5 (log-power-like range);
  * scaler: mean = linspace(-60,-20,64), std = linspace(8,15,64)  (scaler.json format,
    mri2speech_code/preprocess_rtmri_data.py:192-194);
  * weights: the modules' own seeded default init (seed 1234 = config_custom.json:9); an optional
    "scaled" variant randomises BatchNorm statistics so that BN folding is actually exercised.
"""
from __future__ import annotations

import json
from typing import List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SAMPLING_RATE = 11413
HOP = 420


def synthetic_clip(clip_id: int, frames: int, size: int = 256) -> torch.Tensor:
    """(frames, size, size) float32 in [0,1]; deterministic in (clip_id, frames)."""
    g = torch.Generator().manual_seed(1234 + clip_id)
    low = torch.randn(frames, 1, 32, 32, generator=g)
    # temporal + spatial smoothing so consecutive frames are correlated like a real clip
    low = F.avg_pool2d(F.pad(low, (1, 1, 1, 1), mode="replicate"), 3, stride=1)
    k = torch.ones(1, 1, 3) / 3.0
    lt = low.view(frames, -1).t().unsqueeze(1)
    lt = F.conv1d(F.pad(lt, (1, 1), mode="replicate"), k).squeeze(1).t().view(frames, 1, 32, 32)
    up = F.interpolate(lt, size=(size, size), mode="bilinear", align_corners=False)
    x = torch.sigmoid(3.0 * up[:, 0])
    mn = x.amin(dim=(1, 2), keepdim=True)
    mx = x.amax(dim=(1, 2), keepdim=True)
    return ((x - mn) / (mx - mn).clamp_min(1e-12)).contiguous()


def synthetic_clip_u8(clip_id: int, frames: int, size: int = 256) -> torch.Tensor:
    """(frames, size, size) uint8 raw gray frames as a video decoder delivers them: the smooth field of
    ``synthetic_clip`` scaled to a per-frame dynamic range that does NOT span 0..255, so that the device
    min-max normalisation is exercised."""
    x = synthetic_clip(clip_id, frames, size)
    g = torch.Generator().manual_seed(4321 + clip_id)
    lo = torch.randint(0, 40, (frames, 1, 1), generator=g).float()
    hi = torch.randint(150, 256, (frames, 1, 1), generator=g).float()
    return (lo + x * (hi - lo)).round().clamp(0, 255).to(torch.uint8).contiguous()


def synthetic_lengths(n_clips: int, lo: int = 150, hi: int = 600, seed: int = 4321) -> List[int]:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, (n_clips,), generator=g).tolist()


def synthetic_mels(batch: int = 32, frames: int = 256, n_mels: int = 64, seed: int = 2024) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, n_mels, frames, generator=g) * 2.0 - 5.0


def synthetic_scaler(n_mels: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    return (np.linspace(-60.0, -20.0, n_mels).astype(np.float32),
            np.linspace(8.0, 15.0, n_mels).astype(np.float32))


def write_scaler_json(path, n_mels: int = 64) -> None:
    mean, std = synthetic_scaler(n_mels)
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"mean": mean.tolist(), "std": std.tolist(), "count_frames": 1}, f)


def randomize_batchnorm(model: torch.nn.Module, seed: int = 99) -> None:
    """Give every BatchNorm non-trivial affine parameters and running statistics (scaled-init variant)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                n = m.num_features
                m.weight.copy_(1.0 + 0.2 * torch.randn(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(1.0 + 0.3 * torch.rand(n, generator=g))
            elif isinstance(m, torch.nn.Conv2d) and m.bias is not None:
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))

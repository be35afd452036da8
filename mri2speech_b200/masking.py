"""In-memory articulator masking and the alpha-sweep driver of BASELINE.json configs[4].

The reference masks a clip by re-encoding the video (scripts/mask_rtmri_video.py:71-100: decode, multiply every
frame by a soft polygon mask, truncate to uint8, write mp4v) and then runs the inference CLI on the new file.
Here the mask is built once on the host per (preset, alpha) -- same presets (:31-50), same construction (:53-68:
fill the rounded polygon with alpha on a field of ones, Gaussian blur, clip to [alpha, 1]) -- and applied on the
device inside the uint8 ingest of the stem kernel (``m2s_acoustic_forward_u8``): masked = uint8(clip(frame * mask,
0, 255)), truncating like ``ndarray.astype(np.uint8)``.  The lossy mp4v round trip of the reference is skipped
(stated deviation).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class MaskPreset:
    """Polygon in ``base_size`` pixel coordinates (x, y)."""
    name: str
    points: Tuple[Tuple[float, float], ...]
    base_size: Tuple[float, float] = (256.0, 256.0)

    def scaled(self, target_size: Tuple[int, int]) -> np.ndarray:
        """Polygon scaled to ``target_size`` = (width, height)."""
        width, height = target_size
        pts = np.asarray(self.points, dtype=np.float32).copy()
        pts[:, 0] *= width / self.base_size[0]
        pts[:, 1] *= height / self.base_size[1]
        return pts


LIP_MASK = MaskPreset("lip", ((8.0, 84.0), (43.0, 84.0), (45.0, 156.0), (8.0, 156.0)))
TONGUE_MASK = MaskPreset("tongue", ((36.1, 102.7), (63.4, 90.9), (122.7, 111.5), (133.4, 172.2), (47.6, 155.0)))
PRESETS: Dict[str, MaskPreset] = {"lip": LIP_MASK, "tongue": TONGUE_MASK}


def build_mask(shape: Tuple[int, int], polygon: np.ndarray, alpha: float, blur_kernel: int = 11) -> np.ndarray:
    """(H, W) float32 soft mask: ``alpha`` inside the polygon, 1 outside, Gaussian-blurred edge, clipped to
    [alpha, 1].  Host-side, once per (preset, alpha); OpenCV does the rasterisation as in the reference."""
    import cv2
    h, w = shape
    mask = np.ones((h, w), dtype=np.float32)
    cv2.fillConvexPoly(mask, np.round(polygon).astype(np.int32), float(alpha))
    if blur_kernel > 1:
        k = blur_kernel + 1 if blur_kernel % 2 == 0 else blur_kernel
        mask = cv2.GaussianBlur(mask, (k, k), sigmaX=0.0)
    return np.clip(mask, alpha, 1.0).astype(np.float32)


def preset_mask(mask_type: str, alpha: float, shape: Tuple[int, int] = (256, 256), blur_kernel: int = 11) -> np.ndarray:
    if mask_type not in PRESETS:
        raise KeyError(f"unknown mask preset {mask_type!r}; choose from {sorted(PRESETS)}")
    h, w = shape
    return build_mask((h, w), PRESETS[mask_type].scaled((w, h)), alpha, blur_kernel)


def sweep_alphas(steps: int = 11) -> List[float]:
    """alpha 0.0 .. 1.0 in ``steps`` steps (configs[4]: 11)."""
    return [round(i / (steps - 1), 10) for i in range(steps)]


@torch.no_grad()
def masking_sweep(pipe, clips_u8: Sequence[torch.Tensor], mask_types: Iterable[str] = ("lip", "tongue"),
                  alphas: Optional[Sequence[float]] = None, blur_kernel: int = 11, max_batch_frames: int = 4096,
                  dedup_identity: bool = True):
    """Batched re-inference of ``clips_u8`` (list of (T,H,W) uint8 tensors) under every (mask, alpha).

    ``pipe`` is a ``pipeline.MriToSpeech``.  Returns {(mask_type, alpha): [per-clip result dicts]}.  alpha = 1 is
    the identity mask for every preset (clip(blur(ones)) == ones), so with ``dedup_identity`` it is run once."""
    alphas = list(sweep_alphas() if alphas is None else alphas)
    H, W = clips_u8[0].shape[-2:]
    out = {}
    identity = None
    for mt in mask_types:
        for a in alphas:
            if dedup_identity and a >= 1.0:
                if identity is None:
                    identity = pipe.infer(clips_u8, max_batch_frames=max_batch_frames)
                out[(mt, a)] = identity
                continue
            m = torch.from_numpy(preset_mask(mt, a, (H, W), blur_kernel))
            out[(mt, a)] = pipe.infer(clips_u8, max_batch_frames=max_batch_frames, mask=m)
    return out

"""CPU oracle: frame ingest / normalisation and articulator masking.

TEST INFRASTRUCTURE ONLY.  numpy restatement of
  * scripts/run_mri_video_inference.py:34-53  ``_preprocess_frame`` on an already-gray, already-256x256 uint8
    frame: float32 cast, z-score (or mean removal when std == 0), then min-max to [0,1] (zeros when constant);
  * scripts/mask_rtmri_video.py:53-68  ``build_mask`` (OpenCV fill + Gaussian blur + clip);
  * scripts/mask_rtmri_video.py:96-98  masked = uint8(clip(float32(frame) * mask, 0, 255)).
Pinned: tests/golden/ingest_ref.npz holds outputs of the reference's own functions (oracle/gen_golden.py executes
``_preprocess_frame`` extracted from the reference file and imports scripts/mask_rtmri_video.py in the build
container).
"""
from __future__ import annotations

import numpy as np


def preprocess_frame(gray_u8: np.ndarray) -> np.ndarray:
    """(H,W) uint8 -> (H,W) float32 in [0,1], following the reference's operation order exactly."""
    gray = gray_u8.astype(np.float32)
    mean = gray.mean()
    std = gray.std()
    gray = (gray - mean) / std if std > 0 else gray - mean
    lo, hi = gray.min(), gray.max()
    if hi > lo:
        return (gray - lo) / (hi - lo)
    return np.zeros_like(gray)


def preprocess_clip(clip_u8: np.ndarray) -> np.ndarray:
    return np.stack([preprocess_frame(f) for f in clip_u8]).astype(np.float32)


def apply_mask(clip_u8: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """(T,H,W) uint8, (H,W) float32 -> (T,H,W) uint8 (truncating cast, as ndarray.astype(np.uint8))."""
    return (clip_u8.astype(np.float32) * mask[None]).clip(0.0, 255.0).astype(np.uint8)


def build_mask(shape, polygon: np.ndarray, alpha: float, blur_kernel: int) -> np.ndarray:
    import cv2
    h, w = shape
    mask = np.ones((h, w), dtype=np.float32)
    cv2.fillConvexPoly(mask, np.round(polygon).astype(np.int32), alpha)
    if blur_kernel > 1:
        if blur_kernel % 2 == 0:
            blur_kernel += 1
        mask = cv2.GaussianBlur(mask, (blur_kernel, blur_kernel), sigmaX=0.0)
    return np.clip(mask, alpha, 1.0)

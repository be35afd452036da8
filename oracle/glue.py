"""CPU oracle: mel glue between the acoustic model and the vocoder.

TEST INFRASTRUCTURE ONLY.  Follows scripts/run_mri_video_inference.py:160-163
(denormalize_mel), :232-233 (dB -> power -> log with clamp 1e-5), :238-239
(transpose to (1, n_mels, T)); same math in scripts/export_predicted_mels.py:92-98.
"""
from __future__ import annotations

import numpy as np
import torch


def denormalize_mel(pred_norm: torch.Tensor, mean, std) -> torch.Tensor:
    mean_t = torch.as_tensor(np.asarray(mean, dtype=np.float32))
    std_t = torch.as_tensor(np.asarray(std, dtype=np.float32))
    return pred_norm * std_t + mean_t


def db_to_log_power(mel_db: torch.Tensor) -> torch.Tensor:
    mel_power = torch.pow(10.0, mel_db / 10.0)
    return torch.log(torch.clamp(mel_power, min=1e-5))


def mel_glue(pred_norm: torch.Tensor, mean, std):
    """(..., T, n_mels) normalised mel -> (mel_db (...,T,n), mel_log (...,T,n), vocoder input (..., n, T))."""
    mel_db = denormalize_mel(pred_norm, mean, std)
    mel_log = db_to_log_power(mel_db)
    return mel_db, mel_log, mel_log.transpose(-1, -2).contiguous()

"""End-to-end rtMRI -> mel -> waveform runner: ragged batching, utterance sharding, final gather.

Host-side mirror of the glue in the reference CLI (scripts/run_mri_video_inference.py:77-86 load_scaler,
:160-163 denormalize_mel, :227-243 the model chain).  The arithmetic (de-normalise, dB -> log-power,
transpose) runs in ``m2s_mel_glue``; nothing here computes on the CPU.

Multi-GPU: utterances are independent, so ranks take whole clips (longest-processing-time
assignment on frame counts) and there is no collective on the hot path; ``gather_waveforms`` is the only
exchange (lengths all_gather + one padded gather), used by bench.py and the batch CLIs.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def load_scaler(stats_path) -> Tuple[np.ndarray, np.ndarray]:
    """scaler.json {"mean": [n_mels], "std": [n_mels], ...} -> float32 arrays (same errors as the reference)."""
    with open(stats_path, "r", encoding="utf-8") as f:
        stats = json.load(f)
    if "mean" not in stats or "std" not in stats:
        raise KeyError("Scaler JSON must contain 'mean' and 'std' lists")
    mean = np.asarray(stats["mean"], dtype=np.float32)
    std = np.asarray(stats["std"], dtype=np.float32)
    if mean.ndim != 1 or std.ndim != 1:
        raise ValueError("Scaler mean/std must be 1-D lists")
    return mean, std


def mel_glue(pred_norm: torch.Tensor, mean: torch.Tensor, std: torch.Tensor,
             lengths: Optional[torch.Tensor] = None, want_db: bool = True, want_log: bool = True,
             want_voc: bool = True):
    """(B,T,n) normalised mel (cuda) -> (mel_db (B,T,n), mel_log (B,T,n), vocoder input (B,n,T)).

    ``mel_log`` (B,T,n) is already the channels-last operand of the vocoder's conv_pre (rows past ``lengths`` zeroed):
    ``Generator.forward(mel_log, channels_last=True)`` consumes it directly, in which case ``want_voc=False`` skips the
    transposed (B,n,T) copy that only the reference-shaped call needs."""
    _lib.require_device(pred_norm)
    squeeze = pred_norm.dim() == 2
    if squeeze:
        pred_norm = pred_norm.unsqueeze(0)
    pred_norm = pred_norm.contiguous().float()
    B, T, M = pred_norm.shape
    dev = pred_norm.device
    mean = mean.to(dev, torch.float32).contiguous()
    std = std.to(dev, torch.float32).contiguous()
    if mean.numel() != M or std.numel() != M:
        raise ValueError("Scaler mean/std length does not match n_mels")
    mel_db = torch.empty_like(pred_norm) if want_db else None
    mel_log = torch.empty_like(pred_norm) if want_log else None
    voc_in = torch.empty(B, M, T, device=dev, dtype=torch.float32) if want_voc else None
    if lengths is not None:
        lengths = lengths.to(dev, torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().m2s_mel_glue(pred_norm.data_ptr(), mean.data_ptr(), std.data_ptr(), B, T, M,
                                           _lib.ptr(lengths), _lib.ptr(mel_db), _lib.ptr(mel_log),
                                           _lib.ptr(voc_in), _lib.current_stream()))
    if squeeze:
        return (None if mel_db is None else mel_db[0], None if mel_log is None else mel_log[0], voc_in)
    return mel_db, mel_log, voc_in


class MriToSpeech:
    """frames -> normalised mel -> (mel_db, mel_log) -> waveform for a ragged batch of clips."""

    def __init__(self, acoustic_model, generator, mean, std, device: Optional[torch.device] = None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.acoustic = acoustic_model.to(self.device).eval()
        self.generator = generator.to(self.device).eval()
        self.mean = torch.as_tensor(np.asarray(mean, np.float32)).to(self.device)
        self.std = torch.as_tensor(np.asarray(std, np.float32)).to(self.device)
        self.hop = generator.hop

    def reserve(self, max_batch_frames: int = 4096, max_clip_frames: int = 600, height: int = 256, width: int = 256):
        """Pre-size both models' workspaces for the largest padded micro-batch ``infer`` can build, so that no
        allocation (and no implicit device synchronisation) happens once inference has started."""
        nb = max(1, max_batch_frames // max(max_clip_frames, 1))
        frames = max(max_clip_frames, max_batch_frames // nb)
        self.acoustic.reserve(nb, frames, height, width)
        self.generator.reserve(nb, frames)
        # micro-batches of many short clips have the same padded frame count but a larger batch dimension
        self.acoustic.reserve(max_batch_frames // 64 + 1, 64, height, width)
        self.generator.reserve(max_batch_frames // 64 + 1, 64)

    @torch.no_grad()
    def infer_padded(self, frames: torch.Tensor, lengths: Optional[torch.Tensor],
                     mask: Optional[torch.Tensor] = None):
        """frames (B,T,H,W) cuda -- float32 in [0,1], or raw uint8 (min-max normalised on the device, after the
        optional (H,W) articulator ``mask``); lengths int32[B] (cpu or cuda) or None -> dict of padded tensors."""
        pred = self.acoustic(frames, lengths=lengths, mask=mask) if mask is not None else \
            self.acoustic(frames, lengths=lengths)
        # the glue kernel writes conv_pre's operand (mel_log, channels-last) directly: no (B,n,T) tensor, no layout pass
        mel_db, mel_log, _ = mel_glue(pred, self.mean, self.std, lengths, want_voc=False)
        wav = self.generator(mel_log, lengths=lengths, channels_last=True)
        return {"mel_norm": pred, "mel_db": mel_db, "mel_log": mel_log, "audio": wav}

    @torch.no_grad()
    def infer(self, clips: Sequence[torch.Tensor], max_batch_frames: int = 4096,
              mask: Optional[torch.Tensor] = None) -> List[Dict[str, torch.Tensor]]:
        """clips: list of (T_i,H,W) tensors (host or device), all float32 in [0,1] or all raw uint8 (4x fewer
        bytes over PCIe / HBM; normalised on the device).  Returns one dict per clip, in order.

        Clips are sorted by length and packed into micro-batches of at most ``max_batch_frames`` padded
        frames so that padding waste stays small; every clip equals its own B=1 run (ragged parity)."""
        order = sorted(range(len(clips)), key=lambda i: -int(clips[i].shape[0]))
        results: List[Optional[Dict[str, torch.Tensor]]] = [None] * len(clips)
        i = 0
        while i < len(order):
            tmax = int(clips[order[i]].shape[0])
            nb = max(1, min(len(order) - i, max_batch_frames // max(tmax, 1)))
            idx = order[i:i + nb]
            lens = [int(clips[j].shape[0]) for j in idx]
            H, W = clips[idx[0]].shape[-2:]
            dtype = torch.uint8 if clips[idx[0]].dtype == torch.uint8 else torch.float32
            batch = torch.zeros(nb, tmax, H, W, device=self.device, dtype=dtype)
            for b, j in enumerate(idx):
                batch[b, :lens[b]].copy_(clips[j], non_blocking=True)
            out = self.infer_padded(batch, torch.tensor(lens, dtype=torch.int32), mask=mask)
            for b, j in enumerate(idx):
                n = lens[b]
                results[j] = {"mel_norm": out["mel_norm"][b, :n], "mel_db": out["mel_db"][b, :n],
                              "mel_log": out["mel_log"][b, :n], "audio": out["audio"][b, 0, :n * self.hop]}
            i += nb
        return results  # type: ignore[return-value]


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time assignment of clip indices to ranks (cost is linear in frames)."""
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(lengths)), key=lambda k: (-int(lengths[k]), k)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return shards


def gather_waveforms(local: Sequence[torch.Tensor], local_ids: Sequence[int], dst: int = 0):
    """Final gather of ragged waveforms to rank ``dst`` (the only collective of the path).

    Exchanges (id, length) pairs with all_gather, then one gather of max-padded buffers.  Works on the
    NCCL backend (device tensors) and on gloo (CPU tensors, used by the world_size-2 tests).
    Returns {clip_id: waveform} on ``dst`` and None elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = local[0].device if len(local) else torch.device("cpu")
    if dist.get_backend() == "nccl" and dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    # (id, length) pairs: built on the host, one transfer, one all_gather of a max-sized table
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts_host = torch.stack(counts).cpu().view(-1).tolist()
    max_n = max(max(counts_host), 1)
    meta_host = torch.full((max_n, 2), -1, dtype=torch.int64)
    for k, (cid, w) in enumerate(zip(local_ids, local)):
        meta_host[k, 0], meta_host[k, 1] = int(cid), int(w.numel())
    meta = meta_host.to(dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas_host = torch.stack(metas).cpu()
    max_len = max(int(metas_host[:, :, 1].max().item()), 1)
    buf = torch.zeros(max_n, max_len, dtype=torch.float32, device=dev)
    for k, w in enumerate(local):
        buf[k, : w.numel()] = w.reshape(-1).to(dev)
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        table = metas_host[r].tolist()
        for k in range(counts_host[r]):
            cid, n = table[k]
            out[cid] = bufs[r][k, :n]
    return out

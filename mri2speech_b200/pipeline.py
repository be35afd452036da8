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

    # Clips whose BiLSTM runs as ONE launch: the recurrence is a chain of dependent steps (6.6 us per step at B = 10,
    # ~2 us more per further 8 utterances), so one launch over 64 clips costs a third of seven launches over ~10 each.
    RNN_GROUP_CLIPS = 64

    def reserve(self, max_batch_frames: int = 4096, max_clip_frames: int = 600, height: int = 256, width: int = 256):
        """Pre-size both models' workspaces for the largest micro-batch ``infer`` can build, so that no
        allocation (and no implicit device synchronisation) happens once inference has started."""
        nb = max(1, max_batch_frames // max(max_clip_frames, 1))
        frames = max(max_clip_frames, max_batch_frames // nb)
        self.acoustic.reserve(nb, frames, height, width)
        self.generator.reserve(nb, frames)
        # micro-batches of many short clips have the same padded frame count but a larger batch dimension
        self.acoustic.reserve(max_batch_frames // 64 + 1, 64, height, width)
        self.generator.reserve(max_batch_frames // 64 + 1, 64)
        # the recurrence runs once per group of up to RNN_GROUP_CLIPS clips (see infer)
        self.acoustic.reserve(self.RNN_GROUP_CLIPS, max_clip_frames, height, width)

    @torch.no_grad()
    def infer_padded(self, frames: torch.Tensor, lengths: Optional[torch.Tensor],
                     mask: Optional[torch.Tensor] = None):
        """frames (B,T,H,W) cuda -- float32 in [0,1], or raw uint8 (min-max normalised on the device, after the
        optional (H,W) articulator ``mask``); lengths int32[B] (cpu or cuda) or None -> dict of padded tensors."""
        pred = self.acoustic(frames, lengths=lengths, mask=mask) if mask is not None else \
            self.acoustic(frames, lengths=lengths)
        return self._vocode(pred, lengths)

    def _vocode(self, pred: torch.Tensor, lengths: Optional[torch.Tensor]):
        # the glue kernel writes conv_pre's operand (mel_log, channels-last) directly: no (B,n,T) tensor, no layout pass
        mel_db, mel_log, _ = mel_glue(pred, self.mean, self.std, lengths, want_voc=False)
        wav = self.generator(mel_log, lengths=lengths, channels_last=True)
        return {"mel_norm": pred, "mel_db": mel_db, "mel_log": mel_log, "audio": wav}

    @staticmethod
    def plan_micro_batches(lengths: Sequence[int], max_batch_frames: int, ramp: bool = False) -> List[List[int]]:
        """Clip indices sorted by length (longest first) and cut into micro-batches of at most ``max_batch_frames``
        PADDED frames (batch x longest clip: what the recurrence and the vocoder run on).  ``ramp``: the first two
        micro-batches get a quarter / half of the budget -- for clips that still have to cross PCIe, so that the GPU
        starts after a quarter of a micro-batch's copy instead of a whole one (nothing hides the first copy)."""
        order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
        plan, i = [], 0
        while i < len(order):
            tmax = int(lengths[order[i]])
            budget = max_batch_frames
            if ramp and len(plan) < 2:
                budget = max(max_batch_frames // (4 >> len(plan)), tmax)
            nb = max(1, min(len(order) - i, budget // max(tmax, 1)))
            plan.append(order[i:i + nb])
            i += nb
        return plan

    @torch.no_grad()
    def infer(self, clips: Sequence[torch.Tensor], max_batch_frames: int = 4096,
              mask: Optional[torch.Tensor] = None, audio_to_host: bool = False) -> List[Dict[str, torch.Tensor]]:
        """clips: list of (T_i,H,W) tensors (host or device), all float32 in [0,1] or all raw uint8 (4x fewer
        bytes over PCIe / HBM; normalised on the device).  Returns one dict per clip, in order.

        Clips are sorted by length and cut into micro-batches of at most ``max_batch_frames`` padded frames.  A
        micro-batch is fed PACKED -- its clips' frames one after the other in a staging buffer, no zero-filled
        (B, Tmax, H, W) tensor -- through ``m2s_acoustic_forward_packed``; the staging buffers are double-buffered
        and filled on a copy stream, so the host-to-device copies of micro-batch k+1 overlap the kernels of k.
        ``audio_to_host``: the waveforms are returned in pinned host memory (one device-to-host copy per micro-batch
        on a third stream, overlapping the next micro-batch; the call returns once they have landed).
        Every clip equals its own B=1 run (ragged parity)."""
        if not len(clips):
            return []
        lens_all = [int(c.shape[0]) for c in clips]
        plan = self.plan_micro_batches(lens_all, max_batch_frames, ramp=not clips[0].is_cuda)
        H, W = clips[0].shape[-2:]
        dtype = torch.uint8 if clips[0].dtype == torch.uint8 else torch.float32
        dev = self.device
        results: List[Optional[Dict[str, torch.Tensor]]] = [None] * len(clips)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            cap = max(sum(lens_all[j] for j in idx) for idx in plan)
            nbuf = 2 if len(plan) > 1 else 1
            key = (cap, H, W, dtype, nbuf)
            if getattr(self, "_stage_key", None) is None or self._stage_key[1:4] != key[1:4] or \
                    self._stage_key[0] < cap or self._stage_key[4] < nbuf:
                self._stage = [torch.empty(cap, H, W, device=dev, dtype=dtype) for _ in range(nbuf)]
                self._stage_key = key
                self._copy_stream = torch.cuda.Stream(dev)
                self._d2h_stream = torch.cuda.Stream(dev)
            stage, copy_s, d2h_s = self._stage, self._copy_stream, self._d2h_stream
            filled = [torch.cuda.Event() for _ in plan]
            consumed: List[Optional[torch.cuda.Event]] = [None] * len(plan)
            copy_s.wait_stream(main)     # the staging buffers may still be read by kernels of a previous call

            def fill(k):
                buf = stage[k % len(stage)]
                with torch.cuda.stream(copy_s):
                    if k >= len(stage) and consumed[k - len(stage)] is not None:
                        copy_s.wait_event(consumed[k - len(stage)])
                    off = 0
                    for j in plan[k]:
                        n = lens_all[j]
                        buf[off:off + n].copy_(clips[j], non_blocking=True)
                        off += n
                    filled[k].record(copy_s)

            host_done = []
            fill(0)
            # Micro-batches are encoded one after the other (bounded work buffers, H2D of k+1 under the kernels of k) into
            # slices of one feature tensor per GROUP of micro-batches; the BiLSTM + head then run once over the group's
            # clips, and the vocoder runs micro-batch by micro-batch again (D2H of k under the kernels of k+1).
            groups, cur, cur_clips = [], [], 0
            for k, idx in enumerate(plan):
                if cur and cur_clips + len(idx) > self.RNN_GROUP_CLIPS:
                    groups.append(cur)
                    cur, cur_clips = [], 0
                cur.append(k)
                cur_clips += len(idx)
            if cur:
                groups.append(cur)
            feat_dim = self.acoustic.cnn.out_channels
            next_fill = 1
            for grp in groups:
                g_clips = [j for k in grp for j in plan[k]]
                g_lens = [lens_all[j] for j in g_clips]
                g_tmax = max(g_lens)
                feats = torch.empty(len(g_clips), g_tmax, feat_dim, device=dev, dtype=torch.float32)
                pos = 0
                for k in grp:
                    idx = plan[k]
                    if next_fill < len(plan):
                        fill(next_fill)
                        next_fill += 1
                    lens = [lens_all[j] for j in idx]
                    total = sum(lens)
                    main.wait_event(filled[k])
                    self.acoustic.encode_packed(stage[k % len(stage)][:total], torch.tensor(lens, dtype=torch.int32), g_tmax,
                                                feats[pos:pos + len(idx)], mask=mask)
                    consumed[k] = torch.cuda.Event()
                    consumed[k].record(main)
                    pos += len(idx)
                g_lens_t = torch.tensor(g_lens, dtype=torch.int32)
                pred_all = self.acoustic.rnn_head(feats, g_lens_t)
                pos = 0
                for k in grp:
                    idx = plan[k]
                    lens = [lens_all[j] for j in idx]
                    lens_t = torch.tensor(lens, dtype=torch.int32)
                    pred = pred_all[pos:pos + len(idx), :max(lens)]
                    if not pred.is_contiguous():
                        pred = pred.contiguous()
                    pos += len(idx)
                    out = self._vocode(pred, lens_t)
                    audio = out["audio"]
                    if audio_to_host:
                        ready = torch.cuda.Event()
                        ready.record(main)
                        host = torch.empty(audio.shape, dtype=audio.dtype, pin_memory=True)
                        with torch.cuda.stream(d2h_s):
                            d2h_s.wait_event(ready)
                            host.copy_(audio, non_blocking=True)
                            audio.record_stream(d2h_s)
                            ev = torch.cuda.Event()
                            ev.record(d2h_s)
                        host_done.append(ev)
                        audio = host
                    for b, j in enumerate(idx):
                        n = lens[b]
                        results[j] = {"mel_norm": out["mel_norm"][b, :n], "mel_db": out["mel_db"][b, :n],
                                      "mel_log": out["mel_log"][b, :n], "audio": audio[b, 0, :n * self.hop]}
            for ev in host_done:
                ev.synchronize()
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


def gather_planned(local_flat: torch.Tensor, samples_per_rank: Sequence[int], dst: int = 0, cache: Optional[dict] = None):
    """Final gather when the plan is known to every rank (the batch runner shards a clip list whose lengths all ranks
    hold): rank r contributes ``samples_per_rank[r]`` samples, its clips' waveforms concatenated in shard order.  ONE
    gather of max-padded flat buffers -- no metadata exchange, no host round trip, nothing blocks the stream (LPT keeps
    the per-rank totals within one clip of each other, so the padding is negligible).  ``cache``: dict reused across
    calls for the send / receive buffers.  Returns the per-rank flat waveforms (views) on ``dst``, None elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if len(samples_per_rank) != world:
        raise ValueError(f"plan has {len(samples_per_rank)} ranks, world size is {world}")
    if int(local_flat.numel()) != int(samples_per_rank[rank]):
        raise ValueError(f"rank {rank} holds {local_flat.numel()} samples, the plan says {samples_per_rank[rank]}")
    max_n = max(max(int(n) for n in samples_per_rank), 1)
    cache = cache if cache is not None else {}
    dev = local_flat.device
    send = cache.get(("send", max_n, dev))
    if send is None:
        send = cache[("send", max_n, dev)] = torch.zeros(max_n, dtype=torch.float32, device=dev)
    send[: local_flat.numel()].copy_(local_flat.reshape(-1))
    recv = None
    if rank == dst:
        recv = cache.get(("recv", max_n, dev))
        if recv is None:
            recv = cache[("recv", max_n, dev)] = [torch.empty(max_n, dtype=torch.float32, device=dev) for _ in range(world)]
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    return [recv[r][: int(samples_per_rank[r])] for r in range(world)]

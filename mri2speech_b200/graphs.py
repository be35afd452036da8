"""CUDA-graph replay of fixed-shape forwards.

Small batches are launch-bound: the reference CLI's own case (one 150-frame clip, BASELINE.json configs[0]) is 61 vocoder
+ 111 encoder launches of 5-40 us each.  ``m2s_*_forward`` allocates nothing and never synchronises, so a whole forward
captures into one CUDA graph; replaying it removes the per-launch host cost (tensor-map encoding, driver calls) and the
gaps between kernels.  Shapes, the workspace and the input / output buffers are frozen at capture time; ``__call__``
copies the new input into the static buffer, replays, and returns the static output (clone it to keep it).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib


class GraphedForward:
    """``fn(static_input) -> tensor`` captured once for one input shape / dtype."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor, warmup: int = 2):
        _lib.require_device(example)
        self._fn = fn
        self.static_in = example.clone()
        stream = torch.cuda.Stream(device=example.device)
        stream.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(warmup):            # plans, workspaces and lazy kernel attributes are set up eagerly
                self._fn(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(stream)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._fn(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"graph captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, got "
                             f"{tuple(x.shape)} {x.dtype}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


def graph_generator(generator, example_mel: torch.Tensor) -> GraphedForward:
    """Graph of ``Generator.forward`` for mels shaped like ``example_mel`` (B, num_mels, T)."""
    return GraphedForward(lambda m: generator(m), example_mel)


def graph_acoustic(model, example_frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> GraphedForward:
    """Graph of ``OTNLikeCNNBiLSTM.forward`` for full-length batches shaped like ``example_frames`` (B,T,H,W), float32
    or uint8 (ragged batches -- ``lengths=`` -- upload a host table per call and are not capturable)."""
    if mask is not None:
        mask = mask.to(example_frames.device, torch.float32).contiguous()
        return GraphedForward(lambda f: model(f, mask=mask), example_frames)
    return GraphedForward(lambda f: model(f), example_frames)

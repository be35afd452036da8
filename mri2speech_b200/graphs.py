"""CUDA-graph replay of fixed-shape forwards.

Small batches are launch-bound: the reference CLI's own case (one 150-frame clip, BASELINE.json configs[0]) is 61 vocoder
+ 111 encoder launches of 5-40 us each.  ``m2s_*_forward`` allocates nothing and never synchronises (ragged batches included:
the frame map is built on the device), so a whole forward captures into one CUDA graph; replaying it removes the per-launch host cost (tensor-map encoding, driver calls) and the
gaps between kernels.  Shapes, the workspace and the input / output buffers are frozen at capture time; ``__call__``
copies the new input into the static buffer, replays, and returns the static output (clone it to keep it).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib


class GraphedForward:
    """``fn(static_input) -> tensor`` captured once for one input shape / dtype.

    A captured graph holds RAW device pointers: the owning module's workspace and the libm2s handle's packed weights.
    ``owner`` (the Generator / acoustic module behind ``fn``) is therefore watched: while the graph is alive the module
    refuses to re-allocate its workspace (``_graph_pins``), and if its plan changes anyway -- a parameter update,
    ``remove_weight_norm`` or ``.to()`` rebuilds the handle -- the next call re-captures before replaying, so a replay
    never touches freed memory."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor, warmup: int = 2, owner=None):
        _lib.require_device(example)
        self._fn = fn
        self._owner = owner
        self._warmup = warmup
        self.static_in = example.clone()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.captures = 0
        self._capture()

    def _token(self):
        o = self._owner
        if o is None:
            return None
        ws = getattr(o, "_workspace", None)
        return (getattr(o, "_handle", None), None if ws is None else ws.data_ptr(), o._current_key())

    def _capture(self):
        o = self._owner
        if o is not None and self.graph is not None:
            o._graph_pins -= 1           # the old graph is dropped below: the module may size its workspace again
        self.graph = None
        dev = self.static_in.device
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(self._warmup):      # plans, workspaces and lazy kernel attributes are set up eagerly
                self._fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            self.static_out = self._fn(self.static_in)
        self.graph = graph
        self.captures += 1
        if o is not None:
            o._graph_pins += 1
        self._pinned = self._token()

    def release(self):
        """Drop the graph and let the owner manage its workspace freely again."""
        if self.graph is not None and self._owner is not None:
            self._owner._graph_pins -= 1
        self.graph = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.graph is None:
            raise _lib.M2SError("this graph has been released")
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"graph captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, got "
                             f"{tuple(x.shape)} {x.dtype}")
        if self._token() != self._pinned:
            self._capture()                     # the plan behind the captured pointers changed: never replay it
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


def graph_generator(generator, example_mel: torch.Tensor) -> GraphedForward:
    """Graph of ``Generator.forward`` for mels shaped like ``example_mel`` (B, num_mels, T)."""
    return GraphedForward(lambda m: generator(m), example_mel, owner=generator)


def graph_acoustic(model, example_frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> GraphedForward:
    """Graph of ``OTNLikeCNNBiLSTM.forward`` for full-length batches shaped like ``example_frames`` (B,T,H,W), float32
    or uint8.  (Ragged forwards no longer synchronise or upload a host table -- the frame map is built on the device --
    but ``lengths`` are launch parameters, so a graph is specific to one set of lengths; this helper captures the
    full-length case.)"""
    if mask is not None:
        mask = mask.to(example_frames.device, torch.float32).contiguous()
        return GraphedForward(lambda f: model(f, mask=mask), example_frames, owner=model)
    return GraphedForward(lambda f: model(f), example_frames, owner=model)

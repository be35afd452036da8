"""Drop-in HiFi-GAN ``Generator`` whose forward runs on libm2s (sm_100a).

Mirrors the reference module API (models.py:11-55 ResBlock1, :88-140 Generator;
utils.py:22-35 init_weights / get_padding): same constructor argument ``h``, same
attribute names (``conv_pre``, ``ups``, ``resblocks``, ``conv_post``,
``num_kernels``, ``num_upsamples``, ``h``), same ``state_dict`` keys
(``weight_g``/``weight_v`` or plain ``weight`` once weight-norm is removed) and the
same seeded default initialisation, so a reference checkpoint loads strictly.

The nn.Modules here only HOLD parameters.  ``Generator.forward`` hands the
state_dict to ``m2s_generator_create`` once (weight-norm folding, polyphase
re-layout and TMA-friendly packing happen in C++) and then calls
``m2s_generator_forward``; no arithmetic is done in Python and there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn
from torch.nn import Conv1d, ConvTranspose1d
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _lib

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    # reference utils.py:34-35 -- twice the stock HiFi-GAN value; with the keep-first-L trim in
    # ResBlock1.forward it makes every ResBlock conv causal.  Kept for attribute parity only.
    return int(kernel_size * dilation - dilation)


def init_weights(m, mean=0.0, std=0.01):
    # reference utils.py:22-25 (a no-op on weight-normed layers, but it consumes RNG draws, so it is
    # replayed here to keep seeded initialisation identical to the reference).
    if m.__class__.__name__.find("Conv") != -1:
        m.weight.data.normal_(mean, std)


def _wn(module):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return weight_norm(module)


class ResBlock1(nn.Module):
    """Parameter holder for one multi-receptive-field branch (3 dilated + 3 plain causal convs)."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = nn.ModuleList([
            _wn(Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])
        self.convs1.apply(init_weights)
        self.convs2 = nn.ModuleList([
            _wn(Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1)))
            for _ in dilation])
        self.convs2.apply(init_weights)

    def forward(self, x):
        raise _lib.M2SError("ResBlock1 is fused into Generator.forward on the sm_100a path; call the Generator")

    def remove_weight_norm(self):
        for layer in list(self.convs1) + list(self.convs2):
            remove_weight_norm(layer)


class ResBlock2(nn.Module):
    """Parameter holder for the light multi-receptive-field branch of configs with ``"resblock": "2"`` (reference
    models.py:58-85): two causal dilated convs, each x = x + c(lrelu(x))."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = nn.ModuleList([
            _wn(Conv1d(channels, channels, kernel_size, 1, dilation=dilation[m], padding=get_padding(kernel_size, dilation[m])))
            for m in range(2)])
        self.convs.apply(init_weights)

    def forward(self, x):
        raise _lib.M2SError("ResBlock2 is fused into Generator.forward on the sm_100a path; call the Generator")

    def remove_weight_norm(self):
        for layer in self.convs:
            remove_weight_norm(layer)


class Generator(nn.Module):
    def __init__(self, h, precision: Optional[str] = None):
        super().__init__()
        self.h = h
        self.resblock_type = 1 if str(_cfg(h, "resblock")) == "1" else 2   # models.py:95
        rates = list(_cfg(h, "upsample_rates"))
        ksizes = list(_cfg(h, "upsample_kernel_sizes"))
        rb_k = list(_cfg(h, "resblock_kernel_sizes"))
        rb_d = list(_cfg(h, "resblock_dilation_sizes"))
        c0 = _cfg(h, "upsample_initial_channel")
        self.num_kernels = len(rb_k)
        self.num_upsamples = len(rates)
        self.conv_pre = Conv1d(_cfg(h, "num_mels"), c0, 7, 1, padding=0)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(rates, ksizes)):
            self.ups.append(_wn(ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(rb_k, rb_d):
                self.resblocks.append(ResBlock1(h, ch, k, d) if self.resblock_type == 1 else ResBlock2(h, ch, k, d))
        self.conv_post = _wn(Conv1d(ch, 1, 7, 1, padding=0))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)
        self.precision = precision or _lib.DEFAULT_PRECISION
        self._handle: Optional[int] = None
        self._handle_key = None
        self._workspace: Optional[torch.Tensor] = None
        self._graph_pins = 0   # live CUDA graphs holding this module's workspace / handle pointers (graphs.py)
        self.hop = 1
        for u in rates:
            self.hop *= int(u)

    # -- libm2s plumbing -----------------------------------------------------------------
    def _state_key(self):
        key = [self.precision]
        for p in self.parameters():
            key.append((p.data_ptr(), p._version))
        return tuple(key)

    def _current_key(self):
        return self._state_key()

    def _ensure_workspace(self, need: int, dev) -> None:
        if self._workspace is not None and self._workspace.numel() >= need and self._workspace.device == dev:
            return
        if self._graph_pins > 0 and self._workspace is not None:
            raise _lib.M2SError("this Generator's workspace is referenced by a captured CUDA graph and would have to "
                                "grow: call reserve() for the largest shape before capturing, or release the graph")
        self._workspace = None
        self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)

    def _config(self) -> _lib.GeneratorConfig:
        h = self.h
        cfg = _lib.GeneratorConfig()
        cfg.num_mels = int(_cfg(h, "num_mels"))
        cfg.upsample_initial_channel = int(_cfg(h, "upsample_initial_channel"))
        cfg.num_upsamples = self.num_upsamples
        for i, (u, k) in enumerate(zip(_cfg(h, "upsample_rates"), _cfg(h, "upsample_kernel_sizes"))):
            cfg.upsample_rates[i] = int(u)
            cfg.upsample_kernel_sizes[i] = int(k)
        cfg.num_kernels = self.num_kernels
        for j, (k, ds) in enumerate(zip(_cfg(h, "resblock_kernel_sizes"), _cfg(h, "resblock_dilation_sizes"))):
            cfg.resblock_kernel_sizes[j] = int(k)
            need = 3 if self.resblock_type == 1 else 2
            if (len(ds) != 3) if self.resblock_type == 1 else (len(ds) < 2):
                raise _lib.M2SError(f"ResBlock{self.resblock_type} needs {need} dilations per kernel size")
            for m, d in enumerate(list(ds)[:need]):
                cfg.resblock_dilations[j][m] = int(d)
        cfg.precision = _lib.PRECISIONS[self.precision]
        cfg.resblock = self.resblock_type
        return cfg

    def refresh(self):
        """(Re)build the device-side plan from the current parameters."""
        self._release()
        arr, n, keep = _lib.state_dict_to_tensors(self.state_dict())
        cfg = self._config()
        handle = C.c_void_p()
        _lib.check(_lib.lib().m2s_generator_create(C.byref(cfg), arr, n, C.byref(handle)))
        del keep
        self._handle = handle.value
        self._handle_key = self._state_key()

    def _release(self):
        if getattr(self, "_handle", None):
            _lib.lib().m2s_generator_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def launches_per_forward(self) -> int:
        return int(_lib.lib().m2s_generator_launches(self._handle)) if self._handle else 0

    def reserve(self, batch: int, frames: int) -> None:
        """Size the workspace for (batch, frames) up front: a growing workspace is re-allocated (a synchronising
        cudaFree + cudaMalloc) the first time each larger ragged batch arrives."""
        dev = next(self.parameters()).device
        with torch.cuda.device(dev):
            if self._handle is None or self._handle_key != self._state_key():
                self.refresh()
            self._ensure_workspace(int(_lib.lib().m2s_generator_workspace_bytes(self._handle, batch, frames)), dev)

    # -- the drop-in call -----------------------------------------------------------------
    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                channels_last: bool = False) -> torch.Tensor:
        """(B, num_mels, T) [or (num_mels, T)] float32 cuda -> (B, 1, T * prod(upsample_rates)).

        ``lengths`` (int32 cuda, optional extension): valid mel frames per utterance of a
        zero-padded batch; every utterance then equals its own B=1 result.
        ``channels_last`` (extension): x is (B, T, num_mels) -- conv_pre's operand layout, e.g. the ``mel_log`` output
        of ``pipeline.mel_glue`` (rows past ``lengths`` must be zero) -- and is consumed without a layout pass.
        """
        _lib.require_device(x)
        if x.dim() == 2:
            x = x.unsqueeze(0)
        x = x.contiguous().float()
        if channels_last:
            B, T, M = x.shape
        else:
            B, M, T = x.shape
        with torch.cuda.device(x.device):
            if self._handle is None or self._handle_key != self._state_key():
                self.refresh()
            self._ensure_workspace(int(_lib.lib().m2s_generator_workspace_bytes(self._handle, B, T)), x.device)
            out = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=x.device)
            if lengths is not None:
                lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
            fwd = _lib.lib().m2s_generator_forward_btc if channels_last else _lib.lib().m2s_generator_forward
            _lib.check(fwd(
                self._handle, x.data_ptr(), B, T, _lib.ptr(lengths), out.data_ptr(),
                self._workspace.data_ptr(), self._workspace.numel(), _lib.current_stream()))
        return out

    def remove_weight_norm(self):
        # The reference's Generator.remove_weight_norm raises on the un-normed conv_pre (models.py:139);
        # callers use the best-effort per-module loop instead (run_mri_video_inference.py:105-115).
        for layer in self.ups:
            remove_weight_norm(layer)
        for block in self.resblocks:
            block.remove_weight_norm()
        remove_weight_norm(self.conv_post)


def _cfg(h, key):
    return h[key] if isinstance(h, dict) else getattr(h, key)

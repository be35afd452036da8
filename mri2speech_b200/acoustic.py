"""Drop-in ``build_acoustic_model`` / ``OTNLikeCNNBiLSTM`` whose forward runs on libm2s (sm_100a).

Mirrors the reference module API (mri2speech_code/mri_acoustic_model.py:15-156): the same constructor
signature, the attributes external code reaches into (``.cnn.backbone``, ``.cnn.out_channels``,
``.cnn.gap``, ``.rnn.lstm``, ``.rnn.dropout``, ``.head``, ``.n_mels`` -- scripts/mri_gradcam_formant.py:
155-165,221-225) and the same ``state_dict`` keys (timm ``tf_efficientnetv2_b2`` features_only naming under
``cnn.backbone.``, ``rnn.lstm.*``, ``head.*``) so reference checkpoints load with ``strict=False`` and
report (missing, unexpected) the same way.  ``MRIAcousticModel`` is an alias (the name BASELINE.json uses).

The reference builds its backbone with ``timm.create_model`` (mri_acoustic_model.py:28-34); timm is not a
dependency here: the modules below only HOLD parameters under timm's names, the arithmetic is in CUDA.
``forward`` hands the state_dict to ``m2s_acoustic_create`` once (BatchNorm folding, the 3->1 stem fold
and the weight packing happen in C++) and calls ``m2s_acoustic_forward``.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-3  # TF default used by every tf_* timm model

# timm arch_def of tf_efficientnetv2_b2 after the 1.1 width / 1.2 depth multipliers:
# (block, repeats, stride, expand, out_channels, squeeze-excite)
_STAGES = (
    ("cn", 2, 1, 1, 16, False),
    ("er", 3, 2, 4, 32, False),
    ("er", 3, 2, 4, 56, False),
    ("ir", 4, 2, 4, 104, True),
    ("ir", 6, 1, 6, 120, True),
    ("ir", 10, 2, 6, 208, True),
)
_STEM = 32


def _conv(cin, cout, k, stride=1, groups=1, bias=False):
    # padding is TF "same" in the reference (timm Conv2dSame); the value stored here is never used
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=bias)


def _bn(ch):
    return nn.BatchNorm2d(ch, eps=BN_EPS)


class _ConvBnAct(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = _conv(cin, cout, 3, stride)
        self.bn1 = _bn(cout)


class _EdgeResidual(nn.Module):
    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        mid = cin * expand
        self.conv_exp = _conv(cin, mid, 3, stride)
        self.bn1 = _bn(mid)
        self.conv_pwl = _conv(mid, cout, 1)
        self.bn2 = _bn(cout)


class _SqueezeExcite(nn.Module):
    def __init__(self, ch, rd):
        super().__init__()
        self.conv_reduce = _conv(ch, rd, 1, bias=True)
        self.conv_expand = _conv(rd, ch, 1, bias=True)


class _InvertedResidual(nn.Module):
    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        mid = cin * expand
        self.conv_pw = _conv(cin, mid, 1)
        self.bn1 = _bn(mid)
        self.conv_dw = _conv(mid, mid, 3, stride, groups=mid)
        self.bn2 = _bn(mid)
        self.se = _SqueezeExcite(mid, int(round(cin * 0.25)))
        self.conv_pwl = _conv(mid, cout, 1)
        self.bn3 = _bn(cout)


class _EfficientNetV2B2Features(nn.Module):
    """Parameter holder with timm's module tree for tf_efficientnetv2_b2 (features_only)."""

    def __init__(self):
        super().__init__()
        self.conv_stem = _conv(3, _STEM, 3, 2)
        self.bn1 = _bn(_STEM)
        stages = []
        cin = _STEM
        for kind, reps, stride, expand, cout, _se in _STAGES:
            blocks = []
            for b in range(reps):
                st = stride if b == 0 else 1
                if kind == "cn":
                    blocks.append(_ConvBnAct(cin, cout, st))
                elif kind == "er":
                    blocks.append(_EdgeResidual(cin, cout, st, expand))
                else:
                    blocks.append(_InvertedResidual(cin, cout, st, expand))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.num_features = cin
        self._init_weights()

    def _init_weights(self):
        # timm efficientnet_init_weights (goog init): conv ~ N(0, sqrt(2 / fan_out)), BN gamma=1 beta=0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        raise _lib.M2SError("the backbone is fused into OTNLikeCNNBiLSTM.forward on the sm_100a path")


class GlobalAvgPool(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise _lib.M2SError("GlobalAvgPool is fused into the encoder on the sm_100a path")


class EffNetV2B2Backbone(nn.Module):
    def __init__(self, pretrained: bool = False):
        super().__init__()
        if pretrained:
            raise _lib.M2SError("cnn_pretrained=True needs timm + network access; load a checkpoint instead")
        self.backbone = _EfficientNetV2B2Features()
        self.out_channels = self.backbone.num_features
        self.gap = GlobalAvgPool()


class BiLSTMSumMerge(nn.Module):
    def __init__(self, in_dim: int, hidden_size: int = 640, dropout: float = 0.0):
        super().__init__()
        self.lstm = nn.LSTM(input_size=in_dim, hidden_size=hidden_size, num_layers=1, batch_first=True,
                            bidirectional=True, dropout=0.0)
        self.dropout = nn.Dropout(dropout)


class OTNLikeCNNBiLSTM(nn.Module):
    """(B,T,1,H,W) | (B,T,H,W) float32 cuda -> (B,T,n_mels) normalised mel (inference only)."""

    def __init__(self, n_mels: int = 64, cnn_pretrained: bool = False, rnn_hidden: int = 640, dropout: float = 0.5,
                 use_checkpoint: bool = False, ckpt_segments: int = 2, use_reentrant: bool = False,
                 precision: Optional[str] = None):
        super().__init__()
        if rnn_hidden <= 0 or rnn_hidden % 32:
            # 640 (the reference's default, mri_acoustic_model.py:148) runs the tuned cluster recurrence; any other multiple
            # of 32 the generic cooperative kernel of csrc/lstm_sm100.cu
            raise _lib.M2SError(f"rnn_hidden={rnn_hidden}: the sm_100a BiLSTM recurrence needs a positive multiple of 32 "
                                "(refused here rather than at the first forward)")
        self.n_mels = n_mels
        self.use_checkpoint = use_checkpoint
        self.ckpt_segments = ckpt_segments
        self.use_reentrant = use_reentrant
        self.cnn = EffNetV2B2Backbone(pretrained=cnn_pretrained)
        self.rnn = BiLSTMSumMerge(in_dim=self.cnn.out_channels, hidden_size=rnn_hidden, dropout=dropout)
        self.head = nn.Linear(rnn_hidden, n_mels)
        self.rnn_hidden = rnn_hidden
        self.precision = precision or _lib.DEFAULT_PRECISION
        self._handle: Optional[int] = None
        self._handle_key = None
        self._workspace: Optional[torch.Tensor] = None
        self._graph_pins = 0   # live CUDA graphs holding this module's workspace / handle pointers (graphs.py)

    # -- libm2s plumbing -----------------------------------------------------------------
    def _state_key(self, hw):
        key = [self.precision, hw]
        for p in list(self.parameters()) + list(self.buffers()):
            key.append((p.data_ptr(), p._version))
        return tuple(key)

    def _current_key(self):
        return self._state_key(self._handle_key[1] if self._handle_key else (256, 256))

    def refresh(self, height: int = 256, width: int = 256):
        self._release()
        arr, n, keep = _lib.state_dict_to_tensors(self.state_dict())
        cfg = _lib.AcousticConfig()
        cfg.n_mels = self.n_mels
        cfg.rnn_hidden = self.rnn_hidden
        cfg.height, cfg.width = height, width
        cfg.precision = _lib.PRECISIONS[self.precision]
        handle = C.c_void_p()
        _lib.check(_lib.lib().m2s_acoustic_create(C.byref(cfg), arr, n, C.byref(handle)))
        del keep
        self._handle = handle.value
        self._handle_key = self._state_key((height, width))

    def _release(self):
        if getattr(self, "_handle", None):
            _lib.lib().m2s_acoustic_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _prepare(self, x: torch.Tensor, batch: int, frames: int, hw):
        if self.training:
            raise _lib.M2SError("the sm_100a path is inference-only: call .eval() first "
                                "(training / activation checkpointing are out of scope)")
        if self._handle is None or self._handle_key != self._state_key(hw):
            self.refresh(*hw)
        need = int(_lib.lib().m2s_acoustic_workspace_bytes(self._handle, batch, frames))
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != x.device:
            if self._graph_pins > 0 and self._workspace is not None:
                raise _lib.M2SError("this model's workspace is referenced by a captured CUDA graph and would have to "
                                    "grow: call reserve() for the largest shape before capturing, or release the graph")
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=x.device)

    def reserve(self, batch: int, frames: int, height: int = 256, width: int = 256) -> None:
        """Size the workspace for (batch, frames) up front (see Generator.reserve)."""
        dev = next(self.parameters()).device
        with torch.cuda.device(dev):
            self._prepare(torch.empty(0, device=dev), batch, frames, (height, width))

    def launches_per_forward(self) -> int:
        return int(_lib.lib().m2s_acoustic_launches(self._handle)) if self._handle else 0

    # -- the drop-in call -----------------------------------------------------------------
    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x float32: frames already normalised to [0,1] (the reference call, mri_acoustic_model.py:116-136).
        x uint8: raw gray frames; the per-frame min-max of run_mri_video_inference.py:34-53 and, with ``mask``
        (H,W) float32, the articulator masking of mask_rtmri_video.py:96-98 run fused into the stem load."""
        _lib.require_device(x)
        if x.dim() == 5:
            if x.size(2) != 1:
                raise ValueError(f"expected a single grayscale channel, got {tuple(x.shape)}")
            x = x[:, :, 0]
        if x.dim() != 4:
            raise ValueError(f"expected (B,T,1,H,W) or (B,T,H,W), got {tuple(x.shape)}")
        raw = x.dtype == torch.uint8
        if mask is not None and not raw:
            raise ValueError("mask applies to raw uint8 frames (it is applied before normalisation)")
        x = x.contiguous() if raw else x.contiguous().float()
        B, T, H, W = x.shape
        with torch.cuda.device(x.device):
            self._prepare(x, B, T, (H, W))
            out = torch.empty(B, T, self.n_mels, dtype=torch.float32, device=x.device)
            lens_dev = lens_host = None
            if lengths is not None:
                lens_host = lengths.detach().to("cpu", torch.int32).contiguous()
                lens_dev = lens_host.to(x.device)
            lens_host_ptr = lens_host.data_ptr() if lens_host is not None else None
            if raw:
                if mask is not None:
                    if tuple(mask.shape) != (H, W):
                        raise ValueError(f"Mask shape {tuple(mask.shape)} != frame shape {(H, W)}")
                    mask = mask.to(x.device, torch.float32).contiguous()
                _lib.check(_lib.lib().m2s_acoustic_forward_u8(
                    self._handle, x.data_ptr(), _lib.ptr(mask), B, T, _lib.ptr(lens_dev), lens_host_ptr,
                    out.data_ptr(), self._workspace.data_ptr(), self._workspace.numel(), _lib.current_stream()))
            else:
                _lib.check(_lib.lib().m2s_acoustic_forward(
                    self._handle, x.data_ptr(), B, T, _lib.ptr(lens_dev), lens_host_ptr, out.data_ptr(),
                    self._workspace.data_ptr(), self._workspace.numel(), _lib.current_stream()))
        return out

    def forward_packed(self, frames: torch.Tensor, lengths: torch.Tensor, max_frames: Optional[int] = None,
                       mask: Optional[torch.Tensor] = None, lengths_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Ragged batch without input padding (extension; the reference runs one clip per call): ``frames``
        (sum(lengths), H, W) cuda, float32 in [0,1] or raw uint8 -- clip after clip; ``lengths`` int32[B] on the HOST
        (``lengths_dev``: the same values already on the device, optional).  Returns (B, max_frames, n_mels) with
        rows past lengths[b] zero; every clip equals its own B=1 run."""
        _lib.require_device(frames)
        if frames.dim() != 3:
            raise ValueError(f"expected packed frames (sum T, H, W), got {tuple(frames.shape)}")
        raw = frames.dtype == torch.uint8
        if mask is not None and not raw:
            raise ValueError("mask applies to raw uint8 frames (it is applied before normalisation)")
        frames = frames.contiguous() if raw else frames.contiguous().float()
        lens_host = lengths.detach().to("cpu", torch.int32).contiguous()
        B = int(lens_host.numel())
        total, H, W = frames.shape
        if int(lens_host.sum()) != total:
            raise ValueError(f"sum(lengths)={int(lens_host.sum())} != packed frames {total}")
        T = int(max_frames) if max_frames is not None else int(lens_host.max())
        with torch.cuda.device(frames.device):
            self._prepare(frames, B, T, (H, W))
            out = torch.empty(B, T, self.n_mels, dtype=torch.float32, device=frames.device)
            lens_dev = lengths_dev if lengths_dev is not None else lens_host.to(frames.device, non_blocking=True)
            if mask is not None:
                if tuple(mask.shape) != (H, W):
                    raise ValueError(f"Mask shape {tuple(mask.shape)} != frame shape {(H, W)}")
                mask = mask.to(frames.device, torch.float32).contiguous()
            _lib.check(_lib.lib().m2s_acoustic_forward_packed(
                self._handle, frames.data_ptr(), int(raw), _lib.ptr(mask), B, T, lens_dev.data_ptr(),
                lens_host.data_ptr(), out.data_ptr(), self._workspace.data_ptr(), self._workspace.numel(),
                _lib.current_stream()))
        return out

    def encode_packed(self, frames: torch.Tensor, lengths: torch.Tensor, max_frames: int, out: torch.Tensor,
                      mask: Optional[torch.Tensor] = None, lengths_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Encoder half of ``forward_packed``: packed frames (sum(lengths), H, W) -> ``out`` (B, max_frames, 208) float32
        cuda, contiguous (rows past lengths[b] zero).  ``rnn_head`` is the other half: a batch runner encodes micro-batch
        after micro-batch into slices of one feature tensor and runs the recurrence once over all of its clips."""
        _lib.require_device(frames)
        raw = frames.dtype == torch.uint8
        if mask is not None and not raw:
            raise ValueError("mask applies to raw uint8 frames (it is applied before normalisation)")
        frames = frames.contiguous() if raw else frames.contiguous().float()
        lens_host = lengths.detach().to("cpu", torch.int32).contiguous()
        B = int(lens_host.numel())
        total, H, W = frames.shape
        if int(lens_host.sum()) != total:
            raise ValueError(f"sum(lengths)={int(lens_host.sum())} != packed frames {total}")
        if tuple(out.shape) != (B, int(max_frames), self.cnn.out_channels) or not out.is_contiguous() or \
                out.dtype != torch.float32 or out.device != frames.device:
            raise ValueError(f"out must be a contiguous float32 ({B}, {max_frames}, {self.cnn.out_channels}) tensor on {frames.device}")
        with torch.cuda.device(frames.device):
            self._prepare(frames, B, int(max_frames), (H, W))
            lens_dev = lengths_dev if lengths_dev is not None else lens_host.to(frames.device, non_blocking=True)
            if mask is not None:
                if tuple(mask.shape) != (H, W):
                    raise ValueError(f"Mask shape {tuple(mask.shape)} != frame shape {(H, W)}")
                mask = mask.to(frames.device, torch.float32).contiguous()
            _lib.check(_lib.lib().m2s_acoustic_encode_packed(
                self._handle, frames.data_ptr(), int(raw), _lib.ptr(mask), B, int(max_frames), lens_dev.data_ptr(),
                lens_host.data_ptr(), out.data_ptr(), self._workspace.data_ptr(), self._workspace.numel(),
                _lib.current_stream()))
        return out

    def encode_frames(self, frames: torch.Tensor) -> torch.Tensor:
        """(N,H,W) float32 cuda -> (N,208): the time-distributed CNN of _cnn_time_distributed."""
        _lib.require_device(frames)
        frames = frames.contiguous().float()
        N, H, W = frames.shape
        with torch.cuda.device(frames.device):
            self._prepare(frames, 1, N, (H, W))
            out = torch.empty(N, self.cnn.out_channels, dtype=torch.float32, device=frames.device)
            _lib.check(_lib.lib().m2s_acoustic_encode(self._handle, frames.data_ptr(), N, out.data_ptr(),
                                                      self._workspace.data_ptr(), self._workspace.numel(),
                                                      _lib.current_stream()))
        return out

    def rnn_head(self, feats: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,T,208) float32 cuda -> (B,T,n_mels): BiLSTM (sum merge) + head."""
        _lib.require_device(feats)
        feats = feats.contiguous().float()
        B, T, _ = feats.shape
        with torch.cuda.device(feats.device):
            self._prepare(feats, B, T, (256, 256) if self._handle_key is None else self._handle_key[1])
            out = torch.empty(B, T, self.n_mels, dtype=torch.float32, device=feats.device)
            lens_dev = lens_host = None
            if lengths is not None:
                lens_host = lengths.detach().to("cpu", torch.int32).contiguous()
                lens_dev = lens_host.to(feats.device)
            _lib.check(_lib.lib().m2s_acoustic_rnn_head(
                self._handle, feats.data_ptr(), B, T, _lib.ptr(lens_dev),
                lens_host.data_ptr() if lens_host is not None else None, out.data_ptr(),
                self._workspace.data_ptr(), self._workspace.numel(), _lib.current_stream()))
        return out


MRIAcousticModel = OTNLikeCNNBiLSTM


def build_acoustic_model(n_mels: int = 64, cnn_pretrained: bool = False, rnn_hidden: int = 640,
                         dropout: float = 0.5, use_checkpoint: bool = False, ckpt_segments: int = 2,
                         use_reentrant: bool = False, precision: Optional[str] = None) -> nn.Module:
    return OTNLikeCNNBiLSTM(n_mels=n_mels, cnn_pretrained=cnn_pretrained, rnn_hidden=rnn_hidden, dropout=dropout,
                            use_checkpoint=use_checkpoint, ckpt_segments=ckpt_segments,
                            use_reentrant=use_reentrant, precision=precision)

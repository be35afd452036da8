"""Scaled-init weights for parity tests (SURVEY.md 8d "scaled-init variant", 7.1).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Default random init makes this path's tensors tiny or flat: the encoder's activations shrink layer by layer under
identity BatchNorm statistics (normalised mel abs-max ~0.06), and the vocoder's output is almost pure DC (mean 0.10,
std 0.006).  Both flatter the north-star gates (mel max-abs 1e-3, waveform SNR 40 dB).  A trained checkpoint is not
like that: BatchNorm statistics keep every layer's activations O(1), the mel head spans the z-scored range, and the
waveform swings through tanh's range.  These helpers put seeded random weights into that regime, in place:

  * acoustic model: random BatchNorm affine parameters, running statistics CALIBRATED on a few frames (set to the batch
    statistics, as training would leave them), SE biases randomised, and the mel head rescaled so that the normalised
    mel of the calibration clip has unit std;
  * Generator: data-dependent per-layer gains on ``weight_g`` -- every upsampler's output has unit std, every ResBlock
    branch contributes at half its input's std, conv_post's pre-tanh output has std 0.5 and zero mean.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .acoustic import acoustic_forward, encoder_forward
from .vocoder import LRELU_SLOPE, causal_conv1d, fold_weight_norm


def calibration_frames(n_clips: int = 8) -> torch.Tensor:
    """Two frames from each of ``n_clips`` synthetic clips (ids 100+): (2 n_clips, 256, 256) float32."""
    from mri2speech_b200 import synth
    return torch.cat([synth.synthetic_clip(100 + i, 3)[1:] for i in range(n_clips)])


def scale_acoustic(model: torch.nn.Module, calib_frames: torch.Tensor, seed: int = 99) -> None:
    """model: the product's (or any timm-named) OTNLikeCNNBiLSTM on the CPU; calib_frames (N,H,W) float32 in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                n = m.num_features
                m.weight.copy_(1.0 + 0.2 * torch.randn(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
            elif isinstance(m, torch.nn.Conv2d) and m.bias is not None:
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        sd = model.state_dict()   # aliases the parameters / buffers
        encoder_forward(sd, calib_frames.unsqueeze(1), calibrate=True)
        mel = acoustic_forward(sd, calib_frames.unsqueeze(0).unsqueeze(2))
        sd["head.weight"].mul_(1.0 / mel.std().clamp_min(1e-6))
        sd["head.bias"].copy_(0.3 * torch.randn(sd["head.bias"].shape, generator=g))


def scale_generator(gen: torch.nn.Module, h, calib_mel: torch.Tensor) -> None:
    """gen: Generator with weight-norm parameters (``weight_g`` / ``weight_v``), ResBlock1 config; calib_mel (1,n_mels,T)
    log-mel as the glue produces it.  Rescales ``weight_g`` (and conv_post's bias) in place."""
    sd = gen.state_dict()
    rb_k = list(h["resblock_kernel_sizes"])
    rb_d = [list(d) for d in h["resblock_dilation_sizes"]]
    nk = len(rb_k)
    with torch.no_grad():
        x = F.conv1d(F.pad(calib_mel, (0, 6)), fold_weight_norm(sd, "conv_pre"), sd["conv_pre.bias"])
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            def up(x_in):
                return F.conv_transpose1d(F.leaky_relu(x_in, LRELU_SLOPE), fold_weight_norm(sd, f"ups.{i}"),
                                          sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            sd[f"ups.{i}.weight_g"].mul_(1.0 / up(x).std().clamp_min(1e-6))
            x = up(x)
            xs = None
            for j in range(nk):
                r = x
                p = f"resblocks.{i * nk + j}"
                for m, d in enumerate(rb_d[j]):
                    def branch(r_in):
                        t = causal_conv1d(F.leaky_relu(r_in, LRELU_SLOPE), fold_weight_norm(sd, f"{p}.convs1.{m}"),
                                          sd[f"{p}.convs1.{m}.bias"], d)
                        return t
                    sd[f"{p}.convs1.{m}.weight_g"].mul_(r.std() / branch(r).std().clamp_min(1e-6))
                    t1 = branch(r)
                    def second(t_in):
                        return causal_conv1d(F.leaky_relu(t_in, LRELU_SLOPE), fold_weight_norm(sd, f"{p}.convs2.{m}"),
                                             sd[f"{p}.convs2.{m}.bias"], 1)
                    sd[f"{p}.convs2.{m}.weight_g"].mul_(0.5 * r.std() / second(t1).std().clamp_min(1e-6))
                    r = r + second(t1)
                xs = r if xs is None else xs + r
            x = xs / nk
        def post(x_in):
            return F.conv1d(F.pad(F.leaky_relu(x_in, 0.01), (0, 6)), fold_weight_norm(sd, "conv_post"), sd["conv_post.bias"])
        sd["conv_post.weight_g"].mul_(0.5 / post(x).std().clamp_min(1e-6))
        sd["conv_post.bias"].sub_(post(x).mean())

"""Second, independent CPU statement of the frame-CNN encoder, assembled from torchvision's own EfficientNet blocks.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists: the encoder's arithmetic is timm==1.0.21 ``tf_efficientnetv2_b2`` (reference call site
mri2speech_code/mri_acoustic_model.py:28-36,46), which is neither vendored nor installed, so ``oracle/acoustic.py``
is a restatement that nothing pins.  This module does NOT pin it to timm either; it removes a different risk: that the
restatement and the CUDA path share one author's misreading of a block.  Here the blocks are torchvision 0.26's
``FusedMBConv`` / ``MBConv`` / ``SqueezeExcitation`` (written by other people, for EfficientNetV2-S/M/L), configured
with the B2 widths / depths, fed the SAME state_dict (timm parameter names), with the one thing torchvision does
differently -- symmetric padding -- replaced by TF "same" padding (stride-2 3x3 on an even input: 0 left/top, 1
right/bottom).  What torchvision decides on its own and the restatement must agree with:
  * the skip rule (stride 1 and in == out) and that the skip is added after the last op of the block;
  * FusedMBConv with expand 1 = ONE 3x3 conv + BN + SiLU (timm ``cn``); otherwise 3x3 expand + BN + SiLU, 1x1 project
    + BN, no activation (timm ``er``);
  * MBConv = 1x1 expand + BN + SiLU, depthwise 3x3 + BN + SiLU, SE, 1x1 project + BN (timm ``ir``);
  * SE = spatial mean -> 1x1 (bias) -> SiLU -> 1x1 (bias) -> sigmoid -> scale, squeeze width = block input // 4;
  * BatchNorm in eval mode with eps 1e-3.
``tests/test_oracle_acoustic.py`` asserts both statements agree to 1e-5 on BN-randomised weights.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
from torchvision.models.efficientnet import FusedMBConv, FusedMBConvConfig, MBConv, MBConvConfig
from torchvision.ops.misc import Conv2dNormActivation

BN_EPS = 1e-3
# tf_efficientnetv2_b2 after the 1.1 width / 1.2 depth multipliers
STAGES = (
    ("fused", 1, 3, 1, 32, 16, 2),
    ("fused", 4, 3, 2, 16, 32, 3),
    ("fused", 4, 3, 2, 32, 56, 3),
    ("mb", 4, 3, 2, 56, 104, 4),
    ("mb", 6, 3, 1, 104, 120, 6),
    ("mb", 6, 3, 2, 120, 208, 10),
)


def _tf_same(cna: Conv2dNormActivation) -> None:
    """Give the Conv2d inside a Conv2dNormActivation TF "same" padding for the even sizes met on this path."""
    conv = cna[0]
    k, s = conv.kernel_size[0], conv.stride[0]
    if k == 1:
        return
    total = k - s if s > 1 else k - 1          # even input: ceil(i/s)*s - s + k - i  = k - s
    lo = total // 2
    conv.padding = (0, 0)
    cna[0] = nn.Sequential(nn.ZeroPad2d((lo, total - lo, lo, total - lo)), conv)


def _copy_bn(bn: nn.BatchNorm2d, sd, p):
    bn.weight.copy_(sd[p + ".weight"]); bn.bias.copy_(sd[p + ".bias"])
    bn.running_mean.copy_(sd[p + ".running_mean"]); bn.running_var.copy_(sd[p + ".running_var"])


def _conv_of(cna):
    c = cna[0]
    return c[1] if isinstance(c, nn.Sequential) else c


def build_encoder_tv(sd: dict, prefix: str = "cnn.backbone.") -> nn.Module:
    """torchvision-block network carrying the timm-named weights of ``sd``."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    norm = partial(nn.BatchNorm2d, eps=BN_EPS)
    layers = []
    with torch.no_grad():
        stem = Conv2dNormActivation(3, 32, kernel_size=3, stride=2, norm_layer=norm, activation_layer=nn.SiLU)
        _conv_of(stem).weight.copy_(sd["conv_stem.weight"])
        _copy_bn(stem[1], sd, "bn1")
        _tf_same(stem)
        layers.append(stem)
        for s, (kind, expand, k, stride, cin, cout, reps) in enumerate(STAGES):
            for b in range(reps):
                p = f"blocks.{s}.{b}"
                st = stride if b == 0 else 1
                ci = cin if b == 0 else cout
                if kind == "fused":
                    blk = FusedMBConv(FusedMBConvConfig(expand, k, st, ci, cout, 1), 0.0, norm)
                    seq = blk.block
                    if expand == 1:
                        _conv_of(seq[0]).weight.copy_(sd[p + ".conv.weight"]); _copy_bn(seq[0][1], sd, p + ".bn1")
                    else:
                        _conv_of(seq[0]).weight.copy_(sd[p + ".conv_exp.weight"]); _copy_bn(seq[0][1], sd, p + ".bn1")
                        _conv_of(seq[1]).weight.copy_(sd[p + ".conv_pwl.weight"]); _copy_bn(seq[1][1], sd, p + ".bn2")
                    _tf_same(seq[0])
                else:
                    blk = MBConv(MBConvConfig(expand, k, st, ci, cout, 1), 0.0, norm)
                    seq = blk.block
                    _conv_of(seq[0]).weight.copy_(sd[p + ".conv_pw.weight"]); _copy_bn(seq[0][1], sd, p + ".bn1")
                    _conv_of(seq[1]).weight.copy_(sd[p + ".conv_dw.weight"]); _copy_bn(seq[1][1], sd, p + ".bn2")
                    se = seq[2]
                    se.fc1.weight.copy_(sd[p + ".se.conv_reduce.weight"]); se.fc1.bias.copy_(sd[p + ".se.conv_reduce.bias"])
                    se.fc2.weight.copy_(sd[p + ".se.conv_expand.weight"]); se.fc2.bias.copy_(sd[p + ".se.conv_expand.bias"])
                    _conv_of(seq[3]).weight.copy_(sd[p + ".conv_pwl.weight"]); _copy_bn(seq[3][1], sd, p + ".bn3")
                    _tf_same(seq[1])
                layers.append(blk)
    return nn.Sequential(*layers).eval()


def encoder_forward_tv(sd: dict, frames: torch.Tensor, prefix: str = "cnn.backbone.") -> torch.Tensor:
    """(N,1,H,W) or (N,H,W) float32 -> (N,208): mri_acoustic_model.py:39-48 on the torchvision-block network."""
    x = frames
    if x.dim() == 3:
        x = x.unsqueeze(1)
    if x.size(1) == 1:
        x = x.repeat(1, 3, 1, 1)
    net = build_encoder_tv(sd, prefix)
    with torch.no_grad():
        return net(x).mean(dim=(2, 3))

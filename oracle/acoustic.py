"""CPU oracle: OTNLikeCNNBiLSTM forward, restated functionally over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites followed:
  * mri2speech_code/mri_acoustic_model.py:39-48   EffNetV2B2Backbone.forward
    (repeat 1->3 channels, ``backbone(x)[-1]``, spatial mean)
  * mri2speech_code/mri_acoustic_model.py:67-72   BiLSTMSumMerge.forward
    (nn.LSTM bidirectional, fwd+bwd SUM, dropout = identity in eval)
  * mri2speech_code/mri_acoustic_model.py:105-136 time-distributed CNN, head

PARITY UNPINNED for the encoder: its arithmetic is timm==1.0.21's
``tf_efficientnetv2_b2`` (features_only) which is neither vendored in the reference
nor installed here.  Restated from the published topology:
  stem conv3x3 s2 3->32, BN(eps 1e-3), SiLU;
  cn_r2_k3_s1_e1_c16 | er_r3_k3_s2_e4_c32 | er_r3_k3_s2_e4_c56 |
  ir_r4_k3_s2_e4_c104_se0.25 | ir_r6_k3_s1_e6_c120_se0.25 | ir_r10_k3_s2_e6_c208_se0.25
  TF "same" padding (stride 2 on even input: pad 0 left/top, 1 right/bottom).
Parameter names are timm's so that a real checkpoint would load.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3

# (block type, repeats, stride, expand, out channels, has SE)
STAGES = (
    ("cn", 2, 1, 1, 16, False),
    ("er", 3, 2, 4, 32, False),
    ("er", 3, 2, 4, 56, False),
    ("ir", 4, 2, 4, 104, True),
    ("ir", 6, 1, 6, 120, True),
    ("ir", 10, 2, 6, 208, True),
)
STEM_CH = 32
FEAT_CH = 208


def _bn(sd, p, x, act, calibrate=False):
    if calibrate:
        # scaled-init variant (oracle/scaled_init.py): running statistics := the statistics of THIS batch, as training
        # would have left them, so that every layer's activations are O(1) (in place: sd aliases the module's buffers)
        sd[p + ".running_mean"].copy_(x.mean((0, 2, 3)))
        sd[p + ".running_var"].copy_(x.var((0, 2, 3), unbiased=False).clamp_min(1e-6))
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                     sd[p + ".weight"], sd[p + ".bias"], training=False, eps=BN_EPS)
    return F.silu(y) if act else y


def _conv_same(x, w, stride, groups=1):
    """TF 'same' padding for a kxk conv (timm Conv2dSame semantics)."""
    k = w.shape[-1]
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / stride) - 1) * stride + (k - 1) + 1 - ih, 0)
    pw = max((math.ceil(iw / stride) - 1) * stride + (k - 1) + 1 - iw, 0)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    return F.conv2d(x, w, None, stride=stride, groups=groups)


def encoder_forward(sd: dict, frames: torch.Tensor, prefix: str = "cnn.backbone.", calibrate: bool = False) -> torch.Tensor:
    """(N,1,H,W) or (N,H,W) float32 -> (N,208) features (global-average-pooled last stage).
    ``calibrate``: overwrite every BatchNorm's running statistics with this batch's (see ``_bn``)."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = frames
    if x.dim() == 3:
        x = x.unsqueeze(1)
    if x.size(1) == 1:
        x = x.repeat(1, 3, 1, 1)
    x = _bn(sd, "bn1", _conv_same(x, sd["conv_stem.weight"], 2), True, calibrate)
    cin = STEM_CH
    for s, (kind, reps, stride, _exp, cout, _se) in enumerate(STAGES):
        for b in range(reps):
            p = f"blocks.{s}.{b}"
            st = stride if b == 0 else 1
            skip = (st == 1 and cin == cout)
            inp = x
            if kind == "cn":
                x = _bn(sd, p + ".bn1", _conv_same(x, sd[p + ".conv.weight"], st), True, calibrate)
            elif kind == "er":
                x = _bn(sd, p + ".bn1", _conv_same(x, sd[p + ".conv_exp.weight"], st), True, calibrate)
                x = _bn(sd, p + ".bn2", F.conv2d(x, sd[p + ".conv_pwl.weight"]), False, calibrate)
            else:
                x = _bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv_pw.weight"]), True, calibrate)
                x = _bn(sd, p + ".bn2",
                        _conv_same(x, sd[p + ".conv_dw.weight"], st, groups=x.shape[1]), True, calibrate)
                se = x.mean((2, 3), keepdim=True)
                se = F.silu(F.conv2d(se, sd[p + ".se.conv_reduce.weight"], sd[p + ".se.conv_reduce.bias"]))
                se = F.conv2d(se, sd[p + ".se.conv_expand.weight"], sd[p + ".se.conv_expand.bias"])
                x = x * torch.sigmoid(se)
                x = _bn(sd, p + ".bn3", F.conv2d(x, sd[p + ".conv_pwl.weight"]), False, calibrate)
            if skip:
                x = x + inp
            cin = cout
    return x.mean(dim=(2, 3))


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One LSTM direction, PyTorch semantics (gate order i,f,g,o; h0=c0=0).  x: (T, C) -> (T, H)."""
    T = x.shape[0]
    H = w_hh.shape[1]
    h = torch.zeros(H, dtype=x.dtype)
    c = torch.zeros(H, dtype=x.dtype)
    out = torch.zeros(T, H, dtype=x.dtype)
    gin = x @ w_ih.t() + b_ih + b_hh
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        z = gin[t] + w_hh @ h
        i, f, g, o = z.split(H)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out


def bilstm_head_forward(sd: dict, feats: torch.Tensor, explicit: bool = False) -> torch.Tensor:
    """(B,T,208) -> (B,T,n_mels).  ``explicit`` uses the hand-rolled recurrence instead of nn.LSTM."""
    w = {k: sd["rnn.lstm." + k] for k in (
        "weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
        "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse")}
    H = w["weight_hh_l0"].shape[1]
    if explicit:
        ys = []
        for b in range(feats.shape[0]):
            yf = lstm_direction(feats[b], w["weight_ih_l0"], w["weight_hh_l0"],
                                w["bias_ih_l0"], w["bias_hh_l0"], False)
            yb = lstm_direction(feats[b], w["weight_ih_l0_reverse"], w["weight_hh_l0_reverse"],
                                w["bias_ih_l0_reverse"], w["bias_hh_l0_reverse"], True)
            ys.append(yf + yb)
        y = torch.stack(ys)
    else:
        lstm = torch.nn.LSTM(feats.shape[-1], H, 1, batch_first=True, bidirectional=True)
        lstm.load_state_dict(w)
        with torch.no_grad():
            y2, _ = lstm(feats)
        yf, yb = y2.chunk(2, dim=-1)
        y = yf + yb
    return F.linear(y, sd["head.weight"], sd["head.bias"])


def acoustic_forward(sd: dict, x: torch.Tensor, lengths=None) -> torch.Tensor:
    """(B,T,1,H,W)|(B,T,H,W) -> (B,T,n_mels) normalised mel.

    With ``lengths`` each utterance is run on its own first len_i frames (the
    ragged parity target "each utterance equals its own B=1 run"; the reference
    itself has no mask argument, mri_acoustic_model.py:116-136) and the padded
    tail of the output is zero.
    """
    B, T = x.shape[:2]
    with torch.no_grad():
        f = encoder_forward(sd, x.reshape(B * T, *x.shape[2:])).view(B, T, -1)
        if lengths is None:
            return bilstm_head_forward(sd, f)
        out = torch.zeros(B, T, sd["head.weight"].shape[0])
        for b, ln in enumerate(lengths):
            if ln > 0:
                out[b, :ln] = bilstm_head_forward(sd, f[b:b + 1, :ln])[0]
        return out

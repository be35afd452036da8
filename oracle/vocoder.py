"""CPU oracle: HiFi-GAN Generator forward, restated functionally over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference's
own ``models.Generator`` by ``oracle/gen_golden.py`` / ``tests/test_oracle_vocoder.py``.

Follows the reference's arithmetic, not its code structure:
  * models.py:113-131   Generator.forward
  * models.py:35-49     ResBlock1.forward (keep-first-L trimming)
  * utils.py:34-35      get_padding = k*d - d  (twice the stock HiFi-GAN value)

Consequences restated here explicitly:
  * every ResBlock conv is CAUSAL: (k-1)*d zeros on the left, no look-ahead;
  * conv_pre / conv_post are ANTI-causal: 6 zeros on the right (F.pad(x,(0,6)));
  * the activation before conv_post uses slope 0.01 (F.leaky_relu default), all
    others 0.1;
  * weight-norm (old style, dim=0): W = g * v / ||v||_2 over dims (1,2), so g is
    per C_out for Conv1d and per C_in for ConvTranspose1d.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # models.py:8


def fold_weight_norm(sd: dict, prefix: str, dtype=None) -> torch.Tensor:
    """Return the effective conv weight for ``prefix`` (e.g. ``ups.0``).

    Accepts both checkpoint flavours: ``weight_g``/``weight_v``
    (weight-norm present) or plain ``weight`` (weight-norm removed / conv_pre).
    """
    if prefix + ".weight" in sd:
        w = sd[prefix + ".weight"]
    else:
        g = sd[prefix + ".weight_g"].double()
        v = sd[prefix + ".weight_v"].double()
        norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        w = (g * v / norm)
    return w.to(dtype if dtype is not None else torch.float32)


def causal_conv1d(x, w, b, dilation):
    """out[t] = b + sum_j w[:,:,j] x[t-(k-1-j)*d]; zeros for negative time."""
    k = w.shape[-1]
    return F.conv1d(F.pad(x, ((k - 1) * dilation, 0)), w, b, dilation=dilation)


def generator_forward(sd: dict, h, mel: torch.Tensor, lengths=None, dtype=torch.float32):
    """(B, num_mels, T) -> (B, 1, T*prod(upsample_rates)).

    ``lengths`` (optional, per-utterance mel frames) reproduces "each utterance
    equals its own B=1 run" on a zero-padded batch by zeroing activations past
    len_i*(prod u so far) at the conv_pre input, every ups input and the
    conv_post input (SURVEY section 7 "Ragged batches").
    """
    ups_rates = list(h["upsample_rates"])
    ups_k = list(h["upsample_kernel_sizes"])
    rb_k = list(h["resblock_kernel_sizes"])
    rb_d = [list(d) for d in h["resblock_dilation_sizes"]]
    nk = len(rb_k)
    rb2 = str(h["resblock"]) != "1"   # models.py:95: ResBlock1 if h.resblock == '1' else ResBlock2
    if mel.dim() == 2:
        mel = mel.unsqueeze(0)
    x = mel.to(dtype)

    def mask(x, scale):
        if lengths is None:
            return x
        t = torch.arange(x.shape[-1]).view(1, 1, -1)
        ln = (torch.as_tensor(lengths).view(-1, 1, 1) * scale)
        return x * (t < ln).to(x.dtype)

    get = lambda name: sd[name].to(dtype)
    scale = 1
    x = mask(x, scale)
    x = F.conv1d(F.pad(x, (0, 6)), fold_weight_norm(sd, "conv_pre", dtype), get("conv_pre.bias"))
    for i, (u, k) in enumerate(zip(ups_rates, ups_k)):
        x = mask(F.leaky_relu(x, LRELU_SLOPE), scale)
        x = F.conv_transpose1d(x, fold_weight_norm(sd, f"ups.{i}", dtype), get(f"ups.{i}.bias"),
                               stride=u, padding=(k - u) // 2)
        scale *= u
        xs = None
        for j in range(nk):
            r = x
            p = f"resblocks.{i * nk + j}"
            if rb2:
                # ResBlock2 (models.py:58-80): two units x = x + c(lrelu(x)) with dilation[0], dilation[1]; the conv pads
                # (k-1)*d on both sides (utils.py:34-35) and :74-78 keeps the first L outputs -> causal, like ResBlock1
                for m in range(2):
                    t = causal_conv1d(F.leaky_relu(r, LRELU_SLOPE),
                                      fold_weight_norm(sd, f"{p}.convs.{m}", dtype), get(f"{p}.convs.{m}.bias"), rb_d[j][m])
                    r = r + t
            else:
                for m, d in enumerate(rb_d[j]):
                    t = causal_conv1d(F.leaky_relu(r, LRELU_SLOPE),
                                      fold_weight_norm(sd, f"{p}.convs1.{m}", dtype), get(f"{p}.convs1.{m}.bias"), d)
                    t = causal_conv1d(F.leaky_relu(t, LRELU_SLOPE),
                                      fold_weight_norm(sd, f"{p}.convs2.{m}", dtype), get(f"{p}.convs2.{m}.bias"), 1)
                    r = r + t
            xs = r if xs is None else xs + r
        x = xs / nk
    x = mask(F.leaky_relu(x, 0.01), scale)
    x = F.conv1d(F.pad(x, (0, 6)), fold_weight_norm(sd, "conv_post", dtype), get("conv_post.bias"))
    return torch.tanh(x)


def snr_db(ref: torch.Tensor, test: torch.Tensor, remove_mean: bool = False) -> float:
    """10 log10(sum ref^2 / sum (ref-test)^2); optionally after removing ref's DC."""
    ref = ref.double().flatten()
    test = test.double().flatten()
    if remove_mean:
        m = ref.mean()
        ref = ref - m
        test = test - m
    num = (ref * ref).sum()
    den = ((ref - test) ** 2).sum().clamp_min(1e-300)
    return float(10.0 * torch.log10(num / den))

"""Time the fused ResBlock-pair kernel on one vocoder-stage shape with the four combinations of residual source
(fp32 tensor | split-fp16 planes) and output (fp32 + fp16 | split-fp16 planes): which side of the split stream costs.

    python tools/pair_bench.py [B L C k dilation]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib


def main():
    B, L, C, k, dil = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (32, 107520, 32, 3, 1)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, L, C, generator=g).half().cuda()
    x_lo = (torch.randn(B, L, C, generator=g) * 0.2).half().cuda()
    res32 = x.float() + x_lo.float() / 2048.0
    w1 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    w2 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    b1 = torch.randn(C, generator=g).cuda()
    b2 = torch.randn(C, generator=g).cuda()
    o32 = torch.zeros(B, L, C, device="cuda")
    ohi = torch.zeros(B, L, C, device="cuda", dtype=torch.float16)
    olo = torch.zeros(B, L, C, device="cuda", dtype=torch.float16)
    variants = {
        "res fp32  -> fp32 + fp16": dict(res=res32, out=o32, out16=ohi),
        "res split -> fp32 + fp16": dict(res_hi=x, res_lo=x_lo, out=o32, out16=ohi),
        "res fp32  -> hi + lo    ": dict(res=res32, out16=ohi, out16_lo=olo, want_d32=False),
        "res split -> hi + lo    ": dict(res_hi=x, res_lo=x_lo, out16=ohi, out16_lo=olo, want_d32=False),
        "res split -> hi only    ": dict(res_hi=x, res_lo=x_lo, out16=ohi, want_d32=False),
        "no res    -> hi only    ": dict(out16=ohi, want_d32=False),
    }
    for name, kw in variants.items():
        best = 1e9
        try:
            for _ in range(3):
                _lib.profile(True)
                _lib.resblock_pair_fwd(x, w1, b1, dil, w2, b2, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1, **kw)
                ms, _ = _lib.profile_read()
                _lib.profile(False)
                best = min(best, ms[0])
        except _lib.M2SError as e:   # e.g. the lo plane is only written by the split programs
            _lib.profile(False)
            print(f"B={B} L={L} C={C} k={k} d={dil}  {name}: unsupported ({e})", flush=True)
            continue
        print(f"B={B} L={L} C={C} k={k} d={dil}  {name}: {best * 1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()

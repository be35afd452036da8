"""Timeline of CTA 0 of the fused ResBlock-pair kernel on one vocoder-stage shape (debug tool; run on a B200).

    python tools/trace_pair.py [B L C k dilation [res: none|fp32|split]]

Per iteration (cycles relative to the first stamp): when the epilogue warps got conv1's accumulators / finished
epilogue 1 / got conv2's accumulators / finished epilogue 2, when the MMA warp had issued conv1 / saw the T tile /
had issued conv2, and when the producer issued the x tile.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib

NAMES = ["acc1", "epi1", "acc2", "epi2", "C1iss", "Tfull", "C2iss", "Aiss"]


def main():
    a = sys.argv[1:]
    B, L, C, k, dil = [int(v) for v in a[:5]] if len(a) >= 5 else (32, 107520, 32, 3, 1)
    mode = a[5] if len(a) > 5 else "fp32"
    tiles = 24
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, L, C, generator=g).half().cuda()
    x_lo = (torch.randn(B, L, C, generator=g) * 0.2).half().cuda()
    w1 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    w2 = (torch.randn(k, C, C, generator=g) / (C * k) ** 0.5).cuda()
    b1 = torch.randn(C, generator=g).cuda()
    b2 = torch.randn(C, generator=g).cuda()
    ohi = torch.zeros(B, L, C, device="cuda", dtype=torch.float16)
    olo = torch.zeros(B, L, C, device="cuda", dtype=torch.float16)
    kw = dict(res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1, out16=ohi)
    if mode == "fp32":      # fp32 residual -> fp32 + fp16 outputs (the default fp16 build)
        kw["res"] = x.float() + x_lo.float() / 2048.0
    elif mode == "split":   # (hi, lo) residual -> (hi, lo) outputs (M2S_SPLIT_RES=1)
        kw.update(res_hi=x, res_lo=x_lo, out16_lo=olo, want_d32=False)
    else:
        kw["want_d32"] = False
    buf = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
    _lib.resblock_pair_fwd(x, w1, b1, dil, w2, b2, **kw)
    _lib.check(_lib.lib().m2s_debug_trace(buf.data_ptr(), tiles))
    _lib.resblock_pair_fwd(x, w1, b1, dil, w2, b2, **kw)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().m2s_debug_trace(None, 0))
    t = buf.cpu().view(tiles, 8)
    t0 = int(t[t > 0].min())
    print(f"--- fused pair B={B} L={L} C={C} k={k} d={dil} residual={mode}: CTA 0, cycles since its first stamp")
    print("iter " + " ".join(f"{n:>8s}" for n in NAMES))
    for i in range(tiles):
        print(f"{i:4d} " + " ".join(f"{int(v) - t0:8d}" if int(v) > 0 else "       -" for v in t[i]))
    d = (t[8:20, 3] - t[7:19, 3]).float()
    print(f"period (epilogue 2 done -> next): {d.mean().item():.0f} cycles; epilogue 1 {(t[8:20, 1] - t[8:20, 0]).float().mean().item():.0f}, "
          f"wait for acc2 {(t[8:20, 2] - t[9:21, 1]).float().mean().item():.0f}, epilogue 2 {(t[8:20, 3] - t[8:20, 2]).float().mean().item():.0f}")


if __name__ == "__main__":
    main()

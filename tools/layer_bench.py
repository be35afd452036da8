"""Time ONE conv-engine launch of a given shape in isolation (CUDA events around the kernel, engine profile ring),
optionally with probe switches, to find what bounds a layer:  dbg bit0 skip global stores, bit1 skip TMEM loads,
bit2 skip the SMEM transpose, bit3 skip MMA issue.

    python tools/layer_bench.py B L C N k dilation [half=0|1] [res=0|1] [out=32|16|both] [dbg...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib


def run(B, L, C, N, k, dil, half, res, out, dbg, reps=3):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(B, L, C, generator=g).cuda()
    if half:
        a = a.half()
    w = (torch.randn(k, N, C, generator=g) / (C * k) ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    r = torch.randn(B, L, N, generator=g).cuda() if res else None
    d32 = torch.zeros(B, L, N, device="cuda")
    d16 = torch.zeros(B, L, N, device="cuda", dtype=torch.float16) if out in ("16", "both") else None
    shifts = [-(k - 1 - j) * dil for j in range(k)]
    _lib.set_knob("dbg", dbg)
    best = 1e9
    for _ in range(reps):
        _lib.profile(True)
        _lib.conv_fwd(a, w, shifts, L, bias=bias, res=r, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1,
                      out=d32, out16=d16, want_d32=out in ("32", "both"))
        ms, fl = _lib.profile_read()
        _lib.profile(False)
        best = min(best, ms[0])
    _lib.set_knob("dbg", 0)
    return best * 1e3, fl[0] / (best * 1e-3) / 1e12


if __name__ == "__main__":
    B, L, C, N, k, dil = [int(v) for v in sys.argv[1:7]]
    half = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    res = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    out = sys.argv[9] if len(sys.argv) > 9 else "32"
    dbgs = [int(v) for v in sys.argv[10:]] or [0]
    for dbg in dbgs:
        us, tf = run(B, L, C, N, k, dil, half, res, out, dbg)
        print(f"B={B} L={L} C={C} N={N} k={k} d={dil} half={half} res={res} out={out} dbg={dbg}: {us:8.1f} us  {tf:7.1f} TF/s",
              flush=True)

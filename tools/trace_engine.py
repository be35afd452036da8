"""Timeline of CTA 0 of the conv engine on one vocoder-shaped conv (debug tool; run on a B200)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib


def run(B, L, C, k, d, res=False, tiles=40, knobs=None, N=None, act=None, shifts=None, half=False, outs="32"):
    for kk, v in (knobs or {}).items():
        _lib.set_knob(kk, v)
    N = N or C
    act = _lib.ACT_LRELU if act is None else act
    a = torch.randn(B, L, C, device="cuda")
    if half:
        a = a.half()
    w = torch.randn(k, N, C, device="cuda") / (C * k) ** 0.5
    bias = torch.randn(N, device="cuda")
    r = torch.randn(B, L, N, device="cuda") if res else None
    shifts = shifts or [-(k - 1 - j) * d for j in range(k)]
    out = torch.empty(B, L, N, device="cuda")
    buf = torch.zeros(tiles * 17, dtype=torch.int64, device="cuda")
    out16 = torch.empty(B, L, N, device="cuda", dtype=torch.float16) if outs in ("16", "both") else None
    kw = dict(bias=bias, res=r, res_inv_slope=10.0 if res else 1.0, act=act, act_slope=0.1, out=out, out16=out16,
              want_d32=outs in ("32", "both"))
    _lib.conv_fwd(a, w, shifts, L, **kw)
    _lib.check(_lib.lib().m2s_debug_trace(buf.data_ptr(), tiles))
    _lib.conv_fwd(a, w, shifts, L, **kw)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().m2s_debug_trace(None, 0))
    micro = buf.cpu()[tiles * 9:].view(tiles, 8)
    t = buf.cpu()[:tiles * 9].view(tiles, 9)
    valid = int((t[:, 8] > 0).sum())
    t = t[:valid]
    tiles = valid
    t0 = int(t[0, 0])
    print(f"--- B={B} L={L} C={C} N={N} k={k} d={d} res={res} act={act} half={half} out={outs} knobs={knobs}: CTA0 ran {tiles} tiles")
    for i in range(min(tiles, 4)):
        print(f"{i:4d} | " + " ".join(f"{int(v) - t0:8d}" for v in t[i, 0:3]) + " | " +
              " ".join(f"{int(v) - t0:8d}" for v in t[i, 3:6]) + " | " + " ".join(f"{int(v) - t0:8d}" for v in t[i, 6:9]))
    s0 = 1 if tiles > 2 else 0
    n = max(tiles - 1 - s0, 1)
    per_tile = (int(t[tiles - 1, 8]) - int(t[s0, 8])) / n
    f = lambda x: float(x.float().mean())
    print(f"steady-state cycles/tile: {per_tile:.0f}; epilogue busy {f(t[s0:, 8] - t[s0:, 7]):.0f}; "
          f"epilogue wait {f(t[s0:, 7] - t[s0:, 6]):.0f}; mma issue {f(t[s0:, 5] - t[s0:, 4]):.0f}; "
          f"mma wait acc {f(t[s0:, 4] - t[s0:, 3]):.0f}; producer a_empty wait {f(t[s0:, 1] - t[s0:, 0]):.0f}; "
          f"producer issue {f(t[s0:, 2] - t[s0:, 1]):.0f}; total {int(t[tiles-1, 8]) - t0} cycles")
    m = micro[s0:tiles]   # intra-unit stamps: only written by builds before the 16-warp epilogue (kept for old logs)
    if int((m[:, 4] > 0).sum()) > 0:
        m = m[m[:, 4] > 0]
        print(f"   first unit of the tile (warp 2): TMEM wait {f(m[:, 1] - m[:, 0]):.0f}; park in staging {f(m[:, 2] - m[:, 1]):.0f}; "
              f"first row done {f(m[:, 3] - m[:, 2]):.0f}; rows 1-7 + stores {f(m[:, 4] - m[:, 3]):.0f}; unit total {f(m[:, 4] - m[:, 0]):.0f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "vocoder16":
    # narrow vocoder layers with fp16 operands on the single-CTA kernel: where do the cycles of a tile go?
    for dbg in (0, 1, 8, 9):
        run(32, 107520, 32, 11, 1, half=True, outs="16", knobs={"pair": 0, "dbg": dbg})
    for dbg in (0, 9):
        run(32, 107520, 32, 11, 1, half=True, outs="both", res=True, knobs={"pair": 0, "dbg": dbg})
        run(32, 53760, 64, 3, 1, half=True, outs="16", knobs={"pair": 0, "dbg": dbg})
        run(32, 17920, 128, 3, 1, half=True, outs="16", knobs={"pair": 0, "dbg": dbg})
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "micro":
    S = _lib.ACT_SILU
    run(1, 65536, 120, 1, 0, N=720, act=S, shifts=[0], half=True, outs="16", knobs={"pair": 0})        # SiLU, fp16 out
    run(32, 17920, 128, 3, 1, half=True, outs="16", knobs={"pair": 0})                                  # lrelu, fp16 out
    run(32, 17920, 128, 3, 1, half=True, outs="both", res=True, knobs={"pair": 0})                      # RB, both outs
    run(32, 53760, 64, 3, 1, half=True, outs="both", res=True, knobs={"pair": 0})
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "encoder16":
    # encoder IR-stage GEMMs with fp16 operands on the single-CTA kernel
    S = _lib.ACT_SILU
    for dbg in (0, 1, 9):
        run(1, 65536, 120, 1, 0, N=720, act=S, shifts=[0], half=True, outs="16", knobs={"pair": 0, "dbg": dbg})   # stage 4 expand
    for dbg in (0, 9):
        run(1, 65536, 720, 1, 0, N=120, act=_lib.ACT_NONE, shifts=[0], half=True, outs="both", res=True,
            knobs={"pair": 0, "dbg": dbg})                                                                       # stage 4 project
        run(1, 16384, 208, 1, 0, N=1248, act=S, shifts=[0], half=True, outs="16", knobs={"pair": 0, "dbg": dbg})  # stage 5 expand
    sys.exit(0)

if __name__ == "__main__":
    S = _lib.ACT_SILU
    for dbg in (0, 1, 2, 4, 7, 16, 23):
        run(1, 16384, 208, 1, 0, N=1248, act=S, shifts=[0], knobs={"dbg": dbg})       # stage 5 expand

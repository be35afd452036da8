"""Timeline of CTA 0 of the conv engine on one vocoder-shaped conv (debug tool; run on a B200)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib


def run(B, L, C, k, d, res=False, tiles=40, knobs=None):
    for kk, v in (knobs or {}).items():
        _lib.set_knob(kk, v)
    a = torch.randn(B, L, C, device="cuda")
    w = torch.randn(k, C, C, device="cuda") / (C * k) ** 0.5
    bias = torch.randn(C, device="cuda")
    r = torch.randn(B, L, C, device="cuda") if res else None
    shifts = [-(k - 1 - j) * d for j in range(k)]
    out = torch.empty(B, L, C, device="cuda")
    buf = torch.zeros(tiles * 9, dtype=torch.int64, device="cuda")
    _lib.conv_fwd(a, w, shifts, L, bias=bias, res=r, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1, out=out)
    _lib.check(_lib.lib().m2s_debug_trace(buf.data_ptr(), tiles))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    _lib.conv_fwd(a, w, shifts, L, bias=bias, res=r, res_inv_slope=10.0, act=_lib.ACT_LRELU, act_slope=0.1, out=out)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().m2s_debug_trace(None, 0))
    t = buf.cpu().view(tiles, 9)
    t0 = int(t[0, 0])
    print(f"--- B={B} L={L} C={C} k={k} d={d} res={res} knobs={knobs}: {e0.elapsed_time(e1)*1e3:.0f} us (incl. weight packing)")
    print("tile |  P:start  a_empty  issued |  M:start acc_empty  mma_done |  E:start acc_full  done   (cycles since t0)")
    for i in range(min(tiles, 6)):
        print(f"{i:4d} | " + " ".join(f"{int(v) - t0:8d}" for v in t[i, 0:3]) + " | " +
              " ".join(f"{int(v) - t0:8d}" for v in t[i, 3:6]) + " | " + " ".join(f"{int(v) - t0:8d}" for v in t[i, 6:9]))
    n = tiles - 4 if tiles > 4 else 1
    per_tile = (int(t[tiles - 1, 8]) - int(t[3, 8])) / n
    print(f"steady-state cycles/tile: {per_tile:.0f}; epilogue busy {float((t[4:, 8] - t[4:, 7]).float().mean()):.0f}; "
          f"epilogue wait {float((t[4:, 7] - t[4:, 6]).float().mean()):.0f}; mma issue {float((t[4:, 5] - t[4:, 4]).float().mean()):.0f}; "
          f"mma wait acc {float((t[4:, 4] - t[4:, 3]).float().mean()):.0f}; producer a_empty wait {float((t[4:, 1] - t[4:, 0]).float().mean()):.0f}; "
          f"producer issue {float((t[4:, 2] - t[4:, 1]).float().mean()):.0f}")


if __name__ == "__main__":
    run(32, 107520, 32, 3, 1, tiles=40)
    run(32, 107520, 32, 3, 1, res=True, tiles=40)
    run(32, 107520, 32, 11, 5, tiles=40)
    run(32, 53760, 64, 7, 3, tiles=40)
    run(32, 17920, 128, 7, 3, tiles=12)
    run(32, 2560, 256, 11, 5, tiles=4)

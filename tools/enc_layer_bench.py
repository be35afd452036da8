"""Time one encoder-shaped 3x3 stride-1 conv (padded NHWC rows, 9 row-shifted taps, image-border mask, SiLU, fp16 in /
fp16 out) on the conv engine in isolation, with probe switches and kernel-choice knobs.

    python tools/enc_layer_bench.py [frames H W C_in N]
dbg bits: 1 skip global stores, 2 skip TMEM loads, 4 skip the SMEM transpose, 8 skip MMA issue.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib


def main():
    a = sys.argv[1:]
    F, H, W, C, N = [int(v) for v in a[:5]] if len(a) >= 5 else (1024, 64, 64, 32, 128)
    pitch = W + 2
    rows = (H + 2) * pitch
    g = torch.Generator().manual_seed(1)
    x = torch.randn(F, rows, C, generator=g).half().cuda()
    w = (torch.randn(9, N, C, generator=g) / (9 * C) ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    shifts = [(dy - 1) * pitch + (dx - 1) for dy in range(3) for dx in range(3)]
    out16 = torch.zeros(F, rows, N, device="cuda", dtype=torch.float16)
    dummy = torch.empty(1, rows, 1, device="cuda")   # conv_fwd wants an fp32 `out` to size d_batch_rows; it is not written
    first = pitch + 1                       # first interior pixel
    l_out = (H - 1) * pitch + W             # rows first .. last interior pixel
    flops = 2.0 * F * H * W * 9 * C * N

    def run(knobs, dbg):
        for k, v in knobs.items():
            _lib.set_knob(k, v)
        _lib.set_knob("dbg", dbg)
        best = 1e9
        for _ in range(3):
            _lib.profile(True)
            args, _ = _lib.conv_fwd(x[:, first:], w, shifts, l_out, bias=bias, act=_lib.ACT_SILU, out16=out16,
                                    want_d32=False, out=dummy, pitch_mask=(pitch, 1, H + 1, 1, W + 1), d_row_offset=first,
                                    d_rows=rows, _return_args=True)
            # A rows before `first` must be addressable: describe A from the frame start with shifted taps instead
            args.a = x.data_ptr(); args.a_rows = rows; args.a_batch_rows = rows
            for i, s in enumerate(shifts):
                args.shift[i] = s + first
            _lib.check(_lib.lib().m2s_conv_fwd(_lib.C.byref(args), _lib.IMPL_TCGEN05, _lib.current_stream()))
            ms, _f = _lib.profile_read()
            _lib.profile(False)
            best = min(best, ms[0])
        _lib.set_knob("dbg", 0)
        return best * 1e3

    for name, knobs in (("default", {}), ("single-CTA kernel", {"pair": 0}), ("pair forced", {"pair": 2}),
                        ("n_tile_max 64", {"pair": 1, "n_tile_max": 64}), ("n_tile_max 256", {"n_tile_max": 256})):
        for dbg in (0, 1, 8, 9):
            try:
                us = run(knobs, dbg)
                print(f"F={F} {H}x{W} C={C} N={N}  {name:18s} dbg={dbg}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
            except _lib.M2SError as e:
                print(f"{name} dbg={dbg}: {e}", flush=True)
        _lib.set_knob("pair", 1); _lib.set_knob("n_tile_max", 128)


if __name__ == "__main__":
    main()

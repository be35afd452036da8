"""Hardware probe for the tcgen05 conv engine (run on a B200 via gpurun).

Each variant runs in its own subprocess (a trapped kernel poisons the CUDA context), compares the
engine against an fp32 torch reference on the GPU and prints one line per case.  Usage:
    python tools/probe_engine.py            # driver: all variants -> stdout
    python tools/probe_engine.py --variant KNOBS_JSON
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: (B, L, Cin, N, shifts, l_out)
    "gemm_c32_n32": (2, 300, 32, 32, [0], 300),
    "gemm_c64_n64": (2, 300, 64, 64, [0], 300),
    "gemm_c128_n256": (1, 1000, 128, 256, [0], 1000),
    "gemm_c16_n16": (2, 300, 16, 16, [0], 300),
    "gemm_c208_n512": (1, 257, 208, 512, [0], 257),
    "shift8_c32": (2, 300, 32, 32, [-8, 0], 300),
    "shift1_c32": (2, 300, 32, 32, [-1, 0], 300),
    "k3d1_c32": (2, 300, 32, 32, [-2, -1, 0], 300),
    "k3d1_c64_n64": (3, 700, 64, 64, [-2, -1, 0], 700),
    "k7d3_c128": (2, 1000, 128, 128, [-18, -15, -12, -9, -6, -3, 0], 1000),
    "k11d5_c256": (2, 1500, 256, 256, [-50, -45, -40, -35, -30, -25, -20, -15, -10, -5, 0], 1500),
    "pre_k7_c64_n512": (2, 100, 64, 512, [0, 1, 2, 3, 4, 5, 6], 100),
    "poly3_c64_n320": (2, 500, 64, 320, [1, 0, -1], 500),
}


def reference(a, w, shifts, l_out):
    import torch
    B, L, C = a.shape
    out = torch.zeros(B, l_out, w.shape[1], device=a.device, dtype=torch.float64)
    lo = -min(min(shifts), 0)
    hi = max(max(shifts), 0) + max(l_out - L, 0)
    ap = torch.nn.functional.pad(a.double(), (0, 0, lo, hi))
    for j, s in enumerate(shifts):
        out += ap[:, lo + s: lo + s + l_out, :] @ w[j].double().t()
    return out


def run_variant(knobs):
    from mri2speech_b200 import _lib
    for k, v in knobs.items():
        _lib.set_knob(k, v)
    for case in CASES:
        try:
            run_one(knobs, case)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"case": case, "knobs": knobs, "FAILED": str(exc)[:300]}), flush=True)


def run_one(knobs, case):
    import torch
    from mri2speech_b200 import _lib
    B, L, C, N, shifts, l_out = CASES[case]
    g = torch.Generator(device="cpu").manual_seed(7)
    a = torch.randn(B, L, C, generator=g).cuda()
    w = (torch.randn(len(shifts), N, C, generator=g) / (C * len(shifts)) ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = reference(a, w, shifts, l_out) + bias.double()
    out = {}
    for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_TCGEN05)):
        d = _lib.conv_fwd(a, w, shifts, l_out, impl=impl, bias=bias)
        torch.cuda.synchronize()
        err = (d.double() - ref).abs().max().item()
        out[name] = err
    out["ref_absmax"] = ref.abs().max().item()
    print(json.dumps({"case": case, "knobs": knobs, **out}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--variant":
        run_variant(json.loads(sys.argv[2]))
        return
    variants = [
        {"a_per_tap": 1, "msub": 1},
        {"a_per_tap": 1, "msub": 2},
        {"a_per_tap": 0, "base_offset_mode": 0, "msub": 1},
        {"a_per_tap": 0, "base_offset_mode": 1, "msub": 1},
        {"a_per_tap": 0, "base_offset_mode": 0, "msub": 2},
        {"a_per_tap": 0, "base_offset_mode": 0, "msub": 1, "tmap_tf32": 1},
    ]
    for kn in variants:
        try:
            r = subprocess.run([sys.executable, __file__, "--variant", json.dumps(kn)],
                               capture_output=True, text=True, timeout=240)
            print(r.stdout.strip(), flush=True)
            if r.returncode != 0:
                print(json.dumps({"knobs": kn, "EXIT": r.returncode, "stderr": r.stderr[-600:]}), flush=True)
        except subprocess.TimeoutExpired as exc:
            print(json.dumps({"knobs": kn, "FAILED": "timeout", "stdout": (exc.stdout or b"")[-600:].decode("utf-8", "replace")}), flush=True)


if __name__ == "__main__":
    main()

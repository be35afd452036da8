"""Encoder-only timing (CUDA events) for a few M2S_MBCONV settings: python tools/enc_time.py [frames] [modes...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200.acoustic import build_acoustic_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
modes = sys.argv[2:] or ["0", "1", "2", "3"]
torch.manual_seed(1234)
ac = build_acoustic_model(precision="fp16").cuda().eval()
frames = torch.rand(n, 256, 256, device="cuda")
for mode in modes:
    env = dict(kv.split("=") for kv in mode.split(",") if "=" in kv)
    mb = mode.split(",")[0]
    os.environ["M2S_MBCONV"] = mb
    for k, v in env.items():
        os.environ[k] = v
    ac.refresh()
    for _ in range(2):
        f = ac.encode_frames(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        f = ac.encode_frames(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"M2S_MBCONV={mode}: {ms:.2f} ms / {n} frames = {1e3 * ms / n:.2f} us/frame, launches {ac.launches_per_forward()}, "
          f"feat absmax {float(f.abs().max()):.4f}", flush=True)
    for k in env:
        os.environ.pop(k, None)

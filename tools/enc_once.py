"""One encoder pass per M2S_MBCONV mode (for ncu launch lists): python tools/enc_once.py [frames] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200.acoustic import build_acoustic_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
os.environ["M2S_MBCONV"] = sys.argv[2] if len(sys.argv) > 2 else "3"
torch.manual_seed(1234)
ac = build_acoustic_model(precision="fp16").cuda().eval()
frames = torch.rand(n, 256, 256, device="cuda")
for _ in range(2):
    f = ac.encode_frames(frames)
torch.cuda.synchronize()
print("ok", f.shape, float(f.abs().max()))

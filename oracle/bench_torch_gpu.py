"""Library baseline on the B200: the oracle's torch-functional restatement run ON THE GPU (cuDNN / cuBLAS,
PyTorch's default allow_tf32 for convs) -- BASELINE.md section 4 "second reference point".

TEST / MEASUREMENT INFRASTRUCTURE ONLY (lives under oracle/ on purpose): not imported by the product path.

    python -m oracle.bench_torch_gpu
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mri2speech_b200 import synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.vocoder import Generator
from oracle.acoustic import encoder_forward
from oracle.vocoder import generator_forward


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    h = json.load(open(os.path.join(ROOT, "config_custom.json")))
    torch.manual_seed(1234)
    gen_sd = {k: v.cuda() for k, v in Generator(h).state_dict().items()}
    ac_sd = {k: v.cuda() for k, v in build_acoustic_model().state_dict().items()}
    out = {"allow_tf32_cudnn": torch.backends.cudnn.allow_tf32, "allow_tf32_matmul": torch.backends.cuda.matmul.allow_tf32}
    with torch.no_grad():
        mel = synth.synthetic_mels(32, 256).cuda()
        ms = timed(lambda: generator_forward(gen_sd, h, mel))
        out["vocoder_config2_ms"] = ms
        out["vocoder_config2_audio_s_per_s"] = 32 * 256 * 420 / 11413 / (ms * 1e-3)
        frames = torch.rand(256, 1, 256, 256, device="cuda")
        ms = timed(lambda: encoder_forward(ac_sd, frames))
        out["encoder_256_frames_ms"] = ms
        out["encoder_us_per_frame"] = ms * 1e3 / 256
        torch.backends.cudnn.benchmark = True
        ms = timed(lambda: encoder_forward(ac_sd, frames))
        out["encoder_us_per_frame_cudnn_benchmark"] = ms * 1e3 / 256
        ms = timed(lambda: generator_forward(gen_sd, h, mel))
        out["vocoder_config2_ms_cudnn_benchmark"] = ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()

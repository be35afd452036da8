"""Quick per-stage device timing (CUDA events) of the three model stages on synthetic clips."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib, synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.vocoder import Generator


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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
    h = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config_custom.json")))
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision=prec).cuda().eval()
    gen = Generator(h, precision=prec).cuda().eval()
    frames = torch.rand(B * T, 256, 256, device="cuda")
    feats = torch.randn(B, T, 208, device="cuda") * 0.1
    mel = synth.synthetic_mels(B, T).cuda()
    t_enc = timed(lambda: ac.encode_frames(frames))
    _lib.profile(True)
    ac.encode_frames(frames)
    ms, fl = _lib.profile_read()
    _lib.profile(False)
    t_rnn = timed(lambda: ac.rnn_head(feats))
    t_voc = timed(lambda: gen(mel))
    audio_s = B * T * 420 / 11413
    print(json.dumps({"precision": prec, "B": B, "T": T, "audio_s": audio_s, "encoder_ms": t_enc, "encoder_engine_ms": sum(ms),
                      "encoder_engine_launches": len(ms), "encoder_engine_tflops": sum(fl) / (sum(ms) * 1e-3) / 1e12,
                      "rnn_head_ms": t_rnn, "vocoder_ms": t_voc,
                      "e2e_audio_s_per_s": audio_s / ((t_enc + t_rnn + t_voc) * 1e-3),
                      "us_per_frame_encoder": t_enc * 1e3 / (B * T)}))
    per = len(ms) // max(1, (B * T + 255) // 256)
    print("engine launches of the first chunk (us, TF/s):", [(i, round(ms[i] * 1e3), round(fl[i] / ms[i] / 1e9)) for i in range(min(per, 60))])
    # slowest engine launches of the encoder
    order = sorted(range(len(ms)), key=lambda i: -ms[i])[:12]
    print("top encoder engine launches (idx, ms, TF/s):", [(i, round(ms[i], 3), round(fl[i] / ms[i] / 1e9, 1)) for i in order])


if __name__ == "__main__":
    main()

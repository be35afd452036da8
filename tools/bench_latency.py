"""Single-clip latency (BASELINE.json configs[0]: one 150-frame clip, B = 1): eager launches vs CUDA-graph replay.

    python tools/bench_latency.py [frames=150] [precision=fp16]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.graphs import graph_acoustic, graph_generator
from mri2speech_b200.vocoder import Generator


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    h = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config_custom.json")))
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision=prec).cuda().eval()
    gen = Generator(h, precision=prec).cuda().eval()
    clip = synth.synthetic_clip_u8(0, T).unsqueeze(0).cuda()
    mel = synth.synthetic_mels(1, T).cuda()
    with torch.no_grad():
        res = {"frames": T, "precision": prec, "audio_s": T * 420 / 11413,
               "acoustic_eager_ms": timed(lambda: ac(clip)), "vocoder_eager_ms": timed(lambda: gen(mel))}
        ga, gg = graph_acoustic(ac, clip), graph_generator(gen, mel)
        res["acoustic_graph_ms"] = timed(lambda: ga(clip))
        res["vocoder_graph_ms"] = timed(lambda: gg(mel))
    res["eager_total_ms"] = res["acoustic_eager_ms"] + res["vocoder_eager_ms"]
    res["graph_total_ms"] = res["acoustic_graph_ms"] + res["vocoder_graph_ms"]
    res["realtime_factor_graph"] = res["audio_s"] / (res["graph_total_ms"] * 1e-3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

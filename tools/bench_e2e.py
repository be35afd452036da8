"""End-to-end rtMRI -> mel -> wav throughput on ragged synthetic clips (BASELINE.json configs[2]-like).

    python tools/bench_e2e.py [n_clips=64] [max_batch_frames=4096]

Frames are generated on the device (uniform noise, per-frame min-max like the reference's loader) -- this
measures the device-resident pipeline (MriToSpeech.infer), device timing with CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.pipeline import MriToSpeech
from mri2speech_b200.vocoder import Generator


def main():
    n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mbf = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    h = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config_custom.json")))
    torch.manual_seed(1234)
    ac = build_acoustic_model()
    gen = Generator(h)
    mean, std = synth.synthetic_scaler()
    pipe = MriToSpeech(ac, gen, mean, std)
    lens = synth.synthetic_lengths(n_clips)
    g = torch.Generator(device="cuda").manual_seed(1)
    clips = []
    for ln in lens:
        x = torch.rand(ln, 256, 256, device="cuda", generator=g)
        mn, mx = x.amin((1, 2), keepdim=True), x.amax((1, 2), keepdim=True)
        clips.append((x - mn) / (mx - mn))
    frames = sum(lens)
    audio_s = frames * 420 / 11413
    pipe.infer(clips[:4], max_batch_frames=mbf)
    torch.cuda.synchronize()
    best = 1e18
    for _ in range(2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        out = pipe.infer(clips, max_batch_frames=mbf)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert all(o["audio"].shape[0] == ln * 420 for o, ln in zip(out, lens))
    print(json.dumps({"workload": f"e2e ragged: {n_clips} clips, {frames} frames ({audio_s:.1f} s audio), "
                                  f"max_batch_frames={mbf}", "ms": best, "audio_s_per_s": audio_s / (best * 1e-3),
                      "us_per_frame": best * 1e3 / frames,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()

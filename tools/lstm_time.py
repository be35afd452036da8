"""Device time of rnn_head (input projection + BiLSTM recurrence + head) on a configs[2]-shaped ragged batch.
Usage: python tools/lstm_time.py [B] [Tmax] [precision]   (M2S_LSTM_CLUSTER=0 selects the grid-barrier kernel)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri2speech_b200 import _lib
from mri2speech_b200.acoustic import build_acoustic_model


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision=prec).cuda().eval()
    g = torch.Generator().manual_seed(4321)
    lens = torch.randint(min(150, T), T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T
    feats = torch.randn(B, T, 208, device="cuda") * 0.5
    out = None
    for ragged in (True, False):
        args = (feats, lens) if ragged else (feats,)
        for _ in range(3):
            out = ac.rnn_head(*args)
        torch.cuda.synchronize()
        _lib.profile(True)
        ac.rnn_head(*args)
        ms, fl = _lib.profile_read()
        _lib.profile(False)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ac.rnn_head(*args)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"B": B, "T": T, "ragged": ragged, "precision": prec, "rnn_head_ms": best,
                          "launch_ms(inproj, recurrence, head)": [round(x, 4) for x in ms],
                          "cluster": os.environ.get("M2S_LSTM_CLUSTER", "1"),
                          "us_per_step": ms[1] * 1e3 / T if len(ms) > 1 else None,
                          "checksum": float(out.double().abs().sum())}))


if __name__ == "__main__":
    main()

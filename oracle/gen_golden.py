"""Generate the committed golden vectors under tests/golden/ (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Vocoder goldens come from the REFERENCE's own ``models.Generator``
(/root/reference/models.py, imported via oracle/ref_import.py) with its seeded default init
(seed 1234 = config_custom.json:9); the product re-creates identical weights from the same seed, so
only inputs and outputs are stored.  Acoustic-model goldens come from the oracle restatement
(oracle/acoustic.py; encoder parity is UNPINNED, see oracle/__init__.py).

    python -m oracle.gen_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def vocoder_goldens():
    from oracle import ref_import
    from oracle.vocoder import generator_forward
    g_ref, h = ref_import.reference_generator(1234)
    gen = torch.Generator().manual_seed(2024)
    mel = torch.randn(2, 64, 24, generator=gen) * 2.0 - 5.0
    with torch.no_grad():
        wav = g_ref(mel)
    np.savez_compressed(os.path.join(GOLDEN, "vocoder_ref_seed1234_b2_t24.npz"),
                        mel=mel.numpy(), wav=wav.numpy())
    # ragged: each utterance through the reference at B=1 on its own valid frames
    lens = [24, 13]
    outs = []
    with torch.no_grad():
        for b, ln in enumerate(lens):
            outs.append(g_ref(mel[b:b + 1, :, :ln])[0, 0].numpy())
    np.savez_compressed(os.path.join(GOLDEN, "vocoder_ref_seed1234_ragged.npz"),
                        mel=mel.numpy(), lens=np.asarray(lens, np.int32), wav0=outs[0], wav1=outs[1])
    # cross-check the restatement while we are here
    y = generator_forward(g_ref.state_dict(), h, mel)
    assert (y - wav).abs().max().item() < 1e-6
    print("vocoder goldens written")


def acoustic_goldens():
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from oracle.acoustic import acoustic_forward, encoder_forward
    from oracle.glue import mel_glue
    torch.manual_seed(1234)
    model = build_acoustic_model()
    sd = model.state_dict()
    clip = synth.synthetic_clip(0, 6)  # (T, 256, 256)
    with torch.no_grad():
        feats = encoder_forward(sd, clip.unsqueeze(1))
        mel = acoustic_forward(sd, clip.unsqueeze(0).unsqueeze(2))
    mean, std = synth.synthetic_scaler()
    mel_db, mel_log, _ = mel_glue(mel[0], mean, std)
    np.savez_compressed(os.path.join(GOLDEN, "acoustic_oracle_seed1234_clip0_t6.npz"),
                        feats=feats.numpy(), mel_norm=mel.numpy(), mel_db=mel_db.numpy(), mel_log=mel_log.numpy())
    print("acoustic goldens written")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    vocoder_goldens()
    if "--vocoder-only" not in sys.argv:
        acoustic_goldens()

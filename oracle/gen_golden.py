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
    # config with "resblock": "2" (models.py:58-80, :95): same seed, ResBlock2 branches with dilations (1, 3)
    g2, h2 = ref_import.reference_generator(1234, resblock="2", resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])
    lens = [24, 13]
    with torch.no_grad():
        wav2 = g2(mel)
        solo = g2(mel[1:2, :, :13])[0, 0]
    np.savez_compressed(os.path.join(GOLDEN, "vocoder_ref_seed1234_resblock2.npz"),
                        mel=mel.numpy(), wav=wav2.numpy(), lens=np.asarray(lens, np.int32), wav1_ragged=solo.numpy())
    y2 = generator_forward(g2.state_dict(), h2, mel)
    assert (y2 - wav2).abs().max().item() < 1e-6
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


def _reference_preprocess_frame():
    """``_preprocess_frame`` exactly as the reference defines it (scripts/run_mri_video_inference.py:34-53).
    The script itself cannot be imported (soundfile / matplotlib at module import), so the function is cut out
    of the file with ``ast`` and executed with cv2 / numpy in scope."""
    import ast
    import cv2
    path = os.path.join("/root/reference", "scripts", "run_mri_video_inference.py")
    src = open(path, "r", encoding="utf-8").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "_preprocess_frame")
    scope = {"np": np, "cv2": cv2}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), scope)
    return scope["_preprocess_frame"]


def _reference_mask_module():
    import importlib.util
    path = os.path.join("/root/reference", "scripts", "mask_rtmri_video.py")
    spec = importlib.util.spec_from_file_location("_ref_mask_rtmri_video", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod  # dataclass needs the module registered
    spec.loader.exec_module(mod)
    return mod


def ingest_goldens():
    """Outputs of the REFERENCE's frame normalisation and mask construction on synthetic uint8 frames."""
    from mri2speech_b200 import synth
    from oracle import ingest
    pre = _reference_preprocess_frame()
    ref_mask = _reference_mask_module()
    clip = synth.synthetic_clip_u8(7, 3).numpy()
    clip[2] = 93  # constant frame -> zeros (:50-53)
    norm = np.stack([pre(f) for f in clip]).astype(np.float32)
    out = {"clip_u8": clip, "norm": norm}
    presets = {"lip": ref_mask.LIP_MASK, "tongue": ref_mask.TONGUE_MASK}
    for name, preset in presets.items():
        for alpha in (0.0, 0.3, 1.0):
            m = ref_mask.build_mask((256, 256), preset.scaled((256, 256)), alpha, 11)
            out[f"mask_{name}_{alpha}"] = m.astype(np.float32)
    m = out["mask_tongue_0.3"]
    masked = (clip.astype(np.float32) * m[None]).clip(0.0, 255.0).astype(np.uint8)  # mask_rtmri_video.py:96-98
    out["masked_tongue_0.3"] = masked
    out["masked_norm_tongue_0.3"] = np.stack([pre(f) for f in masked]).astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "ingest_ref.npz"), **out)
    # cross-check the restatement against the reference while we are here
    assert np.array_equal(ingest.preprocess_clip(clip), norm)
    assert np.array_equal(ingest.apply_mask(clip, m), masked)
    print("ingest goldens written")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    if "--ingest-only" in sys.argv:
        ingest_goldens()
        sys.exit(0)
    vocoder_goldens()
    if "--vocoder-only" not in sys.argv:
        acoustic_goldens()
        ingest_goldens()

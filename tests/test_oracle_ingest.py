"""CPU: the ingest / masking oracle against the reference-generated golden vectors (tests/golden/ingest_ref.npz,
made by oracle/gen_golden.py from the reference's own _preprocess_frame and mask_rtmri_video.build_mask), and the
product's host-side mask construction against the same vectors."""
import os

import numpy as np
import pytest

from oracle import ingest
from tests.util import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ingest_ref.npz"))


def test_preprocess_matches_reference(gold):
    out = ingest.preprocess_clip(gold["clip_u8"])
    assert np.array_equal(out, gold["norm"])
    assert out[2].max() == 0.0  # constant frame -> zeros (run_mri_video_inference.py:50-53)
    assert out[:2].min() == 0.0 and out[:2].max() == 1.0


def test_zscore_then_minmax_is_plain_minmax(gold):
    """The device ingest computes (u - min) / (max - min) directly; the reference's z-score first is an affine map."""
    clip = gold["clip_u8"][:2].astype(np.float32)
    lo = clip.min(axis=(1, 2), keepdims=True)
    hi = clip.max(axis=(1, 2), keepdims=True)
    assert np.abs((clip - lo) / (hi - lo) - gold["norm"][:2]).max() < 5e-7


def test_apply_mask_matches_reference(gold):
    m = gold["mask_tongue_0.3"]
    assert np.array_equal(ingest.apply_mask(gold["clip_u8"], m), gold["masked_tongue_0.3"])
    assert np.array_equal(ingest.preprocess_clip(gold["masked_tongue_0.3"]), gold["masked_norm_tongue_0.3"])


@pytest.mark.parametrize("name", ["lip", "tongue"])
@pytest.mark.parametrize("alpha", [0.0, 0.3, 1.0])
def test_build_mask_oracle_and_product(gold, name, alpha):
    from mri2speech_b200 import masking
    ref = gold[f"mask_{name}_{alpha}"]
    preset = masking.PRESETS[name]
    poly = preset.scaled((256, 256))
    assert np.array_equal(ingest.build_mask((256, 256), poly, alpha, 11), ref)
    assert np.array_equal(masking.preset_mask(name, alpha), ref)
    assert ref.min() >= alpha and ref.max() <= 1.0
    if alpha == 1.0:
        assert np.all(ref == 1.0)  # identity mask: the sweep runs it once


def test_sweep_alphas():
    from mri2speech_b200 import masking
    a = masking.sweep_alphas()
    assert len(a) == 11 and a[0] == 0.0 and a[-1] == 1.0 and abs(a[3] - 0.3) < 1e-12
    with pytest.raises(KeyError):
        masking.preset_mask("jaw", 0.5)


def test_cli_frame_helpers_match_reference_preprocessing(gold):
    """scripts/run_mri_video_inference.py keeps the reference's helper names (_preprocess_frame, frames_to_tensor);
    its normalisation must reproduce the reference's own _preprocess_frame output (the golden), constant frames and
    already-sized gray frames included."""
    import sys
    import torch
    sys.path.insert(0, os.path.normpath(os.path.join(GOLDEN, "..", "..", "scripts")))
    import run_mri_video_inference as cli
    frames = gold["clip_u8"]
    ref = gold["norm"]            # the reference's own _preprocess_frame, frame by frame (oracle/gen_golden.py)
    try:
        import cv2  # noqa: F401  (_preprocess_frame converts / resizes with OpenCV before normalising)
        helpers = (cli._normalise, cli._preprocess_frame)
    except ImportError:
        helpers = (cli._normalise,)
    for i in range(frames.shape[0]):
        for fn in helpers:
            got = fn(frames[i])
            assert got.dtype == np.float32 and np.array_equal(got, ref[i])
    flat = np.full((256, 256), 37, np.uint8)
    assert np.array_equal(cli._normalise(flat), np.zeros((256, 256), np.float32))
    t = cli.frames_to_tensor(torch.from_numpy(frames[:3]))
    assert tuple(t.shape) == (1, 3, 1, 256, 256)
    assert tuple(cli.frames_to_tensor(torch.from_numpy(frames[:3]), use_channel=False).shape) == (1, 3, 256, 256)
    with pytest.raises(ValueError):
        cli.frames_to_tensor(torch.zeros(4, 4))

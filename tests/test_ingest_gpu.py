"""GPU: fused uint8 ingest (per-frame min-max in the stem load) and in-memory articulator masking against the
reference-generated golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ingest_ref.npz"))


def _model(precision):
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    m = build_acoustic_model(precision=precision)
    synth.randomize_batchnorm(m)
    return m.cuda().eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 3e-4), ("fp16", 3e-4)])
def test_u8_ingest_equals_reference_preprocessing(gold, precision, tol):
    """forward(uint8 frames) == forward(frames normalised by the reference's _preprocess_frame)."""
    m = _model(precision)
    u8 = torch.from_numpy(gold["clip_u8"]).cuda().unsqueeze(0)
    ref_in = torch.from_numpy(gold["norm"]).cuda().unsqueeze(0)
    with torch.no_grad():
        a = m(u8)
        b = m(ref_in)
    assert a.shape == (1, 3, 64)
    assert (a - b).abs().max().item() < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 3e-4), ("fp16", 3e-4)])
def test_masked_ingest_equals_reference_masking(gold, precision, tol):
    """forward(uint8, mask) == forward(reference-normalised frames of the reference-masked clip)."""
    m = _model(precision)
    u8 = torch.from_numpy(gold["clip_u8"]).cuda().unsqueeze(0)
    mask = torch.from_numpy(gold["mask_tongue_0.3"])
    ref_in = torch.from_numpy(gold["masked_norm_tongue_0.3"]).cuda().unsqueeze(0)
    with torch.no_grad():
        a = m(u8, mask=mask)
        b = m(ref_in)
        c = m(u8)
    assert (a - b).abs().max().item() < tol
    assert (a - c).abs().max().item() > 10 * tol  # the mask does change the prediction


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_masked_ragged_batch_against_oracle(precision):
    """Ragged uint8 batch with a mask through the pipeline == oracle (mask -> normalise -> model) per clip."""
    from mri2speech_b200 import masking, synth
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.vocoder import Generator
    from oracle import ingest
    from oracle.acoustic import acoustic_forward
    from tests.util import load_config
    torch.manual_seed(1234)
    gen = Generator(load_config(), precision=precision)
    ac = _model(precision)
    mean, std = synth.synthetic_scaler()
    clips = [synth.synthetic_clip_u8(3, 5), synth.synthetic_clip_u8(4, 3)]
    mask = masking.preset_mask("lip", 0.2)
    pipe = MriToSpeech(ac, gen, mean, std)
    out = pipe.infer(clips, mask=torch.from_numpy(mask))
    sd = {k: v.detach().cpu() for k, v in ac.state_dict().items()}
    for clip, o in zip(clips, out):
        x = torch.from_numpy(ingest.preprocess_clip(ingest.apply_mask(clip.numpy(), mask)))
        ref = acoustic_forward(sd, x.unsqueeze(0).unsqueeze(2))[0]
        assert (o["mel_norm"].cpu() - ref).abs().max().item() < 1e-3
        assert o["audio"].shape[0] == clip.shape[0] * 420


def test_mask_argument_errors():
    m = _model("tf32")
    f32 = torch.rand(1, 2, 256, 256, device="cuda")
    with pytest.raises(ValueError):
        m(f32, mask=torch.ones(256, 256))
    u8 = torch.zeros(1, 2, 256, 256, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        m(u8, mask=torch.ones(128, 256))
    out = m(u8)  # constant frames -> zeros in, finite out
    assert torch.isfinite(out).all()

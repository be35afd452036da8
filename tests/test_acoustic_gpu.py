"""OTNLikeCNNBiLSTM on the B200 path vs the CPU oracle (oracle/acoustic.py).

Gate (BASELINE.json north_star): normalised mel within 1e-3 max-abs of the fp32 CPU path.  Random-init
weights make features / mel tiny, which flatters an absolute gate, so relative errors
against the reference abs-max are asserted as well, and a BatchNorm-randomised variant is run too.
The encoder oracle restates timm's tf_efficientnetv2_b2 (absent dependency): parity UNPINNED."""
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN

pytestmark = pytest.mark.gpu


def _model(precision, randomize_bn=False):
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    torch.manual_seed(1234)
    m = build_acoustic_model(precision=precision)
    if randomize_bn:
        synth.randomize_batchnorm(m)
    return m.cuda().eval()


def _cpu_sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 1e-3)])
def test_bilstm_head_ragged(precision, tol):
    from oracle.acoustic import bilstm_head_forward
    m = _model(precision)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(3, 40, 208, generator=g) * 0.5
    lens = [40, 7, 23]
    got = m.rnn_head(feats.cuda(), torch.tensor(lens, dtype=torch.int32)).cpu()
    sd = _cpu_sd(m)
    for b, ln in enumerate(lens):
        ref = bilstm_head_forward(sd, feats[b:b + 1, :ln])[0]
        err = (got[b, :ln] - ref).abs().max().item()
        assert err < tol, (b, err)
        if ln < 40:
            assert got[b, ln:].abs().max().item() == 0.0


def test_bilstm_long_sequence_t600():
    """600 recurrent steps (the longest clip of config 3), B=2."""
    from oracle.acoustic import bilstm_head_forward
    m = _model("tf32")
    feats = torch.randn(2, 600, 208, generator=torch.Generator().manual_seed(8)) * 0.5
    got = m.rnn_head(feats.cuda()).cpu()
    ref = bilstm_head_forward(_cpu_sd(m), feats)
    assert (got - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
@pytest.mark.parametrize("B,T", [(1, 37), (9, 30), (20, 48), (33, 21), (64, 16), (100, 9)])
def test_bilstm_cluster_groups(precision, B, T):
    """The cluster recurrence (lstm_cluster_sm100.cu): one 16-CTA cluster per (direction, group of up to 8 / 16 / 24
    utterances).  The batch sizes walk through the one-, two- and three-tile kernels, partly filled groups, and more groups
    than the device holds at once (100 clips); lengths include 1 and T."""
    from oracle.acoustic import bilstm_head_forward
    m = _model(precision)
    g = torch.Generator().manual_seed(100 + B)
    feats = torch.randn(B, T, 208, generator=g) * 0.5
    lens = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T
    if B > 2:
        lens[B - 1] = 1
    got = m.rnn_head(feats.cuda(), lens).cpu()
    sd = _cpu_sd(m)
    for b in range(B):
        ln = int(lens[b])
        ref = bilstm_head_forward(sd, feats[b:b + 1, :ln])[0]
        err = (got[b, :ln] - ref).abs().max().item()
        assert err < 1e-3, (b, ln, err)
        if ln < T:
            assert got[b, ln:].abs().max().item() == 0.0


@pytest.mark.parametrize("hidden", [256, 672])
def test_bilstm_other_hidden_sizes(hidden):
    """--rnn-hidden other than the reference's 640 (any multiple of 32): the generic cooperative recurrence of
    csrc/lstm_sm100.cu (a warp per hidden unit, fp32), with the in-projection and the head on the engine as usual."""
    from mri2speech_b200.acoustic import build_acoustic_model
    from oracle.acoustic import bilstm_head_forward
    torch.manual_seed(1234)
    m = build_acoustic_model(rnn_hidden=hidden, precision="fp16").cuda().eval()
    g = torch.Generator().manual_seed(hidden)
    B, T = 11, 29
    feats = torch.randn(B, T, 208, generator=g) * 0.5
    lens = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T
    got = m.rnn_head(feats.cuda(), lens).cpu()
    sd = _cpu_sd(m)
    for b in range(B):
        ln = int(lens[b])
        ref = bilstm_head_forward(sd, feats[b:b + 1, :ln])[0]
        err = (got[b, :ln] - ref).abs().max().item()
        assert err < 1e-3, (hidden, b, ln, err)
        if ln < T:
            assert got[b, ln:].abs().max().item() == 0.0


def test_bilstm_grid_barrier_fallback():
    """M2S_LSTM_CLUSTER=0 (read once per process, hence the subprocess): the tensor-core grid-barrier recurrence of
    lstm_sm100.cu, which is what a device without 16-CTA clusters gets, against torch.nn.LSTM on a ragged batch."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from mri2speech_b200.acoustic import build_acoustic_model\n"
        "from oracle.acoustic import bilstm_head_forward\n"
        "torch.manual_seed(1234)\n"
        "m = build_acoustic_model(precision='fp16').cuda().eval()\n"
        "g = torch.Generator().manual_seed(77)\n"
        "feats = torch.randn(20, 40, 208, generator=g) * 0.5\n"
        "lens = torch.randint(1, 41, (20,), generator=g, dtype=torch.int32); lens[0] = 40\n"
        "got = m.rnn_head(feats.cuda(), lens).cpu()\n"
        "sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}\n"
        "err = max((got[b, :int(lens[b])] - bilstm_head_forward(sd, feats[b:b + 1, :int(lens[b])])[0]).abs().max().item() for b in range(20))\n"
        "print('ERR', err)\n"
        "assert err < 1e-3, err\n")
    env = dict(os.environ, M2S_LSTM_CLUSTER="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ERR" in out.stdout


@pytest.mark.parametrize("precision,rel_tol", [("fp32", 2e-4), ("tf32", 5e-3), ("fp16", 5e-3)])
@pytest.mark.parametrize("randomize_bn", [False, True])
def test_encoder_features(precision, rel_tol, randomize_bn):
    from mri2speech_b200 import synth
    from oracle.acoustic import encoder_forward
    m = _model(precision, randomize_bn)
    clip = synth.synthetic_clip(3, 5)
    with torch.no_grad():
        ref = encoder_forward(_cpu_sd(m), clip.unsqueeze(1))
    got = m.encode_frames(clip.cuda()).cpu()
    assert got.shape == (5, 208)
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"[{precision} bn_random={randomize_bn}] feature rel err {rel:.2e} (abs-max {ref.abs().max():.3e})")
    assert rel < rel_tol


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp16"])
def test_forward_ragged_vs_oracle(precision):
    from mri2speech_b200 import synth
    from oracle.acoustic import acoustic_forward
    m = _model(precision, randomize_bn=True)
    clips = torch.stack([synth.synthetic_clip(0, 6), synth.synthetic_clip(1, 6)])  # (2,6,256,256)
    lens = [6, 4]
    with torch.no_grad():
        ref = acoustic_forward(_cpu_sd(m), clips.unsqueeze(2), lengths=lens)
        got = m(clips.unsqueeze(2).cuda(), lengths=torch.tensor(lens, dtype=torch.int32)).cpu()
    err = (got - ref).abs().max().item()
    print(f"[{precision}] mel max-abs err {err:.2e} (mel abs-max {ref.abs().max():.3e})")
    assert err < (1e-4 if precision == "fp32" else 1e-3)
    assert err / ref.abs().max().item() < (1e-3 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_golden_acoustic_fixture(precision):
    from mri2speech_b200 import synth
    z = np.load(os.path.join(GOLDEN, "acoustic_oracle_seed1234_clip0_t6.npz"))
    m = _model(precision)
    clip = synth.synthetic_clip(0, 6)
    with torch.no_grad():
        mel = m(clip.unsqueeze(0).unsqueeze(2).cuda()).cpu().numpy()
        feats = m.encode_frames(clip.cuda()).cpu().numpy()
    assert np.abs(mel - z["mel_norm"]).max() < 1e-3
    assert np.abs(feats - z["feats"]).max() / np.abs(z["feats"]).max() < 5e-3


def test_training_mode_and_cpu_refused():
    from mri2speech_b200._lib import M2SError
    m = _model("tf32")
    with pytest.raises(M2SError):
        m(torch.zeros(1, 2, 1, 256, 256))
    m.train()
    with pytest.raises(M2SError):
        m(torch.zeros(1, 2, 1, 256, 256).cuda())

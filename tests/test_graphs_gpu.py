"""GPU: CUDA-graph replay of the fixed-shape forwards equals the eager call bit for bit (same kernels, same order)."""
import pytest
import torch

from tests.util import load_config

pytestmark = pytest.mark.gpu


def test_generator_graph_replay_matches_eager():
    from mri2speech_b200.graphs import graph_generator
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    gen = Generator(load_config(), precision="fp16").cuda().eval()
    g = torch.Generator().manual_seed(5)
    mel_a = (torch.randn(1, 64, 150, generator=g) * 2 - 5).cuda()
    mel_b = (torch.randn(1, 64, 150, generator=g) * 2 - 5).cuda()
    graphed = graph_generator(gen, mel_a)
    with torch.no_grad():
        ref_a, ref_b = gen(mel_a).clone(), gen(mel_b).clone()
    assert torch.equal(graphed(mel_a).clone(), ref_a)
    assert torch.equal(graphed(mel_b).clone(), ref_b)
    assert torch.equal(graphed(mel_a), ref_a)                       # replays are repeatable
    with pytest.raises(ValueError):
        graphed(torch.zeros(1, 64, 151, device="cuda"))


def test_acoustic_graph_replay_matches_eager():
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from mri2speech_b200.graphs import graph_acoustic
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision="fp16").cuda().eval()
    synth.randomize_batchnorm(ac)
    clip = synth.synthetic_clip_u8(2, 6).unsqueeze(0).cuda()
    other = synth.synthetic_clip_u8(3, 6).unsqueeze(0).cuda()
    graphed = graph_acoustic(ac, clip)
    with torch.no_grad():
        ref, ref2 = ac(clip).clone(), ac(other).clone()
    assert torch.equal(graphed(clip).clone(), ref)
    assert torch.equal(graphed(other).clone(), ref2)


def test_graph_survives_plan_changes_and_pins_the_workspace():
    """a captured graph holds raw pointers into the module's workspace and the handle's packed weights.
    A larger eager forward must not free the captured workspace; a parameter update (which rebuilds the handle) must
    lead to a re-capture, never to a replay on freed memory."""
    from mri2speech_b200._lib import M2SError
    from mri2speech_b200.graphs import graph_generator
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    gen = Generator(load_config(), precision="fp16").cuda().eval()
    mel = (torch.randn(1, 64, 40, generator=torch.Generator().manual_seed(5)) * 2 - 5).cuda()
    graphed = graph_generator(gen, mel)
    with torch.no_grad():
        ref = gen(mel).clone()
        with pytest.raises(M2SError):                      # would have to grow the workspace the graph points into
            gen(torch.zeros(2, 64, 200, device="cuda"))
    assert torch.equal(graphed(mel), ref)                  # ... and the graph is intact
    with torch.no_grad():
        gen.conv_pre.bias.add_(0.25)                       # plan change: the handle is rebuilt on the next forward
        ref2 = gen(mel).clone()
    assert not torch.equal(ref2, ref)
    out2 = graphed(mel).clone()
    assert graphed.captures == 2 and torch.equal(out2, ref2)
    graphed.release()
    with torch.no_grad():
        assert gen(torch.zeros(2, 64, 200, device="cuda")).shape == (2, 1, 200 * 420)   # free to grow again

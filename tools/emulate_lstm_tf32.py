"""CPU emulation of the BiLSTM recurrence with h_{t-1} and W_hh rounded to tf32 every step (what the tensor-core step of
csrc/lstm_sm100.cu does), against exact fp32, on default-init and scaled-init weights: python tools/emulate_lstm_tf32.py"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from mri2speech_b200.acoustic import build_acoustic_model
from oracle.scaled_init import calibration_frames, scale_acoustic

def tf32(x):
    u = x.contiguous().view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF
    return u.view(torch.float32)

def lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse, rnd):
    B,T,_ = x.shape; H = w_hh.shape[1]
    gin = x @ w_ih.t() + b_ih + b_hh
    h = torch.zeros(B,H); c = torch.zeros(B,H); out = torch.zeros(B,T,H)
    W = rnd(w_hh)
    rng = range(T-1,-1,-1) if reverse else range(T)
    for t in rng:
        z = gin[:,t] + rnd(h) @ W.t()
        i,f,g,o = z.chunk(4,1)
        c = torch.sigmoid(f)*c + torch.sigmoid(i)*torch.tanh(g)
        h = torch.sigmoid(o)*torch.tanh(c)
        out[:,t] = h
    return out

def head(sd, feats, rnd):
    p="rnn.lstm."
    f = lstm_dir(feats, sd[p+"weight_ih_l0"], sd[p+"weight_hh_l0"], sd[p+"bias_ih_l0"], sd[p+"bias_hh_l0"], False, rnd)
    b = lstm_dir(feats, sd[p+"weight_ih_l0_reverse"], sd[p+"weight_hh_l0_reverse"], sd[p+"bias_ih_l0_reverse"], sd[p+"bias_hh_l0_reverse"], True, rnd)
    return (f+b) @ sd["head.weight"].t() + sd["head.bias"]

torch.manual_seed(1234)
for name in ("default","scaled"):
    torch.manual_seed(1234)
    m = build_acoustic_model()
    if name=="scaled": scale_acoustic(m, calibration_frames(8))
    sd = {k:v.detach().clone() for k,v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    scale = 0.05 if name=="default" else 3.0   # feature magnitudes seen in the two regimes
    feats = torch.randn(2, 400, 208, generator=g).abs() * scale
    with torch.no_grad():
        ref = head(sd, feats, lambda x: x)
        got = head(sd, feats, tf32)
    print(name, "mel abs-max %.3f, tf32-recurrence max-abs err %.2e" % (float(ref.abs().max()), float((got-ref).abs().max())))

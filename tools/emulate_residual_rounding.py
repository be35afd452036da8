"""CPU emulation of the fp16 build's rounding points in the frame-CNN encoder (oracle topology, BN folded):
  mode 1 = fp16 GEMM operands + fp32 residual stream (the round-1 data flow), mode 2 = the residual stream in fp16 too.
Reproduces the errors measured on the B200 (scaled init: features 6.8e-3 relative, mel 5.2e-2) and shows what the fp16
residual stream adds (default init: nothing measurable; scaled init: +18 % / +7 %): python tools/emulate_residual_rounding.py"""
import sys, math, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import acoustic as A
from oracle.acoustic import STAGES, STEM_CH, BN_EPS, _conv_same
from mri2speech_b200 import synth
from mri2speech_b200.acoustic import build_acoustic_model

def q(x): return x.half().float()

def fold(sd, conv, bn):
    w = sd[conv + ".weight"]
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
    t = sd[bn + ".bias"] - sd[bn + ".running_mean"] * s
    return w * s.view(-1,1,1,1), t

def enc(sd, frames, mode):
    # mode 0: exact fp32; 1: fp16 operands, fp32 residual stream (current build); 2: fp16 residual stream too
    prefix="cnn.backbone."
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    R = (lambda x: x) if mode == 0 else q
    x = frames.unsqueeze(1).repeat(1,3,1,1)
    w,t = fold(sd,"conv_stem","bn1")
    x = F.silu(_conv_same(x, w, 2) + t.view(1,-1,1,1))
    cin = STEM_CH
    for s,(kind,reps,stride,_e,cout,_se) in enumerate(STAGES):
        for b in range(reps):
            p=f"blocks.{s}.{b}"; st = stride if b==0 else 1
            skip = (st==1 and cin==cout)
            inp = x
            xo = R(x)   # operand copy
            if mode == 2: inp = xo
            if kind=="cn":
                w,t=fold(sd,p+".conv",p+".bn1"); x=F.silu(_conv_same(xo,R(w),st)+t.view(1,-1,1,1))
            elif kind=="er":
                w,t=fold(sd,p+".conv_exp",p+".bn1"); e=R(F.silu(_conv_same(xo,R(w),st)+t.view(1,-1,1,1)))
                w,t=fold(sd,p+".conv_pwl",p+".bn2"); x=F.conv2d(e,R(w))+t.view(1,-1,1,1)
            else:
                w,t=fold(sd,p+".conv_pw",p+".bn1"); e=R(F.silu(F.conv2d(xo,R(w))+t.view(1,-1,1,1)))
                w,t=fold(sd,p+".conv_dw",p+".bn2"); d=F.silu(_conv_same(e,w,st,groups=e.shape[1])+t.view(1,-1,1,1))
                se=d.mean((2,3),keepdim=True); d=R(d)
                se=F.silu(F.conv2d(se,sd[p+".se.conv_reduce.weight"],sd[p+".se.conv_reduce.bias"]))
                se=torch.sigmoid(F.conv2d(se,sd[p+".se.conv_expand.weight"],sd[p+".se.conv_expand.bias"]))
                d=R(d*se)
                w,t=fold(sd,p+".conv_pwl",p+".bn3"); x=F.conv2d(d,R(w))+t.view(1,-1,1,1)
            if skip: x = x + inp
            cin=cout
    return x.mean(dim=(2,3))

def run(name, m):
    sd={k:v.detach().clone() for k,v in m.state_dict().items()}
    clip=synth.synthetic_clip(0,8)
    with torch.no_grad():
        f0=enc(sd,clip,0); f1=enc(sd,clip,1); f2=enc(sd,clip,2)
        ref=A.encoder_forward(sd,clip.unsqueeze(1))
        mel=[A.bilstm_head_forward(sd,f[None])[0] for f in (f0,f1,f2)]
    sc=f0.abs().max()
    print(name,"folded-exact vs oracle",float((f0-ref).abs().max()/sc))
    print(name,"feat rel err: fp16 operands %.2e ; + fp16 residual stream %.2e"%(float((f1-f0).abs().max()/sc),float((f2-f0).abs().max()/sc)))
    print(name,"mel max-abs err: %.2e ; %.2e   (mel absmax %.2f)"%(float((mel[1]-mel[0]).abs().max()),float((mel[2]-mel[0]).abs().max()),float(mel[0].abs().max())))

torch.manual_seed(1234)
m=build_acoustic_model(); synth.randomize_batchnorm(m); run("default-init+BN", m.eval())
from oracle.scaled_init import calibration_frames, scale_acoustic
torch.manual_seed(1234)
m=build_acoustic_model(); scale_acoustic(m, calibration_frames(8)); run("scaled-init", m.eval())

"""CPU: libm2s.so loads and exports every symbol include/m2s.h declares; host-side error behaviour."""
import os
import re

import pytest
import torch

from tests.util import ROOT, load_config


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "m2s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(m2s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mri2speech_b200 import _lib
    lib = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/m2s.h but not exported"
    assert b"sm_100a" in lib.m2s_version()
    for s in _lib.EXPORTS:
        assert s in syms


def test_ctypes_struct_sizes_match_header_layout():
    import ctypes as C
    from mri2speech_b200 import _lib
    # m2s_conv_args: spot-check offsets that the C side relies on (pointer-aligned, no surprises)
    assert _lib.ConvArgs.shift.offset == 40 and _lib.ConvArgs.shift.size == 64
    assert _lib.ConvArgs.w.offset == 104
    assert C.sizeof(_lib.Tensor) == 8 + 8 + 8 + 32
    assert C.sizeof(_lib.GeneratorConfig) == 4 * (3 + 8 + 8 + 1 + 8 + 24 + 1 + 1)   # ... precision, resblock


def test_cpu_tensors_are_refused_loudly():
    from mri2speech_b200._lib import M2SError
    from mri2speech_b200.vocoder import Generator
    from mri2speech_b200.acoustic import build_acoustic_model
    g = Generator(load_config()).eval()
    with pytest.raises(M2SError, match="no CPU fallback"):
        g(torch.zeros(1, 64, 4))
    m = build_acoustic_model().eval()
    with pytest.raises(M2SError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 1, 256, 256))


def test_generator_accepts_reference_checkpoint_formats():
    """strict load of weight_g/weight_v keys; best-effort weight-norm removal (run_mri_video_inference.py:96-115)."""
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    g = Generator(load_config())
    ckpt = {"generator": {k: v.clone() for k, v in g.state_dict().items()}}
    g2 = Generator(load_config())
    g2.load_state_dict(ckpt["generator"])  # strict
    from torch.nn.utils import remove_weight_norm
    for module in list(g2.ups) + [g2.conv_post]:
        remove_weight_norm(module)
    for res in g2.resblocks:
        res.remove_weight_norm()
    with pytest.raises(ValueError):
        remove_weight_norm(g2.conv_pre)  # conv_pre never had weight-norm (models.py:94)
    assert "ups.0.weight" in g2.state_dict() and g2.num_kernels == 3 and g2.num_upsamples == 4
    assert g2.h is not None and hasattr(g2, "conv_post")


def test_acoustic_load_state_dict_reports_missing_and_unexpected():
    from mri2speech_b200.acoustic import build_acoustic_model, MRIAcousticModel, OTNLikeCNNBiLSTM
    assert MRIAcousticModel is OTNLikeCNNBiLSTM
    m = build_acoustic_model(n_mels=64, cnn_pretrained=False, rnn_hidden=640, dropout=0.5, use_checkpoint=False,
                             ckpt_segments=2, use_reentrant=False)
    sd = dict(m.state_dict())
    sd.pop("head.bias")
    sd["extra.key"] = torch.zeros(1)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert missing == ["head.bias"] and unexpected == ["extra.key"]
    assert hasattr(m.cnn, "backbone") and hasattr(m.cnn, "gap") and hasattr(m.rnn, "lstm") and hasattr(m.rnn, "dropout")


def test_ctypes_structs_match_the_header_as_the_c_compiler_sees_it(tmp_path):
    """Compile include/m2s.h with gcc (plain C: the header is the boundary a maintainer binds against) and compare
    sizeof / offsetof of every struct that crosses the boundary with the ctypes mirrors in mri2speech_b200/_lib.py."""
    import ctypes as C
    import shutil
    import subprocess
    from mri2speech_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    checks = {
        "m2s_conv_args": (_lib.ConvArgs, ["a", "shift", "w", "n", "d", "d_row_offset", "bias", "res", "res_inv_slope",
                                          "accum", "out_scale", "act_slope", "lens", "len_scale", "pitch", "j_hi",
                                          "a_half", "d16", "d16_lo", "res_hi", "res_lo"]),
        "m2s_generator_config": (_lib.GeneratorConfig, ["num_mels", "num_upsamples", "upsample_rates", "num_kernels",
                                                        "resblock_dilations", "precision", "resblock"]),
        "m2s_tensor": (_lib.Tensor, ["name", "data", "ndim", "shape"]),
        "m2s_acoustic_config": (_lib.AcousticConfig, ["n_mels", "rnn_hidden", "height", "width", "precision"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "m2s.h"', 'int main(void) {']
    for struct, (_, fields) in checks.items():
        lines.append(f'  printf("{struct} %zu\\n", sizeof({struct}));')
        for f in fields:
            lines.append(f'  printf("{struct}.{f} %zu\\n", offsetof({struct}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    seen = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for struct, (mirror, fields) in checks.items():
        assert int(seen[struct]) == C.sizeof(mirror), struct
        for f in fields:
            assert int(seen[f"{struct}.{f}"]) == getattr(mirror, f).offset, f"{struct}.{f}"

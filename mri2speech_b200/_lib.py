"""ctypes binding of libm2s.so (the C ABI declared in include/m2s.h).

There is deliberately no fallback: if the shared library is missing or the
device is not an sm_100 part, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libm2s.so")
_lib: Optional[C.CDLL] = None

M2S_MAX_TAPS = 16
M2S_MAX_UPS = 8
M2S_MAX_RBK = 8
PREC_TF32, PREC_FP32, PREC_FP16 = 0, 1, 2
# The build the drop-in modules and CLIs run when no precision is given: fp16 operands (tcgen05 kind::f16; the same
# 10-bit mantissa as tf32, fp32 accumulation and fp32 residual streams) -- what bench.py measures.  Parity evidence:
# tests/test_scaled_init_gpu.py, tests/test_bench_paths_gpu.py.  M2S_PRECISION=tf32|fp32 selects the other builds.
DEFAULT_PRECISION = os.environ.get("M2S_PRECISION", "fp16")

PRECISIONS = {"tf32": PREC_TF32, "fp32": PREC_FP32, "fp16": PREC_FP16}
ACT_NONE, ACT_LRELU, ACT_SILU = 0, 1, 2
MASK_NONE, MASK_LEN, MASK_PITCH = 0, 1, 2
IMPL_TCGEN05, IMPL_SIMT = 0, 1


class M2SError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_batch_rows", C.c_int64), ("a_rows", C.c_int32), ("a_ld", C.c_int32),
        ("c_in", C.c_int32), ("batch", C.c_int32), ("l_out", C.c_int32), ("taps", C.c_int32),
        ("shift", C.c_int32 * M2S_MAX_TAPS),
        ("w", C.c_void_p), ("n", C.c_int32),
        ("d", C.c_void_p), ("d_batch_rows", C.c_int64), ("d_ld", C.c_int32), ("d_row_offset", C.c_int32),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("res_ld", C.c_int32), ("res_inv_slope", C.c_float),
        ("res_after_act", C.c_int32),
        ("accum", C.c_void_p), ("accum_ld", C.c_int32), ("out_scale", C.c_float),
        ("act", C.c_int32), ("act_slope", C.c_float),
        ("mask_mode", C.c_int32), ("lens", C.c_void_p), ("len_scale", C.c_int32),
        ("pitch", C.c_int32), ("i_lo", C.c_int32), ("i_hi", C.c_int32), ("j_lo", C.c_int32), ("j_hi", C.c_int32),
        ("a_half", C.c_int32), ("d16", C.c_void_p),
        ("d16_lo", C.c_void_p), ("res_hi", C.c_void_p), ("res_lo", C.c_void_p),
    ]


class GeneratorConfig(C.Structure):
    _fields_ = [
        ("num_mels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * M2S_MAX_UPS), ("upsample_kernel_sizes", C.c_int32 * M2S_MAX_UPS),
        ("num_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * M2S_MAX_RBK),
        ("resblock_dilations", (C.c_int32 * 3) * M2S_MAX_RBK), ("precision", C.c_int32),
        ("resblock", C.c_int32),
    ]


class AcousticConfig(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("rnn_hidden", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("precision", C.c_int32)]


EXPORTS = (
    "m2s_version", "m2s_last_error_string", "m2s_device_check", "m2s_conv_fwd", "m2s_resblock_pair_fwd",
    "m2s_generator_create", "m2s_generator_destroy", "m2s_generator_workspace_bytes", "m2s_generator_forward",
    "m2s_generator_forward_btc",
    "m2s_generator_launches",
    "m2s_acoustic_create", "m2s_acoustic_destroy", "m2s_acoustic_workspace_bytes", "m2s_acoustic_forward",
    "m2s_acoustic_forward_u8", "m2s_acoustic_forward_packed", "m2s_acoustic_encode_packed", "m2s_acoustic_encode", "m2s_acoustic_rnn_head", "m2s_acoustic_launches", "m2s_mel_glue",
)


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load libm2s.so (once).  Raises M2SError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise M2SError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C mri2speech_b200/csrc).  There is no CPU or PyTorch fallback.")
    L = C.CDLL(_LIB_PATH)
    L.m2s_version.restype = C.c_char_p
    L.m2s_last_error_string.restype = C.c_char_p
    L.m2s_device_check.argtypes = [C.c_int]
    L.m2s_conv_fwd.argtypes = [C.POINTER(ConvArgs), C.c_int, C.c_void_p]
    L.m2s_resblock_pair_fwd.argtypes = [C.POINTER(ConvArgs), C.POINTER(ConvArgs), C.c_void_p]
    L.m2s_debug_set_knob.argtypes = [C.c_char_p, C.c_int]
    L.m2s_debug_trace.argtypes = [C.c_void_p, C.c_int32]
    L.m2s_debug_profile.argtypes = [C.c_int]
    L.m2s_debug_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int32,
                                         C.POINTER(C.c_int32)]
    L.m2s_debug_profile_tags.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]
    L.m2s_debug_launch_count.argtypes = [C.c_int]
    L.m2s_debug_launch_count.restype = C.c_longlong
    L.m2s_generator_create.argtypes = [C.POINTER(GeneratorConfig), C.POINTER(Tensor), C.c_int32,
                                       C.POINTER(C.c_void_p)]
    L.m2s_generator_destroy.argtypes = [C.c_void_p]
    L.m2s_generator_destroy.restype = None
    L.m2s_generator_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.m2s_generator_workspace_bytes.restype = C.c_size_t
    L.m2s_generator_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_generator_forward_btc.argtypes = L.m2s_generator_forward.argtypes
    L.m2s_generator_launches.argtypes = [C.c_void_p]
    missing = [name for name in EXPORTS if not hasattr(L, name)]
    if missing:
        raise M2SError(f"{_LIB_PATH} is stale: missing exports {missing}; rebuild it")
    _bind_acoustic(L)
    _lib = L
    return L


def _bind_acoustic(L) -> None:
    L.m2s_acoustic_create.argtypes = [C.POINTER(AcousticConfig), C.POINTER(Tensor), C.c_int32,
                                      C.POINTER(C.c_void_p)]
    L.m2s_acoustic_destroy.argtypes = [C.c_void_p]
    L.m2s_acoustic_destroy.restype = None
    L.m2s_acoustic_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.m2s_acoustic_workspace_bytes.restype = C.c_size_t
    L.m2s_acoustic_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_acoustic_forward_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_acoustic_forward_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_acoustic_encode_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_acoustic_encode_packed.restype = C.c_int
    L.m2s_acoustic_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]
    L.m2s_acoustic_rnn_head.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.m2s_acoustic_launches.argtypes = [C.c_void_p]
    L.m2s_mel_glue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def check(status: int) -> None:
    if status != 0:
        msg = lib().m2s_last_error_string().decode("utf-8", "replace")
        raise M2SError(f"libm2s error {status}: {msg}")


def require_device(t: torch.Tensor) -> None:
    """No CPU fallback: fail loudly on anything that is not a CUDA float32 tensor."""
    if not t.is_cuda:
        raise M2SError("mri2speech_b200 runs on sm_100 CUDA devices only; got a CPU tensor "
                       "(there is no CPU fallback -- move the module and its inputs to cuda)")
    check(lib().m2s_device_check(t.device.index if t.device.index is not None else -1))


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def state_dict_to_tensors(sd: Dict[str, torch.Tensor]):
    """Host float32 views of a state_dict as an m2s_tensor array.  Returns (array, n, keepalive)."""
    keep = []
    items = []
    for name, t in sd.items():
        if not torch.is_tensor(t) or not t.is_floating_point():
            continue
        h = t.detach().to("cpu", torch.float32).contiguous()
        if h.dim() > 4:
            continue
        nm = name.encode("utf-8")
        keep.append((nm, h))
        mt = Tensor()
        mt.name = nm
        mt.data = h.data_ptr()
        mt.ndim = h.dim()
        for i, s in enumerate(h.shape):
            mt.shape[i] = int(s)
        items.append(mt)
    arr = (Tensor * len(items))(*items)
    return arr, len(items), keep


def set_knob(name: str, value: int) -> None:
    check(lib().m2s_debug_set_knob(name.encode(), int(value)))


def conv_fwd(a: torch.Tensor, w: torch.Tensor, shifts: Sequence[int], l_out: int, *, impl: int = IMPL_TCGEN05,
             a_rows: Optional[int] = None, bias=None, res=None, res_inv_slope: float = 1.0, res_after_act: bool = False, accum=None,
             out_scale: float = 1.0, act: int = ACT_NONE, act_slope: float = 0.0,
             lens=None, len_scale: int = 1, pitch_mask=None, d_row_offset: int = 0,
             d_rows: Optional[int] = None, out: Optional[torch.Tensor] = None,
             out16: Optional[torch.Tensor] = None, want_d32: bool = True, out16_lo: Optional[torch.Tensor] = None,
             res_hi: Optional[torch.Tensor] = None, res_lo: Optional[torch.Tensor] = None, _return_args: bool = False):
    """Test helper around m2s_conv_fwd.  a: (B, L_in, C) cuda fp32 (tf32 operands) or fp16 (kind::f16 operands);
    w: (taps, N, C) cuda fp32.  ``out16``: optional fp16 second output (same shape as the fp32 one).
    ``out16_lo`` / ``res_hi`` + ``res_lo``: the split-fp16 residual stream (include/m2s.h)."""
    require_device(a)
    B, L_in, Cin = a.shape
    taps, N, Cw = w.shape
    assert Cw == Cin and taps == len(shifts)
    d_rows = d_rows if d_rows is not None else l_out + d_row_offset
    d = out if out is not None else torch.zeros(B, d_rows, N, device=a.device, dtype=torch.float32)
    args = ConvArgs()
    args.a_half = int(a.dtype == torch.float16)
    args.d16 = ptr(out16)
    args.d16_lo = ptr(out16_lo); args.res_hi = ptr(res_hi); args.res_lo = ptr(res_lo)
    args.a = a.data_ptr(); args.a_batch_rows = L_in; args.a_rows = a_rows if a_rows is not None else L_in
    args.a_ld = Cin; args.c_in = Cin; args.batch = B; args.l_out = l_out; args.taps = taps
    for i, s in enumerate(shifts):
        args.shift[i] = int(s)
    args.w = w.data_ptr(); args.n = N
    args.d = d.data_ptr() if want_d32 else None
    args.d_batch_rows = d.shape[1]; args.d_ld = N; args.d_row_offset = d_row_offset
    args.bias = ptr(bias); args.res = ptr(res); args.res_ld = N; args.res_inv_slope = res_inv_slope
    args.res_after_act = int(res_after_act)
    args.accum = ptr(accum); args.accum_ld = N; args.out_scale = out_scale
    args.act = act; args.act_slope = act_slope
    if lens is not None:
        args.mask_mode = MASK_LEN; args.lens = lens.data_ptr(); args.len_scale = len_scale
    elif pitch_mask is not None:
        args.mask_mode = MASK_PITCH
        args.pitch, args.i_lo, args.i_hi, args.j_lo, args.j_hi = [int(v) for v in pitch_mask]
    if _return_args:
        return args, d
    check(lib().m2s_conv_fwd(C.byref(args), impl, current_stream()))
    return d


def resblock_pair_fwd(x16: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, dilation: int, w2: torch.Tensor,
                      b2: torch.Tensor, *, slope: float = 0.1, **epilogue2):
    """Test helper around m2s_resblock_pair_fwd: conv2(lrelu(conv1(x) + b1)) with conv2's fused epilogue
    (``epilogue2``: the keyword arguments of conv_fwd -- res, accum, out_scale, act, lens, out16, want_d32 ...)."""
    B, L, Cch = x16.shape
    k1, k2 = w1.shape[0], w2.shape[0]
    a1, _ = conv_fwd(x16, w1, [-(k1 - 1 - j) * dilation for j in range(k1)], L, bias=b1, act=ACT_LRELU,
                     act_slope=slope, _return_args=True)
    dummy = torch.empty(B, 1, w2.shape[2], device=x16.device, dtype=torch.float16)
    a2, d = conv_fwd(dummy, w2, [-(k2 - 1 - j) for j in range(k2)], L, bias=b2, a_rows=L, _return_args=True, **epilogue2)
    a2.a_batch_rows = L
    check(lib().m2s_resblock_pair_fwd(C.byref(a1), C.byref(a2), current_stream()))
    return d


def profile(enable: bool) -> None:
    check(lib().m2s_debug_profile(int(enable)))


PROFILE_TAGS = {0: "other", 1: "encoder_gemm", 2: "encoder_simt", 3: "bilstm", 4: "vocoder_gemm", 5: "vocoder_simt"}


def profile_read(cap: int = 262144, with_tags: bool = False):
    """Per-launch (ms, executed flops[, tag]) of every bracketed launch since the last read."""
    ms = (C.c_float * cap)()
    fl = (C.c_double * cap)()
    n = C.c_int32(0)
    check(lib().m2s_debug_profile_read(ms, fl, cap, C.byref(n)))
    if not with_tags:
        return list(ms[: n.value]), list(fl[: n.value])
    tg = (C.c_int32 * cap)()
    nt = C.c_int32(0)
    check(lib().m2s_debug_profile_tags(tg, cap, C.byref(nt)))
    return list(ms[: n.value]), list(fl[: n.value]), list(tg[: nt.value])


def launch_count(reset: bool = False) -> int:
    """Kernels launched by libm2s in this process since the last reset."""
    return int(lib().m2s_debug_launch_count(int(reset)))

// Shared host/device definitions for libm2s (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <mutex>
#include <string>
#include "../../include/m2s.h"

namespace m2s {

// ---- error plumbing --------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define M2S_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::m2s::fail(M2S_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                         cudaGetErrorString(_e));                                           \
  } while (0)

#define M2S_TRY(expr)             \
  do {                            \
    int _s = (expr);              \
    if (_s != M2S_OK) return _s;  \
  } while (0)

// One-time setup that CUDA keeps PER DEVICE (cudaFuncSetAttribute: function attributes belong to the device's copy of
// the kernel): run `f` once for each device a call site is reached on.  Thread-safe.
class PerDeviceOnce {
  std::mutex mu_;
  uint64_t done_ = 0;

 public:
  template <class F>
  int run(F&& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu_);
    if (dev >= 0 && dev < 64 && ((done_ >> dev) & 1)) return M2S_OK;
    const int s = f();
    if (s == M2S_OK && dev >= 0 && dev < 64) done_ |= 1ull << dev;
    return s;
  }
};

// ---- the conv problem as the kernels see it ----------------------------------
// D[b, q + d_row_offset, n] = epi(sum_j sum_c A[b, q + shift[j], c] * W[j][n][c])
struct Epilogue {
  const float* bias;
  const float* res;
  const void* res_hi;  // split-fp16 residual (res == nullptr): value = float(hi) + float(lo), __half planes
  const void* res_lo;
  const float* accum;
  const int32_t* lens;
  int res_ld;
  int accum_ld;
  float res_inv_slope;
  float out_scale;
  float act_slope;
  int act;
  int res_after_act;
  int mask_mode;
  int len_scale;
  int pitch, i_lo, i_hi, j_lo, j_hi;
  int res_half;          // `res` points at __half data (res_ld in elements): the encoder's fp16 residual stream (internal)
  unsigned pitch_magic;  // floor(2^32 / pitch) + 1: row / pitch == __umulhi(row, pitch_magic) for row * pitch < 2^32
                         // (filled in by the launchers: finalize_epilogue)
};

struct ConvProblem {
  const float* a;      // fp32 rows, or __half rows when a_half != 0 (a_ld / c_in count elements either way)
  long long a_batch_rows;
  int a_rows, a_ld, c_in;
  int batch, l_out;
  int taps;
  int shift[M2S_MAX_TAPS];
  int n;
  float* d;            // fp32 output (may be null when d16 is set)
  long long d_batch_rows;
  int d_ld, d_row_offset;
  Epilogue epi;
  int a_half;          // A operand is fp16 (tcgen05 kind::f16); 0 = fp32 rounded to tf32 by TMA
  void* d16;           // optional second output, fp16, indexed like d (same d_ld in elements); null = none
  void* d16_lo;        // optional lo plane of the split-fp16 pair (d16 = hi): fp16(v - float(fp16(v))); needs d16
  // Per-tap K windows (internal; tap_ksteps = 0: off).  When set, tap j contracts only K-steps kofs[j] .. kofs[j] +
  // tap_ksteps - 1 (16 fp16 elements each) of the single K block -- of the A row AND of its weight block.  This is how a
  // conv over P consecutive pixels per GEMM row (N = P * C_out: the narrow first stage of the encoder) reads pixel
  // P q + t: the GEMM row is the P pixels' channels side by side (one 64 / 128-byte row), tap (.., t) is row shift
  // floor(t / P) and the K window of pixel t mod P; the weight blocks carry zeros outside their window.
  int tap_ksteps;
  int kofs[M2S_MAX_TAPS];
};

inline ConvProblem problem_from_args(const m2s_conv_args& a) {
  ConvProblem p{};
  p.a = a.a; p.a_batch_rows = a.a_batch_rows; p.a_rows = a.a_rows; p.a_ld = a.a_ld; p.c_in = a.c_in;
  p.batch = a.batch; p.l_out = a.l_out; p.taps = a.taps;
  for (int i = 0; i < M2S_MAX_TAPS; ++i) p.shift[i] = a.shift[i];
  p.n = a.n; p.d = a.d; p.d_batch_rows = a.d_batch_rows; p.d_ld = a.d_ld; p.d_row_offset = a.d_row_offset;
  p.epi.bias = a.bias; p.epi.res = a.res; p.epi.res_ld = a.res_ld; p.epi.res_inv_slope = a.res_inv_slope; p.epi.res_after_act = a.res_after_act;
  p.epi.accum = a.accum; p.epi.accum_ld = a.accum_ld; p.epi.out_scale = a.out_scale; p.epi.act = a.act;
  p.epi.act_slope = a.act_slope; p.epi.mask_mode = a.mask_mode;
  p.epi.lens = a.lens; p.epi.len_scale = a.len_scale; p.epi.pitch = a.pitch; p.epi.i_lo = a.i_lo;
  p.epi.i_hi = a.i_hi; p.epi.j_lo = a.j_lo; p.epi.j_hi = a.j_hi;
  p.a_half = a.a_half; p.d16 = a.d16; p.d16_lo = a.d16_lo; p.epi.res_hi = a.res_hi; p.epi.res_lo = a.res_lo;
  return p;
}

#ifdef __CUDACC__
// Round-to-nearest (ties away) fp32 -> tf32, result kept in an fp32 register.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// Row validity for the fused output mask. `drow` = q + d_row_offset (row inside batch item b).
__device__ __forceinline__ bool epi_row_valid(const Epilogue& e, int b, int drow) {
  if (e.mask_mode == M2S_MASK_LEN) {
    return drow < __ldg(e.lens + b) * e.len_scale;
  } else if (e.mask_mode == M2S_MASK_PITCH) {
    int i = drow / e.pitch, j = drow - i * e.pitch;
    return i >= e.i_lo && i < e.i_hi && j >= e.j_lo && j < e.j_hi;
  }
  return true;
}

// One output element.  `roff` = element offset of (row, n) for ld == 1 scaling done by caller:
// res/accum are addressed as base + row_index*ld + n by the caller, who passes the loaded values.
__device__ __forceinline__ float epi_apply(const Epilogue& e, float acc, float bias, float res, float accum,
                                           bool valid) {
  float v = acc + bias;
  const float r = e.res ? (res >= 0.f ? res : res * e.res_inv_slope) : 0.f;
  if (!e.res_after_act) v += r;
  if (e.accum) v += accum;
  v *= e.out_scale;
  if (e.act == M2S_ACT_LRELU) v = v >= 0.f ? v : v * e.act_slope;
  else if (e.act == M2S_ACT_SILU) v = v / (1.f + __expf(-v));
  if (e.res_after_act) v += r;
  if (!valid) v = 0.f;
  return v;
}
#endif  // __CUDACC__

// ---- engine entry points (host) ---------------------------------------------
// Pre-packed weights for the tcgen05 engine: [n_tile][cblock][tap][n_tile_rows][32 floats, 128B-swizzled].
enum { PACK_FP32 = 0, PACK_TF32 = 1, PACK_FP16 = 2 };

struct PackedWeights {
  // operand format of the tcgen05 layouts: 0 = tf32 (32 floats per 128-byte SMEM row), 1 = fp16 with 64 halves per
  // 128-byte row (SWIZZLE_128B, c_in > 32), 2 = fp16 with 32 halves per 64-byte row (SWIZZLE_64B, c_in <= 32)
  int half = 0;
  int kblock = 32;        // channels per SMEM row
  int row_bytes = 128;
  float* dev = nullptr;   // packed, tcgen05 layout (fp16 layouts are stored in the same allocation type)
  float* plain = nullptr; // [taps][n][c_in] plain (SIMT path / tests)
  int n = 0, c_in = 0, taps = 0;
  int n_tile = 0, n_tiles = 0, cblocks = 0;
  size_t packed_floats = 0;
  // CTA-pair layout (cta_group::2, wide layers only): rows of 32 floats, unswizzled,
  // [n_tile][cblock][half][tap][n_tile/2]; nullptr when the layer is too narrow
  float* dev_pair = nullptr;
  int n_tile_pair = 0, n_tiles_pair = 0;
};

// Choose N tiling for the tcgen05 engine (n_tile multiple of 16, <= 256).
void choose_n_tiling(int n, int* n_tile, int* n_tiles);
// Geometry of a stride-2 3x3 conv input for the fused EdgeResidual kernel's space-to-depth loads (fused_er_sm100.cu):
// the zero-bordered NHWC frame has h x w pixels of c channels, frame_rows = (h + 2) * (w + 2) rows per frame.
struct ErS2d {
  int h, w, c;
  long long frame_rows;
};

// Pack host weights [taps][n][c_in] (already folded) into both device layouts.  mode = PACK_FP32 (values kept),
// PACK_TF32 (RNE-rounded to tf32) or PACK_FP16 (fp16 operands for tcgen05 kind::f16; `plain` keeps the rounded fp32).
int pack_weights(const float* host_w, int taps, int n, int c_in, int mode, PackedWeights* out);
inline int pack_weights(const float* host_w, int taps, int n, int c_in, bool tf32_round, PackedWeights* out) {
  return pack_weights(host_w, taps, n, c_in, tf32_round ? PACK_TF32 : PACK_FP32, out);
}
void free_weights(PackedWeights* w);

int conv_tcgen05(const ConvProblem& p, const PackedWeights& w, cudaStream_t stream);
int conv_tcgen05_pair(const ConvProblem& p, const PackedWeights& w, cudaStream_t stream);
// Fused ResBlock pair (conv1 -> leaky-ReLU -> conv2 -> epilogue) with the intermediate kept in SMEM (fp16 build).
bool resblock_pair_supported(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2,
                             const PackedWeights& w2);
int resblock_pair_fused(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2,
                        cudaStream_t stream);
// Launch accounting (bench.py): every launch site brackets its kernel(s) with profile_before / profile_after.  The pair
// always counts the launch (m2s_debug_launch_count) and, when the probe is on, records two CUDA events on the launching
// stream plus the executed flops and the current tag (which part of the path the launch belongs to).
enum { PROF_OTHER = 0, PROF_ENC_GEMM = 1, PROF_ENC_SIMT = 2, PROF_RNN = 3, PROF_VOC_GEMM = 4, PROF_VOC_SIMT = 5 };
int profile_before(cudaStream_t stream);
int profile_after(cudaStream_t stream, double flops, int kernels = 1);
void profile_set_tag(int tag);
int profile_read_tags(int32_t* tags, int cap);
long long launch_count(bool reset);
int conv_simt(const ConvProblem& p, const float* w_plain, cudaStream_t stream);

// debug / probe knobs for the tcgen05 engine (environment-driven, read once)
struct EngineKnobs {
  int base_offset_mode = 0;  // 0: base_offset=0 ; 1: (start>>7)&7
  int msub = 0;              // 0 = auto, else force 1 or 2
  int tmap_tf32 = 1;         // encode the A tensor map as TFLOAT32 (TMA rounds fp32 -> tf32 on load)
  int max_ctas = 0;          // 0 = #SMs
  int a_per_tap = 0;         // 1: reload the A tile per tap (no row-shifted descriptors; fallback)
  unsigned long long* trace = nullptr;  // debug timeline buffer (device), trace_tiles x 9 stamps of CTA 0
  int trace_tiles = 0;
  int dbg = 0;
  int pair = 1;              // CTA-pair (cta_group::2) kernel: 0 never, 1 heuristic, 2 whenever packed
  int pair_min_n = 32;       // narrowest layer whose weights are also packed for the CTA-pair kernel
  int n_tile_max = 128;      // N columns per tile once N exceeds it (weights are packed accordingly)
  int fuse_max_n = 64;       // widest layer the fused ResBlock-pair kernel takes (N = 128 works but is slower than
                             // two CTA-pair launches: one accumulator buffer each, no cta_group::2 weight split)
};
EngineKnobs& engine_knobs();

int sm_count();
int profile_enable(int on);
int profile_read(float* ms, double* flops, int cap, int* n_out);

}  // namespace m2s

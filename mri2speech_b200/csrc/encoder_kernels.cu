// CUDA-core kernels of the frame-CNN encoder (the parts that are not GEMM-shaped).
//
// Reference: timm tf_efficientnetv2_b2 (features_only) as called by EffNetV2B2Backbone.forward
// (mri2speech_code/mri_acoustic_model.py:28-48); topology restated in SURVEY.md 8a-1.
//   * stem 3x3 s2 conv (the 3 identical input channels folded into 1) + BN + SiLU
//   * im2col for the two stride-2 3x3 convs (TF "same": pad right/bottom only)
//   * depthwise 3x3 (+BN+SiLU) with the squeeze (spatial sum) fused
//   * squeeze-excite MLP, excite scale, global average pool
// All activations are channels-last.  "Padded" layouts carry a one-pixel zero border so that the
// tensor-core engine can treat a 3x3 stride-1 conv as 9 row-shifted taps over the flattened image.
// Every kernel that touches an activation is a template over its element type: float (tf32 / fp32 builds) or __half
// (fp16 build: activations that are only tensor-core operands live in HBM as fp16, halving the traffic of these
// memory-bound kernels; the arithmetic stays fp32).
#include "m2s_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace m2s {

namespace {

// SiLU.  float activations (tf32 / fp32 builds): exp + divide.  fp16 activations (fp16 build): the one-MUFU form
// h + h * tanh(h), h = v / 2 (tanh.approx.f32, relative error 2^-11 = the rounding of the fp16 store that follows);
// the two-MUFU form makes these memory-bound kernels MUFU-bound (16 MUFU lanes per SM per clock).
template <typename T>
__device__ __forceinline__ float silu_t(float v) {
  if (sizeof(T) == 4) return __fdividef(v, 1.f + __expf(-v));
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float silu(float v) { return v / (1.f + __expf(-v)); }
// packed fp32 pairs (Blackwell FFMA2: two IEEE fp32 FMAs per issued instruction)
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// 4 consecutive channels <-> float4
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v) {
  uint2 u;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(v.w), "f"(v.z));
  *reinterpret_cast<uint2*>(p) = u;
}

// Per-frame min / max of (optionally masked) uint8 frames -> norm[n] = (min, 1 / (max - min)) (scale 0 for a
// constant frame, which then maps to all zeros like run_mri_video_inference.py:50-53).  The reference's z-score
// followed by min-max (:41-51) is an affine map followed by min-max, i.e. plain per-frame min-max.
// Masking (scripts/mask_rtmri_video.py:96-98): masked = uint8(clip(frame * mask, 0, 255)) (truncation).
__device__ __forceinline__ float masked_u8(uint8_t v, float m) {
  return static_cast<float>(static_cast<uint8_t>(fminf(fmaxf(static_cast<float>(v) * m, 0.f), 255.f)));
}

__global__ void __launch_bounds__(256) frame_minmax_kernel(const uint8_t* __restrict__ frames,
                                                           const int32_t* __restrict__ fmap,
                                                           const float* __restrict__ mask, float2* __restrict__ norm,
                                                           int hw) {
  __shared__ float smin[8], smax[8];
  const int n = blockIdx.x;
  const uint8_t* f = frames + static_cast<size_t>(fmap ? fmap[n] : n) * hw;
  float mn = 1e30f, mx = -1e30f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float v = mask ? masked_u8(f[i], mask[i]) : static_cast<float>(f[i]);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { mn = fminf(mn, smin[k]); mx = fmaxf(mx, smax[k]); }
    norm[n] = make_float2(mn, mx > mn ? 1.f / (mx - mn) : 0.f);
  }
}

// frames (n, H, W) -> out padded ((H/2+2) x (W/2+2) rows, 32 ch); one block per padded output row.
// kU8: frames are uint8 and are masked / min-max normalised on the fly (the fused ingest of SURVEY.md 8f-1).
template <bool kU8, typename T>
__global__ void __launch_bounds__(256) stem_kernel(const void* __restrict__ frames_v, const int32_t* __restrict__ fmap,
                                                   const float* __restrict__ mask, const float2* __restrict__ norm,
                                                   T* __restrict__ out, const float* __restrict__ w /*[9][32]*/,
                                                   const float* __restrict__ bias, int H, int W) {
  const int Ho = H / 2, Wo = W / 2, pitch = Wo + 2;
  const int n = blockIdx.y;
  const int i = blockIdx.x;  // padded row 0..Ho+1
  T* orow = out + (static_cast<size_t>(n) * (Ho + 2) * pitch + static_cast<size_t>(i) * pitch) * 32;
  if (i == 0 || i == Ho + 1) {
    for (int k = threadIdx.x; k < pitch * 8; k += blockDim.x) store4(orow + 4 * k, make_float4(0, 0, 0, 0));
    return;
  }
  extern __shared__ float srow[];  // 3 x (W + 1) input pixels
  __shared__ float sw[9 * 32 + 32];
  const int y = i - 1;
  const size_t foff = static_cast<size_t>(fmap ? fmap[n] : n) * H * W;
  float2 nm = make_float2(0.f, 1.f);
  if (kU8) nm = norm[n];
  for (int k = threadIdx.x; k < 3 * (W + 1); k += blockDim.x) {
    const int r = k / (W + 1), x = k % (W + 1);
    const int yy = 2 * y + r;
    float v = 0.f;
    if (yy < H && x < W) {
      const size_t idx = static_cast<size_t>(yy) * W + x;
      if (kU8) {
        const uint8_t u = static_cast<const uint8_t*>(frames_v)[foff + idx];
        v = ((mask ? masked_u8(u, mask[idx]) : static_cast<float>(u)) - nm.x) * nm.y;
      } else {
        v = static_cast<const float*>(frames_v)[foff + idx];
      }
    }
    srow[k] = v;
  }
  for (int k = threadIdx.x; k < 9 * 32 + 32; k += blockDim.x) sw[k] = k < 288 ? w[k] : bias[k - 288];
  __syncthreads();
  // thread -> (x, 8-channel group)
  for (int k = threadIdx.x; k < pitch * 4; k += blockDim.x) {
    const int j = k >> 2, cg = (k & 3) * 8;
    float v[8];
    if (j == 0 || j == Wo + 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = 0.f;
    } else {
      const int x = j - 1;
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = sw[288 + cg + c];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float px = srow[dy * (W + 1) + 2 * x + dx];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = fmaf(px, sw[(dy * 3 + dx) * 32 + cg + c], v[c]);
        }
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = sizeof(T) == 4 ? silu(v[c]) : silu_t<T>(v[c]);
    }
    T* o = orow + static_cast<size_t>(j) * 32 + cg;
    store4(o, make_float4(v[0], v[1], v[2], v[3]));
    store4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
}

// The same layer, R output rows per block: the 2R + 1 input rows are staged once (4 pixels per load), each thread keeps
// the 9 x 8 weights of ITS 8-channel group in registers (72 LDS per block instead of 72 per output pixel) and walks
// R * Wo * 4 (pixel, channel group) items -- a whole number of 256-thread passes, where the one-row kernel above ran a
// third pass for 8 border items.  Needs W % 4 == 0, (H / 2) % R == 0 and frames aligned for 4-pixel loads.
template <bool kU8, typename T, int R>
__global__ void __launch_bounds__(256) stem_rows_kernel(const void* __restrict__ frames_v, const int32_t* __restrict__ fmap,
                                                        const float* __restrict__ mask, const float2* __restrict__ norm,
                                                        T* __restrict__ out, const float* __restrict__ w /*[9][32]*/,
                                                        const float* __restrict__ bias, int H, int W) {
  const int Ho = H / 2, Wo = W / 2, pitch = Wo + 2, Wp = W + 1;
  constexpr int kRowsIn = 2 * R + 1;
  constexpr int kUnits = 32 * sizeof(T) / 16;   // 16-byte vectors per output pixel
  const int n = blockIdx.y;
  const int y0 = blockIdx.x * R;
  const int tid = threadIdx.x;
  T* obase = out + static_cast<size_t>(n) * (Ho + 2) * pitch * 32;
  extern __shared__ float srow[];  // kRowsIn x (W + 1) input pixels (column W = the zero pad)
  __shared__ float sw[9 * 32 + 32];
  for (int k = tid; k < 9 * 32 + 32; k += 256) sw[k] = k < 288 ? w[k] : bias[k - 288];
  const size_t foff = static_cast<size_t>(fmap ? fmap[n] : n) * H * W;
  float2 nm = make_float2(0.f, 1.f);
  if (kU8) nm = norm[n];
  const int w4 = W >> 2;
  // all of a thread's loads are issued before the first conversion / store: with one load per loop iteration the stores
  // waited out an HBM round trip per iteration (a quarter of the kernel's stall samples, profiles/README.md round 2 late)
  constexpr int kStageIters = 6;   // >= ceil(kRowsIn * w4 / 256) for W <= 352 (checked by the launcher)
  uchar4 u8v[kStageIters];
  float4 fv[kStageIters], mv[kStageIters];
#pragma unroll
  for (int it = 0; it < kStageIters; ++it) {
    const int k = tid + it * 256;
    const int r = k / w4, x4 = (k - r * w4) << 2;
    const int yy = 2 * y0 + r;
    u8v[it] = make_uchar4(0, 0, 0, 0);
    fv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    mv[it] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (k < kRowsIn * w4 && yy < H) {
      const size_t idx = static_cast<size_t>(yy) * W + x4;
      if (kU8) {
        u8v[it] = *reinterpret_cast<const uchar4*>(static_cast<const uint8_t*>(frames_v) + foff + idx);
        if (mask) mv[it] = *reinterpret_cast<const float4*>(mask + idx);
      } else {
        fv[it] = *reinterpret_cast<const float4*>(static_cast<const float*>(frames_v) + foff + idx);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kStageIters; ++it) {
    const int k = tid + it * 256;
    if (k >= kRowsIn * w4) break;
    const int r = k / w4, x4 = (k - r * w4) << 2;
    const int yy = 2 * y0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy < H) {
      if (kU8) {
        const uchar4 u = u8v[it];
        float4 f = make_float4(u.x, u.y, u.z, u.w);
        if (mask) {
          const float4 m = mv[it];
          f = make_float4(masked_u8(u.x, m.x), masked_u8(u.y, m.y), masked_u8(u.z, m.z), masked_u8(u.w, m.w));
        }
        v = make_float4((f.x - nm.x) * nm.y, (f.y - nm.x) * nm.y, (f.z - nm.x) * nm.y, (f.w - nm.x) * nm.y);
      } else {
        v = fv[it];
      }
    }
    float* d = srow + r * Wp + x4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  if (tid < kRowsIn) srow[tid * Wp + W] = 0.f;
  __syncthreads();
  const int cg = (tid & 3) * 8;   // 256 and Wo * 4 are multiples of 4: every item of this thread has this channel group
  float wr[9][8], br[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) wr[t][c] = sw[t * 32 + cg + c];
#pragma unroll
  for (int c = 0; c < 8; ++c) br[c] = sw[288 + cg + c];
  const int per_row = Wo * 4;
  if (sizeof(T) == 2) {
    // fp16 build: the same fp32 FMAs, issued two channels at a time (FFMA2 / FMUL2: the loop is bound by the FMA pipe)
    unsigned long long wp[9][4], bp[4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < 4; ++c) wp[t][c] = pack2f(wr[t][2 * c], wr[t][2 * c + 1]);
#pragma unroll
    for (int c = 0; c < 4; ++c) bp[c] = pack2f(br[2 * c], br[2 * c + 1]);
    for (int k = tid; k < R * per_row; k += 256) {
      const int r = k / per_row, x = (k - r * per_row) >> 2;
      unsigned long long v[4] = {bp[0], bp[1], bp[2], bp[3]};
      const float* px0 = srow + (2 * r) * Wp + 2 * x;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float px = px0[dy * Wp + dx];
          const unsigned long long pp = pack2f(px, px);
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = fma2f(pp, wp[dy * 3 + dx][c], v[c]);
        }
      float o[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        unpack2f(v[c], o[2 * c], o[2 * c + 1]);
        o[2 * c] = silu_t<T>(o[2 * c]);
        o[2 * c + 1] = silu_t<T>(o[2 * c + 1]);
      }
      T* op = obase + (static_cast<size_t>(y0 + r + 1) * pitch + x + 1) * 32 + cg;
      store4(op, make_float4(o[0], o[1], o[2], o[3]));
      store4(op + 4, make_float4(o[4], o[5], o[6], o[7]));
    }
  } else {
  for (int k = tid; k < R * per_row; k += 256) {
    const int r = k / per_row, x = (k - r * per_row) >> 2;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = br[c];
    const float* px0 = srow + (2 * r) * Wp + 2 * x;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float px = px0[dy * Wp + dx];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(px, wr[dy * 3 + dx][c], v[c]);
      }
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = sizeof(T) == 4 ? silu(v[c]) : silu_t<T>(v[c]);
    T* o = obase + (static_cast<size_t>(y0 + r + 1) * pitch + x + 1) * 32 + cg;
    store4(o, make_float4(v[0], v[1], v[2], v[3]));
    store4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
  }
  // the zero border: columns 0 and Wo + 1 of these rows, plus padded row 0 / Ho + 1 from the first / last block
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = tid; k < R * 2 * kUnits; k += 256) {
    const int r = k / (2 * kUnits), rem = k - r * 2 * kUnits;
    const int col = rem < kUnits ? 0 : Wo + 1, uu = rem < kUnits ? rem : rem - kUnits;
    reinterpret_cast<float4*>(obase + (static_cast<size_t>(y0 + r + 1) * pitch + col) * 32)[uu] = z4;
  }
  if (blockIdx.x == 0)
    for (int k = tid; k < pitch * kUnits; k += 256) reinterpret_cast<float4*>(obase)[k] = z4;
  if (blockIdx.x == gridDim.x - 1)
    for (int k = tid; k < pitch * kUnits; k += 256)
      reinterpret_cast<float4*>(obase + static_cast<size_t>(Ho + 1) * pitch * 32)[k] = z4;
}

// Zero the rows of a padded layout that the engine's masked epilogue never writes:
// [0, head_rows) and [tail_start, rows_per_frame) of every frame.
// (Type-agnostic: `ld4` = 16-byte vectors per row.)
__global__ void zero_rows_kernel(float* __restrict__ buf, int rows_per_frame, int ld4, int head_rows, int tail_start) {
  const int n = blockIdx.y;
  const int nz = head_rows + (rows_per_frame - tail_start);
  float4* base = reinterpret_cast<float4*>(buf) + static_cast<size_t>(n) * rows_per_frame * ld4;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nz * ld4; k += gridDim.x * blockDim.x) {
    int r = k / ld4;
    const int c = k % ld4;
    if (r >= head_rows) r = tail_start + (r - head_rows);
    base[static_cast<size_t>(r) * ld4 + c] = make_float4(0, 0, 0, 0);
  }
}

// Zero the left / right border pixels of a padded layout's interior rows: pixels (y, pitch - 1) and (y + 1, 0) for
// y = 1 .. H are adjacent in memory.  (The two-pixels-per-row convs of stage 0 run without the engine's border mask.)
__global__ void zero_cols_kernel(float* __restrict__ buf, int rows_per_frame, int ld4, int pitch, int H) {
  const int n = blockIdx.y;
  float4* base = reinterpret_cast<float4*>(buf) + static_cast<size_t>(n) * rows_per_frame * ld4;
  const int per_row = 2 * ld4;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < H * per_row; k += gridDim.x * blockDim.x) {
    const int y = k / per_row + 1, c = k % per_row;
    base[(static_cast<size_t>(y) * pitch + pitch - 1) * ld4 + c] = make_float4(0, 0, 0, 0);
  }
}

// im2col for a 3x3 stride-2 TF-"same" conv.  Input: padded layout (pitch_in = W_in + 2, origin (1,1));
// output rows q = y*(Wo+2) + x (x >= Wo rows are zero), columns (dy*3+dx)*C + c.
// (Type-agnostic copy: `c4n` = 16-byte vectors per pixel = C * sizeof(element) / 16.)
__global__ void im2col_s2_kernel(const float* __restrict__ in, float* __restrict__ col, int Hin, int Win, int c4n) {
  const int Ho = Hin / 2, Wo = Win / 2, pitch_in = Win + 2, pitch_o = Wo + 2;
  const int n = blockIdx.y;
  const size_t total = static_cast<size_t>(Ho) * pitch_o * 9 * c4n;
  const float4* src = reinterpret_cast<const float4*>(in) + static_cast<size_t>(n) * (Hin + 2) * pitch_in * c4n;
  float4* dst = reinterpret_cast<float4*>(col) + static_cast<size_t>(n) * Ho * pitch_o * 9 * c4n;
  for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
       k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c4 = k % c4n;
    const int tap = (k / c4n) % 9;
    const int q = k / (static_cast<size_t>(c4n) * 9);
    const int y = q / pitch_o, x = q % pitch_o;
    float4 v = make_float4(0, 0, 0, 0);
    if (x < Wo) {
      const int dy = tap / 3, dx = tap % 3;
      // input pixel (2y+dy, 2x+dx) lives at padded (2y+dy+1, 2x+dx+1); index Hin / Win hits the zero border
      v = src[(static_cast<size_t>(2 * y + dy + 1) * pitch_in + (2 * x + dx + 1)) * c4n + c4];
    }
    dst[k] = v;
  }
}

// Space-to-depth for a 3x3 stride-2 TF-"same" conv (fp16 build, fused EdgeResidual kernel).  Input: padded layout (pitch
// Win + 2, origin (1, 1)).  Output row q' = y' * (Wo + 2) + x', y' in [0, Ho], x' in [0, Wo + 1]; its 4 C channels are the
// four parity planes (py, px) of the 2x2 pixel block: column (2 py + px) * C + c = in(2 y' + py, 2 x' + px)[c], zero beyond
// the image (the extra row / columns are the conv's bottom / right padding).  The stride-2 conv then reads tap (dy, dx) as
// the K window of plane (dy & 1, dx & 1) at the CONSTANT row shift (dy >> 1) * (Wo + 2) + (dx >> 1): the input crosses
// HBM once more instead of 2.25 times (im2col).  `c16` = 16-byte vectors per pixel.
__global__ void s2d_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int Hin, int Win, int c16) {
  const int Ho = Hin / 2, Wo = Win / 2, pitch_in = Win + 2, pitch_o = Wo + 2;
  const int n = blockIdx.y;
  const int per_row = 4 * c16;
  const size_t total = static_cast<size_t>(Ho + 1) * pitch_o * per_row;
  const uint4* src = in + static_cast<size_t>(n) * (Hin + 2) * pitch_in * c16;
  uint4* dst = out + static_cast<size_t>(n) * total;
  for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
       k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(k % per_row);
    const int q = static_cast<int>(k / per_row);
    const int plane = v / c16, part = v - plane * c16;
    const int y = q / pitch_o, x = q - y * pitch_o;
    const int iy = 2 * y + (plane >> 1), ix = 2 * x + (plane & 1);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (iy < Hin && ix < Win) val = src[(static_cast<size_t>(iy + 1) * pitch_in + (ix + 1)) * c16 + part];
    dst[k] = val;
  }
}

// Depthwise 3x3 + bias (folded BN) + SiLU, with the SE squeeze (per-frame channel sums) fused.
// Input pixel (y, x) is row (y + oy) * pitch_in + (x + ox) of the frame.  TF "same" padding: stride 1 pads 1/1,
// stride 2 pads 0/1 -- both are served by ONE zero-bordered SMEM slab ((Hin+2) x (Win+2) pixels x 32 channels),
// so the tap loop has no boundary checks.  grid = (C / 32, frames); block = 8 channel quads x 32 pixel lanes,
// float4 everywhere (a quarter-warp touches one 128-byte pixel row: coalesced in HBM, conflict-free in SMEM).
template <int kStride, int kF, typename T>
__global__ void __launch_bounds__(256, 3) dwconv_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                        float* __restrict__ sums, const float* __restrict__ w /*[9][C]*/,
                                                        const float* __restrict__ bias, int C, int Hin, int Win,
                                                        int pitch_in, int oy, int ox, int rows_in, int wo_shift,
                                                        int n_frames) {
  extern __shared__ __align__(16) unsigned char slab_raw[];  // [kF][(Hin+2)*(Win+2)][32 channels] of T
  T* slab = reinterpret_cast<T*>(slab_raw);
  __shared__ float4 red[32][8];
  const int Ho = Hin / kStride, Wo = Win / kStride, Wp = Win + 2;
  const int hw = Ho * Wo, spix = (Hin + 2) * Wp;
  const int f0 = blockIdx.y * kF;    // first frame of this block (kF frames share the weights and the launch)
  const int cq = threadIdx.x & 7;    // channel quad inside the 32-channel slab
  const int pl = threadIdx.x >> 3;   // pixel lane 0..31
  const int c = blockIdx.x * 32 + cq * 4;
  const bool c_ok = c < C;           // C is a multiple of 4
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    const bool f_ok = f0 + f < n_frames;
    const T* src = in + static_cast<size_t>(f0 + f) * rows_in * C;
    for (int pix = pl; pix < spix; pix += 32) {
      const int yp = pix / Wp, xp = pix - yp * Wp;
      const bool inside = f_ok && c_ok && yp >= 1 && yp <= Hin && xp >= 1 && xp <= Win;
      store4(slab + (f * spix + pix) * 32 + cq * 4,
             inside ? load4(src + (static_cast<size_t>(yp - 1 + oy) * pitch_in + (xp - 1 + ox)) * C + c) : z4);
    }
  }
  float4 wv[9], b4 = z4;
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = c_ok ? *reinterpret_cast<const float4*>(w + static_cast<size_t>(t) * C + c) : z4;
  if (c_ok) b4 = *reinterpret_cast<const float4*>(bias + c);
  __syncthreads();
  // padded coordinate of tap (dy, dx) for output (y, x): stride 1 -> (y + dy, x + dx); stride 2 -> (2y + dy + 1, 2x + dx + 1)
  constexpr int kOff = kStride == 1 ? 0 : 1;
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    if (f0 + f >= n_frames) break;
    float4 acc_sum = z4;
    for (int p = pl; p < hw; p += 32) {
      const int y = wo_shift >= 0 ? (p >> wo_shift) : p / Wo;
      const int x = p - y * Wo;
      const T* base = slab + (f * spix + (y * kStride + kOff) * Wp + x * kStride + kOff) * 32 + cq * 4;
      float4 v = b4;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4 a = load4(base + (dy * Wp + dx) * 32);
          const float4 ww = wv[dy * 3 + dx];
          v.x = fmaf(a.x, ww.x, v.x); v.y = fmaf(a.y, ww.y, v.y); v.z = fmaf(a.z, ww.z, v.z); v.w = fmaf(a.w, ww.w, v.w);
        }
      v.x = silu_t<T>(v.x); v.y = silu_t<T>(v.y); v.z = silu_t<T>(v.z); v.w = silu_t<T>(v.w);
      if (c_ok) store4(out + (static_cast<size_t>(f0 + f) * hw + p) * C + c, v);
      acc_sum.x += v.x; acc_sum.y += v.y; acc_sum.z += v.z; acc_sum.w += v.w;
    }
    red[pl][cq] = acc_sum;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ch = threadIdx.x;  // channel inside the slab
      float s = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) s += reinterpret_cast<const float*>(&red[k][0])[ch];
      if (blockIdx.x * 32 + ch < C) sums[static_cast<size_t>(f0 + f) * C + blockIdx.x * 32 + ch] = s;
    }
    __syncthreads();
  }
}

// Squeeze-excite in two kernels.
//  (1) se_mlp_kernel: scale[n][c] = sigmoid(W2 silu(W1 mean_n + b1) + b2) for KF frames per block, written over the
//      squeeze sums.  The MLP weights (2 * rd * C floats: 520 KB at C = 1248) are read ONCE per KF frames; the earlier
//      fused kernel re-read them in every block of every frame, which cost more L2 traffic than the tensor it scaled.
//  (2) se_scale_kernel: the streaming multiply, in place; grid = (frames, splits).
template <int KF>
__global__ void __launch_bounds__(512) se_mlp_kernel(float* __restrict__ sums /* in: sums, out: scales */,
                                                     const float* __restrict__ w1 /*[rd][C]*/, const float* __restrict__ b1,
                                                     const float* __restrict__ w2t /*[rd][C]*/, const float* __restrict__ b2,
                                                     int C, int rd, float inv_hw, int n_frames) {
  extern __shared__ float sm[];  // mean[KF][C] + r[KF][rdp], rdp = rd rounded up to 4 (zero padded)
  float* mean = sm;
  float* r = sm + KF * C;
  const int rdp = (rd + 3) & ~3;
  const int n0 = blockIdx.x * KF;
  {
    // the block's KF rows of sums are contiguous: 16-byte loads, several in flight per thread (one scalar load per
    // iteration made this staging loop a chain of ~20-40 exposed L2 round trips: 28 % of the kernel's stall samples)
    const int valid4 = (min(KF, n_frames - n0) * C) >> 2, total4 = (KF * C) >> 2;   // C % 4 == 0
    const float4* src = reinterpret_cast<const float4*>(sums + static_cast<size_t>(n0) * C);
    float4* dst = reinterpret_cast<float4*>(mean);
#pragma unroll 4
    for (int i = threadIdx.x; i < total4; i += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < valid4) v = src[i];
      dst[i] = make_float4(v.x * inv_hw, v.y * inv_hw, v.z * inv_hw, v.w * inv_hw);
    }
  }
  for (int i = threadIdx.x; i < KF * rdp; i += blockDim.x) r[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int c4n = C >> 2;
  // FC1: a warp takes TWO rows of W1 at a time and four k steps per iteration (eight independent 16-byte weight loads in
  // flight: the loop is L2-latency bound); the KF float4 of means a lane reads from SMEM per k serve both rows.
  for (int j0 = 2 * warp; j0 < rd; j0 += 2 * nwarps) {
    const float4* wr[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) wr[q] = reinterpret_cast<const float4*>(w1 + static_cast<size_t>(min(j0 + q, rd - 1)) * C);
    float a[2][KF];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int f = 0; f < KF; ++f) a[q][f] = 0.f;
    for (int k0 = lane; k0 < c4n; k0 += 128) {
      float4 wv[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q)
          wv[i][q] = k0 + 32 * i < c4n ? __ldg(wr[q] + k0 + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + 32 * i < c4n ? k0 + 32 * i : lane;   // (past the end the weights are zero)
#pragma unroll
        for (int f = 0; f < KF; ++f) {
          const float4 mv = reinterpret_cast<const float4*>(mean + f * C)[k];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            a[q][f] = fmaf(wv[i][q].x, mv.x, a[q][f]); a[q][f] = fmaf(wv[i][q].y, mv.y, a[q][f]);
            a[q][f] = fmaf(wv[i][q].z, mv.z, a[q][f]); a[q][f] = fmaf(wv[i][q].w, mv.w, a[q][f]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int f = 0; f < KF; ++f) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[q][f] += __shfl_xor_sync(0xffffffffu, a[q][f], o);
      }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (j0 + q < rd) {
          const float bj = b1[j0 + q];
#pragma unroll
          for (int f = 0; f < KF; ++f) {
            const float v = a[q][f] + bj;
            r[f * rdp + j0 + q] = v / (1.f + expf(-v));
          }
        }
    }
  }
  __syncthreads();
  // FC2: a thread owns channel c; four hidden units per step (one broadcast LDS.128 of r per frame and step)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a[KF];
    const float bc = b2[c];
#pragma unroll
    for (int f = 0; f < KF; ++f) a[f] = bc;
    for (int j = 0; j < rdp; j += 16) {   // sixteen independent weight loads in flight (the loop is L2-latency bound)
      float w[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) w[u] = j + u < rd ? __ldg(w2t + static_cast<size_t>(j + u) * C + c) : 0.f;
#pragma unroll
      for (int f = 0; f < KF; ++f) {
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4)
          if (j + 4 * u4 < rdp) {
            const float4 rr = *reinterpret_cast<const float4*>(r + f * rdp + j + 4 * u4);
            a[f] = fmaf(w[4 * u4], rr.x, a[f]); a[f] = fmaf(w[4 * u4 + 1], rr.y, a[f]);
            a[f] = fmaf(w[4 * u4 + 2], rr.z, a[f]); a[f] = fmaf(w[4 * u4 + 3], rr.w, a[f]);
          }
      }
    }
#pragma unroll
    for (int f = 0; f < KF; ++f)
      if (n0 + f < n_frames) sums[static_cast<size_t>(n0 + f) * C + c] = 1.f / (1.f + expf(-a[f]));
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) se_scale_kernel(T* __restrict__ x, const float* __restrict__ scales, int C, int hw) {
  extern __shared__ float scale[];  // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) scale[c] = scales[static_cast<size_t>(n) * C + c];
  __syncthreads();
  const int c4n = C >> 2;
  // this block's pixels of frame n
  const int p0 = static_cast<int>((static_cast<long long>(hw) * blockIdx.y) / gridDim.y);
  const int p1 = static_cast<int>((static_cast<long long>(hw) * (blockIdx.y + 1)) / gridDim.y);
  T* xf = x + (static_cast<size_t>(n) * hw + p0) * C;
  const int total4 = (p1 - p0) * c4n;
  for (int k = threadIdx.x; k < total4; k += blockDim.x) {
    const int c4 = k % c4n;
    float4 v = load4(xf + 4 * static_cast<size_t>(k));
    const float4 sc = reinterpret_cast<const float4*>(scale)[c4];
    v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
    store4(xf + 4 * static_cast<size_t>(k), v);
  }
}

// feats[dst(n)][c] = mean_p x[n][p][c]
__global__ void gap_kernel(const float* __restrict__ x, const int32_t* __restrict__ fmap, float* __restrict__ feats,
                           int hw, int C, int feat_ld) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < hw; ++p) s += x[(static_cast<size_t>(n) * hw + p) * C + c];
    feats[static_cast<size_t>(fmap ? fmap[n] : n) * feat_ld + c] = s / static_cast<float>(hw);
  }
}

}  // namespace

// `half` selects the activation element type: 0 = float, 1 = __half (`out` / `in` are then __half buffers).
namespace {
constexpr int kStemRows = 8;
// the R-rows kernel when the geometry and the frame alignment allow 4-pixel loads, else the one-row kernel
template <bool kU8, typename T>
int launch_stem(const void* frames, const int32_t* fmap, const float* mask, const float2* norm, T* out, const float* w,
                const float* bias, int n, int H, int W, cudaStream_t st) {
  const size_t frame_bytes = static_cast<size_t>(H) * W * (kU8 ? 1 : 4);
  const bool rows_ok = W % 4 == 0 && (H / 2) % kStemRows == 0 && H % 2 == 0 && (2 * kStemRows + 1) * (W / 4) <= 6 * 256 &&
                       (reinterpret_cast<uintptr_t>(frames) % (kU8 ? 4 : 16)) == 0 && frame_bytes % 16 == 0 &&
                       (!mask || reinterpret_cast<uintptr_t>(mask) % 16 == 0);
  if (rows_ok) {
    dim3 grid(H / 2 / kStemRows, n);
    const size_t sm = (2 * kStemRows + 1) * (W + 1) * sizeof(float);
    stem_rows_kernel<kU8, T, kStemRows><<<grid, 256, sm, st>>>(frames, fmap, mask, norm, out, w, bias, H, W);
  } else {
    dim3 grid(H / 2 + 2, n);
    const size_t sm = 3 * (W + 1) * sizeof(float);
    stem_kernel<kU8, T><<<grid, 256, sm, st>>>(frames, fmap, mask, norm, out, w, bias, H, W);
  }
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}
}  // namespace

int enc_stem(const float* frames, const int32_t* fmap, void* out, int half, const float* w, const float* bias, int n,
             int H, int W, cudaStream_t st) {
  if (half) return launch_stem<false, __half>(frames, fmap, nullptr, nullptr, static_cast<__half*>(out), w, bias, n, H, W, st);
  return launch_stem<false, float>(frames, fmap, nullptr, nullptr, static_cast<float*>(out), w, bias, n, H, W, st);
}

// uint8 ingest: per-frame min-max (after the optional articulator mask) fused into the stem load.
int enc_stem_u8(const uint8_t* frames, const int32_t* fmap, const float* mask, float2* norm, void* out, int half,
                const float* w, const float* bias, int n, int H, int W, cudaStream_t st) {
  frame_minmax_kernel<<<n, 256, 0, st>>>(frames, fmap, mask, norm, H * W);
  M2S_CUDA_OK(cudaGetLastError());
  if (half) return launch_stem<true, __half>(frames, fmap, mask, norm, static_cast<__half*>(out), w, bias, n, H, W, st);
  return launch_stem<true, float>(frames, fmap, mask, norm, static_cast<float*>(out), w, bias, n, H, W, st);
}

// `ld` counts elements of `esize` bytes (rows are multiples of 16 bytes).
int enc_zero_rows(void* buf, int esize, int n, int rows_per_frame, int ld, int head_rows, int tail_start,
                  cudaStream_t st) {
  const int ld4 = ld * esize / 16;
  const int nz = (head_rows + rows_per_frame - tail_start) * ld4;
  dim3 grid((nz + 255) / 256, n);
  zero_rows_kernel<<<grid, 256, 0, st>>>(static_cast<float*>(buf), rows_per_frame, ld4, head_rows, tail_start);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

int enc_zero_cols(void* buf, int esize, int n, int rows_per_frame, int ld, int pitch, int H, cudaStream_t st) {
  const int ld4 = ld * esize / 16;
  dim3 grid((H * 2 * ld4 + 255) / 256, n);
  zero_cols_kernel<<<grid, 256, 0, st>>>(static_cast<float*>(buf), rows_per_frame, ld4, pitch, H);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

int enc_im2col_s2(const void* in, void* col, int esize, int n, int Hin, int Win, int C, cudaStream_t st) {
  const int c4n = C * esize / 16;
  const size_t total = static_cast<size_t>(Hin / 2) * (Win / 2 + 2) * 9 * c4n;
  dim3 grid(static_cast<unsigned>((total + 255) / 256), n);
  im2col_s2_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(in), static_cast<float*>(col), Hin, Win, c4n);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

// rows per frame of the space-to-depth tensor: (Hin / 2 + 1) * (Win / 2 + 2)
int enc_s2d(const void* in, void* out, int esize, int n, int Hin, int Win, int C, cudaStream_t st) {
  if ((C * esize) % 16) return fail(M2S_ERR_UNSUPPORTED, "space-to-depth: %d channels x %d bytes not a multiple of 16", C, esize);
  const int c16 = C * esize / 16;
  const size_t total = static_cast<size_t>(Hin / 2 + 1) * (Win / 2 + 2) * 4 * c16;
  dim3 grid(static_cast<unsigned>((total + 255) / 256), n);
  s2d_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), Hin, Win, c16);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

namespace {
template <typename T>
int dwconv_launch(const T* in, T* out, float* sums, const float* w, const float* bias, int n, int C, int Hin, int Win,
                  int pitch_in, int oy, int ox, int rows_in, int stride, cudaStream_t st) {
  const size_t slab = static_cast<size_t>(Hin + 2) * (Win + 2) * 32 * sizeof(T);
  if (slab > 200 * 1024) return fail(M2S_ERR_UNSUPPORTED, "depthwise input %dx%d too large for the SMEM slab", Hin, Win);
  if (stride != 1 && stride != 2) return fail(M2S_ERR_UNSUPPORTED, "depthwise stride %d", stride);
  // small images: several frames per block (amortises weight loads / block launch; more pixels per thread)
  const int kf = (stride == 1 && slab * 4 <= 64 * 1024) ? 4 : 1;
  using Fn = void (*)(const T*, T*, float*, const float*, const float*, int, int, int, int, int, int, int, int, int);
  Fn fn = stride == 2 ? dwconv_kernel<2, 1, T> : (kf == 4 ? dwconv_kernel<1, 4, T> : dwconv_kernel<1, 1, T>);
  static PerDeviceOnce attr_once;   // (one per instantiation T)
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(dwconv_kernel<1, 1, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    M2S_CUDA_OK(cudaFuncSetAttribute(dwconv_kernel<1, 4, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    M2S_CUDA_OK(cudaFuncSetAttribute(dwconv_kernel<2, 1, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return M2S_OK;
  }));
  const int wo = Win / stride;
  int wo_shift = -1;
  for (int k = 0; k < 16; ++k)
    if ((1 << k) == wo) wo_shift = k;
  dim3 grid((C + 31) / 32, (n + kf - 1) / kf);
  fn<<<grid, 256, slab * kf, st>>>(in, out, sums, w, bias, C, Hin, Win, pitch_in, oy, ox, rows_in, wo_shift, n);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}
}  // namespace

bool dwconv_tma_supported(int half, int C, int H, int W);
int dwconv_tma(const void* in, void* out, int half, float* sums, const float* w, const float* bias, int n, int C, int H,
               int W, int pitch_in, int rows_in, int oy, int ox, cudaStream_t st);

bool dwconv_s2_tma_supported(int C, int Hin, int Win);
int dwconv_s2_tma(const void* in, void* out, float* sums, const float* w, const float* bias, int n, int C, int Hin, int Win,
                  int pitch_in, int rows_in, int oy, int ox, cudaStream_t st);

int enc_dwconv(const void* in, void* out, int half, float* sums, const float* w, const float* bias, int n, int C, int Hin,
               int Win, int pitch_in, int oy, int ox, int rows_in, int stride, cudaStream_t st) {
  static const bool use_tma = !(std::getenv("M2S_DWCONV_TMA") && std::atoi(std::getenv("M2S_DWCONV_TMA")) == 0);
  if (stride == 1 && use_tma && dwconv_tma_supported(half, C, Hin, Win))
    return dwconv_tma(in, out, half, sums, w, bias, n, C, Hin, Win, pitch_in, rows_in, oy, ox, st);
  const char* s2 = std::getenv("M2S_DWCONV_S2_TMA");   // (read per call: the A/B switch of tests/test_mbconv_gpu.py)
  if (stride == 2 && half && use_tma && !(s2 && std::atoi(s2) == 0) && dwconv_s2_tma_supported(C, Hin, Win))
    return dwconv_s2_tma(in, out, sums, w, bias, n, C, Hin, Win, pitch_in, rows_in, oy, ox, st);
  if (half)
    return dwconv_launch(static_cast<const __half*>(in), static_cast<__half*>(out), sums, w, bias, n, C, Hin, Win,
                         pitch_in, oy, ox, rows_in, stride, st);
  return dwconv_launch(static_cast<const float*>(in), static_cast<float*>(out), sums, w, bias, n, C, Hin, Win, pitch_in,
                       oy, ox, rows_in, stride, st);
}

// Squeeze-excite: batched MLP (scales written over the squeeze sums), then the streaming in-place multiply.
// SE MLP alone: sums[n][C] (squeeze sums over hw pixels) -> excite scales, in place
int enc_se_mlp(float* sums, const float* w1, const float* b1, const float* w2, const float* b2, int n, int C, int rd, int hw,
               cudaStream_t st) {
  // Frames per MLP block: every block reads the whole weight set (2 * rd * C floats: 520 KB at C = 1248) from L2, so with 8
  // frames per block a 2048-frame chunk pulled 133 MB through L2 per launch (50 us: L2-bound).  16 frames per block when
  // the chunk still fills the machine, 8 otherwise.  (M2S_SE_MLP_KF overrides: 8 or 16.)
  static const int kf_env = std::getenv("M2S_SE_MLP_KF") ? std::atoi(std::getenv("M2S_SE_MLP_KF")) : 0;
  const int kf = kf_env == 8 || kf_env == 16 ? kf_env : (n >= 16 * 128 ? 16 : 8);
  const size_t sm_mlp = (static_cast<size_t>(kf) * C + static_cast<size_t>(kf) * ((rd + 3) & ~3)) * sizeof(float);
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(se_mlp_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    M2S_CUDA_OK(cudaFuncSetAttribute(se_mlp_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    return M2S_OK;
  }));
  if (sm_mlp > 100 * 1024) return fail(M2S_ERR_UNSUPPORTED, "squeeze-excite: %d channels exceed the MLP kernel's SMEM", C);
  if (kf == 16)
    se_mlp_kernel<16><<<(n + 15) / 16, 512, sm_mlp, st>>>(sums, w1, b1, w2, b2, C, rd, 1.f / hw, n);
  else
    se_mlp_kernel<8><<<(n + 7) / 8, 512, sm_mlp, st>>>(sums, w1, b1, w2, b2, C, rd, 1.f / hw, n);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

int enc_se_apply(void* x, int half, float* sums, const float* w1, const float* b1, const float* w2, const float* b2,
                 int n, int C, int rd, int hw, cudaStream_t st) {
  M2S_TRY(enc_se_mlp(sums, w1, b1, w2, b2, n, C, rd, hw, st));
  const size_t bytes = static_cast<size_t>(hw) * C * (half ? 2 : 4);
  int splits = static_cast<int>((bytes + 192 * 1024 - 1) / (192 * 1024));  // ~<= 192 KB of the tensor per block
  if (splits < 1) splits = 1;
  if (splits > hw) splits = hw;
  dim3 grid(n, splits);
  const size_t sm = static_cast<size_t>(C) * sizeof(float);
  if (half)
    se_scale_kernel<<<grid, 1024, sm, st>>>(static_cast<__half*>(x), sums, C, hw);
  else
    se_scale_kernel<<<grid, 1024, sm, st>>>(static_cast<float*>(x), sums, C, hw);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

namespace {
// fmap[off_b + t] = b * frames + t for t < lens[b], off_b = sum_{i<b} lens[i]: the compact list of the valid frames of a
// ragged batch, built on the device (no host table, no synchronisation: the forward stays capturable)
__global__ void __launch_bounds__(256) build_fmap_kernel(const int32_t* __restrict__ lens, int frames,
                                                         int32_t* __restrict__ fmap) {
  __shared__ int part[8];
  const int b = blockIdx.x;
  int s = 0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) s += lens[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  int off = 0;
  for (int w = 0; w < 8; ++w) off += part[w];
  const int len = lens[b];
  for (int t = threadIdx.x; t < len; t += blockDim.x) fmap[off + t] = b * frames + t;
}
}  // namespace

int enc_build_fmap(const int32_t* lens, int batch, int frames, int32_t* fmap, cudaStream_t st) {
  build_fmap_kernel<<<batch, 256, 0, st>>>(lens, frames, fmap);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

int enc_gap(const float* x, const int32_t* fmap, float* feats, int n, int hw, int C, int feat_ld, cudaStream_t st) {
  gap_kernel<<<n, 256, 0, st>>>(x, fmap, feats, hw, C, feat_ld);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

}  // namespace m2s

// Depthwise 3x3 stride-1 conv (+ folded BN bias + SiLU, SE squeeze fused) of the InvertedResidual blocks, fed by TMA.
//
// Reference: timm InvertedResidual.conv_dw / bn2 / SqueezeExcite squeeze as called through
// EffNetV2B2Backbone.forward (mri2speech_code/mri_acoustic_model.py:28-48); topology in SURVEY.md 8a-1.
//
// The layer is pure data movement (0.7 % of the encoder's FLOPs), so the design goal is HBM streaming:
//   * persistent CTAs walk (frame group, 128-byte channel slab) work items;
//   * ONE 4-D TMA box per item brings the whole (H+2) x (W+2) x slab halo tile into SMEM -- the box starts at
//     (-1, -1), so the "same" zero padding is the TMA unit's out-of-bounds fill, and channel tails (C % slab != 0) are
//     zero-filled the same way;
//   * two SMEM stages: the TMA of item i+1 is in flight while item i is computed (mbarrier complete_tx);
//   * each thread computes a 1 x 4 strip of outputs for 4 channels: 18 shared loads per 4 outputs instead of 36;
//   * a pixel's slab (128 B) is read / written by 8 or 16 consecutive threads: coalesced in HBM, conflict-free in SMEM.
// Element type T = float (32-channel slabs) or __half (64-channel slabs; the fp16 build).  Arithmetic is fp32.
#include "m2s_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>

namespace m2s {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// float activations (tf32 / fp32 builds): exact-ish SiLU; fp16 activations: one-MUFU form h + h*tanh(h), h = v/2
// (tanh.approx.f32, relative error 2^-11 = the rounding of the fp16 store that follows)
template <typename T>
__device__ __forceinline__ float fsilu2(float v) {
  if (sizeof(T) == 4) return __fdividef(v, 1.f + __expf(-v));
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 lds4(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stg4(__half* p, float4 v) {
  uint2 u;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(v.w), "f"(v.z));
  *reinterpret_cast<uint2*>(p) = u;
}

struct DwParams {
  int C, H, W;          // channels, output (= input) height / width
  int n_frames, kf;     // frames, frames per work item
  int n_cslabs, n_items;
  int ox, oy;           // input pixel (0,0) inside the tensor map's (x, y) space (1 for zero-bordered inputs)
  uint32_t stage_bytes;
};

// grid = persistent CTAs, block = 256.  SMEM: 2 stages x [kf][H+2][W+2][128 B] + barriers.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap, T* __restrict__ out,
                                                         float* __restrict__ sums, const float* __restrict__ w /*[9][C]*/,
                                                         const float* __restrict__ bias, const DwParams prm) {
  constexpr int kSlab = 128 / sizeof(T);   // channels per slab: 32 (float) or 64 (half)
  constexpr int kQuads = kSlab / 4;        // threads per pixel
  constexpr int kLanes = 256 / kQuads;     // strips in flight per pass
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ float4 red[8][kQuads];
  const uint32_t base = (smem_addr(smem) + 127u) & ~127u;
  uint8_t* gen_base = smem + (base - smem_addr(smem));
  const int tid = threadIdx.x;
  const int quad = tid % kQuads;           // channel quad inside the slab
  const int lane_idx = tid / kQuads;       // strip lane
  const int H = prm.H, W = prm.W, Wp = W + 2;
  const int strips_per_row = W >> 2;
  const int strips_per_frame = H * strips_per_row;
  const int strips = prm.kf * strips_per_frame;
  const int hw = H * W;
  const uint32_t bar0 = smem_addr(&bars[0]);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  __syncthreads();

  auto issue = [&](int item, int stage) {
    const int cs = item % prm.n_cslabs;
    const int fg = item / prm.n_cslabs;
    const uint32_t bar = bar0 + 8u * stage;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(prm.stage_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(base + stage * prm.stage_bytes), "l"(&tmap), "r"(bar), "r"(cs * kSlab), "r"(prm.ox - 1), "r"(prm.oy - 1),
        "r"(fg * prm.kf)
        : "memory");
  };

  int stage = 0;
  uint32_t ph0 = 0u, ph1 = 0u;
  if (tid == 0 && static_cast<int>(blockIdx.x) < prm.n_items) issue(blockIdx.x, 0);
  for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
    const int next = item + gridDim.x;
    if (tid == 0 && next < prm.n_items) issue(next, stage ^ 1);   // the other stage was released by the barrier below
    const int cs = item % prm.n_cslabs;
    const int f0 = (item / prm.n_cslabs) * prm.kf;
    const int c = cs * kSlab + quad * 4;
    const bool c_ok = c < prm.C;
    float4 wv[9], b4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t)
      wv[t] = c_ok ? __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(t) * prm.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (c_ok) b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
    {  // wait for this item's tile
      const uint32_t bar = bar0 + 8u * stage;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(stage ? ph1 : ph0)
            : "memory");
      }
      if (stage) ph1 ^= 1u; else ph0 ^= 1u;
    }
    const T* slab = reinterpret_cast<const T*>(gen_base + stage * prm.stage_bytes);
    // strips: lane_idx, lane_idx + kLanes, ...; a strip = (frame f, row y, 4 outputs x0..x0+3)
    float4 fsum = make_float4(0.f, 0.f, 0.f, 0.f);
    int fsum_frame = -1;
    auto flush = [&](int f) {
      // reduce the per-thread partial sums of frame f over the strip lanes: shuffle inside the warp, SMEM across warps
      float4 v = fsum;
#pragma unroll
      for (int o = kQuads; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
      }
      if ((tid & 31) < kQuads) red[tid >> 5][quad] = v;
      __syncthreads();
      if (tid < kQuads) {
        float4 s = red[0][tid];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
          const float4 r = red[k][tid];
          s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w;
        }
        const int cc = cs * kSlab + tid * 4;
        if (cc < prm.C && f0 + f < prm.n_frames)
          *reinterpret_cast<float4*>(sums + static_cast<size_t>(f0 + f) * prm.C + cc) = s;
      }
      __syncthreads();
    };
    // every thread walks the same number of passes so that the block-wide reductions stay aligned
    const int passes = (strips + kLanes - 1) / kLanes;
    for (int ps = 0; ps < passes; ++ps) {
      const int s = ps * kLanes + lane_idx;
      const bool s_ok = s < strips;
      const int f = (ps * kLanes) / strips_per_frame;   // frame of this pass (kLanes divides strips_per_frame)
      if (f != fsum_frame) {
        if (fsum_frame >= 0) flush(fsum_frame);
        fsum = make_float4(0.f, 0.f, 0.f, 0.f);
        fsum_frame = f;
      }
      if (s_ok) {
        const int fs = s / strips_per_frame;
        const int r = s - fs * strips_per_frame;
        const int y = r / strips_per_row;
        const int x0 = (r - y * strips_per_row) << 2;
        const T* p0 = slab + ((static_cast<size_t>(fs) * (H + 2) + y) * Wp + x0) * kSlab + quad * 4;
        float4 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = b4;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float4 a[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) a[j] = lds4(p0 + (dy * Wp + j) * kSlab);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 ww = wv[dy * 3 + dx];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o[k].x = fmaf(a[k + dx].x, ww.x, o[k].x);
              o[k].y = fmaf(a[k + dx].y, ww.y, o[k].y);
              o[k].z = fmaf(a[k + dx].z, ww.z, o[k].z);
              o[k].w = fmaf(a[k + dx].w, ww.w, o[k].w);
            }
          }
        }
        const bool f_ok = f0 + fs < prm.n_frames;
        T* op = out + (static_cast<size_t>(f0 + fs) * hw + y * W + x0) * prm.C + c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[k].x = fsilu2<T>(o[k].x); o[k].y = fsilu2<T>(o[k].y); o[k].z = fsilu2<T>(o[k].z); o[k].w = fsilu2<T>(o[k].w);
          if (c_ok && f_ok) stg4(op + static_cast<size_t>(k) * prm.C, o[k]);
          fsum.x += o[k].x; fsum.y += o[k].y; fsum.z += o[k].z; fsum.w += o[k].w;
        }
      }
    }
    if (fsum_frame >= 0) flush(fsum_frame);   // (ends with __syncthreads: the stage may now be refilled)
    stage ^= 1;
  }
}

// ---- stride 2 (the first block of stages 3 and 5), fp16 activations -------------------------------------------------
// TF "same" padding of a stride-2 3x3 conv on an even-sized input pads only the right / bottom edge: output (y, x) reads
// input rows 2y .. 2y+2, columns 2x .. 2x+2.  One TMA box brings (2 R + 1) input rows x (W + 1) columns x a 64-channel
// slab (column W / row H come from the tensor's zero border or from the TMA unit's out-of-bounds fill); two SMEM stages.
// A WARP owns a 4 x 4 block of outputs and its 32 lanes own the slab's 32 channel pairs: an LDS.32 of the warp reads one
// pixel's 128 bytes (one wavefront), the 9 tap weights of a lane's two channels live in registers, a pixel's 64 outputs
// leave as one 128-byte line.  fp32 arithmetic, bias first, taps in (dy, dx) order (= dwconv_kernel<2, 1, __half>).
struct DwS2Params {
  int C, Ho, Wo;           // channels, output height / width
  int n_frames, kf;        // frames, frames per load (small images)
  int r_out;               // output rows per load
  int n_sub;               // loads per (frame group, slab) item: Ho / r_out
  int n_cslabs, n_items;
  int ox, oy;              // input pixel (0, 0) inside the tensor map's (x, y) space
  uint32_t stage_bytes;
};

__device__ __forceinline__ unsigned long long pk2f(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2u(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// grid = persistent CTAs, block = 256 (8 warps = 8 blocks of 4 x 4 outputs per load)
__global__ void __launch_bounds__(256) dwconv_s2_tma_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out,
                                                            float* __restrict__ sums, const float* __restrict__ w /*[9][C]*/,
                                                            const float* __restrict__ bias, const DwS2Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ float2 red[8][32];
  const uint32_t base = (smem_addr(smem) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wo = prm.Wo, Wi1 = 2 * prm.Wo + 1;                   // box width in pixels
  const int rows_box = 2 * prm.r_out + 1;
  const int hw = prm.Ho * Wo;
  const uint32_t bar0 = smem_addr(&bars[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  __syncthreads();
  // load l = (item, sub): item = (frame group, slab)
  const int my_items = (prm.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int my_loads = my_items * prm.n_sub;
  auto issue = [&](int l) {
    const int item = blockIdx.x + (l / prm.n_sub) * gridDim.x, sub = l % prm.n_sub;
    const int cs = item % prm.n_cslabs, fg = item / prm.n_cslabs;
    const uint32_t bar = bar0 + 8u * (l & 1);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(prm.stage_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(base + (l & 1) * prm.stage_bytes), "l"(&tmap), "r"(bar), "r"(cs * 64), "r"(prm.ox),
        "r"(prm.oy + 2 * sub * prm.r_out), "r"(fg * prm.kf)
        : "memory");
  };
  // this warp's 4 x 4 block inside a load: blocks are laid out (frame of the group, block row, block column)
  const int bpr = Wo >> 2;                            // blocks per output row
  const int bpf = (prm.r_out >> 2) * bpr;             // blocks per frame per load
  const int bf = warp / bpf, bb = warp % bpf;
  const int by = (bb / bpr) * 4, bx = (bb % bpr) * 4; // first output row (inside the load) / column of the block
  const bool warp_on = bf < prm.kf;
  uint32_t ph[2] = {0u, 0u};
  if (tid == 0 && my_loads > 0) issue(0);
  unsigned long long fsum = pk2f(0.f, 0.f);
  unsigned long long wv[9], b2 = pk2f(0.f, 0.f);
  int c = 0;
  bool c_ok = false;
  for (int l = 0; l < my_loads; ++l) {
    if (tid == 0 && l + 1 < my_loads) issue(l + 1);   // the other stage was released by the barrier at the end of l - 1
    const int item = blockIdx.x + (l / prm.n_sub) * gridDim.x, sub = l % prm.n_sub;
    const int cs = item % prm.n_cslabs, f0 = (item / prm.n_cslabs) * prm.kf;
    if (sub == 0) {
      c = cs * 64 + 2 * lane;
      c_ok = c < prm.C;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float2 v = make_float2(0.f, 0.f);
        if (c_ok) v = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(t) * prm.C + c));
        wv[t] = pk2f(v.x, v.y);
      }
      float2 v = make_float2(0.f, 0.f);
      if (c_ok) v = __ldg(reinterpret_cast<const float2*>(bias + c));
      b2 = pk2f(v.x, v.y);
      fsum = pk2f(0.f, 0.f);
    }
    {
      const uint32_t bar = bar0 + 8u * (l & 1);
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(ph[l & 1])
            : "memory");
      }
      ph[l & 1] ^= 1u;
    }
    if (warp_on) {
      // window of the block: input rows 2 by .. 2 by + 8, columns 2 bx .. 2 bx + 8 of frame bf of the box
      const uint32_t win = base + (l & 1) * prm.stage_bytes +
                           ((bf * rows_box + 2 * by) * Wi1 + 2 * bx) * 128 + lane * 4;
      unsigned long long res[4][4];
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int k = 0; k < 4; ++k) res[y][k] = b2;
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        unsigned long long a[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          uint32_t v;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(win + (r * Wi1 + j) * 128));
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
          a[j] = pk2f(f.x, f.y);
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int dy = r - 2 * y;
          if (dy < 0 || dy > 2) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) res[y][k] = ffma2u(a[2 * k + dx], wv[dy * 3 + dx], res[y][k]);
        }
      }
      const bool st_ok = c_ok && f0 + bf < prm.n_frames;
      __half* op = out + (static_cast<size_t>(f0 + bf) * hw + (sub * prm.r_out + by) * Wo + bx) * prm.C + c;
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(res[y][k]));
          lo = fsilu2<__half>(lo);
          hi = fsilu2<__half>(hi);
          float slo, shi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(slo), "=f"(shi) : "l"(fsum));
          fsum = pk2f(slo + lo, shi + hi);
          if (st_ok) {
            uint32_t h2;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(hi), "f"(lo));
            *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(y * Wo + k) * prm.C) = h2;
          }
        }
    }
    if (sub == prm.n_sub - 1) {
      // squeeze sums of the item: per frame of the group, over the warps that hold its blocks
      float slo, shi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(slo), "=f"(shi) : "l"(fsum));
      red[warp][lane] = make_float2(slo, shi);
      __syncthreads();
      if (tid < 32 * prm.kf) {
        const int fr = tid >> 5, ln = tid & 31;
        float2 sacc = make_float2(0.f, 0.f);
        for (int k = 0; k < bpf; ++k) {
          const float2 v = red[fr * bpf + k][ln];
          sacc.x += v.x; sacc.y += v.y;
        }
        const int cc = cs * 64 + 2 * ln;
        if (cc < prm.C && f0 + fr < prm.n_frames)
          *reinterpret_cast<float2*>(sums + static_cast<size_t>(f0 + fr) * prm.C + cc) = sacc;
      }
    }
    __syncthreads();   // the stage may now be refilled (and `red` reused)
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn4() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

template <typename T>
int launch(const T* in, T* out, float* sums, const float* w, const float* bias, int n, int C, int H, int W, int pitch_in,
           int rows_in, int oy, int ox, cudaStream_t st) {
  EncodeTiledFn enc = encode_fn4();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  constexpr int kSlab = 128 / sizeof(T);
  constexpr int kLanes = 256 / (kSlab / 4);
  DwParams prm{};
  prm.C = C; prm.H = H; prm.W = W; prm.n_frames = n; prm.ox = ox; prm.oy = oy;
  const int spf = H * (W / 4);  // strips per frame
  // frames per item: fill the strip lanes (small images), keep a stage <= ~52 KB
  int kf = 1;
  const size_t frame_bytes = static_cast<size_t>(H + 2) * (W + 2) * 128;
  while (kf < 8 && spf * kf < 4 * kLanes && (kf * 2) * frame_bytes <= 56 * 1024 && kf * 2 <= n) kf *= 2;
  // the per-frame reduction walks whole passes of kLanes strips: a frame must be a whole number of passes
  if (spf % kLanes != 0)
    return fail(M2S_ERR_UNSUPPORTED, "depthwise %dx%d: strip count does not tile the block", H, W);
  prm.kf = kf;
  prm.n_cslabs = (C + kSlab - 1) / kSlab;
  prm.n_items = prm.n_cslabs * ((n + kf - 1) / kf);
  prm.stage_bytes = static_cast<uint32_t>(kf * frame_bytes);
  const int height_in = rows_in / pitch_in;
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(pitch_in), static_cast<cuuint64_t>(height_in),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * sizeof(T), static_cast<cuuint64_t>(pitch_in) * C * sizeof(T),
                           static_cast<cuuint64_t>(rows_in) * C * sizeof(T)};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kSlab), static_cast<cuuint32_t>(W + 2), static_cast<cuuint32_t>(H + 2),
                       static_cast<cuuint32_t>(kf)};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult cr = enc(&tmap, sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                    const_cast<T*>(in), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "depthwise tensor map failed (%d)", static_cast<int>(cr));
  const size_t smem = 2 * static_cast<size_t>(prm.stage_bytes) + 256;
  static PerDeviceOnce attr_once;   // (one per instantiation T)
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(dwconv_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return M2S_OK;
  }));
  int grid = 2 * sm_count();
  if (grid > prm.n_items) grid = prm.n_items;
  dwconv_tma_kernel<T><<<grid, 256, smem, st>>>(tmap, out, sums, w, bias, prm);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

}  // namespace

// Is the TMA kernel applicable?  (stride 1, width a multiple of 4, C a multiple of 8 (16-byte rows), tile fits SMEM.)
bool dwconv_tma_supported(int half, int C, int H, int W) {
  const int lanes = half ? 16 : 32;
  const int spf = H * (W / 4);
  if (W % 4 || C % 8) return false;
  if (static_cast<size_t>(H + 2) * (W + 2) * 128 > 90 * 1024) return false;
  return spf % lanes == 0;
}

int dwconv_tma(const void* in, void* out, int half, float* sums, const float* w, const float* bias, int n, int C, int H,
               int W, int pitch_in, int rows_in, int oy, int ox, cudaStream_t st) {
  if (half)
    return launch(static_cast<const __half*>(in), static_cast<__half*>(out), sums, w, bias, n, C, H, W, pitch_in, rows_in,
                  oy, ox, st);
  return launch(static_cast<const float*>(in), static_cast<float*>(out), sums, w, bias, n, C, H, W, pitch_in, rows_in, oy,
                ox, st);
}

// stride 2, fp16: blocks of 4 x 4 outputs, 8 per load
bool dwconv_s2_tma_supported(int C, int Hin, int Win) {
  const int Ho = Hin / 2, Wo = Win / 2;
  if (Hin != Win || C % 8 || Wo % 4) return false;
  return Ho == 16 || Ho == 8;
}

int dwconv_s2_tma(const void* in, void* out, float* sums, const float* w, const float* bias, int n, int C, int Hin, int Win,
                  int pitch_in, int rows_in, int oy, int ox, cudaStream_t st) {
  if (!dwconv_s2_tma_supported(C, Hin, Win)) return fail(M2S_ERR_UNSUPPORTED, "stride-2 depthwise %dx%d not supported", Hin, Win);
  EncodeTiledFn enc = encode_fn4();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  DwS2Params prm{};
  prm.C = C; prm.Ho = Hin / 2; prm.Wo = Win / 2; prm.n_frames = n; prm.ox = ox; prm.oy = oy;
  // 8 blocks of 4 x 4 outputs per load: 16 x 16 outputs -> 8 rows of one frame; 8 x 8 outputs -> two whole frames
  if (prm.Ho == 16) { prm.kf = 1; prm.r_out = 8; } else { prm.kf = 2; prm.r_out = 8; }
  prm.n_sub = prm.Ho / prm.r_out;
  prm.n_cslabs = (C + 63) / 64;
  prm.n_items = prm.n_cslabs * ((n + prm.kf - 1) / prm.kf);
  const int rows_box = 2 * prm.r_out + 1, cols_box = Win + 1;
  prm.stage_bytes = static_cast<uint32_t>(prm.kf) * rows_box * cols_box * 128;
  // The map covers the INTERIOR Hin x Win pixels only (base moved to pixel (oy, ox)): column Win and row Hin are then the
  // TMA unit's out-of-bounds zero fill.  (The physical border of a zero-bordered input is NOT zero here: the expand
  // GEMM ran over every row of the padded layout and left silu(bias) in it.)
  const __half* base_in = static_cast<const __half*>(in) + (static_cast<size_t>(oy) * pitch_in + ox) * C;
  prm.ox = 0; prm.oy = 0;
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(Win), static_cast<cuuint64_t>(Hin),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(pitch_in) * C * 2,
                           static_cast<cuuint64_t>(rows_in) * C * 2};
  cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(cols_box), static_cast<cuuint32_t>(rows_box), static_cast<cuuint32_t>(prm.kf)};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base_in), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "stride-2 depthwise tensor map failed (%d)", static_cast<int>(cr));
  const size_t smem = 2 * static_cast<size_t>(prm.stage_bytes) + 256;
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(dwconv_s2_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return M2S_OK;
  }));
  if (smem > 200 * 1024) return fail(M2S_ERR_UNSUPPORTED, "stride-2 depthwise: tile does not fit SMEM");
  int grid = sm_count();
  if (grid > prm.n_items) grid = prm.n_items;
  dwconv_s2_tma_kernel<<<grid, 256, smem, st>>>(tmap, static_cast<__half*>(out), sums, w, bias, prm);
  M2S_CUDA_OK(cudaGetLastError());
  return M2S_OK;
}

}  // namespace m2s

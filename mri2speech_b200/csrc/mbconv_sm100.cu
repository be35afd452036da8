// Fused InvertedResidual (MBConv + squeeze-excite) kernels of the frame-CNN encoder for sm_100a (fp16 build).
//
// Reference: timm InvertedResidual (conv_pw -> bn1 -> SiLU -> conv_dw -> bn2 -> SiLU -> SqueezeExcite -> conv_pwl -> bn3
// [+ shortcut]) as called through EffNetV2B2Backbone.forward (mri2speech_code/mri_acoustic_model.py:28-48); topology in
// SURVEY.md 8a-1 (stages 3-5: 20 blocks, 46 % of the encoder's MACs and 62 % of its time before this file existed).
//
// Two kernels replace five launches (expand GEMM, depthwise, SE MLP, SE scale, project GEMM -> expand_dw, SE MLP, project):
//
//  * mb_expand_dw_kernel -- 1x1 expand GEMM (tcgen05, accumulators in TMEM) whose epilogue IS the depthwise conv: a tile is
//    one 16x16 frame (or two 8x8 frames) x a 64-channel slab.  The epilogue warps apply bias + SiLU to the accumulators and
//    park them as an fp16 (H+2) x (W+2) zero-bordered image in SMEM, run the 3x3 depthwise conv + bias + SiLU from there,
//    reduce the SE squeeze sums of the slab, and hand the 64-channel output tile to ONE TMA store.  The expanded tensor
//    (377 MB per 1024 frames at stage 4) never exists in HBM.
//  * mb_project_kernel -- 1x1 project GEMM with the SE excite scale applied to the A operand in SMEM (8 scaler warps
//    between the TMA producer and the MMA issuer multiply each K block by scale[frame][channel] in place), bias + optional
//    fp32 shortcut in the epilogue, outputs through TMA stores.  The separate read-modify-write pass over the depthwise
//    output (se_scale_kernel) is gone.
#include "engine_device.cuh"
#include <cuda_fp16.h>
#include <cstdlib>
#include <mutex>

namespace m2s {

using namespace engine;

namespace {

// ---- PTX wrappers local to this file -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issued instruction)
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 pk2u(uint32_t lo, uint32_t hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ float lo2(f2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return lo;
}
__device__ __forceinline__ float hi2(f2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return hi;
}
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 fmul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 fadd2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 cvt2(uint32_t h2) {   // fp16 pair -> fp32 pair
  const float2 f = unpack_h2(h2);
  return pk2(f.x, f.y);
}
// SiLU of a pair, the engine's one-MUFU form (fast_silu): h = v / 2, h + h * tanh(h)
__device__ __forceinline__ f2 silu2(f2 v) {
  const f2 h = fmul2(v, pk2(0.5f, 0.5f));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(lo2(h)));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(hi2(h)));
  return ffma2(h, pk2(t0, t1), h);
}

constexpr int kSlab = 64;                  // channels per depthwise slab = one 128-byte fp16 row
constexpr int kPixPitch = 144;             // bytes per pixel of the SMEM image: 128 + 16, so that the row-per-lane writes of
                                           // phase 1 (8 consecutive pixels per quarter-warp) hit 8 different bank groups;
                                           // the 8 threads of a pixel read its 128 bytes contiguously in phase 2
constexpr uint32_t kBTileBytes = 64 * 128; // one K block of a 64-row weight slab

// =====================================================================================================================
// Kernel E: expand 1x1 GEMM -> bias + SiLU -> depthwise 3x3 (zero "same" padding, stride 1) -> bias + SiLU -> squeeze
// =====================================================================================================================
struct ExpandDwParams {
  int n_frames, H, W, hw;
  int c_in, c_mid;
  int kblocks, slabs;
  int fpt;                 // frames per tile (M = fpt * hw = 128 * HALVES rows)
  int n_tiles;
  int nb;                  // weight-slab ring depth
  uint32_t a_bytes;        // kblocks * M * 128
  uint32_t b_stage_bytes;  // kblocks * 8 KB
  uint32_t sp_bytes;       // one SMEM image: fpt * (H+2) * (W+2) * kPixPitch, rounded up to 1 KB
  uint32_t bias_bytes;     // staged bn1 bias vector, c_mid floats rounded up to 1 KB
  const float* bias1;      // [c_mid] folded bn1 bias of the expand conv
  const float* dw_w32;     // [9][c_mid] depthwise weights (bn2 scale folded), fp32
  const float* dw_b;       // [c_mid] folded bn2 bias
  __half* out;             // [n_frames * hw][c_mid] depthwise output
  float* sums;             // [n_frames][c_mid] SE squeeze sums (over the H*W outputs)
  uint32_t idesc;
  uint64_t desc_hi;
};

constexpr int kEThreads = 64 + 512;  // TMA warp, MMA warp, two groups of 8 epilogue / depthwise warps

// HALVES: 128-row MMA halves per tile (2: one 16x16 frame, 1: two 8x8 frames).
//
// Epilogue organisation (what the first version's profile asked for, profiles/README.md round 2): the 16 epilogue warps
// form TWO groups of 8 that take alternate slabs, each with its own accumulator buffer and SMEM image, so one group's
// TMEM / MUFU-heavy phase 1 overlaps the other's SMEM / FMA-heavy phase 2.  In phase 2 a WARP owns a 4 x 4 pixel block and
// its 32 lanes own the slab's 32 channel pairs: an LDS.32 of the warp reads one pixel's 128 bytes (one wavefront, no
// conflicts whatever the pixel pitch), the 9 tap weights of a lane's two channels live in registers for the whole slab, the
// 6 x 6 input window is read once per block (2.25 reads per output instead of 4.5), and a pixel's 64 outputs leave as
// one 128-byte line.
template <int HALVES>
__global__ void __launch_bounds__(kEThreads, 1)
mb_expand_dw_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ ExpandDwParams prm) {
  constexpr int M = 128 * HALVES;
  constexpr int W = HALVES == 2 ? 16 : 8, Wp = W + 2, HW = W * W;
  constexpr int pp_frame = Wp * Wp;
  constexpr int FPT = M / HW;             // frames per tile
  constexpr int BPW = M / 16 / 8;         // 4 x 4 pixel blocks per warp of a group (2 or 1)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + prm.a_bytes;
  const uint32_t sp_base = b_base + prm.nb * prm.b_stage_bytes;   // two images
  const uint32_t bias_base = sp_base + 2 * prm.sp_bytes;          // bn1 bias, all slabs
  const uint32_t red_base = bias_base + prm.bias_bytes;           // [2 groups][8 warps][64] floats
  const uint32_t bar_base = red_base + 2 * 8 * 64 * 4;
  const uint32_t a_full = bar_base, a_empty = bar_base + 8;
  auto b_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 16u + 8u * (4 + s); };
  auto acc_full = [&](int s) { return bar_base + 16u + 8u * (8 + s); };
  auto acc_empty = [&](int s) { return bar_base + 16u + 8u * (10 + s); };
  const uint32_t tmem_slot = bar_base + 16u + 8u * 12;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < prm.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  // the zero borders of the two SMEM images are written once: the epilogue only ever writes interior pixels
  for (uint32_t off = threadIdx.x * 16u; off < 2 * prm.sp_bytes; off += kEThreads * 16u) sts128(sp_base + off, 0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < static_cast<int>(prm.bias_bytes / 4); i += kEThreads) {
    const float v = i < prm.c_mid ? __ldg(prm.bias1 + i) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_base + 4u * i), "f"(v) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int kblocks = prm.kblocks, slabs = prm.slabs;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int sb = 0;
    uint32_t pb = 0, pa = 0;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      mbar_wait(a_empty, pa ^ 1);
      if (elect_one()) {
        mbar_expect_tx(a_full, prm.a_bytes);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(a_base + kb * (M * 128), &tm_x, a_full, kb * 64, tile * M);
      }
      __syncwarp();
      pa ^= 1;
      for (int sl = 0; sl < slabs; ++sl) {
        mbar_wait(b_empty(sb), pb ^ 1);
        if (elect_one()) {
          mbar_expect_tx(b_full(sb), prm.b_stage_bytes);
          for (int kb = 0; kb < kblocks; ++kb)
            tma_load_2d(b_base + sb * prm.b_stage_bytes + kb * kBTileBytes, &tm_w, b_full(sb), kb * 64, sl * kSlab);
        }
        __syncwarp();
        if (++sb == prm.nb) { sb = 0; pb ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sb = 0, acc = 0;
    uint32_t pb = 0, pa = 0, pacc = 0;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      mbar_wait(a_full, pa);
      pa ^= 1;
      for (int sl = 0; sl < slabs; ++sl) {
        mbar_wait(b_full(sb), pb);
        mbar_wait(acc_empty(acc), pacc ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_tile = b_base + sb * prm.b_stage_bytes;
#pragma unroll
          for (int h = 0; h < HALVES; ++h) {
            const uint32_t d = tmem_base + acc * (HALVES * kSlab) + h * kSlab;
            for (int kb = 0; kb < kblocks; ++kb) {
              const int rem = prm.c_in - kb * 64;
              const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
              const uint64_t da = prm.desc_hi | (((a_base + kb * (M * 128) + h * (128 * 128)) & 0x3FFFF) >> 4);
              const uint64_t db = prm.desc_hi | (((b_tile + kb * kBTileBytes) & 0x3FFFF) >> 4);
              mma_f16_k4(d, da, db, prm.idesc, kb ? 1u : 0u, ksteps);
            }
          }
          tc_commit(b_empty(sb));
          tc_commit(acc_full(acc));
          if (sl == slabs - 1) tc_commit(a_empty);
        }
        __syncwarp();
        if (++sb == prm.nb) { sb = 0; pb ^= 1; }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
    }
  } else {
    // ===================== epilogue groups: SiLU -> SMEM image -> depthwise 3x3 -> SiLU -> squeeze + stores =====================
    // All arithmetic is the fp32 arithmetic of the launches this kernel replaces (engine SiLU epilogue, dwconv_tma_kernel:
    // bias first, taps in (dy, dx) order), issued as packed f32x2 instructions (FFMA2 / FMUL2 / FADD2).
    const int grp = (warp - 2) >> 3;        // group 0 / 1: slab sequence numbers n with n % 2 == grp
    const int gw = (warp - 2) & 7;          // warp inside the group
    const int gt = gw * 32 + lane;          // thread inside the group, 0..255
    const int quad = warp & 3;              // TMEM lane quadrant of this warp
    const uint32_t sp_buf = sp_base + grp * prm.sp_bytes;
    const uint32_t red_buf = red_base + grp * (8 * 64 * 4);
    const uint32_t acc_col = tmem_base + grp * (HALVES * kSlab) + (static_cast<uint32_t>(quad * 32) << 16);
    // phase-1 role.  HALVES == 2: warps (quad, half = gw >> 2) drain 32 rows x 64 columns; HALVES == 1: warps (quad, column
    // half = gw >> 2) drain 32 rows x 32 columns.  lane = accumulator row = pixel.
    const int p1_sel = gw >> 2;
    uint32_t sp_w;
    {
      const int row = (HALVES == 2 ? p1_sel * 128 : 0) + quad * 32 + lane;
      const int fr = row / HW, pix = row % HW;
      sp_w = sp_buf + (fr * pp_frame + (pix / W + 1) * Wp + (pix % W) + 1) * kPixPitch + (HALVES == 2 ? 0 : p1_sel * 64);
    }
    const uint32_t p1_tmem = acc_col + (HALVES == 2 ? p1_sel * kSlab : p1_sel * 32);
    const int p1_col = HALVES == 2 ? 0 : p1_sel * 32;
    constexpr int P1_BATCHES = HALVES == 2 ? 2 : 1;   // batches of 32 columns per thread
    // phase-2 role: block b = gw + 8 i of the tile, lane = channel pair
    uint32_t sp_r[BPW];
    int out_row[BPW], blk_frame[BPW];
#pragma unroll
    for (int i = 0; i < BPW; ++i) {
      const int b = gw + 8 * i;
      const int fr = b / (HW / 16), bb = b % (HW / 16);
      const int y0 = (bb / (W / 4)) * 4, x0 = (bb % (W / 4)) * 4;
      blk_frame[i] = fr;
      sp_r[i] = sp_buf + (fr * pp_frame + y0 * Wp + x0) * kPixPitch + lane * 4;
      out_row[i] = fr * HW + y0 * W + x0;
    }
    uint32_t pacc = 0;
    int n = 0;   // slab sequence number of this CTA
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      const int f0 = tile * FPT;
      for (int sl = 0; sl < slabs; ++sl, ++n) {
        if ((n & 1) != grp) continue;
        const int cs = sl * kSlab;
        const int c = cs + 2 * lane;
        const bool c_ok = c < prm.c_mid;
        // the slab's depthwise weights / bias of this lane's channel pair: requested now, used in phase 2
        f2 wv[9], b2;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float2 v = make_float2(0.f, 0.f);
          if (c_ok) v = __ldg(reinterpret_cast<const float2*>(prm.dw_w32 + static_cast<size_t>(t) * prm.c_mid + c));
          wv[t] = pk2(v.x, v.y);
        }
        {
          float2 v = make_float2(0.f, 0.f);
          if (c_ok) v = __ldg(reinterpret_cast<const float2*>(prm.dw_b + c));
          b2 = pk2(v.x, v.y);
        }
        // ---- phase 1: accumulators -> bias + SiLU -> fp16 image in SMEM ----
        mbar_wait(acc_full(grp), pacc);
        tc_fence_after();
#pragma unroll
        for (int bt = 0; bt < P1_BATCHES; ++bt) {
          uint32_t r[32];
          tmem_ld32(p1_tmem + bt * 32, r);
          tmem_ld_wait();
          const uint32_t bsm = bias_base + (cs + p1_col + bt * 32) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // 8 columns -> one 16-byte store
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint4 bv = lds128(bsm + (2 * j + q) * 16);   // (warp-uniform address: broadcast)
              const f2 v0 = silu2(fadd2(pk2u(r[8 * j + 4 * q], r[8 * j + 4 * q + 1]), pk2u(bv.x, bv.y)));
              const f2 v1 = silu2(fadd2(pk2u(r[8 * j + 4 * q + 2], r[8 * j + 4 * q + 3]), pk2u(bv.z, bv.w)));
              o[2 * q] = pack_h2(lo2(v0), hi2(v0));
              o[2 * q + 1] = pack_h2(lo2(v1), hi2(v1));
            }
            sts128(sp_w + bt * 64 + j * 16, o[0], o[1], o[2], o[3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(grp));
        pacc ^= 1;
        named_bar(1 + grp, 256);
        // ---- phase 2: depthwise 3x3 + bias + SiLU from the SMEM image, 4 x 4 outputs per warp-block ----
        f2 fsum = pk2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < BPW; ++i) {
          f2 res[4][4];
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int k = 0; k < 4; ++k) res[y][k] = b2;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            f2 a[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              uint32_t v;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sp_r[i] + (r * Wp + j) * kPixPitch));
              a[j] = cvt2(v);
            }
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              const int dy = r - y;
              if (dy < 0 || dy > 2) continue;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < 4; ++k) res[y][k] = ffma2(a[k + dx], wv[dy * 3 + dx], res[y][k]);
            }
          }
          const bool st_ok = c_ok && f0 + blk_frame[i] < prm.n_frames;
          __half* op = prm.out + (static_cast<size_t>(tile) * M + out_row[i]) * prm.c_mid + c;
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const f2 o = silu2(res[y][k]);
              fsum = fadd2(fsum, o);
              if (st_ok) *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(y * W + k) * prm.c_mid) = pack_h2(lo2(o), hi2(o));
            }
        }
        // squeeze: this warp's sum of the lane's two channels over its block(s); summed across the frame's warps below
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red_buf + (gw * 64 + 2 * lane) * 4), "f"(lo2(fsum)), "f"(hi2(fsum)) : "memory");
        named_bar(1 + grp, 256);
        if (gt < 64 * FPT) {
          const int fr = gt >> 6, ch = gt & 63;
          // 16 x 16: all 8 warps hold blocks of the frame; 8 x 8: blocks 4 fr .. 4 fr + 3 = warps 4 fr .. 4 fr + 3
          const int w0 = HALVES == 2 ? 0 : fr * 4, wn = HALVES == 2 ? 8 : 4;
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < wn; ++k) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(red_buf + ((w0 + k) * 64 + ch) * 4));
            s += v;
          }
          if (cs + ch < prm.c_mid && f0 + fr < prm.n_frames) prm.sums[static_cast<size_t>(f0 + fr) * prm.c_mid + cs + ch] = s;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
// Kernel P: project 1x1 GEMM with the SE scale applied to A in SMEM; bias (+ fp32 shortcut); TMA-store epilogue
// =====================================================================================================================
struct ProjectParams {
  int rows, hw, n_frames;
  int c_mid, c_out, n_pad;
  int kblocks, n_tiles, stages;
  uint32_t a_bytes, stage_bytes;
  const float* scales;   // [n_frames][c_mid] fp32 excite scales (null: plain GEMM)
  const float* bias;     // [c_out]
  const float* res;      // [rows][c_out] fp32 shortcut or null
  const __half* res16;   // ... or the fp16 residual stream (res == null)
  int store32, store16;
  uint32_t idesc;
  uint64_t desc_hi;
};

constexpr int kPScalerWarps = 8;
constexpr int kPThreads = 64 + 32 * kPScalerWarps + 128;   // TMA, MMA, 8 scaler warps, 4 epilogue warps
constexpr int kPMaxStages = 6;
constexpr uint32_t kPStage32 = 32 * 128, kPStage16 = 32 * 64;   // per-warp output staging (fp32 / fp16 unit of 32 x 32)
constexpr uint32_t kPEpiWarpBytes = kPStage32 + 3 * kPStage16;  // ... + two fp16 shortcut tiles

template <int HALVES>
__global__ void __launch_bounds__(kPThreads, 1)
mb_project_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_d32, const __grid_constant__ CUtensorMap tm_d16,
                  const __grid_constant__ CUtensorMap tm_r16, const __grid_constant__ ProjectParams prm) {
  constexpr int M = 128 * HALVES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_base = smem_base;
  const uint32_t st_base = ring_base + prm.stages * prm.stage_bytes;   // 4 warps x (4 KB + 2 KB output staging + 2 x 2 KB shortcut tiles)
  const uint32_t bar_base = st_base + 4 * kPEpiWarpBytes;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto scaled = [&](int s) { return bar_base + 8u * (kPMaxStages + s); };
  auto empty = [&](int s) { return bar_base + 8u * (2 * kPMaxStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (3 * kPMaxStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (3 * kPMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kPMaxStages + 4);
  auto r_full = [&](int w, int b) { return bar_base + 8u * (3 * kPMaxStages + 5 + 2 * w + b); };   // shortcut tile b of epilogue warp w

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_scale = prm.scales != nullptr;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < prm.stages; ++s) { mbar_init(full(s), 1); mbar_init(scaled(s), kPScalerWarps); mbar_init(empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
    for (int w = 0; w < 4; ++w) { mbar_init(r_full(w, 0), 1); mbar_init(r_full(w, 1), 1); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_r16) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_d32) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_d16) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int kblocks = prm.kblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty(s), ph ^ 1);
        if (elect_one()) {
          const uint32_t st = ring_base + s * prm.stage_bytes;
          mbar_expect_tx(full(s), prm.a_bytes + prm.n_pad * 128);
          tma_load_2d(st, &tm_a, full(s), kb * 64, tile * M);
          tma_load_2d(st + prm.a_bytes, &tm_w, full(s), kb * 64, 0);
        }
        __syncwarp();
        if (++s == prm.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int s = 0, acc = 0;
    uint32_t ph = 0, pacc = 0;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      mbar_wait(acc_empty(acc), pacc ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(has_scale ? scaled(s) : full(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = ring_base + s * prm.stage_bytes;
          const int rem = prm.c_mid - kb * 64;
          const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
          const uint64_t db = prm.desc_hi | (((st + prm.a_bytes) & 0x3FFFF) >> 4);
#pragma unroll
          for (int h = 0; h < HALVES; ++h) {
            const uint64_t da = prm.desc_hi | (((st + h * (128 * 128)) & 0x3FFFF) >> 4);
            mma_f16_k4(tmem_base + acc * 256 + h * 128, da, db, prm.idesc, kb ? 1u : 0u, ksteps);
          }
          tc_commit(empty(s));
          if (kb == kblocks - 1) tc_commit(acc_full(acc));
        }
        __syncwarp();
        if (++s == prm.stages) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  } else if (warp < 2 + kPScalerWarps) {
    // ===================== scaler warps: the excite scale, in place in the stage =====================
    // HALVES == 2 (16 x 16 frames: a tile IS a frame): W[n][k] *= scale[frame][k] -- the weight block has at most half the
    // rows of the A tile and one scale row serves the whole block.  HALVES == 1 (8 x 8 frames: two frames per tile):
    // A[row][k] *= scale[frame(row)][k].  Either way one fp16 rounding per product term.  (The first version scaled A
    // everywhere and re-read the scale row for each of its 8 row passes: with the epilogue's row-per-lane shortcut loads
    // that put the LSU data pipe at 72 % of its wavefront rate -- profiles/README.md, round 2 late.)
    if (has_scale) {
      const int t = threadIdx.x - 64;   // 0..255
      const int c8 = t & 7, r0 = t >> 3;  // 16-byte chunk, first row (32 rows per pass)
      int s = 0;
      uint32_t ph = 0;
      constexpr int kSets = HALVES == 2 ? 1 : 4;   // distinct frames a tile can touch (M / hw, hw >= 32)
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        const int row0 = tile * M;
        const int f0 = row0 / prm.hw;
        const int ppf = prm.hw >> 5;                 // 32-row passes per frame
        for (int kb = 0; kb < kblocks; ++kb) {
          const int c = kb * 64 + c8 * 8;
          const bool c_ok = c < prm.c_mid;
          float4 sa[kSets], sb2[kSets];
#pragma unroll
          for (int i = 0; i < kSets; ++i) {
            const int f = f0 + i;
            if (c_ok && f < prm.n_frames && (HALVES == 2 || i * ppf < M / 32)) {
              const float4* sp = reinterpret_cast<const float4*>(prm.scales + static_cast<size_t>(f) * prm.c_mid + c);
              sa[i] = __ldg(sp);
              sb2[i] = __ldg(sp + 1);
            } else {
              sa[i] = sb2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          mbar_wait(full(s), ph);
          const uint32_t st = ring_base + s * prm.stage_bytes;
          if (HALVES == 2) {
            const uint32_t wb = st + prm.a_bytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + i * 32;
              if (r < prm.n_pad) {
                const uint32_t addr = wb + r * 128 + ((c8 ^ (r & 7)) << 4);
                const uint4 v = lds128(addr);
                const float2 v0 = unpack_h2(v.x), v1 = unpack_h2(v.y), v2 = unpack_h2(v.z), v3 = unpack_h2(v.w);
                sts128(addr, pack_h2(v0.x * sa[0].x, v0.y * sa[0].y), pack_h2(v1.x * sa[0].z, v1.y * sa[0].w),
                       pack_h2(v2.x * sb2[0].x, v2.y * sb2[0].y), pack_h2(v3.x * sb2[0].z, v3.y * sb2[0].w));
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < M / 32; ++i) {
              const int r = r0 + i * 32;
              const int fi = i / ppf;              // (ppf is 2 for the 8 x 8 frames this instantiation serves)
              const float4 a4 = fi == 0 ? sa[0] : (fi == 1 ? sa[1 % kSets] : (fi == 2 ? sa[2 % kSets] : sa[3 % kSets]));
              const float4 b4 = fi == 0 ? sb2[0] : (fi == 1 ? sb2[1 % kSets] : (fi == 2 ? sb2[2 % kSets] : sb2[3 % kSets]));
              const uint32_t addr = st + r * 128 + ((c8 ^ (r & 7)) << 4);
              const uint4 v = lds128(addr);
              const float2 v0 = unpack_h2(v.x), v1 = unpack_h2(v.y), v2 = unpack_h2(v.z), v3 = unpack_h2(v.w);
              sts128(addr, pack_h2(v0.x * a4.x, v0.y * a4.y), pack_h2(v1.x * a4.z, v1.y * a4.w),
                     pack_h2(v2.x * b4.x, v2.y * b4.y), pack_h2(v3.x * b4.z, v3.y * b4.w));
            }
          }
          fence_proxy_async();   // the MMA reads the stage through the async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(scaled(s));
          if (++s == prm.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> bias (+ shortcut) -> SMEM staging -> TMA store =====================
    // A warp walks its 32-row x 32-column units of every tile in a fixed order (tile, half, unit).  The fp16 shortcut
    // tile of unit k + 1 is requested by TMA into one of two 2 KB buffers before unit k is processed, and a lane reads its
    // row from SMEM (4 LDS.128): a row per lane straight from global memory touched 32 lines per load instruction and,
    // with the scaler's traffic, saturated the LSU data pipe (profiles/README.md, round 2 late).
    const int ew = warp - (2 + kPScalerWarps);   // 0..3
    const int quad = warp & 3;
    const uint32_t st32 = st_base + ew * kPEpiWarpBytes;
    const uint32_t st16 = st32 + kPStage32;
    const uint32_t rbuf = st16 + kPStage16;          // 2 x kPStage16
    const int units = (prm.c_out + 31) >> 5;
    const bool tma_res = prm.res16 != nullptr;
    int acc = 0;
    uint32_t pacc = 0;
    bool pending = false;
    int k = 0;                                       // running unit index of this warp
    auto request = [&](int tile, int h, int u, int kk) {   // shortcut tile of unit (tile, h, u) -> buffer kk & 1
      if (lane == 0) {
        mbar_expect_tx(r_full(ew, kk & 1), kPStage16);
        tma_load_2d(rbuf + (kk & 1) * kPStage16, &tm_r16, r_full(ew, kk & 1), u * 32, tile * M + h * 128 + quad * 32);
      }
    };
    if (tma_res && static_cast<int>(blockIdx.x) < prm.n_tiles) request(blockIdx.x, 0, 0, 0);
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      mbar_wait(acc_full(acc), pacc);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < HALVES; ++h) {
        const int row_base = tile * M + h * 128 + quad * 32;
        const int row = row_base + lane;
        const bool row_ok = row < prm.rows;
#pragma unroll 1
        for (int u = 0; u < units; ++u, ++k) {
          const int col0 = u * 32;
          if (tma_res) {   // next unit's shortcut tile (its buffer was read two units ago)
            int nu = u + 1, nh = h, nt = tile;
            if (nu == units) { nu = 0; if (++nh == HALVES) { nh = 0; nt += gridDim.x; } }
            if (nt < prm.n_tiles) request(nt, nh, nu, k + 1);
          }
          uint32_t r[32];
          tmem_ld32(tmem_base + acc * 256 + h * 128 + col0 + (static_cast<uint32_t>(quad * 32) << 16), r);
          float4 bv[8], rv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool c_ok = col0 + 4 * j < prm.c_out;   // c_out % 8 == 0
            bv[j] = c_ok ? __ldg(reinterpret_cast<const float4*>(prm.bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            rv[j] = (c_ok && row_ok && prm.res)
                        ? *(reinterpret_cast<const float4*>(prm.res + static_cast<size_t>(row) * prm.c_out + col0) + j)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (tma_res) {   // fp16 residual stream: this lane's 64 bytes of the TMA-loaded tile (out of bounds = zeros)
            mbar_wait(r_full(ew, k & 1), (k >> 1) & 1);
            const uint32_t rb = rbuf + (k & 1) * kPStage16 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 uu = lds128(rb + ((j ^ ((lane >> 1) & 3)) << 4));
              const float2 a = unpack_h2(uu.x), b = unpack_h2(uu.y), c = unpack_h2(uu.z), d = unpack_h2(uu.w);
              rv[2 * j] = make_float4(a.x, a.y, b.x, b.y);
              rv[2 * j + 1] = make_float4(c.x, c.y, d.x, d.y);
            }
          }
          if (pending) {   // the staging buffers are free once the previous unit's stores have read them
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float o0 = __uint_as_float(r[4 * j]) + bv[j].x + rv[j].x;
            const float o1 = __uint_as_float(r[4 * j + 1]) + bv[j].y + rv[j].y;
            const float o2 = __uint_as_float(r[4 * j + 2]) + bv[j].z + rv[j].z;
            const float o3 = __uint_as_float(r[4 * j + 3]) + bv[j].w + rv[j].w;
            if (prm.store32)
              sts128(st32 + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2),
                     __float_as_uint(o3));
            r[4 * j] = pack_h2(o0, o1);       // (the accumulator registers are reused for the fp16 copy)
            r[4 * j + 1] = pack_h2(o2, o3);
          }
          if (prm.store16) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(st16 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), r[8 * j], r[8 * j + 1], r[8 * j + 4], r[8 * j + 5]);
          }
          fence_proxy_async();
          __syncwarp();   // (also: every lane has read its shortcut row before lane 0 requests the tile after next into this buffer)
          if (lane == 0) {
            if (prm.store32) tma_store_2d(&tm_d32, st32, col0, row_base);
            if (prm.store16) tma_store_2d(&tm_d16, st16, col0, row_base);
            tma_store_commit();
          }
          pending = true;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc));
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_mb() {
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

// 2-D row-major tensor (rows x cols of `esize`-byte elements, row pitch ld elements), box (box_cols x box_rows)
int make_map_2d(CUtensorMap* tm, const void* base, int esize, long long rows, int cols, int ld, int box_cols, int box_rows,
                CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn enc = encode_fn_mb();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult cr = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS)
    return fail(M2S_ERR_CUDA, "tensor map (%s) failed (%d): rows=%lld cols=%d ld=%d box=%dx%d", what, static_cast<int>(cr), rows,
                cols, ld, box_cols, box_rows);
  return M2S_OK;
}

uint32_t idesc_f16(int n) {   // D = f32, A = B = f16, both K-major, M = 128
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

}  // namespace

bool mb_expand_dw_supported(int H, int W, int c_in, int c_mid) {
  return ((H == 16 && W == 16) || (H == 8 && W == 8)) && c_in % 8 == 0 && c_mid % 16 == 0 && c_in <= 256;
}

// x: [n*H*W][c_in] fp16; w_exp: [c_mid][c_in] fp16; out: [n*H*W][c_mid] fp16; sums: [n][c_mid]
int mb_expand_dw(const void* x, const void* w_exp, const float* bias1, const void* dw_w16, const float* dw_w32, const float* dw_b, void* out,
                 float* sums, int n, int H, int W, int c_in, int c_mid, cudaStream_t st) {
  if (!mb_expand_dw_supported(H, W, c_in, c_mid))
    return fail(M2S_ERR_UNSUPPORTED, "fused expand + depthwise: %dx%d, c_in=%d, c_mid=%d not supported", H, W, c_in, c_mid);
  if (n <= 0) return M2S_OK;
  const int halves = H == 16 ? 2 : 1;
  const int M = 128 * halves;
  ExpandDwParams prm{};
  prm.n_frames = n; prm.H = H; prm.W = W; prm.hw = H * W; prm.c_in = c_in; prm.c_mid = c_mid;
  prm.kblocks = (c_in + 63) / 64;
  prm.slabs = (c_mid + kSlab - 1) / kSlab;
  prm.fpt = M / prm.hw;
  prm.n_tiles = (n + prm.fpt - 1) / prm.fpt;
  prm.a_bytes = static_cast<uint32_t>(prm.kblocks) * M * 128;
  prm.b_stage_bytes = static_cast<uint32_t>(prm.kblocks) * kBTileBytes;
  prm.sp_bytes = (static_cast<uint32_t>(prm.fpt) * (H + 2) * (W + 2) * kPixPitch + 1023u) & ~1023u;
  prm.bias_bytes = (static_cast<uint32_t>(prm.slabs) * kSlab * 4 + 1023u) & ~1023u;
  prm.bias1 = bias1; prm.dw_w32 = dw_w32; prm.dw_b = dw_b; prm.out = static_cast<__half*>(out); prm.sums = sums;
  prm.idesc = idesc_f16(kSlab);
  prm.desc_hi = make_desc_hi(128);
  const uint32_t fixed = prm.a_bytes + 2 * prm.sp_bytes + prm.bias_bytes + 2 * 8 * 64 * 4 + 256 + 1024;
  int nb = 2;
  while (nb < 4 && fixed + (nb + 1) * prm.b_stage_bytes <= 216 * 1024) ++nb;
  prm.nb = nb;
  uint32_t smem = fixed + nb * prm.b_stage_bytes;
  if (smem > 227 * 1024) return fail(M2S_ERR_UNSUPPORTED, "fused expand + depthwise: tile does not fit SMEM (%u B)", smem);
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM (whole-TMEM allocation)
  CUtensorMap tm_x, tm_w;
  const long long rows = static_cast<long long>(n) * prm.hw;
  M2S_TRY(make_map_2d(&tm_x, x, 2, rows, c_in, c_in, 64, M, CU_TENSOR_MAP_SWIZZLE_128B, "expand A"));
  M2S_TRY(make_map_2d(&tm_w, w_exp, 2, c_mid, c_in, c_in, 64, kSlab, CU_TENSOR_MAP_SWIZZLE_128B, "expand W"));
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const ExpandDwParams);
  static const KernelFn kernels[2] = {mb_expand_dw_kernel<2>, mb_expand_dw_kernel<1>};
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    for (KernelFn k : kernels) M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  int grid = sm_count();
  if (grid > prm.n_tiles) grid = prm.n_tiles;
  M2S_TRY(profile_before(st));
  kernels[halves == 2 ? 0 : 1]<<<grid, kEThreads, smem, st>>>(tm_x, tm_w, prm);
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(st, 2.0 * static_cast<double>(rows) * c_mid * (c_in + 9));
}

bool mb_project_supported(int hw, int c_mid, int c_out) {
  return (hw == 256 || hw == 64) && c_mid % 8 == 0 && c_out % 8 == 0 && c_out <= (hw == 256 ? 128 : 256);
}

// a: [rows][c_mid] fp16 (depthwise output); w: [c_out][c_mid] fp16; scales: [n_frames][c_mid] fp32 or null;
// res / res16: [rows][c_out] shortcut, fp32 or fp16 (at most one; both may be null); d32 / d16: [rows][c_out] outputs
// (either may be null)
int mb_project(const void* a, const void* w, const float* scales, const float* bias, const float* res, const void* res16,
               float* d32, void* d16, int n_frames, int hw, int c_mid, int c_out, cudaStream_t st) {
  if (!mb_project_supported(hw, c_mid, c_out))
    return fail(M2S_ERR_UNSUPPORTED, "fused SE project: hw=%d c_mid=%d c_out=%d not supported", hw, c_mid, c_out);
  if (!d32 && !d16) return fail(M2S_ERR_BAD_ARG, "no output pointer");
  if (n_frames <= 0) return M2S_OK;
  const int halves = hw == 256 ? 2 : 1;
  const int M = 128 * halves;
  ProjectParams prm{};
  prm.rows = n_frames * hw; prm.hw = hw; prm.n_frames = n_frames; prm.c_mid = c_mid; prm.c_out = c_out;
  prm.n_pad = (c_out + 15) / 16 * 16;
  prm.kblocks = (c_mid + 63) / 64;
  prm.n_tiles = (prm.rows + M - 1) / M;
  prm.a_bytes = static_cast<uint32_t>(M) * 128;
  prm.stage_bytes = (prm.a_bytes + static_cast<uint32_t>(prm.n_pad) * 128 + 1023u) & ~1023u;
  prm.scales = scales; prm.bias = bias; prm.res = res; prm.res16 = res ? nullptr : static_cast<const __half*>(res16);
  prm.store32 = d32 != nullptr; prm.store16 = d16 != nullptr;
  prm.idesc = idesc_f16(prm.n_pad);
  prm.desc_hi = make_desc_hi(128);
  const uint32_t fixed = 4 * kPEpiWarpBytes + 512 + 1024;
  int stages = 2;
  while (stages < kPMaxStages && fixed + (stages + 1) * prm.stage_bytes <= 224 * 1024) ++stages;
  prm.stages = stages;
  uint32_t smem = fixed + stages * prm.stage_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;
  CUtensorMap tm_a, tm_w, tm_d32, tm_d16, tm_r16;
  M2S_TRY(make_map_2d(&tm_a, a, 2, prm.rows, c_mid, c_mid, 64, M, CU_TENSOR_MAP_SWIZZLE_128B, "project A"));
  M2S_TRY(make_map_2d(&tm_w, w, 2, c_out, c_mid, c_mid, 64, prm.n_pad, CU_TENSOR_MAP_SWIZZLE_128B, "project W"));
  // (an absent output still gets a valid map -- over the other buffer -- so that the kernel parameters stay well formed)
  M2S_TRY(make_map_2d(&tm_d32, d32 ? static_cast<const void*>(d32) : d16, d32 ? 4 : 2, prm.rows, c_out, c_out, 32, 32,
                      d32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, "project D32"));
  M2S_TRY(make_map_2d(&tm_d16, d16 ? d16 : static_cast<const void*>(d32), d16 ? 2 : 4, prm.rows, c_out, c_out, 32, 32,
                      d16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, "project D16"));
  // the fp16 shortcut as 32 x 32 tiles (a valid map over the output when there is none)
  M2S_TRY(make_map_2d(&tm_r16, prm.res16 ? static_cast<const void*>(prm.res16) : (d16 ? d16 : static_cast<const void*>(d32)),
                      (prm.res16 || d16) ? 2 : 4, prm.rows, c_out, c_out, 32, 32,
                      (prm.res16 || d16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, "project shortcut"));
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const ProjectParams);
  static const KernelFn kernels[2] = {mb_project_kernel<1>, mb_project_kernel<2>};
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    for (KernelFn k : kernels) M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  int grid = sm_count();
  if (grid > prm.n_tiles) grid = prm.n_tiles;
  M2S_TRY(profile_before(st));
  kernels[halves - 1]<<<grid, kPThreads, smem, st>>>(tm_a, tm_w, tm_d32, tm_d16, tm_r16, prm);
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(st, 2.0 * static_cast<double>(prm.rows) * c_mid * c_out);
}

}  // namespace m2s

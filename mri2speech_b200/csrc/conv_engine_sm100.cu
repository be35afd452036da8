// Multi-tap implicit-GEMM convolution engine for sm_100a: TMA -> SMEM (128B swizzle) -> tcgen05.mma
// (kind::tf32, accumulators in TMEM) -> fused epilogue -> global.
//
// One kernel serves every GEMM-shaped layer of the path (reference call sites):
//   * HiFi-GAN ResBlock1 causal dilated Conv1d           models.py:35-49 (72 convs)
//   * conv_pre (anti-causal k7)                          models.py:114-115
//   * ConvTranspose1d as a 3-tap polyphase GEMM          models.py:118
//   * LSTM input projection, mel head                    mri_acoustic_model.py:57-71,135
//   * EfficientNetV2 1x1 convs and stride-1 3x3 convs (as shifted-row taps over a zero-bordered
//     NHWC image flattened to rows)                      mri_acoustic_model.py:28-46 (timm)
//
// D[b, q + d_row_offset, n] = epi( sum_{tap j} sum_c A[b, q + shift[j], c] * W[j][n][c] )
//
// Activations are channels-last fp32, so the contraction (channel) dimension is contiguous and a
// conv tap is a row-shifted view of ONE SMEM tile: the A tile for a 32-channel block is loaded once
// with its halo (TMA zero-fills rows outside [0, a_rows): causal / anti-causal / image-border
// padding for free) and each tap issues tcgen05.mma with the A descriptor start address advanced by
// rel_shift[j] rows.  Weights are pre-packed on the host in the exact swizzled SMEM image and
// streamed per (channel block, tap) with 1-D bulk copies.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0   : TMA producer (one lane)
//   warp 1   : TMEM allocator + MMA issuer (one lane)
//   warps 2-9: epilogue (two per TMEM lane quadrant): tcgen05.ld -> SMEM transpose -> bias / residual /
//              MRF accumulate / scale / activation / mask -> coalesced global stores
#include "m2s_common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace m2s {

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kKBlock = 32;     // tf32 elements per 128-byte swizzle row
constexpr int kRowBytes = 128;
constexpr int kTmemCols = 512;
constexpr int kMaxStagesA = 4;
constexpr int kMaxStagesB = 8;
constexpr uint32_t kSmemBudget = 200 * 1024;  // > 114 KB forces 1 CTA / SM (TMEM is allocated whole)

struct EngineParams {
  ConvProblem p;
  const float* wpacked;
  int n_tile, n_tiles, msub, m_tile;
  int tiles_per_batch, total_tiles;
  int cblocks;
  int a_box_rows, a_nbox, shift_min;
  int na, nb;
  uint32_t a_stage_bytes, b_stage_bytes;
  int nacc, acc_stride;
  uint32_t idesc;
  int base_offset_mode;
  int a_per_tap;
  int rel_shift[M2S_MAX_TAPS];
  int tg;               // taps per weight stage
  uint32_t b_tap_bytes; // bytes of one tap's weight block (n_tile x 128)
  unsigned long long* trace;  // debug: per-role clock64 stamps of CTA 0 (null = off)
  int trace_tiles;
  int dbg;  // debug: bit0 skip global stores, bit1 skip TMEM loads, bit2 skip SMEM transpose, bit3 skip MMA issue
};

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("m2s conv engine: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SMEM matrix descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, int base_offset_mode) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                       // LBO (unused for swizzled K-major) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell) [46,48)
  if (base_offset_mode) d |= static_cast<uint64_t>((saddr >> 7) & 7) << 49;  // base offset [49,52)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B    [61,64)
  return d;
}

// debug timeline: slot = role*3 + k ; layout trace[tile_iter][9]
__device__ __forceinline__ void trace_stamp(const EngineParams& prm, int it, int slot) {
  if (prm.trace && blockIdx.x == 0 && it < prm.trace_tiles) prm.trace[it * 9 + slot] = clock64();
}

// One lane of a converged warp (the compiler then knows the region is warp-uniform and keeps descriptors /
// barrier addresses in uniform registers instead of emitting per-lane R2UR loops).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Four consecutive K-steps (4 x 8 tf32 = one 128-byte swizzle row) of one (sub-tile, tap).
__device__ __forceinline__ void mma_tf32_k4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum0,
                                            int ksteps) {
  mma_tf32(tmem_d, da, db, idesc, accum0);
  if (ksteps > 1) mma_tf32(tmem_d, da + 2, db + 2, idesc, 1u);
  if (ksteps > 2) mma_tf32(tmem_d, da + 4, db + 4, idesc, 1u);
  if (ksteps > 3) mma_tf32(tmem_d, da + 6, db + 6, idesc, 1u);
}

struct EpiConsts {
  float inv_slope, pre_w, post_w, out_scale, act_slope;
};

// Epilogue programs (compile-time): the common cases drop every unused instruction -- with 256-column-wide tiles the
// epilogue is ALU-issue-bound (32K outputs per tile on 8 warps), so instructions per output are what matters.
enum { EPI_FULL = 0, EPI_FULL_SILU = 1, EPI_BIAS = 2, EPI_LRELU = 3, EPI_SILU = 4, EPI_RES = 5 };

__device__ __forceinline__ float fast_silu(float v) { return __fdividef(v, 1.f + __expf(-v)); }

template <int kEpi>
__device__ __forceinline__ float epi_elem(const EpiConsts& c, float acc, float bias, float res, float accum) {
  float v = acc + bias;
  if (kEpi == EPI_BIAS) return v;
  if (kEpi == EPI_LRELU) return v >= 0.f ? v : v * c.act_slope;
  if (kEpi == EPI_SILU) return fast_silu(v);
  if (kEpi == EPI_RES) return v + res;
  const float rt = res >= 0.f ? res : res * c.inv_slope;
  v = fmaf(rt, c.pre_w, v);
  v += accum;
  v *= c.out_scale;
  if (kEpi == EPI_FULL_SILU) v = fast_silu(v);
  else v = v >= 0.f ? v : v * c.act_slope;
  return fmaf(rt, c.post_w, v);
}

// ---- the kernel ----------------------------------------------------------------
// 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (two warps per
// TMEM lane quadrant; a pair splits the (sub-tile, 32-column chunk) units of a tile).
template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv_engine_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ EngineParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + prm.na * prm.a_stage_bytes;
  const uint32_t bar_base = b_base + prm.nb * prm.b_stage_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxStagesA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + kMaxStagesB + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + 2 * kMaxStagesB + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + 2 * kMaxStagesB + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStagesA + 2 * kMaxStagesB + 4);
  const uint32_t stage_base = bar_base + 1024u;  // 8 epilogue warps x 4 KB transpose staging

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvProblem& p = prm.p;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < prm.na; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < prm.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < prm.nacc; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), kEpiWarps); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int taps = p.taps;
  const int cblocks = prm.cblocks;
  const int msub = prm.msub;
  const int n_tile = prm.n_tile;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    const uint32_t a_bytes = prm.a_nbox * prm.a_box_rows * kRowBytes;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % prm.n_tiles;
      const int mt = (tile / prm.n_tiles) % prm.tiles_per_batch;
      const int b = tile / (prm.n_tiles * prm.tiles_per_batch);
      const int q0 = mt * prm.m_tile;
      if (lane == 0) trace_stamp(prm, it, 0);
      for (int cb = 0; cb < cblocks; ++cb) {
        if (!prm.a_per_tap) {
          mbar_wait(a_empty(sa), pa ^ 1);
          if (cb == 0 && lane == 0) trace_stamp(prm, it, 1);
          if (elect_one()) {
            mbar_expect_tx(a_full(sa), a_bytes);
            for (int bx = 0; bx < prm.a_nbox; ++bx)
              tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * kRowBytes, &tmap_a, a_full(sa),
                          cb * kKBlock, q0 + prm.shift_min + bx * prm.a_box_rows, b);
          }
          __syncwarp();
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
        }
        for (int tap0 = 0; tap0 < taps; tap0 += prm.tg) {
          const int cnt = min(prm.tg, taps - tap0);
          if (prm.a_per_tap) {
            mbar_wait(a_empty(sa), pa ^ 1);
            if (elect_one()) {
              mbar_expect_tx(a_full(sa), a_bytes);
              for (int bx = 0; bx < prm.a_nbox; ++bx)
                tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * kRowBytes, &tmap_a, a_full(sa),
                            cb * kKBlock, q0 + p.shift[tap0] + bx * prm.a_box_rows, b);
            }
            __syncwarp();
            if (++sa == prm.na) { sa = 0; pa ^= 1; }
          }
          mbar_wait(b_empty(sb), pb ^ 1);
          if (elect_one()) {
            const uint32_t bytes = cnt * prm.b_tap_bytes;
            mbar_expect_tx(b_full(sb), bytes);
            const float* src =
                prm.wpacked + (static_cast<size_t>((nt * cblocks + cb) * taps + tap0)) * n_tile * kKBlock;
            bulk_load(b_base + sb * prm.b_stage_bytes, src, bytes, b_full(sb));
          }
          __syncwarp();
          if (++sb == prm.nb) { sb = 0; pb ^= 1; }
        }
      }
      if (lane == 0) trace_stamp(prm, it, 2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, pacc = 0;
    const uint64_t desc_hi = make_desc_sw128(0, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
      if (lane == 0) trace_stamp(prm, it, 3);
      mbar_wait(acc_empty(acc), pacc ^ 1);
      tc_fence_after();
      if (lane == 0) trace_stamp(prm, it, 4);
      const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride;
      for (int cb = 0; cb < cblocks; ++cb) {
        const int rem = p.c_in - cb * kKBlock;
        const int ksteps = rem >= kKBlock ? 4 : (rem + 7) / 8;
        if (!prm.a_per_tap) {
          mbar_wait(a_full(sa), pa);
        }
        for (int tap0 = 0; tap0 < taps; tap0 += prm.tg) {
          const int cnt = min(prm.tg, taps - tap0);
          if (prm.a_per_tap) mbar_wait(a_full(sa), pa);
          mbar_wait(b_full(sb), pb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_tile = a_base + sa * prm.a_stage_bytes;
            const uint32_t b_tile = b_base + sb * prm.b_stage_bytes;
            if (!(prm.dbg & 8)) {
              for (int t = 0; t < cnt; ++t) {
                const int row_shift = prm.a_per_tap ? 0 : prm.rel_shift[tap0 + t];
                const uint64_t db = desc_hi | (((b_tile + t * prm.b_tap_bytes) & 0x3FFFF) >> 4);
                const uint32_t first = (cb | tap0 | t) ? 1u : 0u;
                const uint64_t da0 = desc_hi | (((a_tile + row_shift * kRowBytes) & 0x3FFFF) >> 4);
                mma_tf32_k4(tmem_acc, da0, db, prm.idesc, first, ksteps);
                if (msub > 1) mma_tf32_k4(tmem_acc + n_tile, da0 + ((128 * kRowBytes) >> 4), db, prm.idesc, first, ksteps);
              }
            }
            const bool last_group = tap0 + cnt >= taps;
            tc_commit(b_empty(sb));
            if (prm.a_per_tap || last_group) tc_commit(a_empty(sa));
            if (cb == cblocks - 1 && last_group) tc_commit(acc_full(acc));
          }
          __syncwarp();
          if (++sb == prm.nb) { sb = 0; pb ^= 1; }
          if (prm.a_per_tap) {
            if (++sa == prm.na) { sa = 0; pa ^= 1; }
          }
        }
        if (!prm.a_per_tap) {
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
        }
      }
      if (lane == 0) trace_stamp(prm, it, 5);
      if (++acc == prm.nacc) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    // TMEM -> registers (row per thread) -> SMEM transpose (XOR-swizzled, conflict-free) -> coalesced global
    // traffic: 8 lanes cover one 128-byte row segment, a warp instruction covers 4 rows.
    const int ew = warp - 2;      // 0..7
    const int quad = warp & 3;    // TMEM lane quadrant this warp may access
    const int half = ew >> 2;     // which of the two warps of the quadrant
    const Epilogue& e = p.epi;
    const uint32_t stage = stage_base + ew * 4096;  // 32 rows x 128 B
    const int rr0 = lane >> 3, cc = lane & 7;
    const bool has_res = e.res != nullptr, has_acc = e.accum != nullptr;
    EpiConsts ec;
    ec.inv_slope = e.res_inv_slope;
    ec.pre_w = (has_res && !e.res_after_act) ? 1.f : 0.f;
    ec.post_w = (has_res && e.res_after_act) ? 1.f : 0.f;
    ec.out_scale = e.out_scale;
    ec.act_slope = e.act == M2S_ACT_LRELU ? e.act_slope : 1.f;
    const int mask_mode = e.mask_mode;
    const int nchunks = (n_tile + 31) >> 5;
    const int units = msub * nchunks;
    int acc = 0;
    uint32_t pacc = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % prm.n_tiles;
      const int mt = (tile / prm.n_tiles) % prm.tiles_per_batch;
      const int b = tile / (prm.n_tiles * prm.tiles_per_batch);
      const int q0 = mt * prm.m_tile;
      const int n0 = nt * n_tile;
      int len_rows = 0x7fffffff;
      if (mask_mode == M2S_MASK_LEN) len_rows = __ldg(e.lens + b) * e.len_scale;
      const size_t d_base = static_cast<size_t>(b) * p.d_batch_rows + p.d_row_offset;
      if (ew == 0 && lane == 0) trace_stamp(prm, it, 6);
      mbar_wait(acc_full(acc), pacc);
      tc_fence_after();
      if (ew == 0 && lane == 0) trace_stamp(prm, it, 7);
      const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride + (static_cast<uint32_t>(quad * 32) << 16);
      for (int u = half; u < units; u += 2) {
        const int sub = u / nchunks;
        const int c0 = (u - sub * nchunks) << 5;
        const int qw = q0 + sub * 128 + quad * 32;  // first row of this warp's 32-row slab
        uint32_t r[32];
        tmem_ld16(tmem_acc + sub * n_tile + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        if (c0 + 16 < n_tile) tmem_ld16(tmem_acc + sub * n_tile + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        // While the TMEM read is in flight: bias / residual / accumulate loads of this unit (the output may alias
        // them in place, so every load is issued before the first store).  Thread (rr0, cc) owns rows rr0 + 4i,
        // columns n .. n+3; all row predicates reduce to "4i + rr0 < bound".
        const int n = n0 + c0 + cc * 4;
        const bool col_ok = (c0 + cc * 4 < n_tile) && n < p.n;
        const int rows_ok = col_ok ? min(32, p.l_out - qw) : 0;                 // rows that exist
        int rows_valid = 32;                                                    // rows that survive the mask
        if (mask_mode == M2S_MASK_LEN) rows_valid = len_rows - (qw + p.d_row_offset);
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
        const size_t row0 = d_base + qw + rr0;
        float* dptr = p.d + row0 * p.d_ld + n;
        const size_t d_step = static_cast<size_t>(4) * p.d_ld;
        float4 res4[8], acc4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          res4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (kEpi == EPI_FULL || kEpi == EPI_FULL_SILU || kEpi == EPI_RES) {
          if (has_res) {
            const float* rptr = e.res + row0 * e.res_ld + n;
            const size_t r_step = static_cast<size_t>(4) * e.res_ld;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i * 4 + rr0 < rows_ok) res4[i] = *reinterpret_cast<const float4*>(rptr + i * r_step);
          }
        }
        if (kEpi == EPI_FULL || kEpi == EPI_FULL_SILU) {
          if (has_acc) {
            const float* aptr = e.accum + row0 * e.accum_ld + n;
            const size_t a_step = static_cast<size_t>(4) * e.accum_ld;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i * 4 + rr0 < rows_ok) acc4[i] = *reinterpret_cast<const float4*>(aptr + i * a_step);
          }
        }
        // image-border mask: (i, j) = divmod(row, pitch) once, then stepped by 4 rows
        int mi = 0, mj = 0;
        if (mask_mode == M2S_MASK_PITCH) {
          const int drow = qw + rr0 + p.d_row_offset;
          mi = drow / e.pitch;
          mj = drow - mi * e.pitch;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + lane * 128 + ((j ^ (lane & 7)) << 4)),
                       "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        __syncwarp();
        // all 8 rows' (32 independent) element chains are computed unconditionally so the scheduler can interleave
        // them; only the stores are predicated
        float4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rr0;
          float4 a4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w)
                       : "r"(stage + rr * 128 + ((cc ^ (rr & 7)) << 4)));
          o[i].x = epi_elem<kEpi>(ec, a4.x, bias4.x, res4[i].x, acc4[i].x);
          o[i].y = epi_elem<kEpi>(ec, a4.y, bias4.y, res4[i].y, acc4[i].y);
          o[i].z = epi_elem<kEpi>(ec, a4.z, bias4.z, res4[i].z, acc4[i].z);
          o[i].w = epi_elem<kEpi>(ec, a4.w, bias4.w, res4[i].w, acc4[i].w);
        }
        if (mask_mode == M2S_MASK_PITCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool valid = mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi;
            if (!valid) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            mj += 4;
            if (mj >= e.pitch) { mj -= e.pitch; ++mi; }
          }
        } else if (mask_mode == M2S_MASK_LEN) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i * 4 + rr0 >= rows_valid) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i * 4 + rr0 < rows_ok) *reinterpret_cast<float4*>(dptr + i * d_step) = o[i];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) trace_stamp(prm, it, 8);
      if (lane == 0) mbar_arrive(acc_empty(acc));
      if (++acc == prm.nacc) { acc = 0; pacc ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

EngineKnobs& engine_knobs() {
  static EngineKnobs k = [] {
    EngineKnobs x;
    x.base_offset_mode = env_int("M2S_ENGINE_BASE_OFFSET", 0);
    x.msub = env_int("M2S_ENGINE_MSUB", 0);
    x.tmap_tf32 = env_int("M2S_ENGINE_TMAP_TF32", 1);
    x.max_ctas = env_int("M2S_ENGINE_MAX_CTAS", 0);
    x.a_per_tap = env_int("M2S_ENGINE_A_PER_TAP", 0);
    x.n_tile_max = env_int("M2S_ENGINE_NTILE_MAX", 128);
    return x;
  }();
  return k;
}

// ---- optional per-launch timing (bench.py roofline leg) -----------------------------
namespace {
struct ProfileRing {
  std::vector<cudaEvent_t> ev;
  std::vector<double> flops;
  int count = 0;
  bool on = false;
};
ProfileRing& ring() {
  static ProfileRing r;
  return r;
}
}  // namespace

int profile_enable(int on) {
  ProfileRing& r = ring();
  r.on = on != 0;
  r.count = 0;
  r.flops.clear();
  return M2S_OK;
}

int profile_read(float* ms, double* flops, int cap, int* n_out) {
  ProfileRing& r = ring();
  M2S_CUDA_OK(cudaDeviceSynchronize());
  int n = r.count < cap ? r.count : cap;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    M2S_CUDA_OK(cudaEventElapsedTime(&t, r.ev[2 * i], r.ev[2 * i + 1]));
    ms[i] = t;
    if (flops) flops[i] = r.flops[i];
  }
  *n_out = n;
  r.count = 0;
  r.flops.clear();
  return M2S_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

void choose_n_tiling(int n, int* n_tile, int* n_tiles) {
  // smallest number of tiles with n_tile a multiple of 16 and <= 256; prefer an even split.
  const int cap = engine_knobs().n_tile_max;
  int tiles = n > cap ? (n + cap - 1) / cap : 1;
  int nt = (((n + tiles - 1) / tiles) + 15) / 16 * 16;
  *n_tile = nt;
  *n_tiles = tiles;
}

static float host_round_tf32(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return v;
  u += 0x1000u;  // round to nearest, ties away (matches cvt.rna)
  u &= 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

int pack_weights(const float* host_w, int taps, int n, int c_in, bool tf32_round, PackedWeights* out) {
  PackedWeights w;
  w.n = n; w.c_in = c_in; w.taps = taps;
  choose_n_tiling(n, &w.n_tile, &w.n_tiles);
  w.cblocks = (c_in + kKBlock - 1) / kKBlock;
  w.packed_floats = static_cast<size_t>(w.n_tiles) * w.cblocks * taps * w.n_tile * kKBlock;
  std::vector<float> packed(w.packed_floats, 0.f);
  std::vector<float> plain(static_cast<size_t>(taps) * n * c_in);
  for (int j = 0; j < taps; ++j)
    for (int o = 0; o < n; ++o)
      for (int c = 0; c < c_in; ++c) {
        float v = host_w[(static_cast<size_t>(j) * n + o) * c_in + c];
        if (tf32_round) v = host_round_tf32(v);
        plain[(static_cast<size_t>(j) * n + o) * c_in + c] = v;
        const int nt = o / w.n_tile, r = o % w.n_tile, cb = c / kKBlock, cc = c % kKBlock;
        const size_t blk = (static_cast<size_t>(nt * w.cblocks + cb) * taps + j) * w.n_tile * kKBlock;
        // 128B swizzle: 16-byte chunk index ^= (row & 7)
        const int chunk = (cc >> 2) ^ (r & 7);
        packed[blk + static_cast<size_t>(r) * kKBlock + chunk * 4 + (cc & 3)] = v;
      }
  M2S_CUDA_OK(cudaMalloc(&w.dev, packed.size() * sizeof(float)));
  M2S_CUDA_OK(cudaMalloc(&w.plain, plain.size() * sizeof(float)));
  M2S_CUDA_OK(cudaMemcpy(w.dev, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  M2S_CUDA_OK(cudaMemcpy(w.plain, plain.data(), plain.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = w;
  return M2S_OK;
}

void free_weights(PackedWeights* w) {
  if (w->dev) cudaFree(w->dev);
  if (w->plain) cudaFree(w->plain);
  *w = PackedWeights{};
}

int conv_tcgen05(const ConvProblem& p, const PackedWeights& w, cudaStream_t stream) {
  if (p.taps < 1 || p.taps > M2S_MAX_TAPS) return fail(M2S_ERR_BAD_ARG, "taps=%d out of range", p.taps);
  if (p.c_in % 4 || p.a_ld % 4 || p.d_ld % 4 || p.n % 4)
    return fail(M2S_ERR_UNSUPPORTED, "c_in/a_ld/d_ld/n must be multiples of 4 (got %d/%d/%d/%d)", p.c_in, p.a_ld,
                p.d_ld, p.n);
  if ((reinterpret_cast<uintptr_t>(p.a) & 15) || (reinterpret_cast<uintptr_t>(p.d) & 15))
    return fail(M2S_ERR_BAD_ARG, "A and D must be 16-byte aligned");
  if (w.n != p.n || w.c_in != p.c_in || w.taps != p.taps)
    return fail(M2S_ERR_BAD_ARG, "packed weights do not match the problem");
  if (p.batch <= 0 || p.l_out <= 0) return M2S_OK;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const EngineKnobs& knobs = engine_knobs();

  EngineParams prm{};
  prm.p = p;
  prm.wpacked = w.dev;
  prm.n_tile = w.n_tile;
  prm.n_tiles = w.n_tiles;
  prm.cblocks = w.cblocks;
  prm.base_offset_mode = knobs.base_offset_mode;
  prm.dbg = knobs.dbg;
  prm.trace = knobs.trace;
  prm.trace_tiles = knobs.trace_tiles;
  prm.a_per_tap = knobs.a_per_tap;
  int smin = p.shift[0], smax = p.shift[0];
  for (int j = 1; j < p.taps; ++j) { smin = p.shift[j] < smin ? p.shift[j] : smin; smax = p.shift[j] > smax ? p.shift[j] : smax; }
  prm.shift_min = smin;
  for (int j = 0; j < p.taps; ++j) prm.rel_shift[j] = p.shift[j] - smin;
  const int halo = prm.a_per_tap ? 0 : smax - smin;

  // M sub-tiles per CTA: 2 halves the weight traffic per output row; use it when there is enough work.
  int msub = knobs.msub;
  if (msub != 1 && msub != 2) {
    const long long tiles1 = static_cast<long long>(p.batch) * ((p.l_out + 127) / 128) * w.n_tiles;
    // two M sub-tiles halve the weight traffic per output row; keep TWO accumulator buffers so that the
    // epilogue of tile i overlaps the MMAs of tile i+1
    msub = (tiles1 >= 2LL * sm_count() && 4 * w.n_tile <= kTmemCols) ? 2 : 1;
  }
  if (msub * w.n_tile > kTmemCols) msub = 1;
  prm.msub = msub;
  prm.m_tile = 128 * msub;
  prm.tiles_per_batch = (p.l_out + prm.m_tile - 1) / prm.m_tile;
  prm.total_tiles = p.batch * prm.tiles_per_batch * w.n_tiles;
  prm.nacc = (2 * msub * w.n_tile <= kTmemCols) ? 2 : 1;
  prm.acc_stride = prm.nacc == 2 ? kTmemCols / 2 : 0;

  const int a_rows_needed = prm.m_tile + halo;
  prm.a_nbox = (a_rows_needed + 255) / 256;
  prm.a_box_rows = (((a_rows_needed + prm.a_nbox - 1) / prm.a_nbox) + 7) / 8 * 8;
  prm.a_stage_bytes = static_cast<uint32_t>(prm.a_nbox * prm.a_box_rows * kRowBytes);
  prm.a_stage_bytes = (prm.a_stage_bytes + 1023u) & ~1023u;
  prm.b_tap_bytes = static_cast<uint32_t>(w.n_tile * kRowBytes);
  // several taps share one weight stage when the per-tap block is small (amortises barrier round trips)
  int tg = static_cast<int>(32768u / prm.b_tap_bytes);
  if (tg < 1) tg = 1;
  if (tg > p.taps) tg = p.taps;
  if (knobs.a_per_tap) tg = 1;
  prm.tg = tg;
  prm.b_stage_bytes = static_cast<uint32_t>(tg) * prm.b_tap_bytes;
  const uint32_t b_stage_alloc = (prm.b_stage_bytes + 1023u) & ~1023u;
  // stage counts inside the budget: at least 2 A + 2 B
  const uint32_t bar_bytes = 1024 + kEpiWarps * 4096;  // barriers + epilogue staging
  int na = 2, nb = 2;
  auto total = [&](int a, int b) { return a * prm.a_stage_bytes + b * b_stage_alloc + bar_bytes + 1024u; };
  if (total(na, nb) > kSmemBudget + 24 * 1024)
    return fail(M2S_ERR_UNSUPPORTED, "tile does not fit SMEM (A stage %u B, B stage %u B)", prm.a_stage_bytes,
                b_stage_alloc);
  while (nb < kMaxStagesB && nb < 4 && total(na, nb + 1) <= kSmemBudget) ++nb;
  while (na < kMaxStagesA && na < 3 && total(na + 1, nb) <= kSmemBudget) ++na;
  while (nb < kMaxStagesB && total(na, nb + 1) <= kSmemBudget) ++nb;
  prm.na = na;
  prm.nb = nb;
  const uint32_t b_stage_bytes_copy = prm.b_stage_bytes;
  prm.b_stage_bytes = b_stage_bytes_copy;  // bytes actually copied per stage
  // NOTE: stage pitch in SMEM uses the 1024-aligned size
  const uint32_t b_pitch = b_stage_alloc;
  uint32_t smem_bytes = total(na, nb);
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // keep 1 CTA / SM (whole-TMEM allocation)

  // instruction descriptor: D=f32, A=B=tf32, K-major both, N, M=128
  prm.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(w.n_tile >> 3) << 17) |
              (static_cast<uint32_t>(128 >> 4) << 24);

  // tensor map over A: (c_in, a_rows, batch), box (32, a_box_rows, 1), 128B swizzle, OOB -> zero
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p.c_in), static_cast<cuuint64_t>(p.a_rows),
                        static_cast<cuuint64_t>(p.batch)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p.a_ld) * 4ull,
                           static_cast<cuuint64_t>(p.a_batch_rows) * static_cast<cuuint64_t>(p.a_ld) * 4ull};
  if (p.batch == 1) gstride[1] = gstride[0] * static_cast<cuuint64_t>(p.a_rows > 0 ? p.a_rows : 1);
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kKBlock), static_cast<cuuint32_t>(prm.a_box_rows), 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult cr = enc(&tmap, knobs.tmap_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                    const_cast<float*>(p.a), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS)
    return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): c_in=%d rows=%d batch=%d ld=%d box_rows=%d",
                static_cast<int>(cr), p.c_in, p.a_rows, p.batch, p.a_ld, prm.a_box_rows);

  // B stage pitch must equal the copy size for the kernel's addressing: use the aligned pitch everywhere.
  prm.b_stage_bytes = b_pitch;
  // copy size = n_tile*128 which is already a multiple of 2048 for n_tile%16==0 -> equals pitch when n_tile%8==0
  if (b_pitch != static_cast<uint32_t>(tg) * prm.b_tap_bytes)
    return fail(M2S_ERR_UNSUPPORTED, "n_tile=%d gives a non-1024-aligned weight stage", w.n_tile);

  using KernelFn = void (*)(const CUtensorMap, const EngineParams);
  static const KernelFn kernels[6] = {conv_engine_kernel<EPI_FULL>, conv_engine_kernel<EPI_FULL_SILU>,
                                      conv_engine_kernel<EPI_BIAS>, conv_engine_kernel<EPI_LRELU>,
                                      conv_engine_kernel<EPI_SILU>, conv_engine_kernel<EPI_RES>};
  static bool attr_set = false;
  if (!attr_set) {
    for (KernelFn k : kernels)
      M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int epi = p.epi.act == M2S_ACT_SILU ? EPI_FULL_SILU : EPI_FULL;
  const bool plain = !p.epi.accum && p.epi.out_scale == 1.f;
  if (plain && !p.epi.res) epi = p.epi.act == M2S_ACT_SILU ? EPI_SILU : (p.epi.act == M2S_ACT_LRELU ? EPI_LRELU : EPI_BIAS);
  else if (plain && p.epi.act == M2S_ACT_NONE && p.epi.res_inv_slope == 1.f) epi = EPI_RES;
  int grid = knobs.max_ctas > 0 ? knobs.max_ctas : sm_count();
  if (grid > prm.total_tiles) grid = prm.total_tiles;
  ProfileRing& pr = ring();
  if (pr.on) {
    while (static_cast<int>(pr.ev.size()) < 2 * (pr.count + 1)) {
      cudaEvent_t e;
      M2S_CUDA_OK(cudaEventCreate(&e));
      pr.ev.push_back(e);
    }
    M2S_CUDA_OK(cudaEventRecord(pr.ev[2 * pr.count], stream));
  }
  kernels[epi]<<<grid, kThreads, smem_bytes, stream>>>(tmap, prm);
  M2S_CUDA_OK(cudaGetLastError());
  if (pr.on) {
    M2S_CUDA_OK(cudaEventRecord(pr.ev[2 * pr.count + 1], stream));
    pr.flops.push_back(2.0 * p.batch * static_cast<double>(p.l_out) * p.n * p.c_in * p.taps);
    ++pr.count;
  }
  return M2S_OK;
}

}  // namespace m2s

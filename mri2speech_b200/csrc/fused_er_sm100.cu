// Fused EdgeResidual (FusedMBConv) block of the frame-CNN encoder for sm_100a (fp16 build):
//     y = bn2(conv_pwl( silu( bn1(conv_exp_3x3(x)) ) )) [+ x]                       in ONE kernel.
//
// Reference: timm EdgeResidual (conv_exp 3x3 -> bn1 -> SiLU -> conv_pwl 1x1 -> bn2, shortcut when in == out and stride
// 1) as called through EffNetV2B2Backbone.forward (mri2speech_code/mri_acoustic_model.py:28-48); SURVEY.md 8a-1 stages
// 1-2.  The two-launch path writes the 4x-expanded tensor (1.1 GB per 1024 frames at stage 1) to HBM and reads it back;
// here the expanded tile T never leaves the SM:
//
//   TMA (x tile with its 3x3 halo, weights)  ->  tcgen05.mma kind::f16, 9 row-shifted taps  ->  acc1 (TMEM, N1 = 4 C)
//   epilogue warps: acc1 -> + bias -> SiLU -> fp16 -> T tile in SMEM, written directly in the K-major SWIZZLE_128B
//                   operand layout
//   tcgen05.mma (A = T tile, K = N1)  ->  acc2 (TMEM, N2 = C_out)  ->  the engine's fused epilogue (bias, fp32 shortcut,
//   image-border mask) -> fp32 and / or fp16 stores.
//
// Same structure as resblock_pair_kernel (csrc/resblock_pair_sm100.cu), generalised where the block needs it: the two
// convs have different widths (N1 = 128 / 224 in one or two N tiles, N2 = 32 / 64) and different operand formats (x has
// 32 or 56 channels: 64-byte SWIZZLE_64B rows or one 128-byte block; T always 128-byte rows), conv2 has one tap (no halo,
// no recompute), shifts of conv1 are the non-negative dy * pitch + dx of the zero-bordered image.
#include "engine_device.cuh"
#include <cstdlib>
#include <mutex>

namespace m2s {

using namespace engine;

namespace {

struct ErParams {
  ConvProblem p2;          // epilogue / outputs / bias2 of the project conv; p2.l_out, p2.batch, p2.n
  const float* bias1;      // [n1]
  const void* w1;          // packed weights (single-CTA layout): expand [nt1][cb1][9][n_tile1 rows], project [cb2][n_tile2 rows]
  const void* w2;
  int taps1;
  int rel_shift1[M2S_MAX_TAPS];
  int c_in, n1;                  // channels of x / of T
  int n_tile1, n_tiles1;         // N tiling of the expand conv (n_tile1 * n_tiles1 >= n1)
  int n_tile2;                   // N of the project MMA (>= p2.n, multiple of 16)
  int cblocks1, cblocks2;
  int kblock1, row_bytes1;       // operand format of the expand weights (and of x unless K windows are on)
  // operand format of x.  Normally the weights' format; with per-tap K windows (stride-2 blocks over the space-to-depth
  // input, ConvProblem::tap_ksteps) x has 4 c_in channels per row (128-byte rows, cblocks_a K blocks) and tap t contracts
  // K-steps tap_kofs[t] .. + win_ksteps - 1 of K block tap_cb[t] against the whole (c_in-wide) weight block of the tap
  int cblocks_a, kblock_a, row_bytes_a;
  uint64_t desc_hi_a;
  int win_ksteps;                // 0 = off
  int tap_cb[M2S_MAX_TAPS], tap_kofs[M2S_MAX_TAPS];
  // space-to-depth by TMA (s2d_pitch > 0, with K windows): the x tile is `s2d_ny` whole rows of the space-to-depth image
  // (pitch s2d_pitch = W_out + 2), gathered straight from the zero-bordered NHWC input by a 5-D box (ErS2d below)
  int s2d_pitch, s2d_ny;
  uint64_t desc_hi1, desc_hi2;
  uint32_t idesc1, idesc2;
  int x_row0;                    // first x row of a tile relative to q0 (min shift)
  int tiles_per_batch, total_tiles;
  int a_box_rows, a_nbox;
  uint32_t a_stage_bytes, b_stage_bytes, b_tap_bytes1, b_tap_bytes2, t_buf_bytes;
  int na, nb, tg1;
  int nbuf;                      // 1 or 2 buffers of acc1 / acc2 / T; lookahead = nbuf - 1
  int resident;                  // all weights of the block are loaded into SMEM once (no weight ring)
  uint32_t w1_bytes, w2_bytes;   // packed sizes of the two weight tensors
  int tma_epi;                   // epilogue 2 through SMEM tiles: shortcut rows by TMA load, outputs by TMA store (C_out = 32)
};

constexpr uint32_t kRsBytes = 128 * 128;   // fp32 tile of 128 rows x 32 channels (shortcut in, fp32 output out -- in place)
constexpr uint32_t kH16Bytes = 128 * 64;   // fp16 output tile

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_er(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr uint32_t kTkbBytes = 128 * 128;   // one K block (64 channels) of the T tile: 128 rows x 128 bytes

__global__ void __launch_bounds__(kThreads, 1)
fused_er_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tm_r,
                const __grid_constant__ CUtensorMap tm_d32, const __grid_constant__ CUtensorMap tm_d16,
                const __grid_constant__ ErParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t t_base = a_base + prm.na * prm.a_stage_bytes;
  const uint32_t b_base = t_base + prm.nbuf * prm.t_buf_bytes;   // weight ring, or the resident weights (w1 then w2)
  const uint32_t rs_base = b_base + (prm.resident ? prm.w1_bytes + prm.w2_bytes : prm.nb * prm.b_stage_bytes);
  const uint32_t h16_base = rs_base + 2 * kRsBytes;     // (both only when prm.tma_epi)
  const uint32_t bar_base = rs_base + (prm.tma_epi ? 2 * kRsBytes + kH16Bytes : 0u);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxStagesA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kMaxStagesA + kMaxStagesB + s); };
  const uint32_t x_base = bar_base + 8u * (2 * kMaxStagesA + 2 * kMaxStagesB);
  auto acc1_full = [&](int s) { return x_base + 8u * s; };
  auto acc1_empty = [&](int s) { return x_base + 8u * (2 + s); };
  auto acc2_full = [&](int s) { return x_base + 8u * (4 + s); };
  auto acc2_empty = [&](int s) { return x_base + 8u * (6 + s); };
  auto t_full = [&](int s) { return x_base + 8u * (8 + s); };
  auto t_empty = [&](int s) { return x_base + 8u * (10 + s); };
  const uint32_t tmem_slot = x_base + 8u * 12;
  const uint32_t w_full = x_base + 8u * 13;
  auto r_full = [&](int s) { return x_base + 8u * (14 + s); };
  auto r_empty = [&](int s) { return x_base + 8u * (16 + s); };
  const uint32_t bias1_smem = bar_base + 1024u;             // n1 floats (<= 256)
  const uint32_t bias2_smem = bar_base + 2048u;             // p2.n floats (<= 64)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvProblem& p = prm.p2;
  const int nbuf = prm.nbuf;
  const int la = nbuf - 1;  // the expand conv of tile i+la is issued before the project conv of tile i

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < prm.na; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < prm.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < nbuf; ++s) {
      mbar_init(acc1_full(s), 1); mbar_init(acc1_empty(s), kEpiWarps);
      mbar_init(acc2_full(s), 1); mbar_init(acc2_empty(s), kEpiWarps);
      mbar_init(t_full(s), kEpiWarps); mbar_init(t_empty(s), 1);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 4); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias1_smem + 4u * i), "f"(i < prm.n1 ? __ldg(prm.bias1 + i) : 0.f) : "memory");
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias2_smem + 4u * i), "f"((i < p.n && p.epi.bias) ? __ldg(p.epi.bias + i) : 0.f) : "memory");
  const uint32_t bias_smem = bias2_smem;
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n1p = prm.n_tile1 * prm.n_tiles1;   // columns of one acc1 buffer
  const int my_tiles = (prm.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  auto acc1_addr = [&](int buf) { return tmem_base + static_cast<uint32_t>(buf * n1p); };
  auto acc2_addr = [&](int buf) { return tmem_base + static_cast<uint32_t>(nbuf * n1p + buf * prm.n_tile2); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t a_bytes = prm.s2d_pitch ? static_cast<uint32_t>(prm.s2d_ny * prm.s2d_pitch * prm.row_bytes_a)
                                           : prm.a_nbox * prm.a_box_rows * prm.row_bytes_a;
    auto load_b = [&](const uint8_t* src, uint32_t bytes) {
      mbar_wait(b_empty(sb), pb ^ 1);
      if (elect_one()) {
        mbar_expect_tx(b_full(sb), bytes);
        bulk_load(b_base + sb * prm.b_stage_bytes, src, bytes, b_full(sb));
      }
      __syncwarp();
      if (++sb == prm.nb) { sb = 0; pb ^= 1; }
    };
    if (prm.resident) {   // the block's weights: once per CTA
      if (elect_one()) {
        mbar_expect_tx(w_full, prm.w1_bytes + prm.w2_bytes);
        for (uint32_t off = 0; off < prm.w1_bytes; off += 16384u)
          bulk_load(b_base + off, static_cast<const uint8_t*>(prm.w1) + off, min(16384u, prm.w1_bytes - off), w_full);
        for (uint32_t off = 0; off < prm.w2_bytes; off += 16384u)
          bulk_load(b_base + prm.w1_bytes + off, static_cast<const uint8_t*>(prm.w2) + off, min(16384u, prm.w2_bytes - off), w_full);
      }
      __syncwarp();
    }
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        const int tile = blockIdx.x + s * gridDim.x;
        const int b = tile / prm.tiles_per_batch;
        const int q0 = (tile - b * prm.tiles_per_batch) * 128;
        for (int cb = 0; cb < prm.cblocks_a; ++cb) {
          mbar_wait(a_empty(sa), pa ^ 1);
          if (elect_one()) {
            mbar_expect_tx(a_full(sa), a_bytes);
            if (prm.s2d_pitch)   // (pixel pair x channel, row parity, x', y', frame): K block cb = row parity when there are two
              tma_load_5d(a_base + sa * prm.a_stage_bytes, &tmap_x, a_full(sa), 0, cb, 0, q0 / prm.s2d_pitch, b);
            else
            for (int bx = 0; bx < prm.a_nbox; ++bx)
              tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * prm.row_bytes_a, &tmap_x, a_full(sa),
                          cb * prm.kblock_a, q0 + prm.x_row0 + bx * prm.a_box_rows, b);
          }
          __syncwarp();
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
          if (prm.resident) continue;
          for (int nt = 0; nt < prm.n_tiles1; ++nt)
            for (int tap0 = 0; tap0 < prm.taps1; tap0 += prm.tg1) {
              const int cnt = min(prm.tg1, prm.taps1 - tap0);
              load_b(static_cast<const uint8_t*>(prm.w1) +
                         static_cast<size_t>((nt * prm.cblocks1 + cb) * prm.taps1 + tap0) * prm.b_tap_bytes1,
                     cnt * prm.b_tap_bytes1);
            }
        }
        // the tile's shortcut rows, consumed by epilogue 2 (the buffer is released at the end of epilogue 2 of tile s-2)
        if (prm.tma_epi && p.epi.res) {
          mbar_wait(r_empty(s & 1), ((s >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(r_full(s & 1), p.epi.res_half ? kH16Bytes : kRsBytes);
            tma_load_3d(rs_base + (s & 1) * kRsBytes, &tm_r, r_full(s & 1), 0, q0, b);
          }
          __syncwarp();
        }
      }
      if (s >= la && !prm.resident)
        for (int cb = 0; cb < prm.cblocks2; ++cb)
          load_b(static_cast<const uint8_t*>(prm.w2) + static_cast<size_t>(cb) * prm.b_tap_bytes2, prm.b_tap_bytes2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const int ksteps_full1 = prm.row_bytes1 >> 5;
    if (prm.resident) mbar_wait(w_full, 0);
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        // ---- expand: 9 row-shifted taps over the x tile ----
        const int buf = s % nbuf;
        mbar_wait(acc1_empty(buf), ((s / nbuf) & 1) ^ 1);
        tc_fence_after();
        for (int cb = 0; cb < prm.cblocks_a; ++cb) {
          const int rem = prm.c_in - cb * prm.kblock1;
          const int ksteps = rem >= prm.kblock1 ? ksteps_full1 : (rem + 15) >> 4;
          mbar_wait(a_full(sa), pa);
          const uint32_t a_tile = a_base + sa * prm.a_stage_bytes;
          if (prm.win_ksteps) {
            // K windows (resident weights only): the taps whose parity plane lies in K block cb of the x tile
            tc_fence_after();
            int row_off = 0;   // s2d by TMA: the tile starts at the first pixel of an image row, q0 somewhere inside it
            if (prm.s2d_pitch) {
              const int tile = blockIdx.x + s * gridDim.x;
              const int q0 = (tile - (tile / prm.tiles_per_batch) * prm.tiles_per_batch) * 128;
              row_off = q0 - (q0 / prm.s2d_pitch) * prm.s2d_pitch;
            }
            if (elect_one()) {
              for (int nt = 0; nt < prm.n_tiles1; ++nt)
                for (int t = 0; t < prm.taps1; ++t) {
                  if (prm.tap_cb[t] != cb) continue;
                  const uint32_t wt = b_base + static_cast<uint32_t>(nt * prm.taps1 + t) * prm.b_tap_bytes1;
                  const uint64_t db = prm.desc_hi1 | ((wt & 0x3FFFF) >> 4);
                  const uint64_t da = prm.desc_hi_a |
                                      (((a_tile + (row_off + prm.rel_shift1[t]) * prm.row_bytes_a + prm.tap_kofs[t] * 32) & 0x3FFFF) >> 4);
                  mma_f16_k4(acc1_addr(buf) + nt * prm.n_tile1, da, db, prm.idesc1, (cb | t) ? 1u : 0u, prm.win_ksteps);
                }
              tc_commit(a_empty(sa));
              if (cb == prm.cblocks_a - 1) tc_commit(acc1_full(buf));
            }
            __syncwarp();
          } else if (prm.resident) {
            tc_fence_after();
            if (elect_one()) {
              for (int nt = 0; nt < prm.n_tiles1; ++nt)
                for (int t = 0; t < prm.taps1; ++t) {
                  const uint32_t wt = b_base + static_cast<uint32_t>((nt * prm.cblocks1 + cb) * prm.taps1 + t) * prm.b_tap_bytes1;
                  const uint64_t db = prm.desc_hi1 | ((wt & 0x3FFFF) >> 4);
                  const uint64_t da = prm.desc_hi1 | (((a_tile + prm.rel_shift1[t] * prm.row_bytes1) & 0x3FFFF) >> 4);
                  mma_f16_k4(acc1_addr(buf) + nt * prm.n_tile1, da, db, prm.idesc1, (cb | t) ? 1u : 0u, ksteps);
                }
              tc_commit(a_empty(sa));
              if (cb == prm.cblocks1 - 1) tc_commit(acc1_full(buf));
            }
            __syncwarp();
          } else
          for (int nt = 0; nt < prm.n_tiles1; ++nt)
            for (int tap0 = 0; tap0 < prm.taps1; tap0 += prm.tg1) {
              const int cnt = min(prm.tg1, prm.taps1 - tap0);
              mbar_wait(b_full(sb), pb);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_tile = b_base + sb * prm.b_stage_bytes;
                for (int t = 0; t < cnt; ++t) {
                  const uint64_t db = prm.desc_hi1 | (((b_tile + t * prm.b_tap_bytes1) & 0x3FFFF) >> 4);
                  const uint64_t da = prm.desc_hi1 | (((a_tile + prm.rel_shift1[tap0 + t] * prm.row_bytes1) & 0x3FFFF) >> 4);
                  mma_f16_k4(acc1_addr(buf) + nt * prm.n_tile1, da, db, prm.idesc1, (cb | tap0 | t) ? 1u : 0u, ksteps);
                }
                tc_commit(b_empty(sb));
                const bool last = nt == prm.n_tiles1 - 1 && tap0 + cnt >= prm.taps1;
                if (last) tc_commit(a_empty(sa));
                if (last && cb == prm.cblocks1 - 1) tc_commit(acc1_full(buf));
              }
              __syncwarp();
              if (++sb == prm.nb) { sb = 0; pb ^= 1; }
            }
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
        }
      }
      if (s >= la) {
        // ---- project: A = the T tile, K = n1 ----
        const int i = s - la;
        const int buf = i % nbuf;
        mbar_wait(t_full(buf), (i / nbuf) & 1);
        mbar_wait(acc2_empty(buf), ((i / nbuf) & 1) ^ 1);
        tc_fence_after();
        if (prm.resident) {
          if (elect_one()) {
            for (int cb = 0; cb < prm.cblocks2; ++cb) {
              const int rem = prm.n1 - cb * 64;
              const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
              const uint64_t db = prm.desc_hi2 | (((b_base + prm.w1_bytes + cb * prm.b_tap_bytes2) & 0x3FFFF) >> 4);
              const uint64_t da = prm.desc_hi2 | (((t_base + buf * prm.t_buf_bytes + cb * kTkbBytes) & 0x3FFFF) >> 4);
              mma_f16_k4(acc2_addr(buf), da, db, prm.idesc2, cb ? 1u : 0u, ksteps);
            }
            tc_commit(acc2_full(buf));
            tc_commit(t_empty(buf));
          }
          __syncwarp();
        } else
        for (int cb = 0; cb < prm.cblocks2; ++cb) {
          const int rem = prm.n1 - cb * 64;
          const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
          mbar_wait(b_full(sb), pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t db = prm.desc_hi2 | (((b_base + sb * prm.b_stage_bytes) & 0x3FFFF) >> 4);
            const uint64_t da = prm.desc_hi2 | (((t_base + buf * prm.t_buf_bytes + cb * kTkbBytes) & 0x3FFFF) >> 4);
            mma_f16_k4(acc2_addr(buf), da, db, prm.idesc2, cb ? 1u : 0u, ksteps);
            tc_commit(b_empty(sb));
            if (cb == prm.cblocks2 - 1) {
              tc_commit(acc2_full(buf));
              tc_commit(t_empty(buf));
            }
          }
          __syncwarp();
          if (++sb == prm.nb) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3, grp = ew >> 2;   // TMEM lane quadrant / column group of this warp
    const int nchunks = (prm.n1 + kEpiUnitCols - 1) / kEpiUnitCols;
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        // ---- epilogue 1: acc1 -> + bias -> SiLU -> T tile (fp16, SWIZZLE_128B K-major operand layout) ----
        const int buf = s % nbuf;
        mbar_wait(acc1_full(buf), (s / nbuf) & 1);
        mbar_wait(t_empty(buf), ((s / nbuf) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = acc1_addr(buf) + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t tt = t_base + buf * prm.t_buf_bytes;
        const int row = quad * 32 + lane;
        const uint32_t swz = row & 7;
        for (int ci = grp; ci < nchunks; ci += kEpiGroups) {
          const int c0 = ci * kEpiUnitCols;
          uint32_t r[16];
          tmem_ld16(tacc + c0, r);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(bias1_smem + 4u * (c0 + 4 * j)));
            const float v0 = fast_silu(__uint_as_float(r[4 * j]) + b4.x), v1 = fast_silu(__uint_as_float(r[4 * j + 1]) + b4.y);
            const float v2 = fast_silu(__uint_as_float(r[4 * j + 2]) + b4.z), v3 = fast_silu(__uint_as_float(r[4 * j + 3]) + b4.w);
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[2 * j]) : "f"(v1), "f"(v0));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[2 * j + 1]) : "f"(v3), "f"(v2));
          }
          // 16 halves = 32 bytes = two 16-byte chunks of this row inside K block kb
          const int kb = c0 >> 6;
          const int chunk0 = (c0 & 63) >> 3;
          const uint32_t rbase = tt + kb * kTkbBytes + row * 128;
#pragma unroll
          for (int m = 0; m < 2; ++m)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((chunk0 + m) ^ swz) << 4)),
                         "r"(pk[4 * m]), "r"(pk[4 * m + 1]), "r"(pk[4 * m + 2]), "r"(pk[4 * m + 3])
                         : "memory");
        }
        fence_proxy_async();   // generic-proxy writes of T -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_full(buf));
          mbar_arrive(acc1_empty(buf));
        }
      }
      if (s >= la) {
        // ---- epilogue 2: acc2 -> + bias (+ fp32 shortcut) -> image-border mask -> global, a row per thread ----
        // Every warp takes 32 rows x n_tile2 / 4 columns.  The shortcut rows are requested BEFORE the wait for the project
        // MMAs: with the engine's unit epilogue (8 of the 16 warps busy, loads issued after the accumulators arrive) this
        // phase was a 3.5 k-cycle latency chain per tile and bounded the kernel (profiles/README.md round 2).
        const int i = s - la;
        const int tile = blockIdx.x + i * gridDim.x;
        const int b = tile / prm.tiles_per_batch;
        const int q = (tile - b * prm.tiles_per_batch) * 128 + quad * 32 + lane;
        const int buf = i % nbuf;
        const Epilogue& e = p.epi;
        if (prm.tma_epi) {
          // ---- C_out = 32: whole 128-byte rows through SMEM tiles.  The shortcut tile arrived by TMA (requested with the x
          // tile), every thread updates its row's 32 bytes in place (fp32) and writes the fp16 copy, and ONE thread per
          // TMEM lane quadrant hands 32 rows x 128 / 64 bytes to the TMA unit.  (A row per thread straight to global
          // memory touches 32 different lines per warp instruction: the LSU, not the MMAs, bounded the kernel.) ----
          const bool issuer = grp == 0 && lane == 0;
          const int rbuf = i & 1;
          named_bar_er(1 + quad, 128);   // (the issuer comes here after its stores of tile i-1 have read their SMEM tiles)
          if (e.res) mbar_wait(r_full(rbuf), (i >> 1) & 1);
          mbar_wait(acc2_full(buf), (i / nbuf) & 1);
          tc_fence_after();
          uint32_t r[16];
          tmem_ld8(acc2_addr(buf) + (static_cast<uint32_t>(quad * 32) << 16) + grp * 8, r);
          const int row = quad * 32 + lane;
          bool keep = q < p.l_out;
          if (e.mask_mode == M2S_MASK_PITCH) {
            const int drow = q + p.d_row_offset;
            const int mi = drow / e.pitch, mj = drow - mi * e.pitch;
            keep = keep && mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi;
          }
          const uint32_t a32 = rs_base + rbuf * kRsBytes + row * 128;
          float4 rv[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
          if (e.res && e.res_half) {
            // fp16 residual stream: the tile is 128 rows x 64 bytes (SWIZZLE_64B) at the start of the buffer
            uint32_t u0, u1, u2, u3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3)
                         : "r"(rs_base + rbuf * kRsBytes + row * 64 + ((grp ^ ((row >> 1) & 3)) << 4)));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u0));
            const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&u1));
            const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&u2));
            const float2 d2 = __half22float2(*reinterpret_cast<const __half2*>(&u3));
            rv[0] = make_float4(a.x, a.y, b2.x, b2.y);
            rv[1] = make_float4(c2.x, c2.y, d2.x, d2.y);
          } else if (e.res) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rv[m].x), "=f"(rv[m].y), "=f"(rv[m].z), "=f"(rv[m].w)
                           : "r"(a32 + (((2 * grp + m) ^ (row & 7)) << 4)));
          }
          tmem_ld_wait();
          uint32_t hp[4];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(bias_smem + 4u * (grp * 8 + 4 * m)));
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (keep)
              o = make_float4(__uint_as_float(r[4 * m]) + b4.x + rv[m].x, __uint_as_float(r[4 * m + 1]) + b4.y + rv[m].y,
                              __uint_as_float(r[4 * m + 2]) + b4.z + rv[m].z, __uint_as_float(r[4 * m + 3]) + b4.w + rv[m].w);
            if (p.d)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a32 + (((2 * grp + m) ^ (row & 7)) << 4)),
                           "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                           : "memory");
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hp[2 * m]) : "f"(o.y), "f"(o.x));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hp[2 * m + 1]) : "f"(o.w), "f"(o.z));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h16_base + row * 64 + ((grp ^ ((row >> 1) & 3)) << 4)),
                       "r"(hp[0]), "r"(hp[1]), "r"(hp[2]), "r"(hp[3])
                       : "memory");
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty(buf));
          named_bar_er(1 + quad, 128);
          if (issuer) {
            const int q0 = (tile - b * prm.tiles_per_batch) * 128 + quad * 32;
            if (p.d) tma_store_3d(&tm_d32, rs_base + rbuf * kRsBytes + quad * 4096, 0, q0, b);
            if (p.d16) tma_store_3d(&tm_d16, h16_base + quad * 2048, 0, q0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // wait until the TMA unit has read the tiles (a few hundred cycles; the other warps move on): the fp16 tile is
            // single-buffered, and the shortcut buffer can be refilled a whole tile period before it is needed again
            tma_store_wait_read();
            if (e.res) mbar_arrive(r_empty(rbuf));
          }
          continue;
        }
        const int cw = prm.n_tile2 >> 2;                 // 8 or 16 columns per warp
        const int col0 = grp * cw;
        const bool row_ok = q < p.l_out;
        const size_t grow = static_cast<size_t>(b) * p.d_batch_rows + p.d_row_offset + q;
        bool keep = row_ok;
        if (e.mask_mode == M2S_MASK_PITCH) {
          const int drow = q + p.d_row_offset;
          const int mi = drow / e.pitch, mj = drow - mi * e.pitch;
          keep = keep && mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi;
        }
        float4 rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = col0 + 4 * j;
          rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * j < cw && row_ok && c < p.n && e.res) {
            if (e.res_half) {
              const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(e.res) + grow * e.res_ld + c);
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
              const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
              rv[j] = make_float4(a.x, a.y, b2.x, b2.y);
            } else {
              rv[j] = *reinterpret_cast<const float4*>(e.res + grow * e.res_ld + c);
            }
          }
        }
        mbar_wait(acc2_full(buf), (i / nbuf) & 1);
        tc_fence_after();
        const uint32_t tacc = acc2_addr(buf) + (static_cast<uint32_t>(quad * 32) << 16) + col0;
        uint32_t r[16];
        if (cw == 16) tmem_ld16(tacc, r);
        else tmem_ld8(tacc, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = col0 + 4 * j;
          if (4 * j < cw && row_ok && c < p.n) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(bias_smem + 4u * c));
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (keep)
              o = make_float4(__uint_as_float(r[4 * j]) + b4.x + rv[j].x, __uint_as_float(r[4 * j + 1]) + b4.y + rv[j].y,
                              __uint_as_float(r[4 * j + 2]) + b4.z + rv[j].z, __uint_as_float(r[4 * j + 3]) + b4.w + rv[j].w);
            if (p.d) *reinterpret_cast<float4*>(p.d + grow * p.d_ld + c) = o;
            if (p.d16) {
              uint2 h;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(o.y), "f"(o.x));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(o.w), "f"(o.z));
              *reinterpret_cast<uint2*>(static_cast<__half*>(p.d16) + grow * p.d_ld + c) = h;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty(buf));
      }
    }
    if (prm.tma_epi && grp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_er() {
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

}  // namespace

namespace {
int launch_er(const ErParams& prm_in, const ConvProblem& p1, const ConvProblem& p2, uint32_t smem_bytes, cudaStream_t stream,
              const ErS2d* s2d) {
  ErParams prm = prm_in;
  EncodeTiledFn enc = encode_fn_er();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM (whole-TMEM allocation)

  CUtensorMap tmap;
  CUresult cr = CUDA_SUCCESS;
  if (!s2d) {
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p1.c_in), static_cast<cuuint64_t>(p1.a_rows),
                          static_cast<cuuint64_t>(p1.batch)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p1.a_ld) * 2ull,
                             static_cast<cuuint64_t>(p1.a_batch_rows) * static_cast<cuuint64_t>(p1.a_ld) * 2ull};
    if (p1.batch == 1) gstride[1] = gstride[0] * static_cast<cuuint64_t>(p1.a_rows > 0 ? p1.a_rows : 1);
    cuuint32_t box[3] = {static_cast<cuuint32_t>(prm.kblock_a), static_cast<cuuint32_t>(prm.a_box_rows), 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<float*>(p1.a), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE,
             prm.row_bytes_a == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "fused EdgeResidual: x tensor map failed (%d)", static_cast<int>(cr));
  }
  if (s2d) {
    // The zero-bordered NHWC input (pitch W + 2, origin (1, 1)) seen as (pixel pair x channel, row parity, x', y', frame):
    // element (2 x' + px, 2 y' + py).  The dimensions are not in stride order on purpose: a box lands in SMEM as
    // [y'][x'][py][px][c], i.e. one row per space-to-depth pixel with its parity planes side by side in K.  x' = W / 2 and
    // y' = H / 2 (the conv's right / bottom padding) and the pitch's extra column are out of bounds: the TMA unit writes zeros.
    const int C = s2d->c, Wp = s2d->w + 2;
    const cuuint64_t esz = 2;
    cuuint64_t gd5[5] = {static_cast<cuuint64_t>(2 * C), 2ull, static_cast<cuuint64_t>(s2d->w / 2),
                         static_cast<cuuint64_t>(s2d->h / 2), static_cast<cuuint64_t>(p1.batch)};
    cuuint64_t gs5[4] = {static_cast<cuuint64_t>(Wp) * C * esz, 2ull * C * esz, 2ull * Wp * C * esz,
                         static_cast<cuuint64_t>(s2d->frame_rows) * C * esz};
    // one row parity per K block: the box's inner extent (a pixel pair, 4 C bytes) is then exactly the swizzle span.  (A
    // 64-byte inner extent under SWIZZLE_128B -- both parities of C = 16 in one 128-byte row -- does NOT land in the
    // operand layout: measured, tests/test_mbconv_gpu.py mode 196.)
    cuuint32_t bx5[5] = {static_cast<cuuint32_t>(2 * C), 1u, static_cast<cuuint32_t>(prm.s2d_pitch),
                         static_cast<cuuint32_t>(prm.s2d_ny), 1u};
    cuuint32_t es5[5] = {1u, 1u, 1u, 1u, 1u};
    const __half* base = reinterpret_cast<const __half*>(p1.a) + static_cast<size_t>(Wp + 1) * C;
    cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), gd5, gs5, bx5, es5,
             CU_TENSOR_MAP_INTERLEAVE_NONE, prm.row_bytes_a == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "fused EdgeResidual: space-to-depth tensor map failed (%d)", static_cast<int>(cr));
  }
  // epilogue-2 tiles (tma_epi): shortcut / fp32 output / fp16 output as (32 channels, l_out rows, batch) tensors whose row 0
  // is output row q = 0 of a frame (base moved by d_row_offset rows); rows >= l_out are clipped / zero-filled by the TMA unit
  CUtensorMap tm_r = tmap, tm_d32 = tmap, tm_d16 = tmap;
  if (prm.tma_epi) {
    auto map3 = [&](CUtensorMap* tm, const void* base, int esize, long long batch_rows, int ld, int box_rows,
                    CUtensorMapSwizzle sw, const char* what) -> int {
      cuuint64_t gd[3] = {32ull, static_cast<cuuint64_t>(p2.l_out), static_cast<cuuint64_t>(p2.batch)};
      cuuint64_t gs[2] = {static_cast<cuuint64_t>(ld) * esize, static_cast<cuuint64_t>(batch_rows) * ld * esize};
      if (p2.batch == 1) gs[1] = gs[0] * static_cast<cuuint64_t>(p2.l_out);
      cuuint32_t bx[3] = {32u, static_cast<cuuint32_t>(box_rows), 1u};
      cuuint32_t es[3] = {1u, 1u, 1u};
      CUresult r = enc(tm, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                       const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "fused EdgeResidual: %s tensor map failed (%d)", what, static_cast<int>(r));
      return M2S_OK;
    };
    const size_t off = static_cast<size_t>(p2.d_row_offset);
    if (p2.epi.res && p2.epi.res_half)
      M2S_TRY(map3(&tm_r, reinterpret_cast<const __half*>(p2.epi.res) + off * p2.epi.res_ld, 2, p2.d_batch_rows, p2.epi.res_ld, 128,
                   CU_TENSOR_MAP_SWIZZLE_64B, "shortcut (fp16)"));
    else if (p2.epi.res)
      M2S_TRY(map3(&tm_r, p2.epi.res + off * p2.epi.res_ld, 4, p2.d_batch_rows, p2.epi.res_ld, 128, CU_TENSOR_MAP_SWIZZLE_128B, "shortcut"));
    if (p2.d) M2S_TRY(map3(&tm_d32, p2.d + off * p2.d_ld, 4, p2.d_batch_rows, p2.d_ld, 32, CU_TENSOR_MAP_SWIZZLE_128B, "fp32 output"));
    if (p2.d16)
      M2S_TRY(map3(&tm_d16, static_cast<const __half*>(p2.d16) + off * p2.d_ld, 2, p2.d_batch_rows, p2.d_ld, 32, CU_TENSOR_MAP_SWIZZLE_64B, "fp16 output"));
  }
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(fused_er_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  int grid = sm_count();
  if (grid > prm.total_tiles) grid = prm.total_tiles;
  M2S_TRY(profile_before(stream));
  fused_er_kernel<<<grid, kThreads, smem_bytes, stream>>>(tmap, tm_r, tm_d32, tm_d16, prm);
  M2S_CUDA_OK(cudaGetLastError());
  const double rows = static_cast<double>(p2.batch) * p2.l_out;
  const double k1 = prm.win_ksteps ? 16.0 * prm.win_ksteps * p1.taps : static_cast<double>(p1.c_in) * p1.taps;
  return profile_after(stream, 2.0 * rows * p1.n * (k1 + p2.n));
}
}  // namespace

// p1: the expand conv as the engine would run it (a = x fp16, taps / shifts, bias = folded bn1, act = SiLU; its outputs are
//     ignored: T stays on chip).  p2: the project conv's bias, epilogue and outputs (its `a` is ignored; c_in == p1.n).
bool fused_er_supported(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2) {
  if (!p1.a_half || !w1.half || w2.half != 1) return false;                      // T is written with 128-byte rows
  if (p1.epi.act != M2S_ACT_SILU || p2.taps != 1 || p2.shift[0] != 0 || p2.c_in != p1.n) return false;
  if (p1.n > 256 || p1.n % 16 || w2.n_tiles != 1 || (w2.n_tile != 32 && w2.n_tile != 64)) return false;
  if (p2.epi.act != M2S_ACT_NONE || p2.epi.accum || p2.epi.out_scale != 1.f || p2.epi.res_after_act ||
      (p2.epi.res && p2.epi.res_inv_slope != 1.f) || p2.epi.mask_mode == M2S_MASK_LEN || p2.n % 4 || p2.d_ld % 4)
    return false;
  if (p1.batch != p2.batch || p1.l_out != p2.l_out) return false;
  const int epi = choose_epilogue(p2.epi);
  if (epi < 0 || p2.epi.res_hi || p2.d16_lo) return false;
  if (w1.n_tile * w1.n_tiles + w2.n_tile > kTmemCols) return false;
  for (int j = 0; j < p1.taps; ++j)
    if (p1.shift[j] < 0) return false;
  if (p1.tap_ksteps) {
    // K windows: x rows of 4 c_in' channels in 64-channel K blocks, weights c_in' = 16 tap_ksteps wide in one K block
    if (w1.cblocks != 1 || w1.c_in != 16 * p1.tap_ksteps || p1.c_in != 4 * w1.c_in || p1.a_ld != p1.c_in || p1.kofs[0] != 0) return false;
    for (int j = 0; j < p1.taps; ++j)
      if (p1.kofs[j] < 0 || (p1.kofs[j] & 3) + p1.tap_ksteps > 4 || 16 * (p1.kofs[j] + p1.tap_ksteps) > p1.c_in) return false;
  }
  return true;
}

int fused_er(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2,
             cudaStream_t stream, const ErS2d* s2d) {
  if (!fused_er_supported(p1, w1, p2, w2)) return fail(M2S_ERR_UNSUPPORTED, "block not supported by the fused EdgeResidual kernel");
  if (s2d && (!p1.tap_ksteps || s2d->c != w1.c_in || (s2d->c != 16 && s2d->c != 32) || s2d->h % 2 || s2d->w % 2 ||
              s2d->w / 2 + 2 > 256))
    return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: space-to-depth geometry not supported");
  if (p2.batch <= 0 || p2.l_out <= 0) return M2S_OK;
  ErParams prm{};
  prm.p2 = p2;
  finalize_epilogue(&prm.p2.epi, static_cast<long long>(p2.l_out) + p2.d_row_offset);
  prm.bias1 = p1.epi.bias;
  prm.w1 = w1.dev;
  prm.w2 = w2.dev;
  prm.taps1 = p1.taps;
  prm.c_in = p1.c_in;
  prm.n1 = p1.n;
  prm.n_tile1 = w1.n_tile; prm.n_tiles1 = w1.n_tiles;
  prm.n_tile2 = w2.n_tile;
  prm.cblocks1 = w1.cblocks;
  prm.cblocks2 = w2.cblocks;
  prm.kblock1 = w1.kblock; prm.row_bytes1 = w1.row_bytes;
  prm.desc_hi1 = make_desc_hi(w1.row_bytes);
  prm.cblocks_a = w1.cblocks; prm.kblock_a = w1.kblock; prm.row_bytes_a = w1.row_bytes; prm.desc_hi_a = prm.desc_hi1;
  if (p1.tap_ksteps) {
    prm.win_ksteps = p1.tap_ksteps;
    prm.kblock_a = 64; prm.row_bytes_a = 128; prm.desc_hi_a = make_desc_hi(128);
    prm.cblocks_a = (p1.c_in + 63) / 64;
    for (int j = 0; j < p1.taps; ++j) { prm.tap_cb[j] = p1.kofs[j] >> 2; prm.tap_kofs[j] = p1.kofs[j] & 3; }
    if (s2d) {
      // TMA gather: a K block is one row parity = a pixel pair of 2 C channels (64- or 128-byte rows)
      const int kpb = w1.c_in / 8;   // K-steps per block
      prm.kblock_a = 2 * w1.c_in; prm.row_bytes_a = 4 * w1.c_in; prm.desc_hi_a = make_desc_hi(prm.row_bytes_a);
      prm.cblocks_a = 2;
      for (int j = 0; j < p1.taps; ++j) { prm.tap_cb[j] = p1.kofs[j] / kpb; prm.tap_kofs[j] = p1.kofs[j] % kpb; }
    }
  }
  prm.desc_hi2 = make_desc_hi(128);
  prm.idesc1 = (1u << 4) | (static_cast<uint32_t>(prm.n_tile1 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  prm.idesc2 = (1u << 4) | (static_cast<uint32_t>(prm.n_tile2 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  int smin = p1.shift[0], smax = p1.shift[0];
  for (int j = 1; j < p1.taps; ++j) { smin = p1.shift[j] < smin ? p1.shift[j] : smin; smax = p1.shift[j] > smax ? p1.shift[j] : smax; }
  for (int j = 0; j < p1.taps; ++j) prm.rel_shift1[j] = p1.shift[j] - smin;
  prm.x_row0 = smin;
  prm.tiles_per_batch = (p2.l_out + 127) / 128;
  prm.total_tiles = p2.batch * prm.tiles_per_batch;
  const int n1p = prm.n_tile1 * prm.n_tiles1;
  prm.nbuf = (2 * (n1p + prm.n_tile2) <= kTmemCols) ? 2 : 1;

  const int a_rows_needed = 128 + smax - smin;
  prm.a_nbox = (a_rows_needed + 255) / 256;
  prm.a_box_rows = (((a_rows_needed + prm.a_nbox - 1) / prm.a_nbox) + 7) / 8 * 8;
  prm.a_stage_bytes = (static_cast<uint32_t>(prm.a_nbox * prm.a_box_rows * prm.row_bytes_a) + 1023u) & ~1023u;
  if (s2d) {
    // whole image rows: the tile's first row may sit anywhere inside the first one
    prm.s2d_pitch = s2d->w / 2 + 2;
    prm.s2d_ny = (a_rows_needed + prm.s2d_pitch - 1 + prm.s2d_pitch - 1) / prm.s2d_pitch;
    if (prm.s2d_ny > 256) return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: space-to-depth tile too tall");
    prm.a_stage_bytes = (static_cast<uint32_t>(prm.s2d_ny * prm.s2d_pitch * prm.row_bytes_a) + 1023u) & ~1023u;
  }
  prm.t_buf_bytes = prm.cblocks2 * kTkbBytes;
  prm.b_tap_bytes1 = static_cast<uint32_t>(prm.n_tile1 * prm.row_bytes1);
  prm.b_tap_bytes2 = static_cast<uint32_t>(prm.n_tile2 * 128);
  if ((prm.b_tap_bytes1 & 1023u) || (prm.b_tap_bytes2 & 1023u))
    return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: weight blocks not 1 KB aligned (n_tile %d / %d)", prm.n_tile1, prm.n_tile2);
  // SMEM plan: A stages, T buffers, weights (resident when the whole block's fit next to two T buffers, else a ring of
  // tap groups re-streamed per tile), epilogue staging
  const uint32_t fixed = 2048u + 1024u + 1024u;   // barriers + bias1, bias2, alignment slack
  const uint32_t budget = 225u * 1024u;
  prm.w1_bytes = static_cast<uint32_t>(w1.n_tiles * w1.cblocks * p1.taps) * prm.b_tap_bytes1;
  prm.w2_bytes = static_cast<uint32_t>(w2.cblocks) * prm.b_tap_bytes2;
  if (fixed + 2 * prm.a_stage_bytes + prm.nbuf * prm.t_buf_bytes + prm.w1_bytes + prm.w2_bytes <= budget) {
    prm.resident = 1;
    uint32_t used = fixed + 2 * prm.a_stage_bytes + prm.nbuf * prm.t_buf_bytes + prm.w1_bytes + prm.w2_bytes;
    static const bool tma_epi_on = !(std::getenv("M2S_ER_TMA_EPI") && std::atoi(std::getenv("M2S_ER_TMA_EPI")) == 0);
    if (tma_epi_on && p2.n == 32 && prm.n_tile2 == 32 && p2.d_ld == 32 && (!p2.epi.res || p2.epi.res_ld == 32) &&
        !(p2.d && p2.epi.res && p2.epi.res_half) &&   // (an fp16 shortcut tile and an fp32 output tile would share a buffer)
        used + 2 * kRsBytes + kH16Bytes <= budget) {
      prm.tma_epi = 1;
      used += 2 * kRsBytes + kH16Bytes;
    }
    int na_r = 2;
    while (na_r < kMaxStagesA && used + prm.a_stage_bytes <= budget) { ++na_r; used += prm.a_stage_bytes; }
    if (prm.win_ksteps && na_r < prm.cblocks_a) return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: x tile of the K-window mode does not fit SMEM");
    prm.na = na_r;
    prm.nb = 0;
    prm.tg1 = 1;
    prm.b_stage_bytes = 0;
    return launch_er(prm, p1, p2, used + 1024u, stream, s2d);
  }
  if (prm.win_ksteps) return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: the K-window mode needs resident weights");
  int na = 2;
  uint32_t used = fixed + na * prm.a_stage_bytes + prm.nbuf * prm.t_buf_bytes;
  if (used + 2 * prm.b_tap_bytes1 > budget && prm.nbuf == 2) {
    prm.nbuf = 1;
    used = fixed + na * prm.a_stage_bytes + prm.t_buf_bytes;
  }
  if (used + 2 * prm.b_tap_bytes1 > budget) return fail(M2S_ERR_UNSUPPORTED, "fused EdgeResidual: tile does not fit SMEM");
  int tg = static_cast<int>(32768u / prm.b_tap_bytes1);
  if (tg < 1) tg = 1;
  if (tg > p1.taps) tg = p1.taps;
  while (tg > 1 && used + 2u * tg * prm.b_tap_bytes1 > budget) --tg;
  prm.tg1 = tg;
  prm.b_stage_bytes = static_cast<uint32_t>(tg) * prm.b_tap_bytes1;
  if (prm.b_stage_bytes < prm.b_tap_bytes2) prm.b_stage_bytes = prm.b_tap_bytes2;
  int nb = 2;
  while (nb < 4 && used + (nb + 1) * prm.b_stage_bytes <= budget) ++nb;
  used += nb * prm.b_stage_bytes;
  while (na < 3 && used + prm.a_stage_bytes <= budget) { ++na; used += prm.a_stage_bytes; }
  while (nb < kMaxStagesB && used + prm.b_stage_bytes <= budget) { ++nb; used += prm.b_stage_bytes; }
  prm.na = na;
  prm.nb = nb;
  return launch_er(prm, p1, p2, used + 1024u, stream, nullptr);
}

// Do the block's weights stay in SMEM for the whole launch?  (Otherwise they are re-streamed for every 128-row tile:
// 290 KB per tile at stage 2, which makes the fused kernel no faster than two launches -- profiles/README.md round 2.)
bool fused_er_resident(const PackedWeights& w1, const PackedWeights& w2) {
  const size_t w = static_cast<size_t>(w1.n_tiles) * w1.cblocks * w1.taps * w1.n_tile * w1.row_bytes +
                   static_cast<size_t>(w2.cblocks) * w2.n_tile * 128;
  return w <= 96 * 1024;
}

}  // namespace m2s

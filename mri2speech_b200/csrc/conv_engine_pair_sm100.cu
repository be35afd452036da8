// CTA-pair variant of the conv engine: tcgen05.mma.cta_group::2 (M = 256 across two SMs of a TPC).
//
// Why: with one CTA the SS-mode MMA is bound by SMEM operand fetch (measured ~75 B/clk/SM: 4 KB of A + 32*N B of B
// per K=8 step), which caps kind::tf32 at ~55 % of the tensor peak for N = 128.  In pair mode each SM fetches its own
// 128 rows of A but only HALF of the weight tile (N/2 rows): for N = 256 that is 8 KB per 128x256x8 MACs, i.e. the
// MMA becomes math-bound.  Used for the wide layers (N >= 128): HiFi-GAN stages 0-1, transposed convs, conv_pre,
// LSTM input projection, the wide encoder GEMMs.
//
// Layout of a tile: 256 rows x n_tile columns per pair; CTA rank r owns rows [q0 + 128 r, q0 + 128 r + 128) (its A box,
// its 128 TMEM lanes, its epilogue) and weight rows [r * n_tile/2, (r+1) * n_tile/2).  All "full" barriers live in the
// leader (rank 0): both CTAs' TMA loads complete_tx on them (cp.async.bulk.tensor ... .cta_group::2 with the
// leader's shared::cluster barrier address from mapa); "empty" / "accumulator full" signals reach both CTAs through multicast
// tcgen05.commit; the peer's epilogue releases the accumulator with a remote mbarrier arrive.
#include "engine_device.cuh"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace m2s {

using namespace engine;

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {  // arrive on `bar` (same offset) in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_f16_k4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum0,
                                            int ksteps) {
  mma2_f16(tmem_d, da, db, idesc, accum0);
  if (ksteps > 1) mma2_f16(tmem_d, da + 2, db + 2, idesc, 1u);
  if (ksteps > 2) mma2_f16(tmem_d, da + 4, db + 4, idesc, 1u);
  if (ksteps > 3) mma2_f16(tmem_d, da + 6, db + 6, idesc, 1u);
}
__device__ __forceinline__ void mma2_tf32_k4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum0,
                                             int ksteps) {
  mma2_tf32(tmem_d, da, db, idesc, accum0);
  if (ksteps > 1) mma2_tf32(tmem_d, da + 2, db + 2, idesc, 1u);
  if (ksteps > 2) mma2_tf32(tmem_d, da + 4, db + 4, idesc, 1u);
  if (ksteps > 3) mma2_tf32(tmem_d, da + 6, db + 6, idesc, 1u);
}

template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv_engine_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ EngineParams prm) {
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
  const uint32_t stage_base = bar_base + 1024u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const ConvProblem& p = prm.p;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < prm.na; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < prm.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < prm.nacc; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 2 * kEpiWarps); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  cluster_sync_all();  // barrier inits of both CTAs visible before any remote signal / TMA
  const uint32_t bias_smem = stage_bias(p, stage_base);
  if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int taps = p.taps;
  const int cblocks = prm.cblocks;
  const int n_tile = prm.n_tile;
  const int nh = n_tile >> 1;  // weight rows held by each CTA

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; signals land on the LEADER's full barriers) =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t row_bytes = prm.row_bytes;
    const uint32_t a_bytes = prm.a_nbox * prm.a_box_rows * row_bytes;
    for (int tile = pair; tile < prm.total_tiles; tile += npairs) {
      const int nt = tile % prm.n_tiles;
      const int mt = (tile / prm.n_tiles) % prm.tiles_per_batch;
      const int b = tile / (prm.n_tiles * prm.tiles_per_batch);
      const int q0 = mt * prm.m_tile + static_cast<int>(rank) * 128 * prm.msub;
      for (int cb = 0; cb < cblocks; ++cb) {
        mbar_wait(a_empty(sa), pa ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(a_full(sa), 2 * a_bytes);
          for (int bx = 0; bx < prm.a_nbox; ++bx)
            tma2_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * row_bytes, &tmap_a,
                         mapa(a_full(sa), 0), cb * prm.kblock, q0 + prm.shift_min + bx * prm.a_box_rows, b);
        }
        __syncwarp();
        if (++sa == prm.na) { sa = 0; pa ^= 1; }
        for (int tap0 = 0; tap0 < taps; tap0 += prm.tg) {
          mbar_wait(b_empty(sb), pb ^ 1);
          if (elect_one()) {
            if (leader) mbar_expect_tx(b_full(sb), 2 * prm.b_stage_bytes);
            // weight rows: (((nt*cblocks + cb)*2 + rank)*taps + tap0) * nh ; the box always spans tg taps
            const int row = (((nt * cblocks + cb) * 2 + static_cast<int>(rank)) * taps + tap0) * nh;
            tma2_load_2d(b_base + sb * prm.b_stage_bytes, &tmap_w, mapa(b_full(sb), 0), 0, row);
          }
          __syncwarp();
          if (++sb == prm.nb) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      const uint64_t desc_hi = prm.desc_hi;
      const uint32_t row_bytes = prm.row_bytes;
      const int ksteps_full = prm.row_bytes >> 5;
      for (int tile = pair; tile < prm.total_tiles; tile += npairs) {
        mbar_wait(acc_empty(acc), pacc ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride;
        for (int cb = 0; cb < cblocks; ++cb) {
          const int rem = p.c_in - cb * prm.kblock;
          const int ksteps = rem >= prm.kblock ? ksteps_full : (rem + prm.kstep_elems - 1) / prm.kstep_elems;
          mbar_wait(a_full(sa), pa);
          for (int tap0 = 0; tap0 < taps; tap0 += prm.tg) {
            const int cnt = min(prm.tg, taps - tap0);
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_tile = a_base + sa * prm.a_stage_bytes;
              const uint32_t b_tile = b_base + sb * prm.b_stage_bytes;
              for (int t = 0; t < cnt && !(prm.dbg & 8); ++t) {
                const uint64_t db = desc_hi | (((b_tile + t * prm.b_tap_bytes) & 0x3FFFF) >> 4);
                const uint64_t da = desc_hi | (((a_tile + prm.rel_shift[tap0 + t] * row_bytes) & 0x3FFFF) >> 4);
                const uint32_t first = (cb | tap0 | t) ? 1u : 0u;
                if (prm.half) {
                  mma2_f16_k4(tmem_acc, da, db, prm.idesc, first, ksteps);
                  if (prm.msub > 1) mma2_f16_k4(tmem_acc + n_tile, da + ((128 * row_bytes) >> 4), db, prm.idesc, first, ksteps);
                } else {
                  mma2_tf32_k4(tmem_acc, da, db, prm.idesc, first, ksteps);
                  if (prm.msub > 1) mma2_tf32_k4(tmem_acc + n_tile, da + ((128 * row_bytes) >> 4), db, prm.idesc, first, ksteps);
                }
              }
              const bool last_group = tap0 + cnt >= taps;
              tc_commit2(b_empty(sb));
              if (last_group) tc_commit2(a_empty(sa));
              if (cb == cblocks - 1 && last_group) tc_commit2(acc_full(acc));
            }
            __syncwarp();
            if (++sb == prm.nb) { sb = 0; pb ^= 1; }
          }
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
        }
        if (++acc == prm.nacc) { acc = 0; pacc ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (each CTA drains its own 128 TMEM lanes) =====================
    const int ew = warp - 2;
    const EpiWarp epw = make_epi_warp(p.epi, stage_base, ew, warp, lane, prm.dbg, bias_smem);
    int acc = 0;
    uint32_t pacc = 0;
    for (int tile = pair; tile < prm.total_tiles; tile += npairs) {
      const int nt = tile % prm.n_tiles;
      const int mt = (tile / prm.n_tiles) % prm.tiles_per_batch;
      const int b = tile / (prm.n_tiles * prm.tiles_per_batch);
      mbar_wait(acc_full(acc), pacc);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride + (static_cast<uint32_t>(epw.quad * 32) << 16);
      epilogue_tile<kEpi>(p, epw, tmem_acc, b, mt * prm.m_tile + static_cast<int>(rank) * 128 * prm.msub, nt * n_tile,
                          prm.msub, n_tile);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(acc_empty(acc));
        else mbar_arrive_remote(mapa(acc_empty(acc), 0));
      }
      if (++acc == prm.nacc) { acc = 0; pacc ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn2() {
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

// taps per weight stage in pair mode: each CTA's box holds tg * (n_tile/2) rows <= 256
int pair_taps_per_stage(int n_tile, int taps) {
  int tg = 256 / (n_tile / 2);
  if (tg > taps) tg = taps;
  return tg < 1 ? 1 : tg;
}

int conv_tcgen05_pair(const ConvProblem& p, const PackedWeights& w, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn2();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const EngineKnobs& knobs = engine_knobs();
  EngineParams prm{};
  prm.p = p;
  finalize_epilogue(&prm.p.epi, static_cast<long long>(p.l_out) + p.d_row_offset);
  prm.wpacked = w.dev_pair;
  prm.n_tile = w.n_tile_pair;
  prm.n_tiles = w.n_tiles_pair;
  prm.cblocks = w.cblocks;
  prm.half = w.half;
  prm.kblock = w.kblock;
  prm.row_bytes = w.row_bytes;
  prm.kstep_elems = w.half ? 16 : 8;
  prm.desc_hi = make_desc_hi(w.row_bytes);
  int smin = p.shift[0], smax = p.shift[0];
  for (int j = 1; j < p.taps; ++j) { smin = p.shift[j] < smin ? p.shift[j] : smin; smax = p.shift[j] > smax ? p.shift[j] : smax; }
  prm.shift_min = smin;
  for (int j = 0; j < p.taps; ++j) prm.rel_shift[j] = p.shift[j] - smin;
  const int halo = smax - smin;
  // two 128-row sub-tiles per CTA (M = 512 per pair) when the accumulators still double-buffer: halves the halo
  // re-load and the per-tile overhead of narrow-N layers
  int msub = 1;
  if (4 * prm.n_tile <= kTmemCols) {
    const long long tiles2 = static_cast<long long>(p.batch) * ((p.l_out + 511) / 512) * prm.n_tiles;
    if (tiles2 >= sm_count()) msub = 2;
  }
  if (knobs.msub == 1) msub = 1;
  prm.msub = msub;
  prm.m_tile = 256 * msub;  // per pair
  prm.tiles_per_batch = (p.l_out + prm.m_tile - 1) / prm.m_tile;
  prm.total_tiles = p.batch * prm.tiles_per_batch * prm.n_tiles;
  prm.nacc = (2 * msub * prm.n_tile <= kTmemCols) ? 2 : 1;
  prm.acc_stride = prm.nacc == 2 ? kTmemCols / 2 : 0;
  const int a_rows_needed = 128 * msub + halo;
  prm.a_nbox = (a_rows_needed + 255) / 256;
  prm.a_box_rows = (((a_rows_needed + prm.a_nbox - 1) / prm.a_nbox) + 7) / 8 * 8;
  prm.a_stage_bytes = (static_cast<uint32_t>(prm.a_nbox * prm.a_box_rows * w.row_bytes) + 1023u) & ~1023u;
  const int nh = prm.n_tile / 2;
  prm.b_tap_bytes = static_cast<uint32_t>(nh * w.row_bytes);
  prm.tg = pair_taps_per_stage(prm.n_tile, p.taps);
  // a tap's block must keep the 8-row swizzle phase: nh * row_bytes multiple of 1 KB (64-byte rows: nh % 16 == 0)
  while (prm.tg > 1 && ((prm.tg * prm.b_tap_bytes) & 1023u)) --prm.tg;
  prm.b_stage_bytes = static_cast<uint32_t>(prm.tg) * prm.b_tap_bytes;
  if (prm.b_stage_bytes & 1023u) return fail(M2S_ERR_UNSUPPORTED, "pair mode: weight stage not 1 KB aligned");
  const uint32_t bar_bytes = 1024 + kEpiSmemBytes;
  int na = 2, nb = 2;
  auto total = [&](int a, int b) { return a * prm.a_stage_bytes + b * prm.b_stage_bytes + bar_bytes + 1024u; };
  if (total(na, nb) > kSmemBudget + 24 * 1024) return fail(M2S_ERR_UNSUPPORTED, "pair mode: tile does not fit SMEM");
  while (nb < 4 && total(na, nb + 1) <= kSmemBudget) ++nb;
  while (na < 3 && total(na + 1, nb) <= kSmemBudget) ++na;
  while (nb < kMaxStagesB && total(na, nb + 1) <= kSmemBudget) ++nb;
  prm.na = na;
  prm.nb = nb;
  uint32_t smem_bytes = total(na, nb);
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  // instruction descriptor: D=f32, A=B=tf32, K-major both, N = n_tile, M = 256 (two CTAs x 128)
  prm.idesc = (1u << 4) | (w.half ? 0u : ((2u << 7) | (2u << 10))) | (static_cast<uint32_t>(prm.n_tile >> 3) << 17) |
              (static_cast<uint32_t>(256 >> 4) << 24);
  prm.trace = nullptr;
  prm.dbg = knobs.dbg;

  CUtensorMap tmap_a, tmap_w;
  {
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p.c_in), static_cast<cuuint64_t>(p.a_rows),
                          static_cast<cuuint64_t>(p.batch)};
    const cuuint64_t esize = w.half ? 2ull : 4ull;
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p.a_ld) * esize,
                             static_cast<cuuint64_t>(p.a_batch_rows) * static_cast<cuuint64_t>(p.a_ld) * esize};
    if (p.batch == 1) gstride[1] = gstride[0] * static_cast<cuuint64_t>(p.a_rows > 0 ? p.a_rows : 1);
    cuuint32_t box[3] = {static_cast<cuuint32_t>(w.kblock), static_cast<cuuint32_t>(prm.a_box_rows), 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = w.half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                          : (knobs.tmap_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    CUresult cr = enc(&tmap_a, dt, 3, const_cast<float*>(p.a), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      w.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "pair mode: A tensor map failed (%d)", static_cast<int>(cr));
  }
  {
    const cuuint64_t rows = static_cast<cuuint64_t>(w.n_tiles_pair) * w.cblocks * 2 * p.taps * nh;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(w.kblock), rows};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(w.row_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(w.kblock), static_cast<cuuint32_t>(prm.tg * nh)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = enc(&tmap_w, w.half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w.dev_pair,
                      gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      w.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "pair mode: W tensor map failed (%d)", static_cast<int>(cr));
  }

  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const EngineParams);
  static const KernelFn kernels[EPI_COUNT] = {conv_engine_pair_kernel<EPI_FULL>,  conv_engine_pair_kernel<EPI_FULL_SILU>, conv_engine_pair_kernel<EPI_BIAS>,
                                              conv_engine_pair_kernel<EPI_LRELU>, conv_engine_pair_kernel<EPI_SILU>,      conv_engine_pair_kernel<EPI_RES>,
                                              conv_engine_pair_kernel<EPI_RB>,    conv_engine_pair_kernel<EPI_RB_ACC>,
                                              conv_engine_pair_kernel<EPI_RB_S>,  conv_engine_pair_kernel<EPI_RB_ACC_S>};
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    for (KernelFn k : kernels) M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  const int epi = choose_epilogue(p.epi);

  int pairs = (knobs.max_ctas > 0 ? knobs.max_ctas : sm_count()) / 2;
  if (pairs > prm.total_tiles) pairs = prm.total_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  M2S_TRY(profile_before(stream));
  M2S_CUDA_OK(cudaLaunchKernelEx(&cfg, kernels[epi], tmap_a, tmap_w, prm));
  M2S_TRY(profile_after(stream, 2.0 * p.batch * static_cast<double>(p.l_out) * p.n * p.c_in * p.taps));
  return M2S_OK;
}

}  // namespace m2s

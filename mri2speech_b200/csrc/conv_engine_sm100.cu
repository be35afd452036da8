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
// Warp roles (576 threads, 1 CTA / SM, persistent over tiles):
//   warp 0   : TMA producer (one lane)
//   warp 1   : TMEM allocator + MMA issuer (one lane)
//   warps 2-17: epilogue (four per TMEM lane quadrant, 32 x 16 units): tcgen05.ld -> SMEM transpose -> bias / residual /
//              MRF accumulate / scale / activation / mask -> coalesced global stores
#include "engine_device.cuh"
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

namespace m2s {

using namespace engine;

namespace {

// ---- the kernel ----------------------------------------------------------------
// 576 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-17 = epilogue (four warps per
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
  const uint32_t stage_base = bar_base + 1024u;  // 16 epilogue warps x 2 KB transpose staging

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
  const uint32_t bias_smem = stage_bias(p, stage_base);
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
    const uint32_t row_bytes = prm.row_bytes;
    const uint32_t a_bytes = prm.a_nbox * prm.a_box_rows * row_bytes;
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
              tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * row_bytes, &tmap_a, a_full(sa),
                          cb * prm.kblock, q0 + prm.shift_min + bx * prm.a_box_rows, b);
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
                tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * row_bytes, &tmap_a, a_full(sa),
                            cb * prm.kblock, q0 + p.shift[tap0] + bx * prm.a_box_rows, b);
            }
            __syncwarp();
            if (++sa == prm.na) { sa = 0; pa ^= 1; }
          }
          mbar_wait(b_empty(sb), pb ^ 1);
          if (elect_one()) {
            const uint32_t bytes = cnt * prm.b_tap_bytes;
            mbar_expect_tx(b_full(sb), bytes);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(prm.wpacked) +
                                 (static_cast<size_t>((nt * cblocks + cb) * taps + tap0)) * prm.b_tap_bytes;
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
    const uint64_t desc_hi = prm.desc_hi;
    const uint32_t row_bytes = prm.row_bytes;
    const int ksteps_full = prm.row_bytes >> 5;  // a K-step covers 32 bytes of every row
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
      if (lane == 0) trace_stamp(prm, it, 3);
      mbar_wait(acc_empty(acc), pacc ^ 1);
      tc_fence_after();
      if (lane == 0) trace_stamp(prm, it, 4);
      const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride;
      for (int cb = 0; cb < cblocks; ++cb) {
        const int rem = p.c_in - cb * prm.kblock;
        const int ksteps = rem >= prm.kblock ? ksteps_full : (rem + prm.kstep_elems - 1) / prm.kstep_elems;
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
                const uint64_t da0 = desc_hi | (((a_tile + row_shift * row_bytes) & 0x3FFFF) >> 4);
                const uint32_t sub_step = (128 * row_bytes) >> 4;   // next 128-row sub-tile of the A stage
                if (prm.tap_ksteps) {   // per-tap K window (fp16 only): both descriptors advance 32 bytes per K-step
                  const uint32_t ko = 2u * prm.tap_kofs[tap0 + t];
                  for (int sub = 0; sub < msub; ++sub)
                    mma_f16_k4(tmem_acc + sub * n_tile, da0 + ko + sub * sub_step, db + ko, prm.idesc, first, prm.tap_ksteps);
                } else if (prm.half) {
                  for (int sub = 0; sub < msub; ++sub)
                    mma_f16_k4(tmem_acc + sub * n_tile, da0 + sub * sub_step, db, prm.idesc, first, ksteps);
                } else {
                  for (int sub = 0; sub < msub; ++sub)
                    mma_tf32_k4(tmem_acc + sub * n_tile, da0 + sub * sub_step, db, prm.idesc, first, ksteps);
                }
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
    // traffic: 4 lanes cover one 64-byte row segment, a warp instruction covers 8 rows.
    const int ew = warp - 2;  // 0..15
    const EpiWarp epw = make_epi_warp(p.epi, stage_base, ew, warp, lane, prm.dbg, bias_smem);
    int acc = 0;
    uint32_t pacc = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % prm.n_tiles;
      const int mt = (tile / prm.n_tiles) % prm.tiles_per_batch;
      const int b = tile / (prm.n_tiles * prm.tiles_per_batch);
      if (ew == 0 && lane == 0) trace_stamp(prm, it, 6);
      mbar_wait(acc_full(acc), pacc);
      tc_fence_after();
      if (ew == 0 && lane == 0) trace_stamp(prm, it, 7);
      const uint32_t tmem_acc = tmem_base + acc * prm.acc_stride + (static_cast<uint32_t>(epw.quad * 32) << 16);
      epilogue_tile<kEpi>(p, epw, tmem_acc, b, mt * prm.m_tile, nt * n_tile, msub, n_tile);
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
    x.pair = env_int("M2S_ENGINE_PAIR", 1);
    x.pair_min_n = env_int("M2S_ENGINE_PAIR_MIN_N", 32);
    x.fuse_max_n = env_int("M2S_FUSE_MAX_N", 64);
    return x;
  }();
  return k;
}

// ---- optional per-launch timing (bench.py roofline leg) -----------------------------
namespace {
struct ProfileRing {   // debug probe (m2s_debug_profile*): one ring per process, guarded for concurrent launchers
  std::mutex mu;
  std::vector<cudaEvent_t> ev;
  std::vector<double> flops;
  std::vector<int32_t> tags, tags_read;
  int count = 0;
  int tag = 0;
  bool on = false;
};
std::atomic<long long> g_launches{0};
ProfileRing& ring() {
  static ProfileRing r;
  return r;
}
}  // namespace

int profile_enable(int on) {
  ProfileRing& r = ring();
  std::lock_guard<std::mutex> g(r.mu);
  r.on = on != 0;
  r.count = 0;
  r.flops.clear();
  r.tags.clear();
  return M2S_OK;
}

void profile_set_tag(int tag) { ring().tag = tag; }

long long launch_count(bool reset) { return reset ? g_launches.exchange(0) : g_launches.load(); }

// tags of the launches returned by the last profile_read
int profile_read_tags(int32_t* tags, int cap) {
  ProfileRing& r = ring();
  std::lock_guard<std::mutex> g(r.mu);
  const int n = static_cast<int>(r.tags_read.size()) < cap ? static_cast<int>(r.tags_read.size()) : cap;
  for (int i = 0; i < n; ++i) tags[i] = r.tags_read[i];
  return n;
}

int profile_read(float* ms, double* flops, int cap, int* n_out) {
  ProfileRing& r = ring();
  M2S_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(r.mu);
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
  r.tags_read.swap(r.tags);
  r.tags.clear();
  return M2S_OK;
}

int profile_before(cudaStream_t stream) {
  ProfileRing& pr = ring();
  if (!pr.on) return M2S_OK;   // (unsynchronised fast path: the probe is switched on / off between launches)
  std::lock_guard<std::mutex> g(pr.mu);
  while (static_cast<int>(pr.ev.size()) < 2 * (pr.count + 1)) {
    cudaEvent_t e;
    M2S_CUDA_OK(cudaEventCreate(&e));
    pr.ev.push_back(e);
  }
  M2S_CUDA_OK(cudaEventRecord(pr.ev[2 * pr.count], stream));
  return M2S_OK;
}

int profile_after(cudaStream_t stream, double flops, int kernels) {
  g_launches.fetch_add(kernels);
  ProfileRing& pr = ring();
  if (!pr.on) return M2S_OK;
  std::lock_guard<std::mutex> g(pr.mu);
  M2S_CUDA_OK(cudaEventRecord(pr.ev[2 * pr.count + 1], stream));
  pr.flops.push_back(flops);
  pr.tags.push_back(pr.tag);
  ++pr.count;
  return M2S_OK;
}

int sm_count() {   // of the current device (cached per device)
  static std::mutex mu;
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  if (dev < 0 || dev >= 64) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
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

static float host_round_fp16(float v) { return __half2float(__float2half_rn(v)); }

int pack_weights(const float* host_w, int taps, int n, int c_in, int mode, PackedWeights* out) {
  PackedWeights w;
  w.n = n; w.c_in = c_in; w.taps = taps;
  int esize = 4;
  if (mode == PACK_FP16) {
    if (c_in % 8) return fail(M2S_ERR_UNSUPPORTED, "fp16 operands need c_in %% 8 == 0 (got %d)", c_in);
    esize = 2;
    if (c_in <= 32) { w.half = 2; w.kblock = 32; w.row_bytes = 64; }
    else { w.half = 1; w.kblock = 64; w.row_bytes = 128; }
  }
  choose_n_tiling(n, &w.n_tile, &w.n_tiles);
  w.cblocks = (c_in + w.kblock - 1) / w.kblock;
  const size_t tap_bytes = static_cast<size_t>(w.n_tile) * w.row_bytes;
  const size_t packed_bytes = static_cast<size_t>(w.n_tiles) * w.cblocks * taps * tap_bytes;
  w.packed_floats = packed_bytes / 4;
  std::vector<uint8_t> packed(packed_bytes, 0);
  std::vector<float> plain(static_cast<size_t>(taps) * n * c_in);
  auto put = [&](uint8_t* dst, float v) {
    if (esize == 4) std::memcpy(dst, &v, 4);
    else { const __half h = __float2half_rn(v); std::memcpy(dst, &h, 2); }
  };
  for (int j = 0; j < taps; ++j)
    for (int o = 0; o < n; ++o)
      for (int c = 0; c < c_in; ++c) {
        float v = host_w[(static_cast<size_t>(j) * n + o) * c_in + c];
        if (mode == PACK_TF32) v = host_round_tf32(v);
        else if (mode == PACK_FP16) v = host_round_fp16(v);
        plain[(static_cast<size_t>(j) * n + o) * c_in + c] = v;
        const int nt = o / w.n_tile, r = o % w.n_tile, cb = c / w.kblock, cc = c % w.kblock;
        const size_t blk = (static_cast<size_t>(nt * w.cblocks + cb) * taps + j) * tap_bytes;
        // swizzle on (absolute = stage-relative, stages are 1 KB aligned) address bits: 128-byte rows XOR the
        // 16-byte chunk index with (row & 7); 64-byte rows XOR it with ((row >> 1) & 3)
        const int off = cc * esize;
        const int chunk = (off >> 4) ^ (w.row_bytes == 128 ? (r & 7) : ((r >> 1) & 3));
        put(&packed[blk + static_cast<size_t>(r) * w.row_bytes + chunk * 16 + (off & 15)], v);
      }
  std::vector<uint8_t> pair_host;
  if (n >= engine_knobs().pair_min_n && n % 16 == 0 && engine_knobs().pair) {
    // CTA-pair layout: [nt][cb][half][tap][nh rows][row_bytes], unswizzled (TMA applies the swizzle)
    const int tiles = (n + 255) / 256;
    w.n_tile_pair = (((n + tiles - 1) / tiles) + 15) / 16 * 16;
    w.n_tiles_pair = tiles;
    const int nh = w.n_tile_pair / 2;
    std::vector<uint8_t> pk(static_cast<size_t>(tiles) * w.cblocks * 2 * taps * nh * w.row_bytes, 0);
    for (int j = 0; j < taps; ++j)
      for (int o = 0; o < n; ++o)
        for (int c = 0; c < c_in; ++c) {
          const int nt = o / w.n_tile_pair, rr = o % w.n_tile_pair, half = rr / nh, r = rr % nh;
          const int cb = c / w.kblock, cc = c % w.kblock;
          const size_t row = ((static_cast<size_t>(nt * w.cblocks + cb) * 2 + half) * taps + j) * nh + r;
          put(&pk[row * w.row_bytes + static_cast<size_t>(cc) * esize], plain[(static_cast<size_t>(j) * n + o) * c_in + c]);
        }
    pair_host.swap(pk);
  }
  // device copies: on any failure everything allocated so far is released
  auto up = [&](float** dst, const void* src, size_t bytes) -> int {
    M2S_CUDA_OK(cudaMalloc(dst, bytes));
    M2S_CUDA_OK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return M2S_OK;
  };
  int st = M2S_OK;
  if (!pair_host.empty()) st = up(&w.dev_pair, pair_host.data(), pair_host.size());
  if (st == M2S_OK) st = up(&w.dev, packed.data(), packed.size());
  if (st == M2S_OK) st = up(&w.plain, plain.data(), plain.size() * sizeof(float));
  if (st != M2S_OK) {
    free_weights(&w);
    return st;
  }
  *out = w;
  return M2S_OK;
}

void free_weights(PackedWeights* w) {
  if (w->dev) cudaFree(w->dev);
  if (w->dev_pair) cudaFree(w->dev_pair);
  if (w->plain) cudaFree(w->plain);
  *w = PackedWeights{};
}

int conv_tcgen05(const ConvProblem& p, const PackedWeights& w, cudaStream_t stream) {
  if (p.taps < 1 || p.taps > M2S_MAX_TAPS) return fail(M2S_ERR_BAD_ARG, "taps=%d out of range", p.taps);
  if (p.c_in % 4 || p.a_ld % 4 || p.d_ld % 4 || p.n % 4)
    return fail(M2S_ERR_UNSUPPORTED, "c_in/a_ld/d_ld/n must be multiples of 4 (got %d/%d/%d/%d)", p.c_in, p.a_ld,
                p.d_ld, p.n);
  if ((p.a_half != 0) != (w.half != 0))
    return fail(M2S_ERR_BAD_ARG, "operand format mismatch: A is %s but the weights are packed as %s",
                p.a_half ? "fp16" : "fp32", w.half ? "fp16" : "tf32");
  if (p.a_half && (p.c_in % 8 || p.a_ld % 8))
    return fail(M2S_ERR_UNSUPPORTED, "fp16 operands need c_in/a_ld multiples of 8 (got %d/%d)", p.c_in, p.a_ld);
  if (!p.d && !p.d16) return fail(M2S_ERR_BAD_ARG, "no output pointer");
  if ((reinterpret_cast<uintptr_t>(p.a) & 15) || (reinterpret_cast<uintptr_t>(p.d) & 15) ||
      (reinterpret_cast<uintptr_t>(p.d16) & 7) || (reinterpret_cast<uintptr_t>(p.d16_lo) & 7) ||
      (reinterpret_cast<uintptr_t>(p.epi.res_hi) & 15) || (reinterpret_cast<uintptr_t>(p.epi.res_lo) & 15))
    return fail(M2S_ERR_BAD_ARG, "A and D must be 16-byte aligned");
  if (p.d16_lo && !p.d16) return fail(M2S_ERR_BAD_ARG, "d16_lo (lo plane) needs d16 (hi plane)");
  if (p.d16_lo && !lo_output_supported(choose_epilogue(p.epi)))
    return fail(M2S_ERR_UNSUPPORTED, "d16_lo: only the bias + leaky-ReLU program and the split-residual programs write a lo plane");
  if (choose_epilogue(p.epi) < 0 || (p.epi.res_hi && (p.n % 8 || p.epi.res_ld % 8)))
    return fail(M2S_ERR_UNSUPPORTED, "split-fp16 residual: needs res_hi and res_lo, res == NULL, a pre-activation "
                "residual with res_inv_slope >= 1, no activation or a leaky-ReLU with slope in (0, 1], n and res_ld multiples of 8");
  if (w.n != p.n || w.c_in != p.c_in || w.taps != p.taps)
    return fail(M2S_ERR_BAD_ARG, "packed weights do not match the problem");
  if (p.batch <= 0 || p.l_out <= 0) return M2S_OK;
  if (p.tap_ksteps && (!w.half || w.cblocks != 1 || engine_knobs().a_per_tap))
    return fail(M2S_ERR_UNSUPPORTED, "per-tap K windows need fp16 operands in a single K block");
  if (w.dev_pair && !p.tap_ksteps && engine_knobs().pair && !engine_knobs().a_per_tap && !engine_knobs().trace) {
    // CTA-pair (cta_group::2) kernel when the layer is wide, or narrow but bound by the MMA's SMEM operand fetch
    // rather than by HBM (measured model: ~75 B/clk/SM of operand fetch, ~23 B/clk/SM of HBM).
    const long long pair_tiles = static_cast<long long>(p.batch) * ((p.l_out + 255) / 256) * w.n_tiles_pair;
    bool use_pair = p.n >= 128;
    if (!use_pair) {
      const int kstep = w.half ? 16 : 8;
      const double mma_clk =
          static_cast<double>(p.taps) * ((p.c_in + kstep - 1) / kstep) * (4096.0 + 32.0 * w.n_tile) / 75.0;
      const double out_bytes = (p.d ? 4.0 : 0.0) + (p.d16 ? 2.0 : 0.0) + (p.d16_lo ? 2.0 : 0.0) + (p.epi.res || p.epi.res_hi ? 4.0 : 0.0) + (p.epi.accum ? 4.0 : 0.0);
      const double hbm_clk = 128.0 * (p.c_in * (w.half ? 2.0 : 4.0) + static_cast<double>(p.n) * out_bytes) / 23.0;
      use_pair = mma_clk > 1.5 * hbm_clk;
    }
    if (engine_knobs().pair == 2) use_pair = true;
    if (use_pair && pair_tiles >= sm_count() / 4) return conv_tcgen05_pair(p, w, stream);
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const EngineKnobs& knobs = engine_knobs();

  EngineParams prm{};
  prm.p = p;
  finalize_epilogue(&prm.p.epi, static_cast<long long>(p.l_out) + p.d_row_offset);
  prm.wpacked = w.dev;
  prm.n_tile = w.n_tile;
  prm.n_tiles = w.n_tiles;
  prm.cblocks = w.cblocks;
  prm.half = w.half;
  prm.kblock = w.kblock;
  prm.row_bytes = w.row_bytes;
  prm.kstep_elems = w.half ? 16 : 8;
  prm.desc_hi = make_desc_hi(w.row_bytes);
  prm.base_offset_mode = knobs.base_offset_mode;
  prm.dbg = knobs.dbg;
  prm.trace = knobs.trace;
  prm.trace_tiles = knobs.trace_tiles;
  prm.a_per_tap = knobs.a_per_tap;
  int smin = p.shift[0], smax = p.shift[0];
  for (int j = 1; j < p.taps; ++j) { smin = p.shift[j] < smin ? p.shift[j] : smin; smax = p.shift[j] > smax ? p.shift[j] : smax; }
  prm.shift_min = smin;
  const int halo = prm.a_per_tap ? 0 : smax - smin;
  prm.tap_ksteps = p.tap_ksteps;
  for (int j = 0; j < p.taps; ++j) prm.tap_kofs[j] = p.tap_ksteps ? p.kofs[j] : 0;

  // M sub-tiles per CTA: 2 halves the weight traffic per output row; use it when there is enough work.
  int msub = knobs.msub;
  if (msub != 1 && msub != 2 && msub != 4 && msub != 8) {
    // More 128-row sub-tiles per tile = less weight traffic and less halo re-load per output row, fewer barrier round
    // trips per output.  Narrow layers (N <= 32) get up to 8 sub-tiles (1024 rows): the 3x3 convs of the encoder's first
    // stages carry a halo of two image rows (262 rows at 128 x 128), which doubled the A traffic of a 256-row tile.
    // Keep TWO accumulator buffers (the epilogue of tile i overlaps the MMAs of tile i+1), enough tiles to fill the
    // machine twice, and two A stages within the SMEM budget.
    msub = 1;
    for (int cand = 2; cand <= 8; cand *= 2) {
      const long long tiles = static_cast<long long>(p.batch) * ((p.l_out + 128 * cand - 1) / (128 * cand)) * w.n_tiles;
      const size_t a_stage = static_cast<size_t>(128 * cand + halo + 16) * w.row_bytes;
      const size_t tap_b = static_cast<size_t>(w.n_tile) * w.row_bytes;
      const size_t b_stage = std::min<size_t>(static_cast<size_t>(p.taps), std::max<size_t>(1, 32768 / tap_b)) * tap_b;
      if (2 * cand * w.n_tile <= kTmemCols && tiles >= 2LL * sm_count() && 2 * a_stage <= 150 * 1024 &&
          2 * a_stage + 2 * b_stage <= 186 * 1024 &&
          (cand <= 2 || (halo >= 64 && w.n_tile <= 64)))
        msub = cand;
    }
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
  prm.a_stage_bytes = static_cast<uint32_t>(prm.a_nbox * prm.a_box_rows * w.row_bytes);
  prm.a_stage_bytes = (prm.a_stage_bytes + 1023u) & ~1023u;
  for (int j = 0; j < p.taps; ++j) prm.rel_shift[j] = p.shift[j] - smin;
  prm.b_tap_bytes = static_cast<uint32_t>(w.n_tile * w.row_bytes);
  // several taps share one weight stage when the per-tap block is small (amortises barrier round trips)
  int tg = static_cast<int>(32768u / prm.b_tap_bytes);
  if (tg < 1) tg = 1;
  if (tg > p.taps) tg = p.taps;
  if (knobs.a_per_tap) tg = 1;
  prm.tg = tg;
  prm.b_stage_bytes = static_cast<uint32_t>(tg) * prm.b_tap_bytes;
  const uint32_t b_stage_alloc = (prm.b_stage_bytes + 1023u) & ~1023u;
  // stage counts inside the budget: at least 2 A + 2 B
  const uint32_t bar_bytes = 1024 + kEpiSmemBytes;  // barriers + epilogue staging
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

  // instruction descriptor: D=f32, A=B=tf32 (format 2) or f16 (format 0), K-major both, N, M=128
  prm.idesc = (1u << 4) | (w.half ? 0u : ((2u << 7) | (2u << 10))) | (static_cast<uint32_t>(w.n_tile >> 3) << 17) |
              (static_cast<uint32_t>(128 >> 4) << 24);

  // tensor map over A: (c_in, a_rows, batch), box (32, a_box_rows, 1), 128B swizzle, OOB -> zero
  CUtensorMap tmap;
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
  CUresult cr = enc(&tmap, dt, 3, const_cast<float*>(p.a), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    w.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS)
    return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): c_in=%d rows=%d batch=%d ld=%d box_rows=%d",
                static_cast<int>(cr), p.c_in, p.a_rows, p.batch, p.a_ld, prm.a_box_rows);

  // B stage pitch must equal the copy size for the kernel's addressing: use the aligned pitch everywhere.
  prm.b_stage_bytes = b_pitch;
  // copy size = n_tile*128 which is already a multiple of 2048 for n_tile%16==0 -> equals pitch when n_tile%8==0
  if (b_pitch != static_cast<uint32_t>(tg) * prm.b_tap_bytes)
    return fail(M2S_ERR_UNSUPPORTED, "n_tile=%d gives a non-1024-aligned weight stage", w.n_tile);

  using KernelFn = void (*)(const CUtensorMap, const EngineParams);
  static const KernelFn kernels[EPI_COUNT] = {conv_engine_kernel<EPI_FULL>,  conv_engine_kernel<EPI_FULL_SILU>, conv_engine_kernel<EPI_BIAS>,
                                              conv_engine_kernel<EPI_LRELU>, conv_engine_kernel<EPI_SILU>,      conv_engine_kernel<EPI_RES>,
                                              conv_engine_kernel<EPI_RB>,    conv_engine_kernel<EPI_RB_ACC>,
                                              conv_engine_kernel<EPI_RB_S>,  conv_engine_kernel<EPI_RB_ACC_S>};
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    for (KernelFn k : kernels) M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  const int epi = choose_epilogue(p.epi);
  int grid = knobs.max_ctas > 0 ? knobs.max_ctas : sm_count();
  if (grid > prm.total_tiles) grid = prm.total_tiles;
  M2S_TRY(profile_before(stream));
  kernels[epi]<<<grid, kThreads, smem_bytes, stream>>>(tmap, prm);
  M2S_CUDA_OK(cudaGetLastError());
  M2S_TRY(profile_after(stream, 2.0 * p.batch * static_cast<double>(p.l_out) * p.n * p.c_in * p.taps));
  return M2S_OK;
}

}  // namespace m2s

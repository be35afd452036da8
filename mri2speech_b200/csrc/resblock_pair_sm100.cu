// Fused ResBlock1 pair for sm_100a:  y = epi2( conv2( lrelu( conv1(x) + b1 ) ) + b2 , residual ... )   in ONE kernel.
//
// Reference semantics (models.py:36-48): xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x, both convs causal.
// The two-launch path writes lrelu(c1(..)) to HBM (fp16) and reads it back; here that intermediate tile T never
// leaves the SM:
//
//   TMA (x tile with halo, weights)  ->  tcgen05.mma kind::f16  ->  acc1 (TMEM)
//   epilogue warps: acc1 -> +b1 -> leaky-ReLU -> fp16 -> T tile in SMEM, written directly in the K-major swizzled
//                   operand layout (rows before t = 0 forced to zero: conv2's causal padding sees zeros)
//   tcgen05.mma (A = T tile, row-shifted per tap)  ->  acc2 (TMEM)  ->  the engine's fused epilogue (bias, residual via
//   the inverse leaky-ReLU, MRF accumulate, scale, leaky-ReLU, length mask) -> fp32 and/or fp16 stores.
//
// A tile computes M1 = 256 rows of conv1 and keeps the M2 = M1 - (k2 - 1) rows of conv2 whose taps are all inside the
// tile; consecutive tiles advance by M2 (the (k2-1)-row overlap of conv1 is recomputed: <= 4 % for k = 11).
// HBM traffic per element of the pair: x (2 B, fp16 operand copy) + residual (4 B) + outputs (4 B fp32 and / or 2 B fp16),
// against 16 B for two launches; the MMAs of conv1 of tile i+1 overlap both epilogues of tile i (accumulators and the
// T tile are double-buffered whenever 8 * N <= 512 TMEM columns).
//
// Warp roles (576 threads, 1 CTA / SM, persistent): warp 0 TMA producer, warp 1 MMA issuer, warps 2-17 epilogues.
// Operands are fp16 only (the fp16 build); N = C_out <= 128 (one N tile), weights in the single-CTA packed layout.
#include "engine_device.cuh"
#include <mutex>

namespace m2s {

using namespace engine;

namespace {

struct PairParams {
  ConvProblem p2;          // epilogue / outputs / bias2 of conv2; p2.l_out, p2.batch, p2.n
  const float* bias1;      // [n]
  float slope1;            // leaky-ReLU between the two convs
  const void* w1;          // packed weights (single-CTA layout), conv1 / conv2
  const void* w2;
  int taps1, taps2;
  int rel_shift1[M2S_MAX_TAPS];  // rows relative to the first row of the x tile
  int rel_shift2[M2S_MAX_TAPS];  // rows relative to the first row of the T tile (0 .. k2-1)
  int c_in, n;                   // channels of x / of T and of the output (n == n_tile)
  int cblocks1, cblocks2;        // K blocks of conv1 (over c_in) and conv2 (over n)
  int kblock, row_bytes, kstep_elems;
  uint64_t desc_hi;
  uint32_t idesc;
  int m1, m2, msub;              // rows of conv1 per tile, valid rows of conv2 per tile, 128-row sub-tiles
  int halo2;                     // k2 - 1
  int x_row0;                    // first x row of a tile relative to q0:  -(k2-1) + min shift of conv1
  int tiles_per_batch, total_tiles;
  int a_box_rows, a_nbox;
  uint32_t a_stage_bytes, b_stage_bytes, b_tap_bytes, t_kb_bytes, t_buf_bytes;
  int na, nb, tg1, tg2;
  int nbuf;                      // 1 or 2: buffers of acc1 / acc2 / T;  lookahead = nbuf - 1
  int dbg;
  unsigned long long* trace;     // debug timeline of CTA 0 (tools/trace_pair.py): trace[iteration * 8 + slot], null = off
  int trace_tiles;
};

constexpr int kMaxBuf = 2;

// slots: 0 epilogue got acc1, 1 epilogue 1 done, 2 epilogue got acc2, 3 epilogue 2 done (epilogue warp 0);
//        4 conv1 issued, 5 MMA warp got t_full, 6 conv2 issued; 7 producer issued the x tile
__device__ __forceinline__ void pair_stamp(const PairParams& prm, int it, int slot) {
  if (prm.trace && blockIdx.x == 0 && it < prm.trace_tiles) prm.trace[it * 8 + slot] = clock64();
}

template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
resblock_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ PairParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t t_base = a_base + prm.na * prm.a_stage_bytes;
  const uint32_t b_base = t_base + prm.nbuf * prm.t_buf_bytes;
  const uint32_t bar_base = b_base + prm.nb * prm.b_stage_bytes;
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
  const uint32_t bias1_smem = bar_base + 512u;            // n floats (<= 128)
  const uint32_t stage_base = bar_base + 1024u;           // 16 epilogue warps x 2 KB transpose staging

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvProblem& p = prm.p2;
  const int nbuf = prm.nbuf;
  const int la = nbuf - 1;  // conv1 of tile i+la is issued before conv2 of tile i

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < prm.na; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < prm.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < nbuf; ++s) {
      mbar_init(acc1_full(s), 1); mbar_init(acc1_empty(s), kEpiWarps);
      mbar_init(acc2_full(s), 1); mbar_init(acc2_empty(s), kEpiWarps);
      mbar_init(t_full(s), kEpiWarps); mbar_init(t_empty(s), 1);
    }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  for (int i = threadIdx.x; i < prm.n; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias1_smem + 4u * i), "f"(__ldg(prm.bias1 + i)) : "memory");
  const uint32_t bias_smem = stage_bias(p, stage_base);
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n = prm.n;
  const int msub = prm.msub;
  const uint32_t row_bytes = prm.row_bytes;
  const int my_tiles = (prm.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const uint32_t acc_cols = static_cast<uint32_t>(msub * n);  // columns of one accumulator buffer
  auto acc1_addr = [&](int buf) { return tmem_base + static_cast<uint32_t>(buf) * acc_cols; };
  auto acc2_addr = [&](int buf) { return tmem_base + static_cast<uint32_t>(nbuf + buf) * acc_cols; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t a_bytes = prm.a_nbox * prm.a_box_rows * row_bytes;
    auto load_weights = [&](const void* wbase, int cb, int taps, int tg) {
      for (int tap0 = 0; tap0 < taps; tap0 += tg) {
        const int cnt = min(tg, taps - tap0);
        mbar_wait(b_empty(sb), pb ^ 1);
        if (elect_one()) {
          const uint32_t bytes = cnt * prm.b_tap_bytes;
          mbar_expect_tx(b_full(sb), bytes);
          const uint8_t* src = static_cast<const uint8_t*>(wbase) + static_cast<size_t>(cb * taps + tap0) * prm.b_tap_bytes;
          bulk_load(b_base + sb * prm.b_stage_bytes, src, bytes, b_full(sb));
        }
        __syncwarp();
        if (++sb == prm.nb) { sb = 0; pb ^= 1; }
      }
    };
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        const int tile = blockIdx.x + s * gridDim.x;
        const int b = tile / prm.tiles_per_batch;
        const int q0 = (tile - b * prm.tiles_per_batch) * prm.m2;
        for (int cb = 0; cb < prm.cblocks1; ++cb) {
          mbar_wait(a_empty(sa), pa ^ 1);
          if (elect_one()) {
            mbar_expect_tx(a_full(sa), a_bytes);
            for (int bx = 0; bx < prm.a_nbox; ++bx)
              tma_load_3d(a_base + sa * prm.a_stage_bytes + bx * prm.a_box_rows * row_bytes, &tmap_x, a_full(sa),
                          cb * prm.kblock, q0 + prm.x_row0 + bx * prm.a_box_rows, b);
          }
          __syncwarp();
          if (lane == 0 && cb == 0) pair_stamp(prm, s, 7);
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
          load_weights(prm.w1, cb, prm.taps1, prm.tg1);
        }
      }
      if (s >= la) {
        for (int cb = 0; cb < prm.cblocks2; ++cb) load_weights(prm.w2, cb, prm.taps2, prm.tg2);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint64_t desc_hi = prm.desc_hi;
    const int ksteps_full = prm.row_bytes >> 5;
    // one conv over an A tile already in SMEM: `a_tile(cb)` gives the SMEM address of K block cb
    // (`done1` / `done2`: barriers committed by the issuing lane after the last MMA of the conv; 0 = none)
    auto conv = [&](uint32_t tmem_acc, int cblocks, int c_total, int taps, int tg, const int* rel_shift, bool a_from_tma,
                    uint32_t t_tile, uint32_t done1, uint32_t done2) {
      for (int cb = 0; cb < cblocks; ++cb) {
        const int rem = c_total - cb * prm.kblock;
        const int ksteps = rem >= prm.kblock ? ksteps_full : (rem + prm.kstep_elems - 1) / prm.kstep_elems;
        uint32_t a_tile;
        if (a_from_tma) {
          mbar_wait(a_full(sa), pa);
          a_tile = a_base + sa * prm.a_stage_bytes;
        } else {
          a_tile = t_tile + cb * prm.t_kb_bytes;
        }
        for (int tap0 = 0; tap0 < taps; tap0 += tg) {
          const int cnt = min(tg, taps - tap0);
          mbar_wait(b_full(sb), pb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_tile = b_base + sb * prm.b_stage_bytes;
            for (int t = 0; t < cnt && !(prm.dbg & 8); ++t) {
              const uint64_t db = desc_hi | (((b_tile + t * prm.b_tap_bytes) & 0x3FFFF) >> 4);
              const uint64_t da = desc_hi | (((a_tile + rel_shift[tap0 + t] * row_bytes) & 0x3FFFF) >> 4);
              const uint32_t first = (cb | tap0 | t) ? 1u : 0u;
              for (int sub = 0; sub < msub; ++sub)
                mma_f16_k4(tmem_acc + sub * n, da + sub * ((128 * row_bytes) >> 4), db, prm.idesc, first, ksteps);
            }
            tc_commit(b_empty(sb));
            const bool last_group = tap0 + cnt >= taps;
            if (a_from_tma && last_group) tc_commit(a_empty(sa));
            if (last_group && cb == cblocks - 1) {
              tc_commit(done1);
              if (done2) tc_commit(done2);
            }
          }
          __syncwarp();
          if (++sb == prm.nb) { sb = 0; pb ^= 1; }
        }
        if (a_from_tma) {
          if (++sa == prm.na) { sa = 0; pa ^= 1; }
        }
      }
    };
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        const int buf = s % nbuf;
        mbar_wait(acc1_empty(buf), ((s / nbuf) & 1) ^ 1);
        tc_fence_after();
        conv(acc1_addr(buf), prm.cblocks1, prm.c_in, prm.taps1, prm.tg1, prm.rel_shift1, true, 0u, acc1_full(buf), 0u);
        if (lane == 0) pair_stamp(prm, s, 4);
      }
      if (s >= la) {
        const int i = s - la;
        const int buf = i % nbuf;
        mbar_wait(t_full(buf), (i / nbuf) & 1);
        if (lane == 0) pair_stamp(prm, i, 5);
        mbar_wait(acc2_empty(buf), ((i / nbuf) & 1) ^ 1);
        tc_fence_after();
        conv(acc2_addr(buf), prm.cblocks2, n, prm.taps2, prm.tg2, prm.rel_shift2, false, t_base + buf * prm.t_buf_bytes,
             acc2_full(buf), t_empty(buf));
        if (lane == 0) pair_stamp(prm, i, 6);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const EpiWarp epw = make_epi_warp(p.epi, stage_base, ew, warp, lane, prm.dbg, bias_smem);
    const int quad = epw.quad, grp = epw.grp;
    const int nchunks = (n + kEpiUnitCols - 1) / kEpiUnitCols;
    const int kb_shift = prm.kblock == 64 ? 6 : 5;
    for (int s = 0; s < my_tiles + la; ++s) {
      if (s < my_tiles) {
        // ---- epilogue 1: acc1 -> T tile (fp16, swizzled K-major operand layout) ----
        const int tile = blockIdx.x + s * gridDim.x;
        const int b = tile / prm.tiles_per_batch;
        const int q0 = (tile - b * prm.tiles_per_batch) * prm.m2;
        const int buf = s % nbuf;
        mbar_wait(acc1_full(buf), (s / nbuf) & 1);
        mbar_wait(t_empty(buf), ((s / nbuf) & 1) ^ 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) pair_stamp(prm, s, 0);
        const uint32_t tacc = acc1_addr(buf) + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t tt = t_base + buf * prm.t_buf_bytes;
        int sub = 0, ci = grp;   // this warp's (sub-tile, 16-column chunk) units, stepped without divisions
        while (ci >= nchunks) { ci -= nchunks; ++sub; }
        for (; sub < msub; ) {
          const int c0 = ci * kEpiUnitCols;
          const int row = sub * 128 + quad * 32 + lane;   // row of the T tile; time = q0 - halo2 + row
          uint32_t r[16];
          tmem_ld16(tacc + sub * n + c0, r);
          const bool zero_row = q0 - prm.halo2 + row < 0;
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(bias1_smem + 4u * (c0 + 4 * j)));
            float v0 = __uint_as_float(r[4 * j]) + b4.x, v1 = __uint_as_float(r[4 * j + 1]) + b4.y;
            float v2 = __uint_as_float(r[4 * j + 2]) + b4.z, v3 = __uint_as_float(r[4 * j + 3]) + b4.w;
            v0 = fmaxf(v0, v0 * prm.slope1); v1 = fmaxf(v1, v1 * prm.slope1);
            v2 = fmaxf(v2, v2 * prm.slope1); v3 = fmaxf(v3, v3 * prm.slope1);
            if (zero_row) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[2 * j]) : "f"(v1), "f"(v0));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[2 * j + 1]) : "f"(v3), "f"(v2));
          }
          // 16 halves = 32 bytes = two 16-byte chunks of this row inside K block kb
          const int kb = c0 >> kb_shift;
          const int chunk0 = ((c0 & (prm.kblock - 1)) * 2) >> 4;
          const uint32_t swz = row_bytes == 128 ? (row & 7) : ((row >> 1) & 3);
          const uint32_t rbase = tt + kb * prm.t_kb_bytes + row * row_bytes;
          if (c0 < n) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((chunk0 + m) ^ swz) << 4)),
                           "r"(pk[4 * m]), "r"(pk[4 * m + 1]), "r"(pk[4 * m + 2]), "r"(pk[4 * m + 3])
                           : "memory");
          }
          ci += kEpiGroups;
          while (ci >= nchunks) { ci -= nchunks; ++sub; }
        }
        fence_proxy_async();   // generic-proxy writes of T -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_full(buf));
          mbar_arrive(acc1_empty(buf));
          if (ew == 0) pair_stamp(prm, s, 1);
        }
      }
      if (s >= la) {
        // ---- epilogue 2: acc2 -> fused epilogue -> global (rows [q0, q0 + m2) of the tile) ----
        const int i = s - la;
        const int tile = blockIdx.x + i * gridDim.x;
        const int b = tile / prm.tiles_per_batch;
        const int q0 = (tile - b * prm.tiles_per_batch) * prm.m2;
        const int buf = i % nbuf;
        mbar_wait(acc2_full(buf), (i / nbuf) & 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) pair_stamp(prm, i, 2);
        const uint32_t tacc = acc2_addr(buf) + (static_cast<uint32_t>(quad * 32) << 16);
        epilogue_tile<kEpi>(p, epw, tacc, b, q0, 0, msub, n, min(p.l_out, q0 + prm.m2));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty(buf));
        if (ew == 0 && lane == 0) pair_stamp(prm, i, 3);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn3() {
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

// Can the fused pair kernel run this pair?  (fp16 operands, one N tile, square channel count.)
bool resblock_pair_supported(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2,
                             const PackedWeights& w2) {
  if (!p1.a_half || !w1.half || !w2.half || w1.half != w2.half) return false;
  if (p1.n != p2.n || p2.c_in != p1.n || p1.n > engine_knobs().fuse_max_n || p1.n % 32 || w1.n_tiles != 1 ||
      w2.n_tiles != 1)
    return false;
  if (w1.n_tile != p1.n || w2.n_tile != p1.n) return false;
  if (p1.batch != p2.batch || p1.l_out != p2.l_out || p2.d_row_offset != 0) return false;
  if (p2.epi.mask_mode == M2S_MASK_PITCH) return false;
  for (int j = 0; j < p2.taps; ++j)
    if (p2.shift[j] != -(p2.taps - 1 - j)) return false;  // conv2: causal, dilation 1
  int smin = 0;
  for (int j = 0; j < p1.taps; ++j) {
    if (p1.shift[j] > 0) return false;
    smin = p1.shift[j] < smin ? p1.shift[j] : smin;
  }
  return -smin <= 64 && p2.taps <= 16;
}

// p1: conv1 (a = x fp16, shifts, bias = b1, act = leaky-ReLU slope); its outputs are ignored (T stays on chip).
// p2: conv2's taps, bias, epilogue and outputs; its `a` is ignored.
int resblock_pair_fused(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2,
                        cudaStream_t stream) {
  if (!resblock_pair_supported(p1, w1, p2, w2)) return fail(M2S_ERR_UNSUPPORTED, "pair not supported by the fused kernel");
  if (p2.batch <= 0 || p2.l_out <= 0) return M2S_OK;
  EncodeTiledFn enc = encode_fn3();
  if (!enc) return fail(M2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  PairParams prm{};
  prm.p2 = p2;
  finalize_epilogue(&prm.p2.epi, static_cast<long long>(p2.l_out) + p2.d_row_offset);
  prm.bias1 = p1.epi.bias;
  prm.slope1 = p1.epi.act == M2S_ACT_LRELU ? p1.epi.act_slope : 1.f;
  prm.w1 = w1.dev;
  prm.w2 = w2.dev;
  prm.taps1 = p1.taps;
  prm.taps2 = p2.taps;
  prm.c_in = p1.c_in;
  prm.n = p1.n;
  prm.kblock = w1.kblock;
  prm.row_bytes = w1.row_bytes;
  prm.kstep_elems = 16;
  prm.desc_hi = make_desc_hi(w1.row_bytes);
  prm.cblocks1 = w1.cblocks;
  prm.cblocks2 = w2.cblocks;
  // 128-row sub-tiles per tile: 4 (M1 = 512) when all four accumulator buffers still fit the 512 TMEM columns
  prm.msub = (16 * prm.n <= kTmemCols) ? 4 : 2;
  if (engine_knobs().msub == 1 || engine_knobs().msub == 2) prm.msub = 2;
  prm.m1 = 128 * prm.msub;
  prm.halo2 = p2.taps - 1;
  prm.m2 = prm.m1 - prm.halo2;
  int smin = 0;
  for (int j = 0; j < p1.taps; ++j) smin = p1.shift[j] < smin ? p1.shift[j] : smin;
  for (int j = 0; j < p1.taps; ++j) prm.rel_shift1[j] = p1.shift[j] - smin;
  for (int j = 0; j < p2.taps; ++j) prm.rel_shift2[j] = j;
  prm.x_row0 = -prm.halo2 + smin;
  prm.tiles_per_batch = (p2.l_out + prm.m2 - 1) / prm.m2;
  prm.total_tiles = p2.batch * prm.tiles_per_batch;
  prm.nbuf = (4 * prm.msub * prm.n <= kTmemCols) ? 2 : 1;
  prm.dbg = engine_knobs().dbg;
  prm.trace = engine_knobs().trace;
  prm.trace_tiles = engine_knobs().trace_tiles;

  const int a_rows_needed = prm.m1 - smin;
  prm.a_nbox = (a_rows_needed + 255) / 256;
  prm.a_box_rows = (((a_rows_needed + prm.a_nbox - 1) / prm.a_nbox) + 7) / 8 * 8;
  prm.a_stage_bytes = (static_cast<uint32_t>(prm.a_nbox * prm.a_box_rows * prm.row_bytes) + 1023u) & ~1023u;
  prm.t_kb_bytes = (static_cast<uint32_t>((prm.m1 + 16) * prm.row_bytes) + 1023u) & ~1023u;
  prm.t_buf_bytes = prm.cblocks2 * prm.t_kb_bytes;
  prm.b_tap_bytes = static_cast<uint32_t>(prm.n * prm.row_bytes);
  // SMEM plan: A stages, T buffers, weight stages (tap groups), staging
  const uint32_t fixed = 1024u + kEpiSmemBytes + 1024u;
  const uint32_t budget = 225u * 1024u;
  int na = 2;
  uint32_t used = fixed + na * prm.a_stage_bytes + prm.nbuf * prm.t_buf_bytes;
  if (used + 2 * prm.b_tap_bytes > budget && prm.nbuf == 2) {
    prm.nbuf = 1;
    used = fixed + na * prm.a_stage_bytes + prm.t_buf_bytes;
  }
  if (used + 2 * prm.b_tap_bytes > budget) return fail(M2S_ERR_UNSUPPORTED, "fused pair: tile does not fit SMEM");
  const int maxtaps = p1.taps > p2.taps ? p1.taps : p2.taps;
  int tg = static_cast<int>(32768u / prm.b_tap_bytes);
  if (tg < 1) tg = 1;
  if (tg > maxtaps) tg = maxtaps;
  while (tg > 1 && used + 2u * tg * prm.b_tap_bytes > budget) --tg;
  prm.b_stage_bytes = static_cast<uint32_t>(tg) * prm.b_tap_bytes;
  if (prm.b_stage_bytes & 1023u) return fail(M2S_ERR_UNSUPPORTED, "fused pair: weight stage not 1 KB aligned");
  prm.tg1 = tg < p1.taps ? tg : p1.taps;
  prm.tg2 = tg < p2.taps ? tg : p2.taps;
  int nb = 2;
  while (nb < 4 && used + (nb + 1) * prm.b_stage_bytes <= budget) ++nb;
  used += nb * prm.b_stage_bytes;
  while (na < 3 && used + prm.a_stage_bytes <= budget) { ++na; used += prm.a_stage_bytes; }
  while (nb < kMaxStagesB && used + prm.b_stage_bytes <= budget) { ++nb; used += prm.b_stage_bytes; }
  prm.na = na;
  prm.nb = nb;
  const uint32_t smem_bytes = used + 1024u;  // alignment slack
  prm.idesc = (1u << 4) | (static_cast<uint32_t>(prm.n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);

  CUtensorMap tmap;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p1.c_in), static_cast<cuuint64_t>(p1.a_rows),
                        static_cast<cuuint64_t>(p1.batch)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p1.a_ld) * 2ull,
                           static_cast<cuuint64_t>(p1.a_batch_rows) * static_cast<cuuint64_t>(p1.a_ld) * 2ull};
  if (p1.batch == 1) gstride[1] = gstride[0] * static_cast<cuuint64_t>(p1.a_rows > 0 ? p1.a_rows : 1);
  cuuint32_t box[3] = {static_cast<cuuint32_t>(prm.kblock), static_cast<cuuint32_t>(prm.a_box_rows), 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<float*>(p1.a), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    prm.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(M2S_ERR_CUDA, "fused pair: x tensor map failed (%d)", static_cast<int>(cr));

  using KernelFn = void (*)(const CUtensorMap, const PairParams);
  static const KernelFn kernels[EPI_COUNT] = {resblock_pair_kernel<EPI_FULL>,  resblock_pair_kernel<EPI_FULL_SILU>,
                                              resblock_pair_kernel<EPI_BIAS>,  resblock_pair_kernel<EPI_LRELU>,
                                              resblock_pair_kernel<EPI_SILU>,  resblock_pair_kernel<EPI_RES>,
                                              resblock_pair_kernel<EPI_RB>,    resblock_pair_kernel<EPI_RB_ACC>,
                                              resblock_pair_kernel<EPI_RB_S>,  resblock_pair_kernel<EPI_RB_ACC_S>};
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    for (KernelFn k : kernels) M2S_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return M2S_OK;
  }));
  const int epi = choose_epilogue(p2.epi);
  if (epi < 0 || (p2.d16_lo && (!p2.d16 || !lo_output_supported(epi))) || (p2.epi.res_hi && (p2.epi.res_ld % 8 || (reinterpret_cast<uintptr_t>(p2.epi.res_hi) & 15) ||
                                                            (reinterpret_cast<uintptr_t>(p2.epi.res_lo) & 15))))
    return fail(M2S_ERR_UNSUPPORTED, "fused pair: unsupported split-fp16 residual epilogue");
  int grid = engine_knobs().max_ctas > 0 ? engine_knobs().max_ctas : sm_count();
  if (grid > prm.total_tiles) grid = prm.total_tiles;
  M2S_TRY(profile_before(stream));
  kernels[epi]<<<grid, kThreads, smem_bytes, stream>>>(tmap, prm);
  M2S_CUDA_OK(cudaGetLastError());
  const double rows = static_cast<double>(p2.batch) * p2.l_out;
  M2S_TRY(profile_after(stream, 2.0 * rows * p1.n * (static_cast<double>(p1.c_in) * p1.taps + static_cast<double>(p1.n) * p2.taps)));
  return M2S_OK;
}

}  // namespace m2s

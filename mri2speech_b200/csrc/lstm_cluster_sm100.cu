// BiLSTM recurrence on thread-block clusters for sm_100a (tf32 / fp16 builds).
//
// Reference semantics: torch.nn.LSTM(208, 640, 1, batch_first, bidirectional) as called by BiLSTMSumMerge.forward
// (mri2speech_code/mri_acoustic_model.py:57-71); gate order i,f,g,o, h0 = c0 = 0.  The input projection is one
// tensor-core GEMM (acoustic.cu: rnn_head); this kernel is the sequential part:
//   z = gin[t] + W_hh h_{t-1};  c_t = s(z_f) c_{t-1} + s(z_i) tanh(z_g);  h_t = s(z_o) tanh(c_t)
//
// One 16-CTA cluster owns one (direction, group of up to 8 NT utterances), NT = 1, 2 or 3 n8 tiles; clusters never talk to
// each other, so there is no grid-wide barrier (the first version of this path, lstm_sm100.cu, paid ~6 us per step for one
// through L2).  Inside a cluster:
//   * CTA r keeps the W_hh rows of hidden units [40 r, 40 r + 40) -- 160 gate rows x 640, fp16 -- RESIDENT for the
//     whole sequence: 32 (30 for NT = 3) of the 40 K-steps as mma.sync A fragments in SMEM (160 KB, fragment order,
//     conflict-free LDS.128), the rest in registers (SMEM is what limits: W + two h buffers must fit 227 KB);
//   * a step is 10 warps x (16 rows x 8 NT utterances x 640) on mma.sync.m16n8k16 (fp16 in, fp32 accumulate, several
//     independent accumulator sets per tile so the 40 dependent MMAs become short chains);
//   * a warp's tile holds the four gates of 4 hidden units; lanes g and g+4 swap half of their fragments with ONE
//     round of warp shuffles so that every lane ends up with all four gates of (1 unit, NT utterances): the cell state
//     lives in NT registers per lane for the whole sequence;
//   * h_t (fp16, the next step's B operand) goes from a staging tile (40 units x 8 NT utterances) to the h buffer of all
//     16 CTAs with cp.async.bulk shared::cta -> shared::cluster; the copies complete_tx on the DESTINATION's mbarrier,
//     which is the only synchronisation of a step (no cluster barrier inside the loop; two h buffers / two staging
//     tiles make the one-step skew between CTAs safe).  The unrounded fp32 h_t goes to HBM once, for the head GEMM.
// h_{t-1} and W_hh are rounded to fp16 (10-bit mantissa, like the tf32 rounding of lstm_recurrence_mma_kernel; |h| < 1
// and |W_hh| << 1, so fp16's range is not a concern and values below 6e-5 lose at most 3e-8 absolute).
// Measured on B200 (tools/lstm_time.py): 2.1 us per step for NT = 1, +0.95 us per further tile (the legacy mma.sync
// pipe: ~15 cycles per m16n8k16 and SM sub-partition); configs[2]'s 64 ragged clips = 6 clusters of NT = 3: 2.4 ms
// against 11.2 ms for the grid-barrier kernel; one 150-frame clip: 0.31 ms.
#include "m2s_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace m2s {

namespace {

constexpr int kHidden = 640;
constexpr int kCl = 16;                        // CTAs per cluster
constexpr int kClUnits = kHidden / kCl;        // 40 hidden units per CTA = 10 warps x 4
constexpr int kClWarps = 10;
constexpr int kClThreads = 32 * kClWarps;
constexpr int kKSteps = kHidden / 16;          // 40 k16 steps
constexpr int kGroups = kHidden / 8;           // h is stored as 80 groups of 8 hidden units
constexpr int kMaxNT = 3;                      // n8 tiles (8 utterances each) per cluster: 1, 2 or 3

// Per-NT layout.  A group of the h buffer holds 8 NT utterances x 16 B + 32 B of padding (the stride in 16-byte units
// is 2 mod 8: the B-fragment LDS.128 of a quarter warp -- two g, four t4 -- then hits eight different bank groups).
// kreg = k-steps whose A fragments live in registers instead of SMEM: W + two h buffers + two staging tiles <= 227 KB.
template <int NT>
struct ClusterLayout {
  static constexpr int kreg = NT <= 2 ? 8 : 10;
  static constexpr int group_bytes = NT * 128 + 32;
  static constexpr int chunk_bytes = (kClUnits / 8) * group_bytes;   // the 40 units one CTA produces
  static constexpr int hbuf_bytes = kGroups * group_bytes;
  static constexpr int w_bytes = (kKSteps - kreg) * kClWarps * 32 * 16;
  static constexpr int smem_bytes = w_bytes + 2 * hbuf_bytes + 2 * chunk_bytes + 64;
};
static_assert(ClusterLayout<kMaxNT>::smem_bytes <= 227 * 1024, "cluster LSTM: SMEM plan does not fit");

struct ClusterLstmParams {
  const float* gin;      // (B, T, 2*4*H): [dir][gate][unit] pre-activations incl. both biases
  const float* w_hh[2];  // (4H, H) per direction, fp32
  const int32_t* lens;   // device, may be null
  float* hcat;           // (B, T, 2H)
  int batch, frames;
  int utt_per_group;     // utterances [grp * utt_per_group, (grp + 1) * utt_per_group) belong to cluster group grp (<= 8 NT)
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bar_try_wait(uint32_t bar, uint32_t parity) {
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
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  if (bar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!bar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("m2s lstm cluster: mbarrier timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// SMEM of this CTA -> SMEM of a CTA of the cluster; completes `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster_addr, uint32_t src_local, uint32_t bytes,
                                                 uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_local), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void lds128(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ex2.approx / rcp.approx forms (absolute error ~1e-7, three orders below the fp16 rounding of h): the accurate expf /
// tanhf of the fp32 kernel cost ~300 issue slots per lane and step here, a fifth of the step
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// K order.  An MMA's K index is free as long as A and B agree.  The h buffer keeps, per group G of 8 hidden units
// (units 8 G .. 8 G + 7) and utterance, 16 contiguous bytes; lane (g, t4) of k-block kb (two k16 steps) reads the 16 bytes of
// group 4 kb + t4 with ONE LDS.128 and uses halves (0,1) / (2,3) as b0 / b1 of the first step and (4,5) / (6,7) for
// the second; the A fragments are built with the same assignment.
template <int NT>
__global__ void __launch_bounds__(kClThreads, 1) lstm_cluster_kernel(const ClusterLstmParams prm) {
  using L = ClusterLayout<NT>;
  constexpr int kKReg = L::kreg;
  constexpr int kGroupBytes = L::group_bytes, kChunkBytes = L::chunk_bytes, kHBufBytes = L::hbuf_bytes;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t s_w = smem_addr(smem_raw);
  const uint32_t s_h = s_w + L::w_bytes;
  const uint32_t s_stage = s_h + 2 * kHBufBytes;
  const uint32_t s_bar = s_stage + 2 * kChunkBytes;   // two mbarriers: "h buffer i is complete"

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t rank = cluster_rank();
  const int cid = blockIdx.x / kCl;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * prm.utt_per_group;
  const int H = kHidden, H2 = 2 * kHidden, G8 = 8 * kHidden;

  // ---- W_hh slice -> A fragments (fp16): tile row rho = gate * 4 + unit_local, unit = 40 rank + 4 warp + unit_local ----
  uint32_t wreg[kKReg][4];
  {
    const int ul = g & 3;
    const int unit = kClUnits * static_cast<int>(rank) + 4 * warp + ul;
    const float* w = prm.w_hh[dir];
    const float* wlo = w + static_cast<size_t>((g >> 2) * H + unit) * H + 8 * t4;        // tile row g
    const float* whi = w + static_cast<size_t>((2 + (g >> 2)) * H + unit) * H + 8 * t4;  // tile row g + 8
#pragma unroll
    for (int kb = 0; kb < kKSteps / 2; ++kb) {
      const float4 lo0 = __ldg(reinterpret_cast<const float4*>(wlo + 32 * kb));
      const float4 lo1 = __ldg(reinterpret_cast<const float4*>(wlo + 32 * kb + 4));
      const float4 hi0 = __ldg(reinterpret_cast<const float4*>(whi + 32 * kb));
      const float4 hi1 = __ldg(reinterpret_cast<const float4*>(whi + 32 * kb + 4));
      const uint32_t f0[4] = {pack_f16x2(lo0.x, lo0.y), pack_f16x2(hi0.x, hi0.y), pack_f16x2(lo0.z, lo0.w), pack_f16x2(hi0.z, hi0.w)};
      const uint32_t f1[4] = {pack_f16x2(lo1.x, lo1.y), pack_f16x2(hi1.x, hi1.y), pack_f16x2(lo1.z, lo1.w), pack_f16x2(hi1.z, hi1.w)};
      if (2 * kb < kKReg) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { wreg[(2 * kb) % kKReg][i] = f0[i]; wreg[(2 * kb + 1) % kKReg][i] = f1[i]; }
      } else {
        sts128(s_w + static_cast<uint32_t>((((2 * kb - kKReg) * kClWarps + warp) * 32 + lane) * 16), f0);
        sts128(s_w + static_cast<uint32_t>((((2 * kb + 1 - kKReg) * kClWarps + warp) * 32 + lane) * 16), f1);
      }
    }
  }
  // h_{-1} = 0
  for (int i = tid; i < kHBufBytes / 16; i += kClThreads) {
    const uint32_t z[4] = {0u, 0u, 0u, 0u};
    sts128(s_h + 16u * i, z);
  }
  if (tid == 0) {
    bar_init(s_bar, 1);
    bar_init(s_bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- what this lane owns after the gate exchange: hidden unit `unit` for NT utterance slots.  Tiles are taken in
  // pairs (2q, 2q + 1): the low lane (g < 4) keeps both columns of tile 2q, its partner lane ^ 16 those of tile 2q + 1;
  // of an unpaired last tile the low lane keeps column 2 t4, the high lane column 2 t4 + 1 ----
  const bool high = g >= 4;
  const int ul = g & 3;
  const int uc = 4 * warp + ul;                                       // unit inside the CTA
  const int unit = kClUnits * static_cast<int>(rank) + uc;
  int slot[NT], len[NT];
  size_t row0[NT];   // b * frames
#pragma unroll
  for (int p = 0; p < NT; ++p) {
    if (p < (NT & ~1)) slot[p] = 16 * (p >> 1) + (high ? 8 : 0) + 2 * t4 + (p & 1);
    else slot[p] = 8 * (NT - 1) + 2 * t4 + (high ? 1 : 0);
    const int b = b0 + slot[p];
    const bool present = slot[p] < prm.utt_per_group && b < prm.batch;
    len[p] = present ? (prm.lens ? prm.lens[b] : prm.frames) : 0;
    len[p] = len[p] < 0 ? 0 : (len[p] > prm.frames ? prm.frames : len[p]);
    row0[p] = static_cast<size_t>(present ? b : 0) * prm.frames;
  }
  int steps = 0;   // uniform over the cluster: the longest utterance of the group
  for (int i = 0; i < prm.utt_per_group; ++i) {
    const int b = b0 + i;
    int l = b < prm.batch ? (prm.lens ? prm.lens[b] : prm.frames) : 0;
    l = l > prm.frames ? prm.frames : l;
    steps = l > steps ? l : steps;
  }
  float c[NT];
#pragma unroll
  for (int p = 0; p < NT; ++p) c[p] = 0.f;
  const uint32_t stage_off = static_cast<uint32_t>((uc >> 3) * kGroupBytes + (uc & 7) * 2);

  cluster_sync_all();   // every CTA's mbarriers are initialised and its h buffer zeroed before anybody sends

  for (int s = 0; s < steps; ++s) {
    const int cur = s & 1;
    // input projections of this step: requested now, used after the MMAs
    float zin[NT][4];
    bool active[NT];
    int tt[NT];
#pragma unroll
    for (int p = 0; p < NT; ++p) {
      active[p] = s < len[p];
      tt[p] = dir == 0 ? s : len[p] - 1 - s;
#pragma unroll
      for (int q = 0; q < 4; ++q) zin[p][q] = 0.f;
      if (active[p]) {
        const float* gp = prm.gin + (row0[p] + tt[p]) * G8 + dir * 4 * H + unit;
#pragma unroll
        for (int q = 0; q < 4; ++q) zin[p][q] = __ldg(gp + q * H);
      }
    }
    if (s > 0) bar_wait(s_bar + 8u * cur, static_cast<uint32_t>(((s - 1) >> 1) & 1));

    // ---- (16 gate rows of this warp) x (8 NT utterances) x 640 ----
    constexpr int kSets = NT <= 2 ? 4 : 2;   // independent accumulator sets per tile (chains of 40 / kSets MMAs)
    float acc[kSets][NT][4];
#pragma unroll
    for (int a = 0; a < kSets; ++a)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[a][n][0] = acc[a][n][1] = acc[a][n][2] = acc[a][n][3] = 0.f;
    const uint32_t hb = s_h + static_cast<uint32_t>(cur * kHBufBytes + t4 * kGroupBytes + g * 16);
#pragma unroll
    for (int kb = 0; kb < kKSteps / 2; ++kb) {
      uint32_t bq[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) lds128(bq[n], hb + static_cast<uint32_t>(4 * kb * kGroupBytes + n * 128));
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int ks = 2 * kb + h2;
        uint32_t a[4];
        if (ks < kKReg) {
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = wreg[ks % kKReg][i];
        } else {
          lds128(a, s_w + static_cast<uint32_t>((((ks - kKReg) * kClWarps + warp) * 32 + lane) * 16));
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) mma_f16_16x8x16(acc[ks % kSets][n], a, bq[n][2 * h2], bq[n][2 * h2 + 1]);
      }
    }
    float z[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (kSets == 4) z[n][i] = (acc[0][n][i] + acc[1][n][i]) + (acc[2 % kSets][n][i] + acc[3 % kSets][n][i]);
        else z[n][i] = acc[0][n][i] + acc[1][n][i];
      }

    // ---- gate exchange (warp shuffles): a low lane holds z_i (c0, c1) / z_g (c2, c3) of its unit for every tile, its
    // partner lane ^ 16 holds z_f / z_o of the same unit ----
    float zi[NT], zf[NT], zg[NT], zo[NT];
#pragma unroll
    for (int q = 0; q < NT / 2; ++q) {
      float rcv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) rcv[i] = __shfl_xor_sync(0xffffffffu, high ? z[2 * q][i] : z[2 * q + 1][i], 16);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        zi[2 * q + j] = high ? rcv[j] : z[2 * q][j];
        zf[2 * q + j] = high ? z[2 * q + 1][j] : rcv[j];
        zg[2 * q + j] = high ? rcv[2 + j] : z[2 * q][2 + j];
        zo[2 * q + j] = high ? z[2 * q + 1][2 + j] : rcv[2 + j];
      }
    }
    if (NT & 1) {
      constexpr int T = NT - 1;
      const float r0 = __shfl_xor_sync(0xffffffffu, high ? z[T][0] : z[T][1], 16);
      const float r1 = __shfl_xor_sync(0xffffffffu, high ? z[T][2] : z[T][3], 16);
      zi[T] = high ? r0 : z[T][0];
      zf[T] = high ? z[T][1] : r0;
      zg[T] = high ? r1 : z[T][2];
      zo[T] = high ? z[T][3] : r1;
    }
    float hval[NT];
#pragma unroll
    for (int p = 0; p < NT; ++p) {
      const float cn = sigmoid_fast(zf[p] + zin[p][1]) * c[p] + sigmoid_fast(zi[p] + zin[p][0]) * tanh_fast(zg[p] + zin[p][2]);
      const float hn = sigmoid_fast(zo[p] + zin[p][3]) * tanh_fast(cn);
      hval[p] = 0.f;
      if (active[p]) {
        c[p] = cn;
        hval[p] = hn;
        // unrounded: this is the head GEMM's operand
        __stcg(prm.hcat + (row0[p] + tt[p]) * H2 + dir * H + unit, hn);
      }
    }
    if (s + 1 < steps) {
      // ---- h_t (fp16) -> staging tile -> the next h buffer of all 16 CTAs ----
      const uint32_t st = s_stage + static_cast<uint32_t>(cur * kChunkBytes);
#pragma unroll
      for (int p = 0; p < NT; ++p) {
        const unsigned short hb16 = __half_as_ushort(__float2half_rn(hval[p]));
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(st + stage_off + static_cast<uint32_t>(slot[p] * 16)), "h"(hb16) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (warp == 0) {
        const int nxt = cur ^ 1;
        if (lane == 0) bar_expect_tx(s_bar + 8u * nxt, kCl * kChunkBytes);
        if (lane < kCl) {
          const uint32_t dst = map_to_cta(s_h + static_cast<uint32_t>(nxt * kHBufBytes) + rank * kChunkBytes, lane);
          const uint32_t bar = map_to_cta(s_bar + 8u * nxt, lane);
          bulk_copy_to_cta(dst, st, kChunkBytes, bar);
        }
      }
    }
  }
  cluster_sync_all();   // nobody leaves while a peer may still read its staging tile or write its h buffer
}

}  // namespace

using ClusterKernel = void (*)(const ClusterLstmParams);
static ClusterKernel cluster_kernel(int nt) {
  return nt == 1 ? lstm_cluster_kernel<1> : (nt == 2 ? lstm_cluster_kernel<2> : lstm_cluster_kernel<3>);
}
static int cluster_smem(int nt) {
  return nt == 1 ? ClusterLayout<1>::smem_bytes : (nt == 2 ? ClusterLayout<2>::smem_bytes : ClusterLayout<3>::smem_bytes);
}
static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int nt, int clusters, cudaStream_t stream) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(static_cast<unsigned>(kCl * clusters));
  cfg->blockDim = dim3(kClThreads);
  cfg->dynamicSmemBytes = static_cast<size_t>(cluster_smem(nt));
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

// Is the cluster kernel launchable on the current device (16-CTA clusters are a non-portable size), and how many
// clusters can be resident at once?
static int cluster_query(int* max_clusters) {
  int n_min = 1 << 30;
  for (int nt = 1; nt <= kMaxNT; ++nt) {
    M2S_CUDA_OK(cudaFuncSetAttribute(cluster_kernel(nt), cudaFuncAttributeMaxDynamicSharedMemorySize, cluster_smem(nt)));
    M2S_CUDA_OK(cudaFuncSetAttribute(cluster_kernel(nt), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config(&cfg, attr, nt, 8, nullptr);
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, cluster_kernel(nt), &cfg);
    if (e != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    n_min = n < n_min ? n : n_min;
  }
  *max_clusters = n_min;
  return M2S_OK;
}

// Number of 16-CTA clusters of the kernel the current device can hold at once (0: not launchable -> the caller keeps
// the grid-barrier kernel of lstm_sm100.cu).
int lstm_cluster_max_active() {
  static std::mutex mu;
  static int cached[64];
  static uint64_t known = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && ((known >> dev) & 1)) return cached[dev];
  int n = 0;
  if (cluster_query(&n) != M2S_OK) n = 0;
  if (dev >= 0 && dev < 64) { cached[dev] = n; known |= 1ull << dev; }
  if (std::getenv("M2S_DEBUG")) fprintf(stderr, "m2s: lstm cluster kernel: %d active 16-CTA clusters on device %d\n", n, dev);
  return n;
}

// Grouping: a cluster steps through its group's longest utterance and a step costs (almost) the same for 1 or 8 utterances of
// an n8 tile, so the batch is spread over as many clusters as the device holds at once (per direction: half of them),
// with the smallest tile count NT that covers it; beyond 24 utterances per resident cluster the groups run in waves.
int lstm_recurrence_cluster(const float* gin, const float* w_hh_fwd, const float* w_hh_bwd, const int32_t* lens,
                            float* hcat, int batch, int frames, int hidden, cudaStream_t stream) {
  if (hidden != kHidden) return fail(M2S_ERR_UNSUPPORTED, "LSTM recurrence is specialised for hidden=640 (got %d)", hidden);
  if (batch <= 0 || frames <= 0) return M2S_OK;
  const int resident = lstm_cluster_max_active();
  if (resident < 1) return fail(M2S_ERR_UNSUPPORTED, "16-CTA clusters are not launchable on this device");
  const int per_dir = resident / 2 > 0 ? resident / 2 : 1;
  int nt = kMaxNT;
  for (int t = 1; t <= kMaxNT; ++t)
    if ((batch + 8 * t - 1) / (8 * t) <= per_dir) { nt = t; break; }
  const int groups = (batch + 8 * nt - 1) / (8 * nt);
  ClusterLstmParams prm{};
  prm.gin = gin; prm.w_hh[0] = w_hh_fwd; prm.w_hh[1] = w_hh_bwd; prm.lens = lens; prm.hcat = hcat;
  prm.batch = batch; prm.frames = frames;
  prm.utt_per_group = (batch + groups - 1) / groups;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, nt, 2 * groups, stream);
  M2S_CUDA_OK(cudaLaunchKernelEx(&cfg, cluster_kernel(nt), prm));
  return M2S_OK;
}

}  // namespace m2s

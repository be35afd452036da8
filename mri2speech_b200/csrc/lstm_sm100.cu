// Persistent bidirectional LSTM recurrence for sm_100a.
//
// Reference semantics: torch.nn.LSTM(208, 640, 1, batch_first, bidirectional) as called by
// BiLSTMSumMerge.forward (mri2speech_code/mri_acoustic_model.py:57-71): gate order i,f,g,o, h0=c0=0,
//   z = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh;  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)
// The input projection (W_ih x + b_ih + b_hh for all timesteps, both directions) is ONE tensor-core GEMM
// on the conv engine; this kernel does only the sequential part.
//
// Design: one cooperative launch, 2 directions x 64 CTAs.  CTA p of a direction owns 10 hidden units
// (40 gate rows of W_hh); the 40 x 640 weight slice lives in REGISTERS for the whole sequence
// (320 threads x 80 weights; a warp = 4 rows, a lane = 1/32 of k), so a step costs 40 packed FMAs per (thread,
// utterance), 5 coalesced LDS.128 of h and a halving butterfly over the lanes.  h_t is exchanged through the output buffer itself (it is both the
// layer output and the next step's operand, read back through L2 with ld.cg) and a per-direction
// monotonic arrive counter acts as the grid barrier between steps.  Utterances of a ragged batch are
// skipped once s >= len_b; the reverse direction starts at each utterance's own last frame.
#include "m2s_common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace m2s {

namespace {

constexpr int kHidden = 640;
constexpr int kUnits = 10;                 // hidden units per CTA
constexpr int kParts = kHidden / kUnits;   // 64 CTAs per direction
constexpr int kRows = 4 * kUnits;          // 40 gate rows per CTA
constexpr int kWarps = 10;                 // a warp owns 4 gate rows, its 32 lanes split K 32 ways
constexpr int kRowsPerWarp = kRows / kWarps;            // 4
constexpr int kThreads = 32 * kWarps;      // 320
constexpr int kK4 = kHidden / 4 / 32;      // 5 float4 of h (20 k values) per lane
constexpr int kChunk = 16;                 // utterances per SMEM pass

struct LstmParams {
  const float* gin;      // (B, T, 2*4*H): [dir][gate][unit] pre-activations incl. both biases
  const float* w_hh[2];  // (4H, H) per direction
  const int32_t* lens;   // device, may be null
  float* hcat;           // (B, T, 2H): forward h in [0,H), backward h in [H,2H)
  unsigned int* counters;  // [2] arrive counters, zeroed before launch
  int batch, frames, max_len;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2_pk(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Thread mapping (round 2).  The first version gave a thread one gate row and an eighth of K: all 40 rows of a CTA then
// read the same h values, 20 broadcast LDS.128 per utterance per thread, and the step was bound by SMEM wavefronts
// (3.3 us per 8 utterances, whatever the FMA count).  Now a WARP owns 4 gate rows and its 32 lanes split K 32 ways: a
// lane reads its 5 float4 of h once per utterance (the warp's LDS.128 covers 512 contiguous bytes) and uses them for
// all 4 rows; the 4 row sums are reduced over the lanes with a halving butterfly (6 shuffles).
__global__ void __launch_bounds__(kThreads, 1) lstm_recurrence_kernel(LstmParams prm) {
  __shared__ float4 sh_h[kChunk][kHidden / 4];
  __shared__ float sh_z[kChunk][kRows];
  extern __shared__ float sh_c[];  // [batch][kUnits] cell state, lives for the whole sequence

  const int dir = blockIdx.x / kParts;
  const int part = blockIdx.x % kParts;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int H2 = 2 * kHidden, G = 8 * kHidden;

  // weights -> registers: row q of this warp is CTA row r = 4 warp + q = gate * 10 + unit
  unsigned long long w2[kRowsPerWarp][kK4 * 2];   // (k, k+1) pairs of W[grow][(j * 32 + lane) * 4 ..]
  int grow[kRowsPerWarp];
#pragma unroll
  for (int q = 0; q < kRowsPerWarp; ++q) {
    const int r = warp * kRowsPerWarp + q;
    grow[q] = (r / kUnits) * kHidden + part * kUnits + (r % kUnits);
    const float4* wrow = reinterpret_cast<const float4*>(prm.w_hh[dir] + static_cast<size_t>(grow[q]) * kHidden);
#pragma unroll
    for (int j = 0; j < kK4; ++j) {
      const float4 v = __ldg(wrow + j * 32 + lane);
      w2[q][2 * j] = pack2(v.x, v.y);
      w2[q][2 * j + 1] = pack2(v.z, v.w);
    }
  }
  // after the butterfly, lane l holds the sum of row q_of(l) = ((l >> 4) & 1) * 2 + ((l >> 3) & 1); lanes with (l & 7) == 0
  // publish it
  const int my_q = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  const int my_r = warp * kRowsPerWarp + my_q;
  const int my_grow = (my_r / kUnits) * kHidden + part * kUnits + (my_r % kUnits);
  for (int i = tid; i < prm.batch * kUnits; i += kThreads) sh_c[i] = 0.f;
  __syncthreads();

  unsigned int* counter = prm.counters + dir;

  for (int s = 0; s < prm.max_len; ++s) {
    // ---- wait until every CTA of this direction has published step s-1 ----
    if (s > 0) {
      if (tid == 0) {
        const unsigned int target = static_cast<unsigned int>(s) * kParts;
        unsigned int v;
        long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (v < target && clock64() - t0 > 4000000000LL) {
            printf("m2s lstm: grid barrier timeout (block %d step %d, %u < %u)\n", blockIdx.x, s, v, target);
            __trap();
          }
        } while (v < target);
      }
      __syncthreads();
    }
    for (int b0 = 0; b0 < prm.batch; b0 += kChunk) {
      const int nb = min(kChunk, prm.batch - b0);
      // ---- stage h_{t-1} of the chunk into SMEM (zeros at the first step / inactive utterances) ----
      for (int i = tid; i < nb * (kHidden / 4); i += kThreads) {
        const int bb = i / (kHidden / 4), k4 = i % (kHidden / 4);
        const int b = b0 + bb;
        const int len = prm.lens ? prm.lens[b] : prm.frames;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s > 0 && s < len) {
          const int tprev = dir == 0 ? s - 1 : len - s;
          v = __ldcg(reinterpret_cast<const float4*>(prm.hcat + (static_cast<size_t>(b) * prm.frames + tprev) * H2 +
                                                     dir * kHidden) + k4);
        }
        sh_h[bb][k4] = v;
      }
      __syncthreads();
      // ---- W_hh slice . h for every active utterance of the chunk ----
#pragma unroll 2
      for (int bb = 0; bb < nb; ++bb) {
        const int b = b0 + bb;
        const int len = prm.lens ? prm.lens[b] : prm.frames;
        if (s >= len) continue;  // block-uniform
        const int t = dir == 0 ? s : len - 1 - s;
        float gin = 0.f;
        if ((lane & 7) == 0) gin = __ldg(prm.gin + (static_cast<size_t>(b) * prm.frames + t) * G + dir * 4 * kHidden + my_grow);
        unsigned long long acc[kRowsPerWarp][2];
#pragma unroll
        for (int q = 0; q < kRowsPerWarp; ++q) acc[q][0] = acc[q][1] = 0ull;
#pragma unroll
        for (int j = 0; j < kK4; ++j) {
          const float4 h = sh_h[bb][j * 32 + lane];
          const unsigned long long h01 = pack2(h.x, h.y), h23 = pack2(h.z, h.w);
#pragma unroll
          for (int q = 0; q < kRowsPerWarp; ++q) {
            acc[q][0] = ffma2_pk(w2[q][2 * j], h01, acc[q][0]);
            acc[q][1] = ffma2_pk(w2[q][2 * j + 1], h23, acc[q][1]);
          }
        }
        float sum[kRowsPerWarp];
#pragma unroll
        for (int q = 0; q < kRowsPerWarp; ++q) {
          float a, b2, c, d;
          unpack2(acc[q][0], a, b2);
          unpack2(acc[q][1], c, d);
          sum[q] = (a + b2) + (c + d);
        }
        // halving butterfly: after xor 16 a lane keeps rows {0,1} (lanes 0-15) or {2,3} (16-31); after xor 8 one row
        {
          const bool up = (lane & 16) != 0;
          const float s0 = up ? sum[0] : sum[2], s1 = up ? sum[1] : sum[3];   // what the partner keeps
          const float k0 = up ? sum[2] : sum[0], k1 = up ? sum[3] : sum[1];   // what this lane keeps
          const float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
          const float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
          const bool up8 = (lane & 8) != 0;
          float v = (up8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, up8 ? r0 : r1, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if ((lane & 7) == 0) sh_z[bb][my_r] = v + gin;
        }
      }
      __syncthreads();
      // ---- gate math for (utterance, unit) ----
      if (tid < nb * kUnits) {
        const int bb = tid / kUnits, u = tid % kUnits;
        const int b = b0 + bb;
        const int len = prm.lens ? prm.lens[b] : prm.frames;
        if (s < len) {
          const int t = dir == 0 ? s : len - 1 - s;
          const float zi = sh_z[bb][u], zf = sh_z[bb][kUnits + u], zg = sh_z[bb][2 * kUnits + u],
                      zo = sh_z[bb][3 * kUnits + u];
          float c = sh_c[b * kUnits + u];
          c = sigmoidf_acc(zf) * c + sigmoidf_acc(zi) * tanhf(zg);
          float h = sigmoidf_acc(zo) * tanhf(c);
          sh_c[b * kUnits + u] = c;
          // h is stored unrounded: it is both the recurrent state and the head-GEMM operand (the tensor
          // core truncates; rounding here would feed the rounding error back through the recurrence)
          __stcg(prm.hcat + (static_cast<size_t>(b) * prm.frames + t) * H2 + dir * kHidden + part * kUnits + u, h);
        }
      }
      __syncthreads();
    }
    // ---- publish step s ----
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core step (tf32 / fp16 builds).  The same decomposition (2 x 64 CTAs, 40 gate rows each, grid barrier per step),
// but the 16 x 40 x 640 product of a chunk of 16 utterances is warp-level mma.sync.m16n8k8 tf32 (fp32 accumulate): warp w
// owns k in [64 w, 64 w + 64) -- its W_hh fragments (8 k-steps x 5 n-tiles x 2 registers) stay in registers for the whole
// sequence -- and the ten partial products are summed through SMEM in a fixed order.  Per chunk a lane issues 32 LDS + 40
// MMAs instead of 80 LDS.128 + 640 FFMA2 + 96 shuffles.  (A 16-row product per step is far below what tcgen05 is for:
// its M is 64 or 128 and its A operand would have to be gathered row by row, because every utterance of a ragged
// reverse pass sits at its own time index.)  h_{t-1} and W_hh are rounded to tf32 (cvt.rna); on this path's weights that
// changes the mel by 4e-6 (default init) / 5e-5 relative (scaled init): tools/emulate_residual_rounding.py's sibling probe,
// DESIGN.md 3.  The fp32 build keeps the FMA kernel above.
constexpr int kHPitch = kHidden + 4;   // floats per staged h row: A-fragment loads (8 rows x 4 k) hit 32 different banks

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT: m16 tiles (16 utterances each) per pass over the chunk: 2 halves the barriers / staging round trips per step for
// batches above 16 utterances.
template <int MT>
__global__ void __launch_bounds__(kThreads, 1) lstm_recurrence_mma_kernel(LstmParams prm) {
  constexpr int kCh = 16 * MT;                             // utterances per pass
  extern __shared__ float smem_dyn[];
  float* sh_h = smem_dyn;                                  // [kCh][kHPitch] h_{t-1} of the chunk, tf32-rounded
  float* sh_part = sh_h + kCh * kHPitch;                   // [10 warps][kCh][40] partial products
  float* sh_z = sh_part + kWarps * kCh * kRows;            // [kCh][40]
  float* sh_c = sh_z + kCh * kRows;                        // [batch][kUnits] cell state

  const int dir = blockIdx.x / kParts;
  const int part = blockIdx.x % kParts;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int H2 = 2 * kHidden, G = 8 * kHidden;

  // B fragments: W[row n = nt * 8 + g][k = 64 warp + 8 ks + t4 (+ 4)], row r = gate * 10 + unit
  uint32_t wb[8][5][2];
#pragma unroll
  for (int nt = 0; nt < 5; ++nt) {
    const int r = nt * 8 + g;
    const int grow = (r / kUnits) * kHidden + part * kUnits + (r % kUnits);
    const float* wrow = prm.w_hh[dir] + static_cast<size_t>(grow) * kHidden + warp * 64 + t4;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      wb[ks][nt][0] = to_tf32(__ldg(wrow + ks * 8));
      wb[ks][nt][1] = to_tf32(__ldg(wrow + ks * 8 + 4));
    }
  }
  for (int i = tid; i < prm.batch * kUnits; i += kThreads) sh_c[i] = 0.f;
  __syncthreads();

  unsigned int* counter = prm.counters + dir;
  constexpr int kZPerThread = (kCh * kRows + kThreads - 1) / kThreads;   // (utterance, row) sums per thread: 2 or 4

  for (int s = 0; s < prm.max_len; ++s) {
    if (s > 0) {
      if (tid == 0) {
        const unsigned int target = static_cast<unsigned int>(s) * kParts;
        unsigned int v;
        long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (v < target && clock64() - t0 > 4000000000LL) {
            printf("m2s lstm: grid barrier timeout (block %d step %d, %u < %u)\n", blockIdx.x, s, v, target);
            __trap();
          }
        } while (v < target);
      }
      __syncthreads();
    }
    for (int b0 = 0; b0 < prm.batch; b0 += kCh) {
      const int nb = min(kCh, prm.batch - b0);
      // the input projections this thread will add below: requested now, used after the MMAs
      float zin[kZPerThread];
#pragma unroll
      for (int j = 0; j < kZPerThread; ++j) {
        const int i = tid + j * kThreads;
        zin[j] = 0.f;
        if (i < nb * kRows) {
          const int bb = i / kRows, r = i % kRows;
          const int b = b0 + bb;
          const int len = prm.lens ? prm.lens[b] : prm.frames;
          if (s < len) {
            const int t = dir == 0 ? s : len - 1 - s;
            const int grow = (r / kUnits) * kHidden + part * kUnits + (r % kUnits);
            zin[j] = __ldg(prm.gin + (static_cast<size_t>(b) * prm.frames + t) * G + dir * 4 * kHidden + grow);
          }
        }
      }
      // ---- stage h_{t-1} of the chunk (tf32-rounded; zeros at the first step, for finished / absent utterances) ----
      for (int i = tid; i < kCh * (kHidden / 4); i += kThreads) {
        const int bb = i / (kHidden / 4), k4 = i % (kHidden / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bb < nb && s > 0) {
          const int b = b0 + bb;
          const int len = prm.lens ? prm.lens[b] : prm.frames;
          if (s < len) {
            const int tprev = dir == 0 ? s - 1 : len - s;
            v = __ldcg(reinterpret_cast<const float4*>(prm.hcat + (static_cast<size_t>(b) * prm.frames + tprev) * H2 +
                                                       dir * kHidden) + k4);
          }
        }
        float* dst = sh_h + bb * kHPitch + 4 * k4;
        dst[0] = __uint_as_float(to_tf32(v.x)); dst[1] = __uint_as_float(to_tf32(v.y));
        dst[2] = __uint_as_float(to_tf32(v.z)); dst[3] = __uint_as_float(to_tf32(v.w));
      }
      __syncthreads();
      // ---- partial product of this warp's K slice: (16 MT utterances) x (40 rows) ----
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt * 16 >= nb) break;   // (block-uniform)
        float acc[5][4];
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const float* hp = sh_h + (mt * 16 + g) * kHPitch + warp * 64 + t4;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t a[4];
          a[0] = __float_as_uint(hp[ks * 8]);
          a[1] = __float_as_uint(hp[8 * kHPitch + ks * 8]);
          a[2] = __float_as_uint(hp[ks * 8 + 4]);
          a[3] = __float_as_uint(hp[8 * kHPitch + ks * 8 + 4]);
#pragma unroll
          for (int nt = 0; nt < 5; ++nt) mma_tf32_16x8x8(acc[nt], a, wb[ks][nt][0], wb[ks][nt][1]);
        }
        // C fragment: (utt g, row nt*8 + 2 t4 + {0,1}), (utt g + 8, ...)
        float* pp = sh_part + warp * (kCh * kRows) + mt * 16 * kRows;
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
          *reinterpret_cast<float2*>(pp + g * kRows + nt * 8 + 2 * t4) = make_float2(acc[nt][0], acc[nt][1]);
          *reinterpret_cast<float2*>(pp + (g + 8) * kRows + nt * 8 + 2 * t4) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
      __syncthreads();
      // ---- sum the ten K slices (fixed order) + the input projection ----
#pragma unroll
      for (int j = 0; j < kZPerThread; ++j) {
        const int i = tid + j * kThreads;
        if (i < nb * kRows) {
          float z = zin[j];
#pragma unroll
          for (int w = 0; w < kWarps; ++w) z += sh_part[w * (kCh * kRows) + i];
          sh_z[i] = z;
        }
      }
      __syncthreads();
      // ---- gate math for (utterance, unit) ----
      for (int i = tid; i < nb * kUnits; i += kThreads) {
        const int bb = i / kUnits, u = i % kUnits;
        const int b = b0 + bb;
        const int len = prm.lens ? prm.lens[b] : prm.frames;
        if (s < len) {
          const int t = dir == 0 ? s : len - 1 - s;
          const float* zr = sh_z + bb * kRows;
          const float zi = zr[u], zf = zr[kUnits + u], zg = zr[2 * kUnits + u], zo = zr[3 * kUnits + u];
          float c = sh_c[b * kUnits + u];
          c = sigmoidf_acc(zf) * c + sigmoidf_acc(zi) * tanhf(zg);
          const float h = sigmoidf_acc(zo) * tanhf(c);
          sh_c[b * kUnits + u] = c;
          __stcg(prm.hcat + (static_cast<size_t>(b) * prm.frames + t) * H2 + dir * kHidden + part * kUnits + u, h);
        }
      }
      // (no barrier needed here: the next chunk's staging writes sh_h, which nobody reads any more, and the barrier after
      // it orders sh_part / sh_z)
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Any hidden size (H % 32 == 0, H != 640: models trained with another --rnn-hidden).  Not a tuned path: the same
// cooperative structure as lstm_recurrence_kernel (one arrive counter per direction as the grid barrier, h_t exchanged
// through the output buffer), but a WARP owns one hidden unit -- its four gate rows, read from L2 every step -- and its
// lanes split K; an utterance's four row sums are reduced over the lanes with a halving butterfly and lane 0 does the gate
// math, so the cell state lives in SMEM as [batch][units].  fp32 FMAs in every build.
constexpr int kGenUnits = 8;                  // hidden units (= warps) per CTA
constexpr int kGenThreads = 32 * kGenUnits;
constexpr int kGenChunk = 8;                  // utterances staged per pass

__global__ void __launch_bounds__(kGenThreads) lstm_recurrence_generic_kernel(LstmParams prm, int H) {
  extern __shared__ float sh_gen[];           // [kGenChunk][H] h_{t-1} of the chunk, then [batch][kGenUnits] cell state
  float* sh_h = sh_gen;
  float* sh_c = sh_gen + kGenChunk * H;
  const int parts = (H + kGenUnits - 1) / kGenUnits;
  const int dir = blockIdx.x / parts;
  const int part = blockIdx.x % parts;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = part * kGenUnits + warp;
  const bool unit_ok = unit < H;
  const int H2 = 2 * H, G = 8 * H;
  const int k4n = H >> 2;                     // float4 per row
  for (int i = tid; i < prm.batch * kGenUnits; i += kGenThreads) sh_c[i] = 0.f;
  __syncthreads();
  unsigned int* counter = prm.counters + dir;
  const float* w = prm.w_hh[dir];

  for (int s = 0; s < prm.max_len; ++s) {
    if (s > 0) {
      if (tid == 0) {
        const unsigned int target = static_cast<unsigned int>(s) * parts;
        unsigned int v;
        long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (v < target && clock64() - t0 > 4000000000LL) {
            printf("m2s lstm (generic): grid barrier timeout (block %d step %d, %u < %u)\n", blockIdx.x, s, v, target);
            __trap();
          }
        } while (v < target);
      }
      __syncthreads();
    }
    for (int b0 = 0; b0 < prm.batch; b0 += kGenChunk) {
      const int nb = min(kGenChunk, prm.batch - b0);
      for (int i = tid; i < nb * k4n; i += kGenThreads) {
        const int bb = i / k4n, k4 = i - bb * k4n;
        const int b = b0 + bb;
        const int len = prm.lens ? prm.lens[b] : prm.frames;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s > 0 && s < len) {
          const int tprev = dir == 0 ? s - 1 : len - s;
          v = __ldcg(reinterpret_cast<const float4*>(prm.hcat + (static_cast<size_t>(b) * prm.frames + tprev) * H2 + dir * H) + k4);
        }
        reinterpret_cast<float4*>(sh_h)[bb * k4n + k4] = v;
      }
      __syncthreads();
      if (unit_ok) {
        float acc[kGenChunk][4];
#pragma unroll
        for (int bb = 0; bb < kGenChunk; ++bb) acc[bb][0] = acc[bb][1] = acc[bb][2] = acc[bb][3] = 0.f;
        for (int k4 = lane; k4 < k4n; k4 += 32) {
          float4 wv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) wv[q] = __ldg(reinterpret_cast<const float4*>(w + (static_cast<size_t>(q) * H + unit) * H) + k4);
#pragma unroll
          for (int bb = 0; bb < kGenChunk; ++bb) {
            const float4 h = reinterpret_cast<const float4*>(sh_h)[bb * k4n + k4];   // (rows past nb hold stale data: unused)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[bb][q] = fmaf(wv[q].x, h.x, fmaf(wv[q].y, h.y, fmaf(wv[q].z, h.z, fmaf(wv[q].w, h.w, acc[bb][q]))));
          }
        }
#pragma unroll
        for (int bb = 0; bb < kGenChunk; ++bb) {
          if (bb >= nb) break;
          const int b = b0 + bb;
          const int len = prm.lens ? prm.lens[b] : prm.frames;
          float z[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v = acc[bb][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            z[q] = v;
          }
          if (lane == 0 && s < len) {
            const int t = dir == 0 ? s : len - 1 - s;
            const float* gp = prm.gin + (static_cast<size_t>(b) * prm.frames + t) * G + dir * 4 * H + unit;
            const float zi = z[0] + gp[0], zf = z[1] + gp[H], zg = z[2] + gp[2 * H], zo = z[3] + gp[3 * H];
            float c = sh_c[b * kGenUnits + warp];
            c = sigmoidf_acc(zf) * c + sigmoidf_acc(zi) * tanhf(zg);
            sh_c[b * kGenUnits + warp] = c;
            __stcg(prm.hcat + (static_cast<size_t>(b) * prm.frames + t) * H2 + dir * H + unit, sigmoidf_acc(zo) * tanhf(c));
          }
        }
      }
      __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}

}  // namespace

// lstm_cluster_sm100.cu: the recurrence on 16-CTA clusters (W_hh resident in SMEM, h exchanged through DSMEM)
int lstm_cluster_max_active();
int lstm_recurrence_cluster(const float* gin, const float* w_hh_fwd, const float* w_hh_bwd, const int32_t* lens,
                            float* hcat, int batch, int frames, int hidden, cudaStream_t stream);

// hcat must be zero-initialised by the caller where rows past lens[b] are expected to read as zero.
// Dispatch: fp32 build -> FMA kernel (grid barrier); tf32 / fp16 builds -> cluster kernel when the device can launch
// 16-CTA clusters (M2S_LSTM_CLUSTER=0 forces the grid-barrier tensor-core kernel, M2S_LSTM_MMA=0 the FMA kernel).
int lstm_recurrence(const float* gin, const float* w_hh_fwd, const float* w_hh_bwd, const int32_t* lens, float* hcat,
                    unsigned int* counters, int batch, int frames, int max_len, int hidden, bool tensor_cores,
                    cudaStream_t stream) {
  if (batch <= 0 || max_len <= 0) return M2S_OK;
  if (hidden != kHidden) {
    // any other hidden size: the generic cooperative kernel (one warp per hidden unit, weights read from L2 every step)
    if (hidden <= 0 || hidden % 32) return fail(M2S_ERR_UNSUPPORTED, "LSTM hidden size %d: must be a positive multiple of 32", hidden);
    const int parts = (hidden + kGenUnits - 1) / kGenUnits;
    const size_t dyn = (static_cast<size_t>(kGenChunk) * hidden + static_cast<size_t>(batch) * kGenUnits) * sizeof(float);
    if (dyn > 200 * 1024) return fail(M2S_ERR_UNSUPPORTED, "LSTM batch %d / hidden %d too large for one launch", batch, hidden);
    static PerDeviceOnce gen_once;
    M2S_TRY(gen_once.run([&]() -> int {
      M2S_CUDA_OK(cudaFuncSetAttribute(lstm_recurrence_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      return M2S_OK;
    }));
    int per_sm = 0;
    M2S_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_recurrence_generic_kernel, kGenThreads, dyn));
    if (per_sm * sm_count() < 2 * parts)
      return fail(M2S_ERR_UNSUPPORTED, "LSTM hidden size %d needs %d co-resident CTAs, the device holds %d", hidden, 2 * parts, per_sm * sm_count());
    LstmParams gp{};
    gp.gin = gin; gp.w_hh[0] = w_hh_fwd; gp.w_hh[1] = w_hh_bwd; gp.lens = lens; gp.hcat = hcat;
    gp.counters = counters; gp.batch = batch; gp.frames = frames; gp.max_len = max_len;
    M2S_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), stream));
    int h_arg = hidden;
    void* gargs[] = {&gp, &h_arg};
    M2S_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_recurrence_generic_kernel), dim3(2 * parts),
                                            dim3(kGenThreads), gargs, dyn, stream));
    return M2S_OK;
  }
  static const bool mma_on = !(std::getenv("M2S_LSTM_MMA") && std::atoi(std::getenv("M2S_LSTM_MMA")) == 0);
  const bool mma = tensor_cores && mma_on;
  static const bool cluster_on = !(std::getenv("M2S_LSTM_CLUSTER") && std::atoi(std::getenv("M2S_LSTM_CLUSTER")) == 0);
  if (mma && cluster_on && lstm_cluster_max_active() >= 1)
    return lstm_recurrence_cluster(gin, w_hh_fwd, w_hh_bwd, lens, hcat, batch, frames, hidden, stream);
  const size_t cells = static_cast<size_t>(batch) * kUnits * sizeof(float);
  const int mt = batch > 16 ? 2 : 1;   // m16 tiles per pass
  const size_t dyn = mma ? (static_cast<size_t>(16 * mt) * kHPitch + static_cast<size_t>(kWarps + 1) * 16 * mt * kRows) * sizeof(float) + cells
                         : cells;
  if (dyn > (mma ? 200 : 160) * 1024) return fail(M2S_ERR_UNSUPPORTED, "LSTM batch %d too large for one launch", batch);
  LstmParams prm{};
  prm.gin = gin; prm.w_hh[0] = w_hh_fwd; prm.w_hh[1] = w_hh_bwd; prm.lens = lens; prm.hcat = hcat;
  prm.counters = counters; prm.batch = batch; prm.frames = frames; prm.max_len = max_len;
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(lstm_recurrence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    M2S_CUDA_OK(cudaFuncSetAttribute(lstm_recurrence_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    M2S_CUDA_OK(cudaFuncSetAttribute(lstm_recurrence_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return M2S_OK;
  }));
  M2S_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), stream));
  void* args[] = {&prm};
  void* fn = !mma ? reinterpret_cast<void*>(lstm_recurrence_kernel)
                  : (mt == 2 ? reinterpret_cast<void*>(lstm_recurrence_mma_kernel<2>) : reinterpret_cast<void*>(lstm_recurrence_mma_kernel<1>));
  M2S_CUDA_OK(cudaLaunchCooperativeKernel(fn,
                                          dim3(2 * kParts), dim3(kThreads), args, dyn, stream));
  return M2S_OK;
}

}  // namespace m2s

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

}  // namespace

// hcat must be zero-initialised by the caller where rows past lens[b] are expected to read as zero.
int lstm_recurrence(const float* gin, const float* w_hh_fwd, const float* w_hh_bwd, const int32_t* lens, float* hcat,
                    unsigned int* counters, int batch, int frames, int max_len, int hidden,
                    cudaStream_t stream) {
  if (hidden != kHidden) return fail(M2S_ERR_UNSUPPORTED, "LSTM recurrence is specialised for hidden=640 (got %d)", hidden);
  if (batch <= 0 || max_len <= 0) return M2S_OK;
  const size_t dyn = static_cast<size_t>(batch) * kUnits * sizeof(float);
  if (dyn > 64 * 1024) return fail(M2S_ERR_UNSUPPORTED, "LSTM batch %d too large for one launch (max 1638)", batch);
  LstmParams prm{};
  prm.gin = gin; prm.w_hh[0] = w_hh_fwd; prm.w_hh[1] = w_hh_bwd; prm.lens = lens; prm.hcat = hcat;
  prm.counters = counters; prm.batch = batch; prm.frames = frames; prm.max_len = max_len;
  static PerDeviceOnce attr_once;
  M2S_TRY(attr_once.run([&]() -> int {
    M2S_CUDA_OK(cudaFuncSetAttribute(lstm_recurrence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    return M2S_OK;
  }));
  M2S_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), stream));
  void* args[] = {&prm};
  M2S_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_recurrence_kernel), dim3(2 * kParts),
                                          dim3(kThreads), args, dyn, stream));
  return M2S_OK;
}

}  // namespace m2s

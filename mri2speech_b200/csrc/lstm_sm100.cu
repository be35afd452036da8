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
// (320 threads x 80 weights; thread = (row, 1/8 of k)), so a step costs 80 FMAs per (thread, utterance)
// plus a 3-step shuffle reduction.  h_t is exchanged through the output buffer itself (it is both the
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
constexpr int kKSplit = 8;
constexpr int kThreads = kRows * kKSplit;  // 320
constexpr int kWPerThread = kHidden / kKSplit;  // 80
constexpr int kChunk = 8;                  // utterances per SMEM pass

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

__global__ void __launch_bounds__(kThreads, 1) lstm_recurrence_kernel(LstmParams prm) {
  __shared__ float4 sh_h[kChunk][kHidden / 4];
  __shared__ float sh_z[kChunk][kRows];
  extern __shared__ float sh_c[];  // [batch][kUnits] cell state, lives for the whole sequence

  const int dir = blockIdx.x / kParts;
  const int part = blockIdx.x % kParts;
  const int tid = threadIdx.x;
  const int ks = tid & (kKSplit - 1);
  const int r = tid >> 3;                 // 0..39 : gate*10 + unit
  const int gate = r / kUnits, ul = r % kUnits;
  const int grow = gate * kHidden + part * kUnits + ul;  // row in W_hh / gate vector
  const int H2 = 2 * kHidden, G = 8 * kHidden;

  // weights -> registers: w[i*4+j] = W[grow][(i*8+ks)*4 + j]
  float w[kWPerThread];
  {
    const float4* wrow = reinterpret_cast<const float4*>(prm.w_hh[dir] + static_cast<size_t>(grow) * kHidden);
#pragma unroll
    for (int i = 0; i < kWPerThread / 4; ++i) {
      const float4 v = __ldg(wrow + i * kKSplit + ks);
      w[i * 4 + 0] = v.x; w[i * 4 + 1] = v.y; w[i * 4 + 2] = v.z; w[i * 4 + 3] = v.w;
    }
  }
  unsigned long long w2[kWPerThread / 2];   // the same weights as (k, k + 1) pairs
#pragma unroll
  for (int i = 0; i < kWPerThread / 2; ++i) w2[i] = pack2(w[2 * i], w[2 * i + 1]);
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
        if (ks == 0) gin = __ldg(prm.gin + (static_cast<size_t>(b) * prm.frames + t) * G + dir * 4 * kHidden + grow);
        // packed fp32 FMAs (FFMA2: two IEEE FMAs per issued instruction -- the step is bound by the FMA pipe's issue
        // rate once more than a few utterances ride along); four independent pair accumulators
        unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
#pragma unroll
        for (int i = 0; i < kWPerThread / 4; i += 2) {
          const float4 h0 = sh_h[bb][i * kKSplit + ks];
          const float4 h1 = sh_h[bb][(i + 1) * kKSplit + ks];
          a0 = ffma2_pk(w2[i * 2 + 0], pack2(h0.x, h0.y), a0);
          a1 = ffma2_pk(w2[i * 2 + 1], pack2(h0.z, h0.w), a1);
          a2 = ffma2_pk(w2[i * 2 + 2], pack2(h1.x, h1.y), a2);
          a3 = ffma2_pk(w2[i * 2 + 3], pack2(h1.z, h1.w), a3);
        }
        float s0, s1, s2, s3, s4, s5, s6, s7;
        unpack2(a0, s0, s1); unpack2(a1, s2, s3); unpack2(a2, s4, s5); unpack2(a3, s6, s7);
        float acc = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (ks == 0) sh_z[bb][r] = acc + gin;
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

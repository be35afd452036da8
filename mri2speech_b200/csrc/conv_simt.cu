// CUDA-core (fp32 FMA) kernels: the exact-fp32 build of the multi-tap conv (M2S_PREC_FP32, also the
// on-GPU cross-check of the tcgen05 engine), plus the layers that are not GEMM-shaped:
//   * conv_post 32->1 k7 + tanh                       reference models.py:126-129
//   * mel glue (de-normalise, dB -> log-power)        reference scripts/run_mri_video_inference.py:160-163,232-239
//   * (B,C,T) <-> (B,T,C) layout changes at the Generator boundary
#include "m2s_common.cuh"

namespace m2s {

namespace {

constexpr int TM = 64;  // rows per block
constexpr int TN = 64;  // cols per block
constexpr int TK = 32;  // channels per smem step

// Each thread: 4 rows x 4 cols.  256 threads -> 64 x 64 tile.
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvProblem p, const float* __restrict__ w,
                                                        int tiles_per_batch) {
  __shared__ float sa[TK][TM + 1];
  __shared__ float sw[TK][TN + 1];
  const int b = blockIdx.x / tiles_per_batch;
  const int q0 = (blockIdx.x % tiles_per_batch) * TM;
  const int n0 = blockIdx.y * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const float* abase = p.a + static_cast<size_t>(b) * p.a_batch_rows * p.a_ld;
  for (int tap = 0; tap < p.taps; ++tap) {
    const float* wt = w + static_cast<size_t>(tap) * p.n * p.c_in;
    for (int c0 = 0; c0 < p.c_in; c0 += TK) {
      for (int i = threadIdx.x; i < TM * TK; i += 256) {
        const int r = i / TK, c = i % TK;
        const int row = q0 + r + p.shift[tap];
        float v = 0.f;
        if (row >= 0 && row < p.a_rows && c0 + c < p.c_in && q0 + r < p.l_out)
          v = abase[static_cast<size_t>(row) * p.a_ld + c0 + c];
        sa[c][r] = v;
      }
      for (int i = threadIdx.x; i < TN * TK; i += 256) {
        const int r = i / TK, c = i % TK;
        float v = 0.f;
        if (n0 + r < p.n && c0 + c < p.c_in) v = wt[static_cast<size_t>(n0 + r) * p.c_in + c0 + c];
        sw[c][r] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < TK; ++c) {
        float av[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = sa[c][ty * 4 + i]; wv[i] = sw[c][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const Epilogue& e = p.epi;
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= p.l_out) continue;
    const int drow = q + p.d_row_offset;
    const bool valid = epi_row_valid(e, b, drow);
    const size_t row_index = static_cast<size_t>(b) * p.d_batch_rows + drow;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.n) continue;
      const float bias = e.bias ? e.bias[n] : 0.f;
      const float res = e.res ? e.res[row_index * e.res_ld + n] : 0.f;
      const float accum = e.accum ? e.accum[row_index * e.accum_ld + n] : 0.f;
      p.d[row_index * p.d_ld + n] = epi_apply(e, acc[i][j], bias, res, accum, valid);
    }
  }
}

// conv_post: out[b, t] = tanh(bias + sum_{j<k} sum_c W[j][c] * A[b, t+j, c]); rows >= a_rows read as zero.
template <int C, int K>
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                        float bias, float* __restrict__ out, int rows,
                                                        long long a_batch_rows, long long out_batch_stride) {
  constexpr int TT = 256;
  __shared__ float sx[(TT + K - 1) * (C + 1)];
  __shared__ float swt[K * C];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const float* ab = a + static_cast<size_t>(b) * a_batch_rows * C;
  for (int i = threadIdx.x; i < K * C; i += 256) swt[i] = w[i];
  for (int i = threadIdx.x; i < (TT + K - 1) * C; i += 256) {
    const int r = i / C, c = i % C;
    const int t = t0 + r;
    sx[r * (C + 1) + c] = (t < rows) ? ab[static_cast<size_t>(t) * C + c] : 0.f;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= rows) return;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float* xr = sx + (threadIdx.x + j) * (C + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) acc = fmaf(swt[j * C + c], xr[c], acc);
  }
  out[static_cast<size_t>(b) * out_batch_stride + t] = tanhf(acc);
}

// generic fallback (any C, K) for non-default configs
__global__ void conv_post_generic_kernel(const float* __restrict__ a, const float* __restrict__ w, float bias,
                                         float* __restrict__ out, int rows, int C, int K, long long a_batch_rows,
                                         long long out_batch_stride) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows) return;
  const float* ab = a + static_cast<size_t>(b) * a_batch_rows * C;
  float acc = bias;
  for (int j = 0; j < K; ++j) {
    if (t + j >= rows) break;
    for (int c = 0; c < C; ++c) acc = fmaf(w[j * C + c], ab[static_cast<size_t>(t + j) * C + c], acc);
  }
  out[static_cast<size_t>(b) * out_batch_stride + t] = tanhf(acc);
}

// (B, C, T) -> (B, T, C) with zero rows past lens[b]; optional TF32 rounding (operand of conv_pre).
__global__ void bct_to_btc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T,
                                  const int32_t* __restrict__ lens, int round) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = lens ? lens[b] : T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? in[(static_cast<size_t>(b) * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      float v = t < len ? tile[threadIdx.x][i] : 0.f;
      if (round) v = round_tf32(v);
      out[(static_cast<size_t>(b) * T + t) * C + c] = v;
    }
  }
}

// mel glue, one thread per (b, t, m)
__global__ void mel_glue_kernel(const float* __restrict__ pred, const float* __restrict__ mean,
                                const float* __restrict__ stdv, int B, int T, int M,
                                const int32_t* __restrict__ lens, float* __restrict__ mel_db,
                                float* __restrict__ mel_log, float* __restrict__ voc_in) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * T * M;
  if (idx >= total) return;
  const int m = idx % M;
  const int t = (idx / M) % T;
  const int b = idx / (static_cast<size_t>(M) * T);
  const bool valid = !lens || t < lens[b];
  float db = 0.f, lg = 0.f;
  if (valid) {
    db = fmaf(pred[idx], stdv[m], mean[m]);
    // log(clamp(10^(db/10), 1e-5)) evaluated as the reference does (pow, clamp, log) in fp32
    const float pw = powf(10.0f, db / 10.0f);
    lg = logf(fmaxf(pw, 1e-5f));
  }
  if (mel_db) mel_db[idx] = db;
  if (mel_log) mel_log[idx] = lg;
  if (voc_in) voc_in[(static_cast<size_t>(b) * M + m) * T + t] = lg;
}

}  // namespace

int conv_simt(const ConvProblem& p, const float* w_plain, cudaStream_t stream) {
  if (p.batch <= 0 || p.l_out <= 0) return M2S_OK;
  const int tiles_per_batch = (p.l_out + TM - 1) / TM;
  dim3 grid(p.batch * tiles_per_batch, (p.n + TN - 1) / TN);
  M2S_TRY(profile_before(stream));
  conv_simt_kernel<<<grid, 256, 0, stream>>>(p, w_plain, tiles_per_batch);
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(stream, 2.0 * p.batch * static_cast<double>(p.l_out) * p.n * p.c_in * p.taps);
}

int conv_post_tanh(const float* a, const float* w, float bias, float* out, int batch, int rows, int c, int k,
                   long long a_batch_rows, long long out_batch_stride, cudaStream_t stream) {
  if (batch <= 0 || rows <= 0) return M2S_OK;
  profile_set_tag(PROF_VOC_SIMT);
  M2S_TRY(profile_before(stream));
  if (c == 32 && k == 7) {
    dim3 grid((rows + 255) / 256, batch);
    conv_post_kernel<32, 7><<<grid, 256, 0, stream>>>(a, w, bias, out, rows, a_batch_rows, out_batch_stride);
  } else {
    dim3 grid((rows + 255) / 256, batch);
    conv_post_generic_kernel<<<grid, 256, 0, stream>>>(a, w, bias, out, rows, c, k, a_batch_rows, out_batch_stride);
  }
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(stream, 2.0 * batch * static_cast<double>(rows) * c * k);
}

int bct_to_btc(const float* in, float* out, int batch, int c, int t, const int32_t* lens, bool round,
               cudaStream_t stream) {
  if (batch <= 0 || t <= 0) return M2S_OK;
  dim3 grid((t + 31) / 32, (c + 31) / 32, batch);
  profile_set_tag(PROF_VOC_SIMT);
  M2S_TRY(profile_before(stream));
  bct_to_btc_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, out, c, t, lens, round ? 1 : 0);
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(stream, 0.0);
}

int mel_glue(const float* pred, const float* mean, const float* stdv, int batch, int frames, int n_mels,
             const int32_t* lens, float* mel_db, float* mel_log, float* voc_in, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(batch) * frames * n_mels;
  if (!total) return M2S_OK;
  profile_set_tag(PROF_VOC_SIMT);
  M2S_TRY(profile_before(stream));
  mel_glue_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(pred, mean, stdv, batch, frames,
                                                                                 n_mels, lens, mel_db, mel_log,
                                                                                 voc_in);
  M2S_CUDA_OK(cudaGetLastError());
  return profile_after(stream, 0.0);
}

}  // namespace m2s

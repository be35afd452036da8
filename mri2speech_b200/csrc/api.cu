// C-ABI plumbing: errors, device check, the generic conv test entry, mel glue.
#include "m2s_common.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace m2s {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return status;
}

int mel_glue(const float* pred, const float* mean, const float* stdv, int batch, int frames, int n_mels,
             const int32_t* lens, float* mel_db, float* mel_log, float* voc_in, cudaStream_t stream);

}  // namespace m2s

using namespace m2s;

extern "C" const char* m2s_version(void) { return "m2s 0.2.0 (sm_100a; tcgen05 tf32 / f16 conv engine)"; }

extern "C" const char* m2s_last_error_string(void) { return g_error; }

extern "C" int m2s_device_check(int device) {
  int dev = device;
  if (dev < 0) M2S_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  M2S_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  M2S_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return fail(M2S_ERR_DEVICE, "device %d is sm_%d%d; libm2s is built for sm_100a only and has no fallback", dev,
                major, minor);
  return M2S_OK;
}

extern "C" int m2s_debug_set_knob(const char* name, int value) {
  EngineKnobs& k = engine_knobs();
  if (!std::strcmp(name, "base_offset_mode")) k.base_offset_mode = value;
  else if (!std::strcmp(name, "msub")) k.msub = value;
  else if (!std::strcmp(name, "tmap_tf32")) k.tmap_tf32 = value;
  else if (!std::strcmp(name, "max_ctas")) k.max_ctas = value;
  else if (!std::strcmp(name, "a_per_tap")) k.a_per_tap = value;
  else if (!std::strcmp(name, "dbg")) k.dbg = value;
  else if (!std::strcmp(name, "n_tile_max")) k.n_tile_max = value;
  else if (!std::strcmp(name, "pair")) k.pair = value;
  else if (!std::strcmp(name, "pair_min_n")) k.pair_min_n = value;
  else if (!std::strcmp(name, "fuse_max_n")) k.fuse_max_n = value;
  else return fail(M2S_ERR_BAD_ARG, "unknown knob %s", name);
  return M2S_OK;
}

// Test entry: weights arrive as a plain device array [taps][n][c_in]; the tcgen05 path packs them on the fly.
extern "C" int m2s_conv_fwd(const m2s_conv_args* args, int impl, m2s_stream_t stream) {
  if (!args || !args->a || !args->w || (!args->d && !args->d16)) return fail(M2S_ERR_BAD_ARG, "null argument");
  M2S_TRY(m2s_device_check(-1));
  ConvProblem p = problem_from_args(*args);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == M2S_IMPL_SIMT) {
    if (p.a_half || p.d16) return fail(M2S_ERR_UNSUPPORTED, "the CUDA-core path is fp32 only");
    return conv_simt(p, args->w, st);
  }
  if (impl != M2S_IMPL_TCGEN05) return fail(M2S_ERR_BAD_ARG, "unknown impl %d", impl);
  const size_t nw = static_cast<size_t>(p.taps) * p.n * p.c_in;
  std::vector<float> hw(nw);
  M2S_CUDA_OK(cudaMemcpy(hw.data(), args->w, nw * sizeof(float), cudaMemcpyDeviceToHost));
  PackedWeights w;
  M2S_TRY(pack_weights(hw.data(), p.taps, p.n, p.c_in, p.a_half ? PACK_FP16 : PACK_FP32, &w));
  int s = conv_tcgen05(p, w, st);
  cudaError_t e = cudaStreamSynchronize(st);
  free_weights(&w);
  if (s == M2S_OK && e != cudaSuccess) return fail(M2S_ERR_CUDA, "conv kernel failed: %s", cudaGetErrorString(e));
  return s;
}

// Test entry of the fused ResBlock pair kernel: a1 describes conv1 (a = fp16 x, w, shifts, bias, act), a2 conv2
// (w, shifts, bias, epilogue, outputs; its `a` is ignored).  Weights are packed on the fly.
extern "C" int m2s_resblock_pair_fwd(const m2s_conv_args* a1, const m2s_conv_args* a2, m2s_stream_t stream) {
  if (!a1 || !a2 || !a1->a || !a1->w || !a2->w || (!a2->d && !a2->d16)) return fail(M2S_ERR_BAD_ARG, "null argument");
  M2S_TRY(m2s_device_check(-1));
  ConvProblem p1 = problem_from_args(*a1), p2 = problem_from_args(*a2);
  p2.a_half = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PackedWeights w1, w2;
  {
    std::vector<float> hw(static_cast<size_t>(p1.taps) * p1.n * p1.c_in);
    M2S_CUDA_OK(cudaMemcpy(hw.data(), a1->w, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
    M2S_TRY(pack_weights(hw.data(), p1.taps, p1.n, p1.c_in, PACK_FP16, &w1));
  }
  {
    std::vector<float> hw(static_cast<size_t>(p2.taps) * p2.n * p2.c_in);
    M2S_CUDA_OK(cudaMemcpy(hw.data(), a2->w, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
    int s2 = pack_weights(hw.data(), p2.taps, p2.n, p2.c_in, PACK_FP16, &w2);
    if (s2 != M2S_OK) { free_weights(&w1); return s2; }
  }
  int s = resblock_pair_fused(p1, w1, p2, w2, st);
  cudaError_t e = cudaStreamSynchronize(st);
  free_weights(&w1);
  free_weights(&w2);
  if (s == M2S_OK && e != cudaSuccess) return fail(M2S_ERR_CUDA, "fused pair kernel failed: %s", cudaGetErrorString(e));
  return s;
}

extern "C" int m2s_mel_glue(const float* pred_norm, const float* mean, const float* std, int32_t batch, int32_t frames,
                            int32_t n_mels, const int32_t* lengths, float* mel_db, float* mel_log, float* voc_in,
                            m2s_stream_t stream) {
  if (!pred_norm || !mean || !std) return fail(M2S_ERR_BAD_ARG, "null argument");
  return mel_glue(pred_norm, mean, std, batch, frames, n_mels, lengths, mel_db, mel_log, voc_in,
                  reinterpret_cast<cudaStream_t>(stream));
}

// Per-launch timing of the conv engine (CUDA events on the launching stream), for bench.py's roofline leg.
extern "C" int m2s_debug_profile(int enable) { return profile_enable(enable); }
extern "C" int m2s_debug_profile_read(float* ms, double* flops, int32_t cap, int32_t* n) {
  int k = 0;
  int s = profile_read(ms, flops, cap, &k);
  if (n) *n = k;
  return s;
}

extern "C" int m2s_debug_profile_tags(int32_t* tags, int32_t cap, int32_t* n) {
  if (!tags || !n) return fail(M2S_ERR_BAD_ARG, "null argument");
  *n = profile_read_tags(tags, cap);
  return M2S_OK;
}
extern "C" long long m2s_debug_launch_count(int reset) { return launch_count(reset != 0); }

// Debug timeline of CTA 0 of the conv engine: buf = device uint64[tiles * 9] (null disables).
extern "C" int m2s_debug_trace(unsigned long long* buf, int32_t tiles) {
  EngineKnobs& k = engine_knobs();
  k.trace = buf;
  k.trace_tiles = tiles;
  return M2S_OK;
}

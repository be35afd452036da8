// HiFi-GAN Generator forward on the conv engine.
//
// Reference semantics (models.py:113-131, ResBlock1 models.py:35-49, utils.py:34-35):
//   x = conv_pre(pad_right(mel, 6))
//   for each upsample stage i: x = ConvTranspose1d(lrelu(x, .1)); x = mean_j ResBlock1_j(x)
//   x = tanh(conv_post(pad_right(lrelu(x, .01), 6)))
// with every ResBlock conv causal ((k-1)*d zeros on the left).
//
// Data layout in HBM: channels-last fp32 (B, L, C).  Activations are stored POST leaky-ReLU (and
// TF32-rounded) because every consumer is a conv that wants lrelu(x) as its tensor-core operand; the
// residual path recovers x = y >= 0 ? y : y / slope in the epilogue of the conv that needs it.
//
// Buffers (workspace): P, Q ping-pong between stages; R = ResBlock state, T = conv1 output,
// S = MRF running sum.  Per stage:
//   ups:     A=P -> Q = lrelu(convT(P)+b)                       (3-tap polyphase GEMM, N = u*C_out)
//   block j: c1: A=state -> T = lrelu(c1+b); c2: A=T, res=state -> R (pairs 0,1)
//            pair 2: j=0: S = x; j=1: S += x; j=2: P = mask(lrelu((S + x)/3, slope_next))
#include "m2s_common.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace m2s {

int conv_post_tanh(const float* a, const float* w, float bias, float* out, int batch, int rows, int c, int k,
                   long long a_batch_rows, long long out_batch_stride, cudaStream_t stream);
int bct_to_btc(const float* in, float* out, int batch, int c, int t, const int32_t* lens, bool round,
               cudaStream_t stream);

namespace {

struct HostTensor {
  const float* data;
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto s : shape) n *= static_cast<size_t>(s);
    return n;
  }
};

using TensorMap = std::map<std::string, HostTensor>;

TensorMap index_tensors(const m2s_tensor* t, int n) {
  TensorMap m;
  for (int i = 0; i < n; ++i) {
    HostTensor h;
    h.data = t[i].data;
    for (int d = 0; d < t[i].ndim; ++d) h.shape.push_back(t[i].shape[d]);
    m[t[i].name] = h;
  }
  return m;
}

// Effective weight of a (possibly weight-normed) conv: W = g * v / ||v||_2 over dims (1,2) per dim-0 slice.
int folded_weight(const TensorMap& m, const std::string& prefix, std::vector<float>* w, std::vector<int64_t>* shape) {
  auto it = m.find(prefix + ".weight");
  if (it != m.end()) {
    w->assign(it->second.data, it->second.data + it->second.numel());
    *shape = it->second.shape;
    return M2S_OK;
  }
  auto ig = m.find(prefix + ".weight_g");
  auto iv = m.find(prefix + ".weight_v");
  if (ig == m.end() || iv == m.end())
    return fail(M2S_ERR_MISSING_TENSOR, "missing %s.weight or %s.weight_g/weight_v", prefix.c_str(), prefix.c_str());
  const HostTensor& v = iv->second;
  const HostTensor& g = ig->second;
  if (v.shape.size() != 3 || g.numel() != static_cast<size_t>(v.shape[0]))
    return fail(M2S_ERR_BAD_ARG, "%s: unexpected weight_g/weight_v shapes", prefix.c_str());
  const size_t inner = static_cast<size_t>(v.shape[1] * v.shape[2]);
  w->resize(v.numel());
  for (int64_t o = 0; o < v.shape[0]; ++o) {
    double ss = 0.0;
    for (size_t i = 0; i < inner; ++i) ss += static_cast<double>(v.data[o * inner + i]) * v.data[o * inner + i];
    const double s = static_cast<double>(g.data[o]) / std::sqrt(ss);
    for (size_t i = 0; i < inner; ++i) (*w)[o * inner + i] = static_cast<float>(v.data[o * inner + i] * s);
  }
  *shape = v.shape;
  return M2S_OK;
}

int get_bias(const TensorMap& m, const std::string& prefix, int n, std::vector<float>* b) {
  auto it = m.find(prefix + ".bias");
  if (it == m.end()) return fail(M2S_ERR_MISSING_TENSOR, "missing %s.bias", prefix.c_str());
  if (it->second.numel() != static_cast<size_t>(n)) return fail(M2S_ERR_BAD_ARG, "%s.bias has wrong size", prefix.c_str());
  b->assign(it->second.data, it->second.data + n);
  return M2S_OK;
}

struct Layer {
  PackedWeights w;
  float* bias = nullptr;  // device [n]
  int taps = 0;
  int shift[M2S_MAX_TAPS] = {};
};

int upload(const std::vector<float>& h, float** dev) {
  M2S_CUDA_OK(cudaMalloc(dev, h.size() * sizeof(float)));
  M2S_CUDA_OK(cudaMemcpy(*dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return M2S_OK;
}

}  // namespace
}  // namespace m2s

using namespace m2s;

struct m2s_generator {
  m2s_generator_config cfg;
  bool tf32 = true;   // tensor-core build (tf32 or fp16 operands); false = CUDA-core fp32 build
  bool fp16 = false;  // M2S_PREC_FP16: layers with c_in % 8 == 0 run kind::f16 on fp16 copies of the activations
  bool fuse_pairs = true;  // fp16 build: fused ResBlock pair kernel where it applies (M2S_FUSE_PAIRS=0 disables)
  bool split_res = false;  // fp16 build, opt-in (M2S_SPLIT_RES=1): residual stream stored as two fp16 planes (hi = the
                           // operand, lo = fp16(v - hi)) instead of fp32 + an fp16 operand copy.  8 instead of 12 bytes
                           // of DRAM traffic per element of a pair, bit-compatible results -- but measured SLOWER
                           // (profiles/README.md): the epilogue is bound by memory requests in flight, not by bytes.
  bool res16 = false;      // fp16 build (M2S_VOC_RES16): the residual of a ResBlock conv is read from the fp16 operand copy of
                           // the state the conv reads anyway, and states carry no fp32 copy: 6 instead of 12 bytes of DRAM
                           // traffic per element of a pair and a third of the epilogue's memory requests gone (one more
                           // fp16 rounding per pair on the residual chain; the MRF sums stay fp32)
  bool rb2 = false;        // cfg.resblock == 2: ResBlock2 branches (c1 holds their 2 convs each, c2 stays empty)
  int hop = 1;
  Layer pre;
  std::vector<Layer> ups;
  std::vector<int> ups_cout;
  // resblocks[i*num_kernels + j] -> c1[3], c2[3]
  std::vector<Layer> c1, c2;
  float* post_w = nullptr;  // [k][c]
  float post_bias = 0.f;
  int post_k = 0, post_c = 0;
  int launches = 0;
};

namespace {

// Conv1d weight (C_out, C_in, k) -> engine layout [tap][n][c]
int make_conv1d_layer(const TensorMap& m, const std::string& prefix, int c_out, int c_in, int k, bool causal,
                      int dilation, int mode, Layer* L) {
  std::vector<float> w;
  std::vector<int64_t> shape;
  M2S_TRY(folded_weight(m, prefix, &w, &shape));
  if (shape.size() != 3 || shape[0] != c_out || shape[1] != c_in || shape[2] != k)
    return fail(M2S_ERR_BAD_ARG, "%s: weight shape mismatch", prefix.c_str());
  if (k > M2S_MAX_TAPS) return fail(M2S_ERR_UNSUPPORTED, "%s: kernel size %d > %d", prefix.c_str(), k, M2S_MAX_TAPS);
  std::vector<float> e(static_cast<size_t>(k) * c_out * c_in);
  for (int j = 0; j < k; ++j)
    for (int o = 0; o < c_out; ++o)
      for (int c = 0; c < c_in; ++c)
        e[(static_cast<size_t>(j) * c_out + o) * c_in + c] = w[(static_cast<size_t>(o) * c_in + c) * k + j];
  L->taps = k;
  for (int j = 0; j < k; ++j) L->shift[j] = causal ? -(k - 1 - j) * dilation : j * dilation;
  M2S_TRY(pack_weights(e.data(), k, c_out, c_in, mode, &L->w));
  std::vector<float> b;
  M2S_TRY(get_bias(m, prefix, c_out, &b));
  return upload(b, &L->bias);
}

// ConvTranspose1d weight (C_in, C_out, k), stride u, padding p -> polyphase GEMM:
// row q, col r*C_out+co = sum_s sum_ci x[q - s, ci] * W[ci, co, s*u + r + p]
int make_convT_layer(const TensorMap& m, const std::string& prefix, int c_in, int c_out, int k, int u, int mode,
                     Layer* L) {
  std::vector<float> w;
  std::vector<int64_t> shape;
  M2S_TRY(folded_weight(m, prefix, &w, &shape));
  if (shape.size() != 3 || shape[0] != c_in || shape[1] != c_out || shape[2] != k)
    return fail(M2S_ERR_BAD_ARG, "%s: weight shape mismatch", prefix.c_str());
  if ((k - u) % 2) return fail(M2S_ERR_UNSUPPORTED, "%s: (kernel - stride) must be even", prefix.c_str());
  const int p = (k - u) / 2;
  int s_lo = 1 << 30, s_hi = -(1 << 30);
  for (int r = 0; r < u; ++r)
    for (int kk = 0; kk < k; ++kk)
      if ((kk - r - p) % u == 0) {
        const int s = (kk - r - p) / u;
        s_lo = s < s_lo ? s : s_lo;
        s_hi = s > s_hi ? s : s_hi;
      }
  const int taps = s_hi - s_lo + 1;
  if (taps > M2S_MAX_TAPS) return fail(M2S_ERR_UNSUPPORTED, "%s: %d polyphase taps", prefix.c_str(), taps);
  const int n = u * c_out;
  std::vector<float> e(static_cast<size_t>(taps) * n * c_in, 0.f);
  for (int t = 0; t < taps; ++t) {
    const int s = s_lo + t;
    for (int r = 0; r < u; ++r) {
      const int kk = s * u + r + p;
      if (kk < 0 || kk >= k) continue;
      for (int co = 0; co < c_out; ++co)
        for (int ci = 0; ci < c_in; ++ci)
          e[(static_cast<size_t>(t) * n + r * c_out + co) * c_in + ci] = w[(static_cast<size_t>(ci) * c_out + co) * k + kk];
    }
  }
  L->taps = taps;
  for (int t = 0; t < taps; ++t) L->shift[t] = -(s_lo + t);
  M2S_TRY(pack_weights(e.data(), taps, n, c_in, mode, &L->w));
  std::vector<float> b, be(n);
  M2S_TRY(get_bias(m, prefix, c_out, &b));
  for (int r = 0; r < u; ++r)
    for (int co = 0; co < c_out; ++co) be[r * c_out + co] = b[co];
  return upload(be, &L->bias);
}

void free_layer(Layer* L) {
  free_weights(&L->w);
  if (L->bias) cudaFree(L->bias);
  L->bias = nullptr;
}

int run_conv(const m2s_generator* g, const ConvProblem& p, const Layer& L, cudaStream_t st) {
  profile_set_tag(PROF_VOC_GEMM);
  return g->tf32 ? conv_tcgen05(p, L.w, st) : conv_simt(p, L.w.plain, st);
}

// `a` is read as fp16 rows when the layer's weights are packed for kind::f16 (L.w.half), as fp32 rows otherwise.
ConvProblem base_problem(const void* a, int a_rows, int c_in, int batch, long long batch_rows_a, float* d,
                         int d_ld, long long batch_rows_d, int l_out, const Layer& L) {
  ConvProblem p{};
  p.a = static_cast<const float*>(a); p.a_half = L.w.half != 0;
  p.a_batch_rows = batch_rows_a; p.a_rows = a_rows; p.a_ld = c_in; p.c_in = c_in;
  p.batch = batch; p.l_out = l_out; p.taps = L.taps;
  for (int j = 0; j < L.taps; ++j) p.shift[j] = L.shift[j];
  p.n = L.w.n; p.d = d; p.d_batch_rows = batch_rows_d; p.d_ld = d_ld; p.d_row_offset = 0;
  p.epi.bias = L.bias; p.epi.out_scale = 1.f; p.epi.res_inv_slope = 1.f; p.epi.act = M2S_ACT_NONE;
  return p;
}

}  // namespace

extern "C" int m2s_generator_create(const m2s_generator_config* cfg, const m2s_tensor* tensors, int32_t n_tensors,
                                    m2s_generator** out) {
  if (!cfg || !tensors || !out) return fail(M2S_ERR_BAD_ARG, "null argument");
  if (cfg->num_upsamples < 1 || cfg->num_upsamples > M2S_MAX_UPS || cfg->num_kernels < 1 ||
      cfg->num_kernels > M2S_MAX_RBK)
    return fail(M2S_ERR_BAD_ARG, "unsupported num_upsamples/num_kernels");
  if (cfg->num_mels % 4 || (cfg->upsample_initial_channel >> cfg->num_upsamples) % 4)
    return fail(M2S_ERR_UNSUPPORTED, "channel counts must stay multiples of 4");
  M2S_TRY(m2s_device_check(-1));
  TensorMap m = index_tensors(tensors, n_tensors);
  auto* g = new m2s_generator();
  g->cfg = *cfg;
  g->tf32 = cfg->precision != M2S_PREC_FP32;
  g->fp16 = cfg->precision == M2S_PREC_FP16;
  g->rb2 = cfg->resblock == 2;
  if (cfg->resblock < 0 || cfg->resblock > 2) { delete g; return fail(M2S_ERR_BAD_ARG, "resblock must be 1 or 2"); }
  if (const char* f = std::getenv("M2S_FUSE_PAIRS")) g->fuse_pairs = std::atoi(f) != 0;
  if (const char* f = std::getenv("M2S_SPLIT_RES")) g->split_res = g->fp16 && std::atoi(f) != 0;
  g->res16 = g->fp16 && !g->split_res;
  if (const char* f = std::getenv("M2S_VOC_RES16")) g->res16 = g->fp16 && !g->split_res && std::atoi(f) != 0;
  // operand format per layer: fp16 needs 16-byte aligned rows of halves (c_in % 8 == 0); conv_pre reads the fp32 mel
  auto mode_for = [&](int c_in) {
    if (!g->tf32) return static_cast<int>(PACK_FP32);
    return static_cast<int>((g->fp16 && c_in % 8 == 0) ? PACK_FP16 : PACK_TF32);
  };
  int st = M2S_OK;
  auto bail = [&](int s) { m2s_generator_destroy(g); return s; };
  const int c0 = cfg->upsample_initial_channel;
  if ((st = make_conv1d_layer(m, "conv_pre", c0, cfg->num_mels, 7, false, 1, g->tf32 ? PACK_TF32 : PACK_FP32,
                              &g->pre)) != M2S_OK)
    return bail(st);
  g->hop = 1;
  int ch = c0;
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    const int cout = c0 >> (i + 1);
    Layer L;
    if ((st = make_convT_layer(m, "ups." + std::to_string(i), ch, cout, cfg->upsample_kernel_sizes[i],
                               cfg->upsample_rates[i], mode_for(ch), &L)) != M2S_OK)
      return bail(st);
    g->ups.push_back(L);
    g->ups_cout.push_back(cout);
    g->hop *= cfg->upsample_rates[i];
    for (int j = 0; j < cfg->num_kernels; ++j) {
      const std::string p = "resblocks." + std::to_string(i * cfg->num_kernels + j);
      const int k = cfg->resblock_kernel_sizes[j];
      if (g->rb2) {  // models.py:58-80: convs.0 / convs.1, causal like ResBlock1 (utils.py:34-35 + the keep-first-L trim)
        for (int d = 0; d < 2; ++d) {
          Layer a;
          if ((st = make_conv1d_layer(m, p + ".convs." + std::to_string(d), cout, cout, k, true,
                                      cfg->resblock_dilations[j][d], mode_for(cout), &a)) != M2S_OK)
            return bail(st);
          g->c1.push_back(a);
        }
        continue;
      }
      for (int d = 0; d < 3; ++d) {
        Layer a, b;
        if ((st = make_conv1d_layer(m, p + ".convs1." + std::to_string(d), cout, cout, k, true,
                                    cfg->resblock_dilations[j][d], mode_for(cout), &a)) != M2S_OK)
          return bail(st);
        g->c1.push_back(a);
        if ((st = make_conv1d_layer(m, p + ".convs2." + std::to_string(d), cout, cout, k, true, 1, mode_for(cout),
                                    &b)) != M2S_OK)
          return bail(st);
        g->c2.push_back(b);
      }
    }
    ch = cout;
  }
  {
    std::vector<float> w;
    std::vector<int64_t> shape;
    if ((st = folded_weight(m, "conv_post", &w, &shape)) != M2S_OK) return bail(st);
    if (shape.size() != 3 || shape[0] != 1 || shape[1] != ch) return bail(fail(M2S_ERR_BAD_ARG, "conv_post shape"));
    g->post_k = static_cast<int>(shape[2]);
    g->post_c = ch;
    std::vector<float> e(static_cast<size_t>(g->post_k) * ch);
    for (int j = 0; j < g->post_k; ++j)
      for (int c = 0; c < ch; ++c) e[static_cast<size_t>(j) * ch + c] = w[static_cast<size_t>(c) * g->post_k + j];
    if ((st = upload(e, &g->post_w)) != M2S_OK) return bail(st);
    std::vector<float> b;
    if ((st = get_bias(m, "conv_post", 1, &b)) != M2S_OK) return bail(st);
    g->post_bias = b[0];
  }
  for (const Layer& L : g->c1)  // the split stream needs fp16 operands in every stage
    if (!L.w.half) g->split_res = false;
  g->launches = 2 + cfg->num_upsamples * (1 + cfg->num_kernels * (g->rb2 ? 2 : 6)) + 1;
  if (g->fp16 && g->fuse_pairs && !g->rb2) {  // stages whose ResBlock pairs run fused: one launch per pair instead of two
    for (int i = 0; i < cfg->num_upsamples; ++i)
      if (g->ups_cout[i] <= engine_knobs().fuse_max_n && g->ups_cout[i] % 32 == 0) g->launches -= cfg->num_kernels * 3;
  }
  *out = g;
  return M2S_OK;
}

extern "C" void m2s_generator_destroy(m2s_generator* g) {
  if (!g) return;
  free_layer(&g->pre);
  for (auto& L : g->ups) free_layer(&L);
  for (auto& L : g->c1) free_layer(&L);
  for (auto& L : g->c2) free_layer(&L);
  if (g->post_w) cudaFree(g->post_w);
  delete g;
}

extern "C" int m2s_generator_launches(const m2s_generator* g) { return g ? g->launches : 0; }

namespace {
struct GenBuffers {
  size_t mel_floats, p_floats, q_floats;
  size_t total_bytes;
};
GenBuffers plan_buffers(const m2s_generator* g, int batch, int frames) {
  GenBuffers b{};
  b.mel_floats = static_cast<size_t>(batch) * frames * g->cfg.num_mels;
  size_t p = static_cast<size_t>(batch) * frames * g->cfg.upsample_initial_channel;
  size_t q = 0;
  size_t L = frames;
  for (int i = 0; i < g->cfg.num_upsamples; ++i) {
    L *= g->cfg.upsample_rates[i];
    const size_t s = static_cast<size_t>(batch) * L * g->ups_cout[i];
    p = s > p ? s : p;
    q = s > q ? s : q;
  }
  auto al = [](size_t f) { return (f + 63) / 64 * 64; };
  b.mel_floats = al(b.mel_floats);
  b.p_floats = al(p);
  b.q_floats = al(q);
  // P, T, S (fp32-sized; P and T hold fp16 data when their consumer is an fp16 layer) and the stage input Q plus two
  // ResBlock states Ra / Rb (ping-pong: a pair never writes the tensor it reads, the fused pair kernel re-reads a halo
  // of its input).  Split-fp16 residual stream: each of Q, Ra, Rb is a (hi, lo) pair of fp16 planes = one fp32-sized
  // slot; otherwise an fp32 tensor plus an fp16 operand copy = 1.5 slots.
  b.total_bytes = (b.mel_floats + b.p_floats + (g->split_res ? 5 : 6) * b.q_floats + (g->split_res ? 0 : b.q_floats / 2)) * sizeof(float) + 256;
  return b;
}
}  // namespace

extern "C" size_t m2s_generator_workspace_bytes(const m2s_generator* g, int32_t batch, int32_t frames) {
  if (!g || batch <= 0 || frames <= 0) return 0;
  return plan_buffers(g, batch, frames).total_bytes;
}

namespace {
// mel_btc != 0: `mel` already is conv_pre's operand, channels-last (batch, frames, num_mels) with rows past lengths[b]
// zeroed (what m2s_mel_glue writes as mel_log): no layout pass.
int generator_forward_impl(m2s_generator* g, const float* mel, bool mel_btc, int32_t batch, int32_t frames,
                           const int32_t* lengths, float* audio, void* workspace, size_t workspace_bytes,
                           m2s_stream_t stream) {
  if (!g || !mel || !audio) return fail(M2S_ERR_BAD_ARG, "null argument");
  if (batch <= 0 || frames <= 0) return M2S_OK;
  const GenBuffers bufs = plan_buffers(g, batch, frames);
  if (!workspace || workspace_bytes < bufs.total_bytes)
    return fail(M2S_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", bufs.total_bytes, workspace_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const float* M0 = mel_btc ? mel : base;
  float* P = base + bufs.mel_floats;
  float* T = P + bufs.p_floats;
  float* S = T + bufs.q_floats;
  // stage input Q and the two ResBlock states: fp32 tensor + fp16 operand copy, or the (hi, lo) fp16 planes
  struct Stream { float* f32; __half* hi; __half* lo; };
  Stream Qs{}, Rs[2] = {};
  {
    float* cur = S + bufs.q_floats;
    Stream* all[3] = {&Qs, &Rs[0], &Rs[1]};
    for (Stream* x : all) {
      if (g->split_res) {
        x->hi = reinterpret_cast<__half*>(cur);
        x->lo = reinterpret_cast<__half*>(cur + bufs.q_floats / 2);
        cur += bufs.q_floats;
      } else {
        x->f32 = cur;
        x->hi = reinterpret_cast<__half*>(cur + bufs.q_floats);
        cur += bufs.q_floats + bufs.q_floats / 2;
      }
    }
  }
  const m2s_generator_config& cfg = g->cfg;
  const int mask = lengths ? M2S_MASK_LEN : M2S_MASK_NONE;
  // fp16 build: every tensor that is only ever a tensor-core operand (P, T) is stored as fp16; tensors that are also
  // a residual / MRF source (Q, R) are stored as the fp16 operand plane `hi` plus EITHER an fp16 correction plane
  // `lo` = fp16(v - hi) (split stream: hi + lo carries v to ~22 mantissa bits, so nothing accumulates along the residual
  // chain, in 4 instead of 6 bytes per element) OR an fp32 copy; S stays fp32.  An fp32 output goes to `d`, an fp16
  // output to `d16` (+ `d16_lo`).
  auto set_out = [](ConvProblem* p, float* buf32, void* buf16) { p->d = buf32; p->d16 = buf16; };

  // (B, mels, T) -> channels-last, zero past lengths (conv_pre look-ahead must see zeros)
  if (!mel_btc) M2S_TRY(bct_to_btc(mel, base, batch, cfg.num_mels, frames, lengths, false, st));

  int L = frames;
  int ch = cfg.upsample_initial_channel;
  {  // conv_pre -> P = mask(lrelu(conv+b, .1))  (only consumer: ups[0])
    ConvProblem p = base_problem(M0, L, cfg.num_mels, batch, L, P, ch, L, L, g->pre);
    if (g->ups[0].w.half) set_out(&p, nullptr, P);
    p.epi.act = M2S_ACT_LRELU; p.epi.act_slope = 0.1f;
    p.epi.mask_mode = mask; p.epi.lens = lengths; p.epi.len_scale = 1;
    M2S_TRY(run_conv(g, p, g->pre, st));
  }
  int scale = 1;
  for (int i = 0; i < cfg.num_upsamples; ++i) {
    const int u = cfg.upsample_rates[i];
    const int cout = g->ups_cout[i];
    const int per_block = g->rb2 ? 2 : 3;   // convs (ResBlock2) / conv pairs (ResBlock1) per branch
    const bool hs = g->c1[i * cfg.num_kernels * per_block].w.half != 0;  // this stage's ResBlock convs take fp16 operands
    {  // ups: rows = L, N = u*cout; D viewed as (B, L, u*cout) == (B, L*u, cout)
      const bool split = hs && g->split_res;
      ConvProblem p = base_problem(P, L, ch, batch, L, (split || (hs && g->res16)) ? nullptr : Qs.f32, u * cout, L, L, g->ups[i]);
      if (hs) p.d16 = Qs.hi;
      if (split) p.d16_lo = Qs.lo;
      p.epi.act = M2S_ACT_LRELU; p.epi.act_slope = 0.1f;
      M2S_TRY(run_conv(g, p, g->ups[i], st));
    }
    L *= u; scale *= u; ch = cout;
    const bool last_stage = (i + 1 == cfg.num_upsamples);
    const bool next_half = !last_stage && g->ups[i + 1].w.half != 0;  // who reads this stage's output P
    for (int j = 0; j < cfg.num_kernels; ++j) {
      const bool split = hs && g->split_res;
      const Stream* state = &Qs;
      // the epilogue of a branch's LAST conv: MRF bookkeeping -- S = x | S += x | P = mask(lrelu((S + x) / nk, slope_next))
      auto finish_branch = [&](ConvProblem* pl) {
        if (j + 1 < cfg.num_kernels) {
          pl->d = S; pl->d16 = nullptr; pl->d16_lo = nullptr;
          if (j > 0) { pl->epi.accum = S; pl->epi.accum_ld = ch; }
        } else {
          if (next_half) set_out(pl, nullptr, P);
          else { pl->d = P; pl->d16 = nullptr; }
          pl->d16_lo = nullptr;
          if (j > 0) { pl->epi.accum = S; pl->epi.accum_ld = ch; }
          pl->epi.out_scale = 1.f / static_cast<float>(cfg.num_kernels);
          pl->epi.act = M2S_ACT_LRELU; pl->epi.act_slope = last_stage ? 0.01f : 0.1f;
          pl->epi.mask_mode = mask; pl->epi.lens = lengths; pl->epi.len_scale = scale;
        }
      };
      if (g->rb2) {  // ResBlock2: x = x + c_d(lrelu(x)) for the two dilations -- one launch per conv, residual fused
        for (int d = 0; d < 2; ++d) {
          const Stream& out = Rs[d & 1];
          const Layer& l = g->c1[(i * cfg.num_kernels + j) * 2 + d];
          const void* state_op = hs ? static_cast<const void*>(state->hi) : static_cast<const void*>(state->f32);
          const bool r16 = hs && g->res16;
          ConvProblem pc = base_problem(state_op, L, ch, batch, L, (split || r16) ? nullptr : out.f32, ch, L, L, l);
          if (split) { pc.epi.res_hi = state->hi; pc.epi.res_lo = state->lo; }
          else if (r16) { pc.epi.res = reinterpret_cast<const float*>(state->hi); pc.epi.res_half = 1; }
          else pc.epi.res = state->f32;
          pc.epi.res_ld = ch; pc.epi.res_inv_slope = 10.f;
          if (d == 0) {
            pc.epi.act = M2S_ACT_LRELU; pc.epi.act_slope = 0.1f;
            if (hs) pc.d16 = out.hi;
            if (split) pc.d16_lo = out.lo;
          } else {
            finish_branch(&pc);
          }
          M2S_TRY(run_conv(g, pc, l, st));
          state = &out;
        }
        continue;
      }
      for (int d = 0; d < 3; ++d) {
        const Stream& out = Rs[d & 1];
        const void* state_op = hs ? static_cast<const void*>(state->hi) : static_cast<const void*>(state->f32);
        const Layer& l1 = g->c1[(i * cfg.num_kernels + j) * 3 + d];
        const Layer& l2 = g->c2[(i * cfg.num_kernels + j) * 3 + d];
        ConvProblem p1 = base_problem(state_op, L, ch, batch, L, T, ch, L, L, l1);
        if (hs) set_out(&p1, nullptr, T);
        p1.epi.act = M2S_ACT_LRELU; p1.epi.act_slope = 0.1f;
        const bool r16 = hs && g->res16;
        ConvProblem p2 = base_problem(T, L, ch, batch, L, (split || r16) ? nullptr : out.f32, ch, L, L, l2);
        if (split) { p2.epi.res_hi = state->hi; p2.epi.res_lo = state->lo; }
        else if (r16) { p2.epi.res = reinterpret_cast<const float*>(state->hi); p2.epi.res_half = 1; }
        else p2.epi.res = state->f32;
        p2.epi.res_ld = ch; p2.epi.res_inv_slope = 10.f;  // 1 / LRELU_SLOPE
        if (d < 2) {
          p2.epi.act = M2S_ACT_LRELU; p2.epi.act_slope = 0.1f;
          if (hs) p2.d16 = out.hi;
          if (split) p2.d16_lo = out.lo;
        } else {
          finish_branch(&p2);
        }
        // fp16 build, C <= 128: conv1 -> leaky-ReLU -> conv2 in ONE kernel, the intermediate stays in SMEM
        if (hs && g->fuse_pairs && resblock_pair_supported(p1, l1.w, p2, l2.w)) {
          profile_set_tag(PROF_VOC_GEMM);
          M2S_TRY(resblock_pair_fused(p1, l1.w, p2, l2.w, st));
        } else {
          M2S_TRY(run_conv(g, p1, l1, st));
          M2S_TRY(run_conv(g, p2, l2, st));
        }
        state = &out;
      }
    }
  }
  // conv_post + tanh: P holds mask(lrelu(x, .01)) as (B, L, ch)
  M2S_TRY(conv_post_tanh(P, g->post_w, g->post_bias, audio, batch, L, g->post_c, g->post_k, L, L, st));
  return M2S_OK;
}
}  // namespace

extern "C" int m2s_generator_forward(m2s_generator* g, const float* mel, int32_t batch, int32_t frames,
                                     const int32_t* lengths, float* audio, void* workspace, size_t workspace_bytes,
                                     m2s_stream_t stream) {
  return generator_forward_impl(g, mel, false, batch, frames, lengths, audio, workspace, workspace_bytes, stream);
}

extern "C" int m2s_generator_forward_btc(m2s_generator* g, const float* mel_btc, int32_t batch, int32_t frames,
                                         const int32_t* lengths, float* audio, void* workspace,
                                         size_t workspace_bytes, m2s_stream_t stream) {
  return generator_forward_impl(g, mel_btc, true, batch, frames, lengths, audio, workspace, workspace_bytes, stream);
}

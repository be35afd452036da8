// Acoustic model forward: frame-CNN encoder (EfficientNetV2-B2 features) -> BiLSTM (sum merge) -> mel head.
//
// Reference semantics: OTNLikeCNNBiLSTM.forward (mri2speech_code/mri_acoustic_model.py:105-136):
//   f = GAP(backbone(repeat3(x))[-1]) per frame; y = LSTM_fwd(f) + LSTM_bwd(f); out = Linear(y)
// The backbone is timm's tf_efficientnetv2_b2 (un-vendored; topology in SURVEY.md 8a-1).  Eval-mode BatchNorm
// is folded into the conv weights, the 3 identical input channels into the stem, fwd+bwd SUM into the head GEMM
// (K = 2H with the head weight duplicated), b_ih + b_hh into the input-projection GEMM.
//
// Layouts: channels-last fp32.  Stages 0-2 keep a one-pixel zero border ("padded": pitch = W+2) so that a 3x3
// stride-1 conv is 9 row-shifted taps on the conv engine; stages 3-5 are plain (frame, pixel) rows.
#include "m2s_common.cuh"
#include <cmath>
#include <cuda_fp16.h>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace m2s {
int enc_stem(const float* frames, const int32_t* fmap, void* out, int half, const float* w, const float* bias, int n,
             int H, int W, cudaStream_t st);
int enc_stem_u8(const uint8_t* frames, const int32_t* fmap, const float* mask, float2* norm, void* out, int half,
                const float* w, const float* bias, int n, int H, int W, cudaStream_t st);
int enc_zero_rows(void* buf, int esize, int n, int rows_per_frame, int ld, int head_rows, int tail_start,
                  cudaStream_t st);
int enc_zero_cols(void* buf, int esize, int n, int rows_per_frame, int ld, int pitch, int H, cudaStream_t st);
int enc_im2col_s2(const void* in, void* col, int esize, int n, int Hin, int Win, int C, cudaStream_t st);
int enc_s2d(const void* in, void* out, int esize, int n, int Hin, int Win, int C, cudaStream_t st);
int enc_dwconv(const void* in, void* out, int half, float* sums, const float* w, const float* bias, int n, int C, int Hin,
               int Win, int pitch_in, int oy, int ox, int rows_in, int stride, cudaStream_t st);
int enc_se_apply(void* x, int half, float* sums, const float* w1, const float* b1, const float* w2, const float* b2,
                 int n, int C, int rd, int hw, cudaStream_t st);
int enc_gap(const float* x, const int32_t* fmap, float* feats, int n, int hw, int C, int feat_ld, cudaStream_t st);
int enc_build_fmap(const int32_t* lens, int batch, int frames, int32_t* fmap, cudaStream_t st);
bool mb_expand_dw_supported(int H, int W, int c_in, int c_mid);
int mb_expand_dw(const void* x, const void* w_exp, const float* bias1, const void* dw_w16, const float* dw_w32, const float* dw_b,
                 void* out, float* sums, int n, int H, int W, int c_in, int c_mid, cudaStream_t st);
bool mb_project_supported(int hw, int c_mid, int c_out);
int mb_project(const void* a, const void* w, const float* scales, const float* bias, const float* res, const void* res16,
               float* d32, void* d16, int n_frames, int hw, int c_mid, int c_out, cudaStream_t st);
bool fused_er_supported(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2);
bool fused_er_resident(const PackedWeights& w1, const PackedWeights& w2);
int fused_er(const ConvProblem& p1, const PackedWeights& w1, const ConvProblem& p2, const PackedWeights& w2,
             cudaStream_t stream, const ErS2d* s2d = nullptr);
int enc_se_mlp(float* sums, const float* w1, const float* b1, const float* w2, const float* b2, int n, int C, int rd, int hw,
               cudaStream_t st);
int lstm_recurrence(const float* gin, const float* w_hh_fwd, const float* w_hh_bwd, const int32_t* lens, float* hcat,
                    unsigned int* counters, int batch, int frames, int max_len, int hidden, bool tensor_cores,
                    cudaStream_t stream);
}  // namespace m2s

using namespace m2s;

namespace {

constexpr float kBnEps = 1e-3f;

struct HT {
  const float* data = nullptr;
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto s : shape) n *= static_cast<size_t>(s);
    return n;
  }
};
using TMap = std::map<std::string, HT>;

int need(const TMap& m, const std::string& name, size_t numel, const HT** out) {
  auto it = m.find(name);
  if (it == m.end()) return fail(M2S_ERR_MISSING_TENSOR, "missing tensor %s", name.c_str());
  if (numel && it->second.numel() != numel)
    return fail(M2S_ERR_BAD_ARG, "%s has %zu elements, expected %zu", name.c_str(), it->second.numel(), numel);
  *out = &it->second;
  return M2S_OK;
}

// BN fold factors: y = s*(x) + t with s = g/sqrt(var+eps), t = b - mean*s
int bn_fold(const TMap& m, const std::string& p, int ch, std::vector<float>* s, std::vector<float>* t) {
  const HT *g, *b, *mu, *var;
  M2S_TRY(need(m, p + ".weight", ch, &g));
  M2S_TRY(need(m, p + ".bias", ch, &b));
  M2S_TRY(need(m, p + ".running_mean", ch, &mu));
  M2S_TRY(need(m, p + ".running_var", ch, &var));
  s->resize(ch);
  t->resize(ch);
  for (int c = 0; c < ch; ++c) {
    const double sc = static_cast<double>(g->data[c]) / std::sqrt(static_cast<double>(var->data[c]) + kBnEps);
    (*s)[c] = static_cast<float>(sc);
    (*t)[c] = static_cast<float>(b->data[c] - mu->data[c] * sc);
  }
  return M2S_OK;
}

// fp32 host values -> fp16 (round to nearest even) device array
int upload_half(const std::vector<float>& h, void** dev) {
  std::vector<uint16_t> v(h.size());
  for (size_t i = 0; i < h.size(); ++i) {
    const __half x = __float2half_rn(h[i]);
    std::memcpy(&v[i], &x, 2);
  }
  M2S_CUDA_OK(cudaMalloc(dev, v.size() * 2));
  M2S_CUDA_OK(cudaMemcpy(*dev, v.data(), v.size() * 2, cudaMemcpyHostToDevice));
  return M2S_OK;
}

int upload(const std::vector<float>& h, float** dev) {
  M2S_CUDA_OK(cudaMalloc(dev, h.size() * sizeof(float)));
  M2S_CUDA_OK(cudaMemcpy(*dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return M2S_OK;
}

struct GemmLayer {
  PackedWeights w;
  float* bias = nullptr;
  void* w16 = nullptr;   // 1x1 layers of the fp16 build: plain [n][c_in] fp16, K-major (TMA operand of the fused MBConv kernels)
  int taps = 1;
  int shift[M2S_MAX_TAPS] = {};
  int tap_ksteps = 0;                // per-tap K windows (pixel-pair layers of stage 0), see ConvProblem::tap_ksteps
  int kofs[M2S_MAX_TAPS] = {};
};

void free_gemm(GemmLayer* L) {
  free_weights(&L->w);
  if (L->bias) cudaFree(L->bias);
  if (L->w16) cudaFree(L->w16);
  L->bias = nullptr;
  L->w16 = nullptr;
}

// conv weight (cout, cin, k, k) + BN -> engine layer.  mode 0: k*k shifted taps (pitch given); mode 1: single tap
// with K = k*k*cin (im2col order (dy*3+dx)*cin + c); 1x1 convs are mode 0 with one tap.
int make_conv_layer(const TMap& m, const std::string& conv, const std::string& bn, int cout, int cin, int k,
                    int pitch, int mode, int pack, GemmLayer* L) {
  const HT* w;
  M2S_TRY(need(m, conv + ".weight", static_cast<size_t>(cout) * cin * k * k, &w));
  std::vector<float> s, t;
  M2S_TRY(bn_fold(m, bn, cout, &s, &t));
  const int kk = k * k;
  std::vector<float> e(static_cast<size_t>(kk) * cout * cin);
  for (int tap = 0; tap < kk; ++tap)
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c) {
        const float v = w->data[(static_cast<size_t>(o) * cin + c) * kk + tap] * s[o];
        if (mode == 0) e[(static_cast<size_t>(tap) * cout + o) * cin + c] = v;
        else e[static_cast<size_t>(o) * kk * cin + static_cast<size_t>(tap) * cin + c] = v;
      }
  if (mode == 0) {
    L->taps = kk;
    for (int tap = 0; tap < kk; ++tap) L->shift[tap] = (tap / k) * pitch + (tap % k);
    M2S_TRY(pack_weights(e.data(), kk, cout, cin, pack, &L->w));
    if (kk == 1 && pack == PACK_FP16) M2S_TRY(upload_half(e, &L->w16));
  } else {
    L->taps = 1;
    L->shift[0] = 0;
    M2S_TRY(pack_weights(e.data(), 1, cout, kk * cin, pack, &L->w));
  }
  return upload(t, &L->bias);
}

// 3x3 stride-1 conv over the zero-bordered layout with TWO consecutive pixels per GEMM row (N = 2 * cout): the convs of
// stage 0 have 16 output channels, and an M128 x N16 MMA fetches 4 KB of A for 32 K MACs -- the tensor core spends its time
// on operand fetches (SURVEY.md 8a-1 "stage 0/1 are large-M/small-N").  The GEMM row Q is pixels 2 Q, 2 Q + 1 side by side
// (2 * cin channels: one 128- or 64-byte row).  Output pixel o = 2 Q + p (p = 0, 1) reads input pixels
// o + (dy - 1) * pitch + (dx - 1) = 2 Q + t with t = (dy - 1) * pitch + s - 1, s = p + dx in 0..3: tap (dy, s) is row shift
// floor(t / 2) and the K window of pixel t mod 2, and its weight block holds W[dy][s - p] for the p whose dx = s - p is a
// real tap (zeros otherwise, and zeros outside the K window: 12 taps for 9, a quarter of the MACs are padding).
int make_pair_conv_layer(const TMap& m, const std::string& conv, const std::string& bn, int cout, int cin, int pitch,
                         int pack, GemmLayer* L) {
  const HT* w;
  M2S_TRY(need(m, conv + ".weight", static_cast<size_t>(cout) * cin * 9, &w));
  std::vector<float> s, t;
  M2S_TRY(bn_fold(m, bn, cout, &s, &t));
  const int taps = 12, n = 2 * cout, k2 = 2 * cin;
  std::vector<float> e(static_cast<size_t>(taps) * n * k2, 0.f);
  for (int dy = 0; dy < 3; ++dy)
    for (int sx = 0; sx < 4; ++sx) {
      const int tap = dy * 4 + sx;
      const int tt = (dy - 1) * pitch + sx - 1;
      const int plane = ((tt % 2) + 2) % 2;
      L->shift[tap] = (tt - plane) / 2;            // floor(tt / 2)
      L->kofs[tap] = plane * (cin / 16);           // K-steps of 16 channels
      for (int p = 0; p < 2; ++p) {
        const int dx = sx - p;
        if (dx < 0 || dx > 2) continue;
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c)
            e[(static_cast<size_t>(tap) * n + p * cout + o) * k2 + plane * cin + c] =
                w->data[(static_cast<size_t>(o) * cin + c) * 9 + dy * 3 + dx] * s[o];
      }
    }
  L->taps = taps;
  L->tap_ksteps = cin / 16;
  M2S_TRY(pack_weights(e.data(), taps, n, k2, pack, &L->w));
  std::vector<float> t2(n);
  for (int i = 0; i < n; ++i) t2[i] = t[i % cout];
  return upload(t2, &L->bias);
}

enum BlockKind { CN, ER, IR };
struct StageDef { BlockKind kind; int reps, stride, expand, cout, se; };
const StageDef kStages[6] = {{CN, 2, 1, 1, 16, 0},  {ER, 3, 2, 4, 32, 0},   {ER, 3, 2, 4, 56, 0},
                             {IR, 4, 2, 4, 104, 1}, {IR, 6, 1, 6, 120, 1}, {IR, 10, 2, 6, 208, 1}};
constexpr int kStem = 32;
constexpr int kFeat = 208;

struct Block {
  BlockKind kind;
  int cin, cout, mid, stride, rd;
  int hin, win;      // input spatial size
  bool in_padded;    // input activation carries a zero border
  bool out_padded;   // output activation carries a zero border
  bool skip;
  GemmLayer conv;    // CN conv / ER conv_exp / IR conv_pw
  GemmLayer pair;    // CN conv with two pixels per GEMM row (fp16 build; empty otherwise)
  GemmLayer s2d;     // stride-2 ER conv_exp over the space-to-depth input: 9 taps = K windows of the 4 parity planes (fp16 build)
  GemmLayer pwl;     // ER / IR projection
  float *dw_w = nullptr, *dw_b = nullptr;                                   // IR depthwise [9][mid], [mid]
  void* dw_w16 = nullptr;                                                   // ... and its fp16 copy (fp16 build)
  float *se_w1 = nullptr, *se_b1 = nullptr, *se_w2 = nullptr, *se_b2 = nullptr;
};

}  // namespace

struct m2s_acoustic {
  m2s_acoustic_config cfg;
  bool tf32 = true;   // tensor-core build (tf32 or fp16 operands); false = CUDA-core fp32 build
  bool fp16 = false;  // M2S_PREC_FP16: encoder GEMMs run kind::f16, operand-only activations live in HBM as fp16
  float *stem_w = nullptr, *stem_b = nullptr;
  std::vector<Block> blocks;
  GemmLayer inproj, head;
  float* w_hh[2] = {nullptr, nullptr};
  // fused block kernels (fp16 build): bit0 = InvertedResidual expand + depthwise + squeeze in one kernel, bit1 = SE scale
  // inside the project GEMM (csrc/mbconv_sm100.cu); bit2 = EdgeResidual 3x3 expand + 1x1 project in one kernel
  // (csrc/fused_er_sm100.cu); bit3 = that kernel also where the weights must be streamed per tile (slower); bit4 = stage 0's
  // 3x3 convs with two pixels per GEMM row (N = 32 instead of 16: per-tap K windows of the conv engine); bit5 = fp16
  // residual stream (no fp32 copies of the block outputs); bit6 = the stride-2 EdgeResidual blocks read a space-to-depth
  // copy of their input (K windows over the 4 parity planes) instead of an im2col matrix.  M2S_MBCONV=0 keeps the
  // unfused launches and the fp32 stream (the A/B reference of tests/).  bit7 = no copy at all: the fused kernel's TMA
  // loads gather the space-to-depth tile from the NHWC input (5-D box).
  int mbconv = 247;
  int chunk = 4096;  // frames per encoder pass (M2S_ENCODER_CHUNK): ~7 MB of work buffers per frame, 28 GB at 4096 (8192 is
                     // another 2 % faster, but bench.py's two builds side by side then peak at 156 of 180 GB).  Mid-round kernels: 14.1 / 12.6 / 12.1 / 12.2 us per frame at 512 / 1024 / 2048 / 4096; with the
                     // late-round kernels the tails weigh more: 11.0 / 10.5 / 10.3 at 1024 / 2048 / 4096 (encoder alone) and
                     // configs[2] end to end 2 907 / 2 968 / 3 023 audio-s/s at 2048 / 4096 / 8192 (fp16 build)
  // per-frame buffer sizes (floats)
  size_t x_floats = 0, e_floats = 0, e2_floats = 0, col_floats = 0;
  int max_mid = 0;
  int launches_encoder = 0;
};

namespace {

void free_block(Block* b) {
  free_gemm(&b->conv);
  free_gemm(&b->pair);
  free_gemm(&b->s2d);
  free_gemm(&b->pwl);
  for (float** p : {&b->dw_w, &b->dw_b, &b->se_w1, &b->se_b1, &b->se_w2, &b->se_b2}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (b->dw_w16) cudaFree(b->dw_w16);
  b->dw_w16 = nullptr;
}

size_t padded_rows(int h, int w) { return static_cast<size_t>(h + 2) * (w + 2); }

int run_gemm(const m2s_acoustic* m, const ConvProblem& p, const GemmLayer& L, cudaStream_t st, int tag = PROF_ENC_GEMM) {
  profile_set_tag(tag);
  return m->tf32 ? conv_tcgen05(p, L.w, st) : conv_simt(p, L.w.plain, st);
}

// `a` holds fp16 rows when the layer is packed for kind::f16 (L.w.half), fp32 rows otherwise.
ConvProblem gemm_problem(const void* a, long long a_batch_rows, int a_rows, int c_in, int batch, int l_out, float* d,
                         long long d_batch_rows, int d_ld, int d_row_offset, const GemmLayer& L) {
  ConvProblem p{};
  p.a = static_cast<const float*>(a); p.a_half = L.w.half != 0; p.a_batch_rows = a_batch_rows; p.a_rows = a_rows; p.a_ld = c_in; p.c_in = c_in;
  p.batch = batch; p.l_out = l_out; p.taps = L.taps;
  for (int j = 0; j < L.taps; ++j) p.shift[j] = L.shift[j];
  p.n = L.w.n; p.d = d; p.d_batch_rows = d_batch_rows; p.d_ld = d_ld; p.d_row_offset = d_row_offset;
  p.epi.bias = L.bias; p.epi.out_scale = 1.f; p.epi.res_inv_slope = 1.f;
  return p;
}

void set_pitch_mask(ConvProblem* p, int h, int w) {
  p->epi.mask_mode = M2S_MASK_PITCH;
  p->epi.pitch = w + 2; p->epi.i_lo = 1; p->epi.i_hi = h + 1; p->epi.j_lo = 1; p->epi.j_hi = w + 1;
}

struct EncBuffers {
  float *x0, *x1, *e, *e2, *col, *sums, *scales;
  float *x0h, *x1h;  // fp16 copies of the block outputs (fp16 build)
  float2* norm;  // per-frame (min, 1/range) of the uint8 ingest
  int32_t* fmap;
};

size_t align64(size_t f) { return (f + 63) / 64 * 64; }

struct AcWorkspace {
  size_t enc_floats, feats_floats, gin_floats, hcat_floats, fmap_ints;
  size_t total_bytes;
};

AcWorkspace plan_ws(const m2s_acoustic* m, int batch, int frames) {
  AcWorkspace w{};
  const size_t nf = static_cast<size_t>(batch) * frames;
  const size_t nc = nf < static_cast<size_t>(m->chunk) ? nf : m->chunk;
  w.enc_floats = align64(nc * m->x_floats) * 2 + align64(nc * m->e_floats) + align64(nc * m->e2_floats) +
                 align64(nc * m->col_floats) + 2 * align64(nc * m->max_mid) + align64(2 * nc) +
                 (m->fp16 ? 2 * align64(nc * m->x_floats / 2 + 64) : 0);
  w.feats_floats = align64(nf * kFeat);
  w.gin_floats = align64(nf * 8 * m->cfg.rnn_hidden);
  w.hcat_floats = align64(nf * 2 * m->cfg.rnn_hidden);
  w.fmap_ints = align64(nf);
  w.total_bytes = (w.enc_floats + w.feats_floats + w.gin_floats + w.hcat_floats + w.fmap_ints + 64) * 4 + 512;
  return w;
}

// One activation tensor of the encoder: the fp32 copy (residual source / GAP input) and/or the fp16 copy (tensor-core
// operand of the fp16 build); either may be absent (null).
struct Act {
  float* f32;
  void* f16;
};

// encoder over `n` frames (compact list).  src_map: compact index -> source frame (null = the frames are already the
// compact list); dst_map: compact index -> feature row (null = compact)
int encode_chunk(const m2s_acoustic* m, const void* frames, bool u8, const float* mask, const int32_t* src_map,
                 const int32_t* dst_map, int n, float* feats, int feat_ld, const EncBuffers& B, cudaStream_t st) {
  const int H = m->cfg.height, W = m->cfg.width;
  const bool h = m->fp16;
  const int esz = h ? 2 : 4;
  Act x{B.x0, B.x0h}, y{B.x1, B.x1h};
  // CUDA-core launches of the encoder, bracketed for the launch accounting (kernels = launches inside the call)
  auto simt = [&](int status_before, auto&& call, int kernels) -> int {
    (void)status_before;
    profile_set_tag(PROF_ENC_SIMT);
    M2S_TRY(profile_before(st));
    M2S_TRY(call());
    return profile_after(st, 0.0, kernels);
  };
  // operand view of an activation / of the scratch tensors e, e2, col (fp16 build: fp16 data in the same buffers)
  auto op = [&](const Act& a) { return h ? static_cast<const void*>(a.f16) : static_cast<const void*>(a.f32); };
  // a GEMM output that is only ever an operand (e, e2): fp16 in the fp16 build
  auto set_operand_out = [&](ConvProblem* p, float* buf) {
    if (h) { p->d = nullptr; p->d16 = buf; } else { p->d = buf; }
  };
  {  // the stem's only consumer is block 0.0 (no skip): operand copy only in the fp16 build
    void* out = h ? x.f16 : static_cast<void*>(x.f32);
    if (u8)
      M2S_TRY(simt(0, [&] { return enc_stem_u8(static_cast<const uint8_t*>(frames), src_map, mask, B.norm, out, h, m->stem_w, m->stem_b, n, H, W, st); }, 2));
    else
      M2S_TRY(simt(0, [&] { return enc_stem(static_cast<const float*>(frames), src_map, out, h, m->stem_w, m->stem_b, n, H, W, st); }, 1));
  }
  for (size_t bi = 0; bi < m->blocks.size(); ++bi) {
    const Block& b = m->blocks[bi];
    const bool last = bi + 1 == m->blocks.size();
    // block output copies: fp32 when the next block adds it as a shortcut (or the GAP reads it), fp16 when the next
    // consumer is a GEMM of the fp16 build
    // fp16 residual stream (mbconv bit 5, fp16 build): a shortcut is read from the fp16 copy the next GEMM reads anyway,
    // and no fp32 copy is written (26 % of the encoder's HBM traffic; the rounding it adds is below the operand rounding
    // already there: tools/emulate_residual_rounding.py)
    const bool res16 = h && (m->mbconv & 32);
    const bool out32 = !h || last || (m->blocks[bi + 1].skip && !res16);
    const bool out16 = h && !last;
    auto set_skip = [&](ConvProblem* p, int ld) {
      p->epi.res = res16 ? reinterpret_cast<const float*>(x.f16) : x.f32;
      p->epi.res_half = res16 ? 1 : 0;
      p->epi.res_ld = ld;
    };
    auto set_block_out = [&](ConvProblem* p) {
      p->d = out32 ? y.f32 : nullptr;
      p->d16 = out16 ? y.f16 : nullptr;
    };
    auto zero_border = [&](int rows, int ld, int head, int tail) -> int {
      if (out32) M2S_TRY(simt(0, [&] { return enc_zero_rows(y.f32, 4, n, rows, ld, head, tail, st); }, 1));
      if (out16) M2S_TRY(simt(0, [&] { return enc_zero_rows(y.f16, 2, n, rows, ld, head, tail, st); }, 1));
      return M2S_OK;
    };
    const int hin = b.hin, win = b.win;
    const int hout = hin / b.stride, wout = win / b.stride;
    if (b.kind == CN) {
      // 3x3 s1 conv on the padded layout -> padded layout, SiLU, (+ skip after the activation)
      const int rows = static_cast<int>(padded_rows(hin, win));
      if (b.pair.w.dev) {
        // two pixels per GEMM row over EVERY pair-row of the padded frame (no offset, no mask: the border pixels come
        // out as garbage and are zeroed below)
        const int prow = rows / 2;
        ConvProblem p = gemm_problem(op(x), prow, prow, 2 * b.cin, n, prow, y.f32, prow, 2 * b.cout, 0, b.pair);
        p.tap_ksteps = b.pair.tap_ksteps;
        for (int j = 0; j < b.pair.taps; ++j) p.kofs[j] = b.pair.kofs[j];
        set_block_out(&p);
        p.epi.act = M2S_ACT_SILU;
        if (b.skip) { set_skip(&p, 2 * b.cin); p.epi.res_after_act = 1; }
        M2S_TRY(run_gemm(m, p, b.pair, st));
        M2S_TRY(zero_border(rows, b.cout, win + 3, hin * (win + 2) + win + 3));
        if (out32) M2S_TRY(simt(0, [&] { return enc_zero_cols(y.f32, 4, n, rows, b.cout, win + 2, hin, st); }, 1));
        if (out16) M2S_TRY(simt(0, [&] { return enc_zero_cols(y.f16, 2, n, rows, b.cout, win + 2, hin, st); }, 1));
        std::swap(x, y);
        continue;
      }
      ConvProblem p = gemm_problem(op(x), rows, rows, b.cin, n, hin * (win + 2), y.f32, rows, b.cout, win + 3, b.conv);
      set_block_out(&p);
      p.epi.act = M2S_ACT_SILU;
      set_pitch_mask(&p, hin, win);
      if (b.skip) { set_skip(&p, b.cin); p.epi.res_after_act = 1; }
      M2S_TRY(run_gemm(m, p, b.conv, st));
      M2S_TRY(zero_border(rows, b.cout, win + 3, hin * (win + 2) + win + 3));
      std::swap(x, y);
    } else if (b.kind == ER) {
      const int rows_in = static_cast<int>(padded_rows(hin, win));
      const int rows_out = static_cast<int>(padded_rows(hout, wout));
      const int lq = hout * (wout + 2);  // rows in the (W+2)-pitch output space
      // expand: 9 row-shifted taps over the zero-bordered input (stride 1), or one tap over the im2col'd input (stride 2)
      ConvProblem p1;
      const GemmLayer* exp_layer = &b.conv;
      ErS2d s2d_geo{hin, win, b.cin, static_cast<long long>(rows_in)};
      const bool s2d_tma = (m->mbconv & 128) != 0;   // the fused kernel's TMA loads gather the space-to-depth tile themselves
      if (b.stride == 2 && b.s2d.w.dev) {
        // space-to-depth view of the input ((hout + 1) x (wout + 2) rows of 4 cin channels), then 9 K-window taps
        const int rows_s2d = (hout + 1) * (wout + 2);
        if (!s2d_tma) M2S_TRY(simt(0, [&] { return enc_s2d(op(x), B.col, esz, n, hin, win, b.cin, st); }, 1));
        p1 = gemm_problem(s2d_tma ? op(x) : static_cast<const void*>(B.col), rows_s2d, rows_s2d, 4 * b.cin, n, lq, B.e, lq, b.mid, 0, b.s2d);
        p1.tap_ksteps = b.s2d.tap_ksteps;
        for (int j = 0; j < b.s2d.taps; ++j) p1.kofs[j] = b.s2d.kofs[j];
        exp_layer = &b.s2d;
      } else if (b.stride == 2) {
        M2S_TRY(simt(0, [&] { return enc_im2col_s2(op(x), B.col, esz, n, hin, win, b.cin, st); }, 1));
        p1 = gemm_problem(B.col, lq, lq, 9 * b.cin, n, lq, B.e, lq, b.mid, 0, b.conv);
      } else {
        p1 = gemm_problem(op(x), rows_in, rows_in, b.cin, n, lq, B.e, lq, b.mid, 0, b.conv);
      }
      set_operand_out(&p1, B.e);
      p1.epi.act = M2S_ACT_SILU;
      ConvProblem p = gemm_problem(B.e, lq, lq, b.mid, n, lq, y.f32, rows_out, b.cout, wout + 3, b.pwl);
      set_block_out(&p);
      set_pitch_mask(&p, hout, wout);
      if (b.skip) set_skip(&p, b.cin);
      bool fused = false;
      if ((m->mbconv & 4) && fused_er_supported(p1, exp_layer->w, p, b.pwl.w) &&
          ((m->mbconv & 8) || fused_er_resident(exp_layer->w, b.pwl.w))) {
        // expand -> SiLU -> 1x1 project in one kernel: the expanded tile stays in SMEM
        profile_set_tag(PROF_ENC_GEMM);
        M2S_TRY(fused_er(p1, exp_layer->w, p, b.pwl.w, st, (exp_layer == &b.s2d && s2d_tma) ? &s2d_geo : nullptr));
        fused = true;
      } else if (exp_layer != &b.conv) {
        return fail(M2S_ERR_UNSUPPORTED, "space-to-depth EdgeResidual block needs the fused kernel");
      } else {
        M2S_TRY(run_gemm(m, p1, b.conv, st));
      }
      if (!fused) M2S_TRY(run_gemm(m, p, b.pwl, st));
      M2S_TRY(zero_border(rows_out, b.cout, wout + 3, hout * (wout + 2) + wout + 3));
      std::swap(x, y);
    } else {
      // IR: 1x1 expand (+SiLU) -> depthwise 3x3 (+SiLU, squeeze) -> SE -> 1x1 project (+ skip).
      // The 1x1 convs have no halo, so all frames of the chunk are flattened into ONE batch item: tiles then
      // span frame boundaries (a per-frame batch would waste 3/4 of every 256-row tile at 8x8 = 64 rows/frame).
      const int rows_in = b.in_padded ? static_cast<int>(padded_rows(hin, win)) : hin * win;
      const int hw = hout * wout;
      const bool fuse_e = (m->mbconv & 1) && b.stride == 1 && !b.in_padded && b.conv.w16 && b.dw_w16 &&
                          mb_expand_dw_supported(hin, win, b.cin, b.mid);
      const bool fuse_p = (m->mbconv & 2) && b.pwl.w16 && mb_project_supported(hw, b.mid, b.cout);
      if (fuse_e) {
        // expand GEMM with the depthwise conv as its epilogue: the expanded tensor stays in SMEM / TMEM
        profile_set_tag(PROF_ENC_GEMM);
        M2S_TRY(mb_expand_dw(x.f16, b.conv.w16, b.conv.bias, b.dw_w16, b.dw_w, b.dw_b, B.e2, B.sums, n, hin, win, b.cin, b.mid, st));
      } else {
        {
          const int rows = n * rows_in;
          ConvProblem p = gemm_problem(op(x), rows, rows, b.cin, 1, rows, B.e, rows, b.mid, 0, b.conv);
          set_operand_out(&p, B.e);
          p.epi.act = M2S_ACT_SILU;
          M2S_TRY(run_gemm(m, p, b.conv, st));
        }
        const int pitch_in = b.in_padded ? win + 2 : win;
        const int o = b.in_padded ? 1 : 0;
        M2S_TRY(simt(0, [&] { return enc_dwconv(B.e, B.e2, h, B.sums, b.dw_w, b.dw_b, n, b.mid, hin, win, pitch_in, o, o, rows_in, b.stride, st); }, 1));
      }
      if (fuse_p) {
        // SE MLP (sums -> scales, in place), then the project GEMM scales its A operand in SMEM
        M2S_TRY(simt(0, [&] { return enc_se_mlp(B.sums, b.se_w1, b.se_b1, b.se_w2, b.se_b2, n, b.mid, b.rd, hw, st); }, 1));
        profile_set_tag(PROF_ENC_GEMM);
        M2S_TRY(mb_project(B.e2, b.pwl.w16, B.sums, b.pwl.bias, (b.skip && !res16) ? x.f32 : nullptr,
                           (b.skip && res16) ? x.f16 : nullptr, out32 ? y.f32 : nullptr, out16 ? y.f16 : nullptr, n, hw, b.mid,
                           b.cout, st));
        std::swap(x, y);
        continue;
      }
      M2S_TRY(simt(0, [&] { return enc_se_apply(B.e2, h, B.sums, b.se_w1, b.se_b1, b.se_w2, b.se_b2, n, b.mid, b.rd, hw, st); }, 2));
      const int rows_out = n * hw;
      ConvProblem p = gemm_problem(B.e2, rows_out, rows_out, b.mid, 1, rows_out, y.f32, rows_out, b.cout, 0, b.pwl);
      set_block_out(&p);
      if (b.skip) set_skip(&p, b.cin);
      M2S_TRY(run_gemm(m, p, b.pwl, st));
      std::swap(x, y);
    }
  }
  const Block& lastb = m->blocks.back();
  const int hw = (lastb.hin / lastb.stride) * (lastb.win / lastb.stride);
  return simt(0, [&] { return enc_gap(x.f32, dst_map, feats, n, hw, kFeat, feat_ld, st); }, 1);
}

// src_map / dst_map: device tables over the compact frame list (see encode_chunk), either may be null
int encode_all(const m2s_acoustic* m, const void* frames, bool u8, const float* mask, const int32_t* src_map,
               const int32_t* dst_map, int n_frames, float* feats, float* enc_base, cudaStream_t st) {
  const int nc = n_frames < m->chunk ? n_frames : m->chunk;
  EncBuffers B{};
  float* p = enc_base;
  B.x0 = p; p += align64(static_cast<size_t>(nc) * m->x_floats);
  B.x1 = p; p += align64(static_cast<size_t>(nc) * m->x_floats);
  B.e = p; p += align64(static_cast<size_t>(nc) * m->e_floats);
  B.e2 = p; p += align64(static_cast<size_t>(nc) * m->e2_floats);
  B.col = p; p += align64(static_cast<size_t>(nc) * m->col_floats);
  B.sums = p; p += align64(static_cast<size_t>(nc) * m->max_mid);
  B.scales = p; p += align64(static_cast<size_t>(nc) * m->max_mid);
  B.norm = reinterpret_cast<float2*>(p); p += align64(2 * static_cast<size_t>(nc));
  B.x0h = B.x1h = nullptr;
  if (m->fp16) {
    B.x0h = p; p += align64(static_cast<size_t>(nc) * m->x_floats / 2 + 64);
    B.x1h = p;
  }
  for (int f0 = 0; f0 < n_frames; f0 += nc) {
    const int n = n_frames - f0 < nc ? n_frames - f0 : nc;
    const size_t fsz = static_cast<size_t>(m->cfg.height) * m->cfg.width * (u8 ? 1 : 4);
    const void* src = src_map ? frames : static_cast<const void*>(static_cast<const char*>(frames) + f0 * fsz);
    float* dst = dst_map ? feats : feats + static_cast<size_t>(f0) * kFeat;
    M2S_TRY(encode_chunk(m, src, u8, mask, src_map ? src_map + f0 : nullptr, dst_map ? dst_map + f0 : nullptr, n, dst,
                         kFeat, B, st));
  }
  return M2S_OK;
}

int rnn_head(const m2s_acoustic* m, const float* feats, int batch, int frames, const int32_t* lens,
             const int32_t* lens_host, float* mel_norm, float* gin, float* hcat, unsigned int* counters,
             cudaStream_t st) {
  const int Hd = m->cfg.rnn_hidden;
  const int rows = batch * frames;
  int max_len = frames;
  if (lens_host) {
    max_len = 0;
    for (int b = 0; b < batch; ++b) {
      if (lens_host[b] < 0 || lens_host[b] > frames) return fail(M2S_ERR_BAD_ARG, "lengths[%d]=%d out of range", b, lens_host[b]);
      max_len = lens_host[b] > max_len ? lens_host[b] : max_len;
    }
  }
  {  // input projection for every timestep and both directions: (rows, 208) x (208, 8H) + (b_ih + b_hh)
    ConvProblem p = gemm_problem(feats, rows, rows, kFeat, 1, rows, gin, rows, 8 * Hd, 0, m->inproj);
    M2S_TRY(run_gemm(m, p, m->inproj, st, PROF_RNN));
  }
  M2S_CUDA_OK(cudaMemsetAsync(hcat, 0, static_cast<size_t>(rows) * 2 * Hd * sizeof(float), st));
  profile_set_tag(PROF_RNN);
  M2S_TRY(profile_before(st));
  M2S_TRY(lstm_recurrence(gin, m->w_hh[0], m->w_hh[1], lens, hcat, counters, batch, frames, max_len, Hd, m->tf32, st));
  M2S_TRY(profile_after(st, 2.0 * 2 * 4 * Hd * static_cast<double>(Hd) * batch * max_len));
  {  // head on [h_fwd | h_bwd] with the weight duplicated: y = W (h_fwd + h_bwd) + b ; rows past lens -> 0
    ConvProblem p = gemm_problem(hcat, frames, frames, 2 * Hd, batch, frames, mel_norm, frames, m->cfg.n_mels, 0, m->head);
    if (lens) { p.epi.mask_mode = M2S_MASK_LEN; p.epi.lens = lens; p.epi.len_scale = 1; }
    M2S_TRY(run_gemm(m, p, m->head, st, PROF_RNN));
  }
  return M2S_OK;
}

}  // namespace

extern "C" int m2s_acoustic_create(const m2s_acoustic_config* cfg, const m2s_tensor* tensors, int32_t n_tensors,
                                   m2s_acoustic** out) {
  if (!cfg || !tensors || !out) return fail(M2S_ERR_BAD_ARG, "null argument");
  if (cfg->height % 32 || cfg->width % 32 || cfg->height <= 0 || cfg->width <= 0)
    return fail(M2S_ERR_UNSUPPORTED, "frame size %dx%d must be a positive multiple of 32", cfg->height, cfg->width);
  if (cfg->n_mels % 4) return fail(M2S_ERR_UNSUPPORTED, "n_mels must be a multiple of 4");
  M2S_TRY(m2s_device_check(-1));
  TMap tm;
  for (int i = 0; i < n_tensors; ++i) {
    HT h;
    h.data = tensors[i].data;
    for (int d = 0; d < tensors[i].ndim; ++d) h.shape.push_back(tensors[i].shape[d]);
    tm[tensors[i].name] = h;
  }
  auto* m = new m2s_acoustic();
  m->cfg = *cfg;
  m->tf32 = cfg->precision != M2S_PREC_FP32;
  m->fp16 = cfg->precision == M2S_PREC_FP16;
  // encoder GEMMs: fp16 operands in the fp16 build (every c_in of the topology is a multiple of 8); the BiLSTM input
  // projection and the head stay on tf32 (fp32 features / hidden states, 0.2 % of the FLOPs)
  const int enc_pack = !m->tf32 ? PACK_FP32 : (m->fp16 ? PACK_FP16 : PACK_TF32);
  if (const char* c = std::getenv("M2S_ENCODER_CHUNK")) m->chunk = std::max(1, std::atoi(c));
  if (const char* c = std::getenv("M2S_MBCONV")) m->mbconv = std::atoi(c);
  if (!m->fp16) m->mbconv = 0;
  int st = M2S_OK;
  auto bail = [&](int s) { m2s_acoustic_destroy(m); return s; };
  const std::string bb = "cnn.backbone.";

  {  // stem: fold the 3 identical input channels and BN
    const HT* w;
    if ((st = need(tm, bb + "conv_stem.weight", static_cast<size_t>(kStem) * 27, &w)) != M2S_OK) return bail(st);
    std::vector<float> s, t;
    if ((st = bn_fold(tm, bb + "bn1", kStem, &s, &t)) != M2S_OK) return bail(st);
    std::vector<float> e(9 * kStem);
    for (int o = 0; o < kStem; ++o)
      for (int tap = 0; tap < 9; ++tap) {
        float v = 0.f;
        for (int c = 0; c < 3; ++c) v += w->data[(static_cast<size_t>(o) * 3 + c) * 9 + tap];
        e[tap * kStem + o] = v * s[o];
      }
    if ((st = upload(e, &m->stem_w)) != M2S_OK || (st = upload(t, &m->stem_b)) != M2S_OK) return bail(st);
  }

  int cin = kStem, h = cfg->height / 2, w = cfg->width / 2;
  bool padded = true;  // the stem writes a padded layout
  m->x_floats = padded_rows(h, w) * kStem;
  int launches = 1;
  for (int s = 0; s < 6; ++s) {
    const StageDef& sd = kStages[s];
    for (int r = 0; r < sd.reps; ++r) {
      Block b{};
      b.kind = sd.kind; b.cin = cin; b.cout = sd.cout; b.stride = r == 0 ? sd.stride : 1;
      b.mid = cin * sd.expand; b.hin = h; b.win = w; b.in_padded = padded;
      b.skip = (b.stride == 1 && cin == sd.cout);
      b.rd = sd.se ? static_cast<int>(std::lround(cin * 0.25)) : 0;
      const std::string p = bb + "blocks." + std::to_string(s) + "." + std::to_string(r);
      const int ho = h / b.stride, wo = w / b.stride;
      if (sd.kind == CN) {
        b.out_padded = true;
        if ((st = make_conv_layer(tm, p + ".conv", p + ".bn1", b.cout, cin, 3, w + 2, 0, enc_pack, &b.conv)) != M2S_OK)
          return bail(st);
        if ((m->mbconv & 16) && (w + 2) % 2 == 0 && ((h + 2) * (w + 2)) % 2 == 0 && b.cout % 8 == 0 && cin % 16 == 0 &&
            (st = make_pair_conv_layer(tm, p + ".conv", p + ".bn1", b.cout, cin, w + 2, enc_pack, &b.pair)) != M2S_OK)
          return bail(st);
        m->x_floats = std::max(m->x_floats, padded_rows(ho, wo) * b.cout);
        launches += b.pair.w.dev ? 3 : 2;
      } else if (sd.kind == ER) {
        b.out_padded = true;
        if ((st = make_conv_layer(tm, p + ".conv_exp", p + ".bn1", b.mid, cin, 3, w + 2, b.stride == 2 ? 1 : 0,
                                  enc_pack, &b.conv)) != M2S_OK)
          return bail(st);
        if ((st = make_conv_layer(tm, p + ".conv_pwl", p + ".bn2", b.cout, b.mid, 1, 0, 0, enc_pack, &b.pwl)) != M2S_OK)
          return bail(st);
        if (b.stride == 2 && enc_pack == PACK_FP16 && (m->mbconv & 64) && (m->mbconv & 4) && cin % 16 == 0 && cin <= 32) {
          // the same conv over the space-to-depth input (enc_s2d): tap (dy, dx) = plane (dy & 1, dx & 1) as a K window of
          // the 4 cin-channel row, at row shift (dy >> 1) * (wo + 2) + (dx >> 1)
          if ((st = make_conv_layer(tm, p + ".conv_exp", p + ".bn1", b.mid, cin, 3, wo + 2, 0, enc_pack, &b.s2d)) != M2S_OK)
            return bail(st);
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            b.s2d.shift[tap] = (dy >> 1) * (wo + 2) + (dx >> 1);
            b.s2d.kofs[tap] = (2 * (dy & 1) + (dx & 1)) * (cin / 16);
          }
          b.s2d.tap_ksteps = cin / 16;
        }
        const size_t lq = static_cast<size_t>(ho) * (wo + 2);
        m->e_floats = std::max(m->e_floats, lq * b.mid);
        if (b.stride == 2) m->col_floats = std::max(m->col_floats, lq * 9 * cin);
        m->x_floats = std::max(m->x_floats, padded_rows(ho, wo) * b.cout);
        launches += (b.stride == 2 ? 4 : 3) - (((m->mbconv & 4) && ((m->mbconv & 8) || b.mid * 9 * cin * 2 <= 90 * 1024)) ? 1 : 0)   // (im2col,) expand, project (one fused kernel), border rows
                    - ((b.s2d.w.dev && (m->mbconv & 128)) ? 1 : 0);   // no im2col / space-to-depth pass
      } else {
        b.out_padded = false;
        if ((st = make_conv_layer(tm, p + ".conv_pw", p + ".bn1", b.mid, cin, 1, 0, 0, enc_pack, &b.conv)) != M2S_OK)
          return bail(st);
        if ((st = make_conv_layer(tm, p + ".conv_pwl", p + ".bn3", b.cout, b.mid, 1, 0, 0, enc_pack, &b.pwl)) != M2S_OK)
          return bail(st);
        {  // depthwise (mid,1,3,3) + bn2 -> [9][mid]
          const HT* dw;
          if ((st = need(tm, p + ".conv_dw.weight", static_cast<size_t>(b.mid) * 9, &dw)) != M2S_OK) return bail(st);
          std::vector<float> s2, t2;
          if ((st = bn_fold(tm, p + ".bn2", b.mid, &s2, &t2)) != M2S_OK) return bail(st);
          std::vector<float> e(static_cast<size_t>(9) * b.mid);
          for (int c = 0; c < b.mid; ++c)
            for (int tap = 0; tap < 9; ++tap) e[static_cast<size_t>(tap) * b.mid + c] = dw->data[c * 9 + tap] * s2[c];
          if ((st = upload(e, &b.dw_w)) != M2S_OK || (st = upload(t2, &b.dw_b)) != M2S_OK) return bail(st);
          if (m->fp16 && (st = upload_half(e, &b.dw_w16)) != M2S_OK) return bail(st);
        }
        {
          const HT *w1, *b1, *w2, *b2;
          if ((st = need(tm, p + ".se.conv_reduce.weight", static_cast<size_t>(b.rd) * b.mid, &w1)) != M2S_OK ||
              (st = need(tm, p + ".se.conv_reduce.bias", b.rd, &b1)) != M2S_OK ||
              (st = need(tm, p + ".se.conv_expand.weight", static_cast<size_t>(b.mid) * b.rd, &w2)) != M2S_OK ||
              (st = need(tm, p + ".se.conv_expand.bias", b.mid, &b2)) != M2S_OK)
            return bail(st);
          auto up = [&](const HT* t, float** d) { return upload(std::vector<float>(t->data, t->data + t->numel()), d); };
          std::vector<float> w2t(static_cast<size_t>(b.rd) * b.mid);  // (mid, rd) -> (rd, mid): coalesced over channels
          for (int c = 0; c < b.mid; ++c)
            for (int j = 0; j < b.rd; ++j) w2t[static_cast<size_t>(j) * b.mid + c] = w2->data[static_cast<size_t>(c) * b.rd + j];
          if ((st = up(w1, &b.se_w1)) != M2S_OK || (st = up(b1, &b.se_b1)) != M2S_OK ||
              (st = upload(w2t, &b.se_w2)) != M2S_OK || (st = up(b2, &b.se_b2)) != M2S_OK)
            return bail(st);
        }
        const size_t rows_in = padded ? padded_rows(h, w) : static_cast<size_t>(h) * w;
        m->e_floats = std::max(m->e_floats, rows_in * b.mid);
        m->e2_floats = std::max(m->e2_floats, static_cast<size_t>(ho) * wo * b.mid);
        m->x_floats = std::max(m->x_floats, static_cast<size_t>(ho) * wo * b.cout);
        m->max_mid = std::max(m->max_mid, b.mid);
        // expand, depthwise, SE MLP, SE scale, project; the fused kernels fold depthwise into expand and scale into project
        const bool fe = (m->mbconv & 1) && b.stride == 1 && !padded && mb_expand_dw_supported(h, w, cin, b.mid);
        const bool fp = (m->mbconv & 2) && mb_project_supported(ho * wo, b.mid, b.cout);
        launches += 5 - (fe ? 1 : 0) - (fp ? 1 : 0);
      }
      m->blocks.push_back(b);
      padded = b.out_padded;
      cin = sd.cout; h = ho; w = wo;
    }
  }
  m->launches_encoder = launches + 1;  // + GAP
  if (cin != kFeat) return bail(fail(M2S_ERR_BAD_ARG, "unexpected feature width %d", cin));

  {  // LSTM input projection (both directions) + recurrent weights + head
    const int Hd = cfg->rnn_hidden;
    const std::string l = "rnn.lstm.";
    const HT *wi[2], *wh[2], *bi[2], *bh[2];
    const char* sfx[2] = {"", "_reverse"};
    for (int d = 0; d < 2; ++d) {
      if ((st = need(tm, l + "weight_ih_l0" + sfx[d], static_cast<size_t>(4) * Hd * kFeat, &wi[d])) != M2S_OK ||
          (st = need(tm, l + "weight_hh_l0" + sfx[d], static_cast<size_t>(4) * Hd * Hd, &wh[d])) != M2S_OK ||
          (st = need(tm, l + "bias_ih_l0" + sfx[d], static_cast<size_t>(4) * Hd, &bi[d])) != M2S_OK ||
          (st = need(tm, l + "bias_hh_l0" + sfx[d], static_cast<size_t>(4) * Hd, &bh[d])) != M2S_OK)
        return bail(st);
    }
    std::vector<float> wcat(static_cast<size_t>(8) * Hd * kFeat), bcat(static_cast<size_t>(8) * Hd);
    for (int d = 0; d < 2; ++d) {
      std::memcpy(wcat.data() + static_cast<size_t>(d) * 4 * Hd * kFeat, wi[d]->data, sizeof(float) * 4 * Hd * kFeat);
      for (int i = 0; i < 4 * Hd; ++i) bcat[d * 4 * Hd + i] = bi[d]->data[i] + bh[d]->data[i];
      if ((st = upload(std::vector<float>(wh[d]->data, wh[d]->data + wh[d]->numel()), &m->w_hh[d])) != M2S_OK)
        return bail(st);
    }
    if ((st = pack_weights(wcat.data(), 1, 8 * Hd, kFeat, m->tf32, &m->inproj.w)) != M2S_OK ||
        (st = upload(bcat, &m->inproj.bias)) != M2S_OK)
      return bail(st);
    const HT *hw, *hb;
    if ((st = need(tm, "head.weight", static_cast<size_t>(cfg->n_mels) * Hd, &hw)) != M2S_OK ||
        (st = need(tm, "head.bias", cfg->n_mels, &hb)) != M2S_OK)
      return bail(st);
    std::vector<float> hcat(static_cast<size_t>(cfg->n_mels) * 2 * Hd);
    for (int o = 0; o < cfg->n_mels; ++o)
      for (int k = 0; k < Hd; ++k) hcat[static_cast<size_t>(o) * 2 * Hd + k] = hcat[static_cast<size_t>(o) * 2 * Hd + Hd + k] = hw->data[o * Hd + k];
    if ((st = pack_weights(hcat.data(), 1, cfg->n_mels, 2 * Hd, m->tf32, &m->head.w)) != M2S_OK ||
        (st = upload(std::vector<float>(hb->data, hb->data + cfg->n_mels), &m->head.bias)) != M2S_OK)
      return bail(st);
  }
  *out = m;
  return M2S_OK;
}

extern "C" void m2s_acoustic_destroy(m2s_acoustic* m) {
  if (!m) return;
  if (m->stem_w) cudaFree(m->stem_w);
  if (m->stem_b) cudaFree(m->stem_b);
  for (auto& b : m->blocks) free_block(&b);
  free_gemm(&m->inproj);
  free_gemm(&m->head);
  for (int d = 0; d < 2; ++d)
    if (m->w_hh[d]) cudaFree(m->w_hh[d]);
  delete m;
}

extern "C" int m2s_acoustic_launches(const m2s_acoustic* m) { return m ? m->launches_encoder + 5 : 0; }

extern "C" size_t m2s_acoustic_workspace_bytes(const m2s_acoustic* m, int32_t batch, int32_t frames) {
  if (!m || batch <= 0 || frames <= 0) return 0;
  return plan_ws(m, batch, frames).total_bytes;
}

namespace {
struct WsPtrs {
  float *enc, *feats, *gin, *hcat;
  int32_t* fmap;
  unsigned int* counters;
};
int carve(const m2s_acoustic* m, int batch, int frames, void* ws, size_t ws_bytes, WsPtrs* o) {
  const AcWorkspace w = plan_ws(m, batch, frames);
  if (!ws || ws_bytes < w.total_bytes)
    return fail(M2S_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total_bytes, ws_bytes);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  o->enc = base;
  o->feats = o->enc + w.enc_floats;
  o->gin = o->feats + w.feats_floats;
  o->hcat = o->gin + w.gin_floats;
  o->fmap = reinterpret_cast<int32_t*>(o->hcat + w.hcat_floats);
  o->counters = reinterpret_cast<unsigned int*>(o->fmap + w.fmap_ints);
  return M2S_OK;
}
}  // namespace

extern "C" int m2s_acoustic_encode(m2s_acoustic* m, const float* frames_dev, int32_t n_frames, float* feats,
                                   void* workspace, size_t workspace_bytes, m2s_stream_t stream) {
  if (!m || !frames_dev || !feats) return fail(M2S_ERR_BAD_ARG, "null argument");
  if (n_frames <= 0) return M2S_OK;
  WsPtrs w;
  M2S_TRY(carve(m, 1, n_frames, workspace, workspace_bytes, &w));
  return encode_all(m, frames_dev, false, nullptr, nullptr, nullptr, n_frames, feats, w.enc,
                    reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int m2s_acoustic_rnn_head(m2s_acoustic* m, const float* feats, int32_t batch, int32_t frames,
                                     const int32_t* lengths, const int32_t* lengths_host, float* mel_norm,
                                     void* workspace, size_t workspace_bytes, m2s_stream_t stream) {
  if (!m || !feats || !mel_norm) return fail(M2S_ERR_BAD_ARG, "null argument");
  if ((lengths == nullptr) != (lengths_host == nullptr))
    return fail(M2S_ERR_BAD_ARG, "lengths and lengths_host must be given together");
  if (batch <= 0 || frames <= 0) return M2S_OK;
  WsPtrs w;
  M2S_TRY(carve(m, batch, frames, workspace, workspace_bytes, &w));
  return rnn_head(m, feats, batch, frames, lengths, lengths_host, mel_norm, w.gin, w.hcat, w.counters,
                  reinterpret_cast<cudaStream_t>(stream));
}

namespace {
// `packed`: frames_dev holds only the valid frames, clip after clip (sum of lengths frames); otherwise the padded
// (batch, frames) layout.  Ragged batches encode the compact list of valid frames either way; the table that scatters
// the features into the padded (batch, frames) layout of the recurrence is built on the device from `lengths`.
int acoustic_forward_impl(m2s_acoustic* m, const void* frames_dev, bool u8, bool packed, const float* mask, int32_t batch,
                          int32_t frames, const int32_t* lengths, const int32_t* lengths_host, float* mel_norm,
                          void* workspace, size_t workspace_bytes, m2s_stream_t stream) {
  if (!m || !frames_dev || !mel_norm) return fail(M2S_ERR_BAD_ARG, "null argument");
  if ((lengths == nullptr) != (lengths_host == nullptr))
    return fail(M2S_ERR_BAD_ARG, "lengths and lengths_host must be given together");
  if (packed && !lengths) return fail(M2S_ERR_BAD_ARG, "a packed batch needs lengths");
  if (batch <= 0 || frames <= 0) return M2S_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WsPtrs w;
  M2S_TRY(carve(m, batch, frames, workspace, workspace_bytes, &w));
  const int total = batch * frames;
  if (lengths_host) {
    long long valid = 0;
    for (int b = 0; b < batch; ++b) {
      if (lengths_host[b] < 0 || lengths_host[b] > frames)
        return fail(M2S_ERR_BAD_ARG, "lengths[%d]=%d out of range", b, lengths_host[b]);
      valid += lengths_host[b];
    }
    M2S_CUDA_OK(cudaMemsetAsync(w.feats, 0, static_cast<size_t>(total) * kFeat * sizeof(float), st));
    if (valid > 0) {
      profile_set_tag(PROF_ENC_SIMT);
      M2S_TRY(profile_before(st));
      M2S_TRY(enc_build_fmap(lengths, batch, frames, w.fmap, st));
      M2S_TRY(profile_after(st, 0.0));
      M2S_TRY(encode_all(m, frames_dev, u8, mask, packed ? nullptr : w.fmap, w.fmap, static_cast<int>(valid), w.feats,
                         w.enc, st));
    }
  } else {
    M2S_TRY(encode_all(m, frames_dev, u8, mask, nullptr, nullptr, total, w.feats, w.enc, st));
  }
  return rnn_head(m, w.feats, batch, frames, lengths, lengths_host, mel_norm, w.gin, w.hcat, w.counters, st);
}
}  // namespace

extern "C" int m2s_acoustic_encode_packed(m2s_acoustic* m, const void* frames_dev, int32_t frames_are_u8, const float* mask,
                                          int32_t batch, int32_t max_frames, const int32_t* lengths,
                                          const int32_t* lengths_host, float* feats, void* workspace,
                                          size_t workspace_bytes, m2s_stream_t stream) {
  if (!m || !frames_dev || !feats || !lengths || !lengths_host) return fail(M2S_ERR_BAD_ARG, "null argument");
  if (mask && !frames_are_u8) return fail(M2S_ERR_BAD_ARG, "the articulator mask applies to raw uint8 frames");
  if (batch <= 0 || max_frames <= 0) return M2S_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WsPtrs w;
  M2S_TRY(carve(m, batch, max_frames, workspace, workspace_bytes, &w));
  long long valid = 0;
  for (int b = 0; b < batch; ++b) {
    if (lengths_host[b] < 0 || lengths_host[b] > max_frames)
      return fail(M2S_ERR_BAD_ARG, "lengths[%d]=%d out of range", b, lengths_host[b]);
    valid += lengths_host[b];
  }
  M2S_CUDA_OK(cudaMemsetAsync(feats, 0, static_cast<size_t>(batch) * max_frames * kFeat * sizeof(float), st));
  if (valid == 0) return M2S_OK;
  profile_set_tag(PROF_ENC_SIMT);
  M2S_TRY(profile_before(st));
  M2S_TRY(enc_build_fmap(lengths, batch, max_frames, w.fmap, st));
  M2S_TRY(profile_after(st, 0.0));
  return encode_all(m, frames_dev, frames_are_u8 != 0, mask, nullptr, w.fmap, static_cast<int>(valid), feats, w.enc, st);
}

extern "C" int m2s_acoustic_forward(m2s_acoustic* m, const float* frames_dev, int32_t batch, int32_t frames,
                                    const int32_t* lengths, const int32_t* lengths_host, float* mel_norm,
                                    void* workspace, size_t workspace_bytes, m2s_stream_t stream) {
  return acoustic_forward_impl(m, frames_dev, false, false, nullptr, batch, frames, lengths, lengths_host, mel_norm,
                               workspace, workspace_bytes, stream);
}

extern "C" int m2s_acoustic_forward_u8(m2s_acoustic* m, const uint8_t* frames_dev, const float* mask, int32_t batch,
                                       int32_t frames, const int32_t* lengths, const int32_t* lengths_host,
                                       float* mel_norm, void* workspace, size_t workspace_bytes, m2s_stream_t stream) {
  return acoustic_forward_impl(m, frames_dev, true, false, mask, batch, frames, lengths, lengths_host, mel_norm, workspace,
                               workspace_bytes, stream);
}

extern "C" int m2s_acoustic_forward_packed(m2s_acoustic* m, const void* frames_dev, int32_t frames_are_u8, const float* mask,
                                           int32_t batch, int32_t max_frames, const int32_t* lengths,
                                           const int32_t* lengths_host, float* mel_norm, void* workspace,
                                           size_t workspace_bytes, m2s_stream_t stream) {
  if (mask && !frames_are_u8) return fail(M2S_ERR_BAD_ARG, "the articulator mask applies to raw uint8 frames");
  return acoustic_forward_impl(m, frames_dev, frames_are_u8 != 0, true, mask, batch, max_frames, lengths, lengths_host,
                               mel_norm, workspace, workspace_bytes, stream);
}

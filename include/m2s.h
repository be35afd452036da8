/*
 * m2s.h -- C ABI of libm2s.so, the sm_100a (B200) implementation of the
 * rtMRI-video -> mel -> waveform inference path of YamaneKoyo/mri-to-speech.
 *
 * The reference is pure Python/PyTorch and has NO plugin / FFI layer (SURVEY.md
 * section 1): its boundary for this path is the nn.Module API
 *   models.Generator.forward                      (reference models.py:113-131)
 *   OTNLikeCNNBiLSTM.forward                      (reference mri2speech_code/mri_acoustic_model.py:116-136)
 *   the mel glue in scripts/run_mri_video_inference.py:160-163,227-239
 * This header is the binding a maintainer puts UNDER those modules (ctypes stub in
 * INTEGRATION.md; our own host-side mirror lives in mri2speech_b200/).  Every
 * entry point takes plain pointers and sizes -- no torch types.
 *
 * Conventions
 *   - all device pointers are on the current CUDA device, float32 unless noted;
 *   - no allocation and no synchronisation inside *_forward: the caller owns
 *     inputs, outputs and the workspace (size from *_workspace_bytes);
 *   - every call returns M2S_OK (0) or a negative m2s_status; the message of the
 *     last failure on the calling thread is m2s_last_error_string();
 *   - calls on distinct streams / distinct handles are thread-safe (one-time kernel attributes and the SM count are
 *     kept per device under a mutex); the m2s_debug_* probes are process-global and single-threaded;
 *   - anything but an sm_100 device is refused (there is no CPU fallback).
 */
#ifndef M2S_H_
#define M2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* m2s_stream_t; /* == cudaStream_t */

typedef enum {
  M2S_OK = 0,
  M2S_ERR_BAD_ARG = -1,
  M2S_ERR_UNSUPPORTED = -2,
  M2S_ERR_CUDA = -3,
  M2S_ERR_DEVICE = -4,
  M2S_ERR_WORKSPACE = -5,
  M2S_ERR_MISSING_TENSOR = -6
} m2s_status;

/* arithmetic of the GEMM-shaped layers */
enum {
  M2S_PREC_TF32 = 0, /* tcgen05 kind::tf32, fp32 accumulate in TMEM                  */
  M2S_PREC_FP32 = 1, /* CUDA-core fp32 FMA kernels (exact-fp32 build, slow)          */
  M2S_PREC_FP16 = 2  /* tcgen05 kind::f16: fp16 operands (10-bit mantissa, the same as tf32), fp32 accumulate,
                        fp32 residual / MRF streams.  A layer runs kind::f16 iff its c_in is a multiple of 8
                        (16-byte rows of halves); conv_pre (fp32 mel operand), the BiLSTM input projection
                        and the mel head stay on tf32.  The Python modules' default since round 2; the encoder
                        then also keeps its residual stream in fp16 */
};

const char* m2s_version(void);
const char* m2s_last_error_string(void);
/* 0 if `device` is an sm_100 part, M2S_ERR_DEVICE otherwise. */
int m2s_device_check(int device);

/* A named host tensor, exactly one state_dict entry of the reference checkpoints
 * (keys as listed in SURVEY.md 8a-1/8a-3/8a-5). */
typedef struct {
  const char* name;
  const float* data; /* host, contiguous, row-major */
  int32_t ndim;
  int64_t shape[4];
} m2s_tensor;

/* ------------------------------------------------------------------------- */
/* Generic multi-tap implicit-GEMM convolution (the engine every GEMM-shaped  */
/* layer of the path runs on).  Exposed for parity tests of the kernel itself.*/
/*   D[b, q + d_row_offset, n] = epi( sum_j sum_c A[b, q + shift[j], c] * W[j][n][c] ) */
/* Rows of A outside [0, a_rows) read as zero.  Activations are channels-last. */
/* ------------------------------------------------------------------------- */
enum { M2S_ACT_NONE = 0, M2S_ACT_LRELU = 1, M2S_ACT_SILU = 2 };
enum { M2S_MASK_NONE = 0, M2S_MASK_LEN = 1, M2S_MASK_PITCH = 2 };
enum { M2S_IMPL_TCGEN05 = 0, M2S_IMPL_SIMT = 1 };
#define M2S_MAX_TAPS 16

typedef struct {
  /* A operand */
  const float* a;
  int64_t a_batch_rows; /* rows between consecutive batch items            */
  int32_t a_rows;       /* valid rows per batch item (zero beyond)         */
  int32_t a_ld;         /* elements per row (multiple of 4)                */
  int32_t c_in;         /* channels contracted per tap (multiple of 4)     */
  int32_t batch;
  int32_t l_out;        /* rows computed per batch item                    */
  int32_t taps;
  int32_t shift[M2S_MAX_TAPS];
  /* weights: plain host-order device array [taps][n][c_in] */
  const float* w;
  int32_t n;
  /* output */
  float* d;
  int64_t d_batch_rows;
  int32_t d_ld;
  int32_t d_row_offset;
  /* fused epilogue: v = acc + bias[n] + inv_lrelu(res) + accum; v *= out_scale; v = act(v); mask
   * (res_after_act != 0: the residual is added after the activation instead: v = act(..) + res) */
  const float* bias;     /* [n] or NULL */
  const float* res;      /* indexed like d (same rows), NULL = none */
  int32_t res_ld;
  float res_inv_slope;   /* res >= 0 ? res : res*res_inv_slope  (1 = plain residual) */
  int32_t res_after_act;
  const float* accum;    /* indexed like d, NULL = none */
  int32_t accum_ld;
  float out_scale;
  int32_t act;
  float act_slope;
  int32_t mask_mode;
  const int32_t* lens;   /* M2S_MASK_LEN: row valid iff (q + d_row_offset) < lens[b]*len_scale */
  int32_t len_scale;
  int32_t pitch, i_lo, i_hi, j_lo, j_hi; /* M2S_MASK_PITCH: (i,j)=divmod(row,pitch) inside the box */
  /* fp16 operands (M2S_IMPL_TCGEN05 only): a holds __half rows (a_ld, c_in in elements, multiples of 8) and the
   * weights are rounded to fp16; d16 = optional second output in fp16, indexed like d (d may then be NULL) */
  int32_t a_half;
  void* d16;
  /* split-fp16 residual stream (fp16 build of the vocoder): a tensor that is both a tensor-core operand and a
   * residual source is stored as TWO fp16 planes, hi = fp16(v) (the operand, == d16) and lo = fp16((v - hi) * 2048),
   * instead of an fp32 copy next to the fp16 operand copy: hi + lo / 2048 carries v to ~22 mantissa bits in 4 instead
   * of 6 bytes (the scale keeps the remainder of small activations out of fp16's subnormal range).
   * d16_lo: optional lo-plane output (needs d16; d is normally NULL then).  res_hi / res_lo: the residual read as
   * float(hi) + float(lo) / 2048 (then res must be NULL); both indexed like res (res_ld elements per row). */
  void* d16_lo;
  const void* res_hi;
  const void* res_lo;
} m2s_conv_args;

int m2s_conv_fwd(const m2s_conv_args* args, int impl, m2s_stream_t stream);
/* Fused ResBlock1 pair (reference models.py:36-48: xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x) in one kernel:
 * conv1 (args1: a = fp16 x, w, causal shifts, bias, leaky-ReLU) feeds conv2 (args2: w, shifts -(k-1)..0, bias, fused
 * epilogue and outputs; args2->a is ignored) through an fp16 tile that stays in shared memory.  fp16 operands only,
 * c_in == n in {32, 64, 128}.  Exposed for parity tests of the kernel itself. */
int m2s_resblock_pair_fwd(const m2s_conv_args* args1, const m2s_conv_args* args2, m2s_stream_t stream);

/* ------------------------------------------------------------------------- */
/* HiFi-GAN Generator (reference models.py:88-131, config_custom.json)        */
/* ------------------------------------------------------------------------- */
#define M2S_MAX_UPS 8
#define M2S_MAX_RBK 8
typedef struct {
  int32_t num_mels;                 /* h.num_mels                      */
  int32_t upsample_initial_channel; /* h.upsample_initial_channel      */
  int32_t num_upsamples;
  int32_t upsample_rates[M2S_MAX_UPS];
  int32_t upsample_kernel_sizes[M2S_MAX_UPS];
  int32_t num_kernels;
  int32_t resblock_kernel_sizes[M2S_MAX_RBK];
  int32_t resblock_dilations[M2S_MAX_RBK][3];
  int32_t precision;                /* M2S_PREC_*                      */
  int32_t resblock;                 /* h.resblock: 0 or 1 = ResBlock1 (3 pairs of convs1 / convs2, 3 dilations),
                                     * 2 = ResBlock2 (models.py:58-85: 2 convs, resblock_dilations[j][0..1])      */
} m2s_generator_config;

typedef struct m2s_generator m2s_generator;

/* Tensors: the Generator state_dict (233 entries for config_custom.json), either
 * weight_g/weight_v pairs or plain .weight (weight-norm removed).  ResBlock2 configs name their convs
 * resblocks.N.convs.M instead of resblocks.N.convs1.M / convs2.M. */
int m2s_generator_create(const m2s_generator_config* cfg, const m2s_tensor* tensors, int32_t n_tensors,
                         m2s_generator** out);
void m2s_generator_destroy(m2s_generator* g);
size_t m2s_generator_workspace_bytes(const m2s_generator* g, int32_t batch, int32_t frames);
/* mel: device (batch, num_mels, frames) -- the layout Generator.forward takes.
 * lengths: device int32[batch] valid mel frames per utterance, or NULL (all = frames).
 * audio:  device (batch, 1, frames * prod(upsample_rates)). */
int m2s_generator_forward(m2s_generator* g, const float* mel, int32_t batch, int32_t frames,
                          const int32_t* lengths, float* audio, void* workspace, size_t workspace_bytes,
                          m2s_stream_t stream);
/* Same forward, fed with conv_pre's operand directly: mel_btc = device (batch, frames, num_mels), channels-last, rows
 * past lengths[b] zero -- exactly the mel_log output of m2s_mel_glue.  This is how the de-normalisation of
 * scripts/run_mri_video_inference.py:160-163,232-239 is fused into the vocoder's conv_pre load: the glue kernel writes
 * the tensor conv_pre's TMA reads, no (B, n_mels, T) tensor and no layout pass in between. */
int m2s_generator_forward_btc(m2s_generator* g, const float* mel_btc, int32_t batch, int32_t frames,
                              const int32_t* lengths, float* audio, void* workspace, size_t workspace_bytes,
                              m2s_stream_t stream);
/* Number of kernels one forward launches (for bench.py's gpu_launches). */
int m2s_generator_launches(const m2s_generator* g);

/* ------------------------------------------------------------------------- */
/* Acoustic model: frame-CNN encoder + BiLSTM (sum merge) + mel head           */
/* (reference mri2speech_code/mri_acoustic_model.py:20-136)                    */
/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t n_mels;
  int32_t rnn_hidden;
  int32_t height, width; /* frame size, 256 x 256 */
  int32_t precision;
} m2s_acoustic_config;

typedef struct m2s_acoustic m2s_acoustic;

int m2s_acoustic_create(const m2s_acoustic_config* cfg, const m2s_tensor* tensors, int32_t n_tensors,
                        m2s_acoustic** out);
void m2s_acoustic_destroy(m2s_acoustic* m);
size_t m2s_acoustic_workspace_bytes(const m2s_acoustic* m, int32_t batch, int32_t frames);
/* frames: device (batch, frames, height, width) float32 in [0,1];
 * lengths: device int32[batch] or NULL; lengths_host: the same values on the host
 * (they size the launches, so that nothing has to be read back; NULL iff lengths is NULL);
 * mel_norm: device (batch, frames, n_mels) normalised mel (rows past lengths[b] are zero). */
int m2s_acoustic_forward(m2s_acoustic* m, const float* frames_dev, int32_t batch, int32_t frames,
                         const int32_t* lengths, const int32_t* lengths_host, float* mel_norm,
                         void* workspace, size_t workspace_bytes, m2s_stream_t stream);
/* Fused ingest: frames are the raw uint8 gray frames (batch, frames, height, width).
 * Replaces the host-side chain of the reference CLI: scripts/run_mri_video_inference.py:34-53
 * (_preprocess_frame: z-score then min-max == per-frame min-max; a constant frame maps to zeros) and, when
 * mask != NULL, scripts/mask_rtmri_video.py:96-98 first (masked = uint8(clip(frame * mask, 0, 255)),
 * truncating).  mask: device float32 (height, width) shared by the batch, or NULL.  The min-max pass and
 * the stem conv read the uint8 frames directly: the float32 frames never exist in HBM. */
int m2s_acoustic_forward_u8(m2s_acoustic* m, const uint8_t* frames_dev, const float* mask, int32_t batch,
                            int32_t frames, const int32_t* lengths, const int32_t* lengths_host,
                            float* mel_norm, void* workspace, size_t workspace_bytes, m2s_stream_t stream);
/* Ragged batch WITHOUT padding on the input side: frames_dev holds only the valid frames, clip after clip --
 * sum(lengths) frames of height x width, float32 in [0,1] (frames_are_u8 == 0) or raw uint8 (frames_are_u8 != 0, with
 * the optional mask, as m2s_acoustic_forward_u8).  This is how a batch runner feeds clips of 150-600 frames without
 * zero-filling and copying a padded (batch, max_frames) tensor (the reference runs one clip per call,
 * scripts/run_mri_video_inference.py:218-243; it has no batched feeder to mirror).  lengths / lengths_host: int32[batch]
 * on the device / on the host (same values; the host copy sizes the launches, so nothing synchronises).
 * mel_norm: device (batch, max_frames, n_mels), rows past lengths[b] are zero.  Workspace as for (batch, max_frames). */
int m2s_acoustic_forward_packed(m2s_acoustic* m, const void* frames_dev, int32_t frames_are_u8, const float* mask,
                                int32_t batch, int32_t max_frames, const int32_t* lengths, const int32_t* lengths_host,
                                float* mel_norm, void* workspace, size_t workspace_bytes, m2s_stream_t stream);
/* Encoder only, ragged batch fed packed (as m2s_acoustic_forward_packed): feats (batch, max_frames, 208) float32, rows
 * past lengths[b] zero.  With m2s_acoustic_rnn_head it splits the forward in two, so that a batch runner can encode
 * micro-batch after micro-batch (bounded work buffers) and run the recurrence ONCE over many clips: the BiLSTM is a
 * chain of dependent steps whose cost per step barely depends on how many utterances ride along. */
int m2s_acoustic_encode_packed(m2s_acoustic* m, const void* frames_dev, int32_t frames_are_u8, const float* mask,
                               int32_t batch, int32_t max_frames, const int32_t* lengths, const int32_t* lengths_host,
                               float* feats, void* workspace, size_t workspace_bytes, m2s_stream_t stream);
/* Encoder only: (n_frames, height, width) -> (n_frames, 208) features. */
int m2s_acoustic_encode(m2s_acoustic* m, const float* frames_dev, int32_t n_frames, float* feats,
                        void* workspace, size_t workspace_bytes, m2s_stream_t stream);
/* BiLSTM + head only: feats (batch, frames, 208) -> mel_norm (batch, frames, n_mels). */
int m2s_acoustic_rnn_head(m2s_acoustic* m, const float* feats, int32_t batch, int32_t frames,
                          const int32_t* lengths, const int32_t* lengths_host, float* mel_norm,
                          void* workspace, size_t workspace_bytes, m2s_stream_t stream);
int m2s_acoustic_launches(const m2s_acoustic* m);

/* ------------------------------------------------------------------------- */
/* Mel glue (reference scripts/run_mri_video_inference.py:160-163,232-239)     */
/*   mel_db  = pred*std + mean                       (batch, frames, n_mels)   */
/*   mel_log = log(clamp(10^(mel_db/10), 1e-5))      (batch, frames, n_mels)   */
/*   voc_in  = mel_log transposed                    (batch, n_mels, frames)   */
/* Rows past lengths[b] are written as zero.  Any output may be NULL.          */
/* ------------------------------------------------------------------------- */
int m2s_mel_glue(const float* pred_norm, const float* mean, const float* std, int32_t batch,
                 int32_t frames, int32_t n_mels, const int32_t* lengths, float* mel_db, float* mel_log,
                 float* voc_in, m2s_stream_t stream);

/* Debug / probe knobs of the tcgen05 engine ("msub", "base_offset_mode", "a_per_tap", "tmap_tf32",
 * "max_ctas").  Not part of the reference-facing surface. */
int m2s_debug_set_knob(const char* name, int value);
/* Per-launch CUDA-event timing of the conv engine: enable, run, then read (synchronises the device).
 * ms[i] = duration of launch i, flops[i] = 2*M*N*K it executed. */
int m2s_debug_profile(int enable);
int m2s_debug_profile_read(float* ms, double* flops, int32_t cap, int32_t* n);
/* Per-launch timing covers every kernel of the path (not only the conv engine): tags[i] of the launches returned by the
 * last m2s_debug_profile_read -- 1 encoder GEMM (tcgen05 engine), 2 encoder CUDA-core kernels (stem, depthwise, SE,
 * pooling ...), 3 BiLSTM (input projection, recurrence, head), 4 vocoder GEMM (engine / fused pair), 5 vocoder CUDA-core
 * kernels (mel glue, layout, conv_post). */
int m2s_debug_profile_tags(int32_t* tags, int32_t cap, int32_t* n);
/* Kernels launched by this library since the last reset (process-wide; bench.py's gpu_launches). */
long long m2s_debug_launch_count(int reset);
/* clock64 timeline of CTA 0 (producer / MMA / epilogue stamps, 9 per tile) into a device buffer; NULL = off. */
int m2s_debug_trace(unsigned long long* buf, int32_t tiles);

#ifdef __cplusplus
}
#endif
#endif /* M2S_H_ */

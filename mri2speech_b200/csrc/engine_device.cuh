// Device-side building blocks shared by the conv engine kernels (single-CTA and CTA-pair variants):
// PTX wrappers (mbarrier, TMA, tcgen05 / TMEM), the 128B-swizzle SMEM descriptor, and the fused epilogue.
#pragma once
#include "m2s_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace m2s {
namespace engine {

constexpr int kEpiWarps = 16;                 // four per TMEM lane quadrant
constexpr int kEpiGroups = kEpiWarps / 4;     // epilogue warps per quadrant: a warp takes every kEpiGroups-th unit
constexpr int kEpiUnitCols = 16;              // an epilogue unit is 32 rows x 16 columns
constexpr uint32_t kEpiStageBytes = 32 * kEpiUnitCols * 4;   // transpose staging per warp
constexpr int kThreads = 64 + 32 * kEpiWarps; // 18 warps: 112 registers per thread
constexpr int kKBlock = 32;     // tf32 elements per 128-byte swizzle row (fp16: EngineParams::kblock = 64 or 32)
constexpr int kRowBytes = 128;  // tf32 / wide fp16 rows; narrow fp16 layers (c_in <= 32) use 64-byte rows (SWIZZLE_64B)
constexpr int kTmemCols = 512;
constexpr int kMaxStagesA = 4;
constexpr int kMaxStagesB = 8;
constexpr uint32_t kSmemBudget = 200 * 1024;  // > 114 KB forces 1 CTA / SM (TMEM is allocated whole)
// The bias vector is staged in SMEM once per kernel when it has at most this many entries (everything but the
// LSTM input projection): a per-unit __ldg of the bias sat on the epilogue's critical path with ~500 cycles of
// exposed latency per 32 x 32 unit (measured with intra-unit clock stamps, profiles/README.md).
constexpr int kBiasSmemFloats = 1280;
constexpr uint32_t kEpiSmemBytes = kEpiWarps * kEpiStageBytes + kBiasSmemFloats * 4;  // transpose staging + staged bias

// all threads of the CTA, before the first __syncthreads
__device__ __forceinline__ uint32_t stage_bias(const ConvProblem& p, uint32_t stage_base) {
  if (p.n > kBiasSmemFloats) return 0u;
  const uint32_t base = stage_base + kEpiWarps * kEpiStageBytes;
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
    const float v = p.epi.bias ? __ldg(p.epi.bias + i) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * i), "f"(v) : "memory");
  }
  return base;
}

struct EngineParams {
  ConvProblem p;
  const float* wpacked;
  int n_tile, n_tiles, msub, m_tile;
  int tiles_per_batch, total_tiles;
  int cblocks;
  int a_box_rows, a_nbox, shift_min;
  int na, nb;
  uint32_t a_stage_bytes, b_stage_bytes;
  int nacc, acc_stride;
  uint32_t idesc;
  int base_offset_mode;
  int a_per_tap;
  int tap_ksteps;       // per-tap K windows (0 = off): tap j contracts K-steps tap_kofs[j] .. + tap_ksteps - 1 of its K block
  int tap_kofs[M2S_MAX_TAPS];
  int rel_shift[M2S_MAX_TAPS];
  int tg;               // taps per weight stage
  uint32_t b_tap_bytes; // bytes of one tap's weight block (n_tile x 128)
  // operand format: half = 0 tf32 (kblock 32, 128-byte rows, K = 8 per MMA), 1 fp16 (kblock 64, 128-byte rows,
  // K = 16), 2 fp16 narrow (kblock 32, 64-byte rows, SWIZZLE_64B).  A K-step always advances 32 bytes.
  int half, kblock, row_bytes, kstep_elems;
  uint64_t desc_hi;           // SMEM matrix descriptor without the start address
  unsigned long long* trace;  // debug: per-role clock64 stamps of CTA 0 (null = off)
  int trace_tiles;
  int dbg;  // debug: bit0 skip global stores, bit1 skip TMEM loads, bit2 skip SMEM transpose, bit3 skip MMA issue
};

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
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
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires) instead
// of polling.  The polling loop of the first version (try_wait + clock read + compare + branch per iteration) was a
// quarter of all issued instructions of resblock_pair_kernel (ncu source page: 5.4 M loop iterations per launch), taken from
// the schedulers the epilogue warps need.
__device__ __forceinline__ bool mbar_try_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_sleep(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("m2s conv engine: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SMEM matrix descriptor (without the start address), K-major, rows of `row_bytes` (128 -> SWIZZLE_128B, 64 ->
// SWIZZLE_64B); 8-row groups are 8 * row_bytes apart.  base_offset stays 0: the hardware swizzles on absolute SMEM
// address bits, so a start address advanced by whole rows needs no correction (profiles/probe_r01.log).
inline uint64_t make_desc_hi(int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}

// debug timeline: slot = role*3 + k ; layout trace[tile_iter][9]
__device__ __forceinline__ void trace_stamp(const EngineParams& prm, int it, int slot) {
  if (prm.trace && blockIdx.x == 0 && it < prm.trace_tiles) prm.trace[it * 9 + slot] = clock64();
}

// One lane of a converged warp (the compiler then knows the region is warp-uniform and keeps descriptors /
// barrier addresses in uniform registers instead of emitting per-lane R2UR loops).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Four consecutive K-steps (4 x 8 tf32 = one 128-byte swizzle row) of one (sub-tile, tap).
__device__ __forceinline__ void mma_tf32_k4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum0,
                                            int ksteps) {
  mma_tf32(tmem_d, da, db, idesc, accum0);
  if (ksteps > 1) mma_tf32(tmem_d, da + 2, db + 2, idesc, 1u);
  if (ksteps > 2) mma_tf32(tmem_d, da + 4, db + 4, idesc, 1u);
  if (ksteps > 3) mma_tf32(tmem_d, da + 6, db + 6, idesc, 1u);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-steps of one (sub-tile, tap) over one SMEM row block, fp16 operands (K = 16 per instruction, 32 bytes per step).
__device__ __forceinline__ void mma_f16_k4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum0,
                                           int ksteps) {
  mma_f16(tmem_d, da, db, idesc, accum0);
  if (ksteps > 1) mma_f16(tmem_d, da + 2, db + 2, idesc, 1u);
  if (ksteps > 2) mma_f16(tmem_d, da + 4, db + 4, idesc, 1u);
  if (ksteps > 3) mma_f16(tmem_d, da + 6, db + 6, idesc, 1u);
}

struct EpiConsts {
  float inv_slope, pre_w, post_w, out_scale, act_slope;
};

// Epilogue programs (compile-time): the common cases drop every unused instruction -- with 256-column-wide tiles the
// epilogue is ALU-issue-bound (32K outputs per tile), so instructions per output are what matters.
// EPI_RB / EPI_RB_ACC are the two ResBlock conv2 programs of the vocoder (residual stored post-leaky-ReLU and recovered
// with min(y, y/slope); leaky-ReLU as max(v, slope*v), which also covers "no activation" with slope = 1).
// EPI_RB_S / EPI_RB_ACC_S: the same two programs with the residual read from the split-fp16 planes (hi + lo).
enum { EPI_FULL = 0, EPI_FULL_SILU = 1, EPI_BIAS = 2, EPI_LRELU = 3, EPI_SILU = 4, EPI_RES = 5, EPI_RB = 6, EPI_RB_ACC = 7,
       EPI_RB_S = 8, EPI_RB_ACC_S = 9, EPI_COUNT = 10 };

// Host side, every launcher, after copying the problem into its kernel parameters: derived epilogue constants.
// `rows` = largest row index the mask is evaluated for (l_out + d_row_offset): the magic number is exact while
// rows * pitch < 2^32; beyond that pitch_magic stays 0 and the kernels divide.
inline void finalize_epilogue(Epilogue* e, long long rows = 0) {
  const bool on = e->mask_mode == M2S_MASK_PITCH && e->pitch > 0;
  const bool exact = on && (rows + 64) * static_cast<long long>(e->pitch) < (1ll << 32);
  e->pitch_magic = exact ? 0xFFFFFFFFu / static_cast<unsigned>(e->pitch) + 1u : 0u;
}

// Host-side choice of the epilogue program for a problem (shared by both kernels).
// -1: a split-fp16 residual with an epilogue the two ResBlock programs do not cover (unsupported).
inline int choose_epilogue(const Epilogue& e);
// ... and the lo-plane output exists only in the programs that produce a split stream
inline bool lo_output_supported(int epi) { return epi == EPI_LRELU || epi == EPI_RB_S || epi == EPI_RB_ACC_S; }
inline int choose_epilogue(const Epilogue& e) {
  const bool lrelu_ok = (e.act == M2S_ACT_LRELU && e.act_slope > 0.f && e.act_slope <= 1.f);
  const bool plain = !e.accum && e.out_scale == 1.f;
  if (e.res_hi) {
    if (e.res || !e.res_lo || e.res_after_act || e.res_inv_slope < 1.f || !(lrelu_ok || e.act == M2S_ACT_NONE)) return -1;
    return plain && lrelu_ok ? EPI_RB_S : EPI_RB_ACC_S;
  }
  if (plain && !e.res) {
    if (e.act == M2S_ACT_SILU) return EPI_SILU;
    if (lrelu_ok) return EPI_LRELU;
    if (e.act == M2S_ACT_NONE) return EPI_BIAS;
  }
  if (plain && e.res && e.act == M2S_ACT_NONE && e.res_inv_slope == 1.f) return EPI_RES;
  if (e.res && !e.res_after_act && e.res_inv_slope >= 1.f && (lrelu_ok || e.act == M2S_ACT_NONE))
    return plain && lrelu_ok ? EPI_RB : EPI_RB_ACC;
  return e.act == M2S_ACT_SILU ? EPI_FULL_SILU : EPI_FULL;
}

// SiLU with ONE MUFU op: v * sigmoid(v) = h + h * tanh(h), h = v / 2 (tanh.approx.f32: max relative error 2^-11, the
// same size as the operand rounding of the next GEMM).  The exp + reciprocal form costs two quarter-rate MUFU ops per
// output and made every SiLU epilogue MUFU-bound (measured: 2000 cycles per 32 x 32 unit per warp, stores or MMAs
// switched off made no difference).
// `hm` = 0.5, or 0 for an output the image-border mask zeroes (the mask then costs nothing: silu(0) = 0).
__device__ __forceinline__ float fast_silu(float v, float hm = 0.5f) {
  const float h = hm * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// Split-fp16 stream: hi = fp16(v), lo = fp16((v - hi) * 2^11).  The scale keeps lo in v's own exponent range: an
// unscaled remainder of a small activation (|v| < 0.1) would land in fp16's subnormals and lose its mantissa.
constexpr float kSplitScale = 2048.f, kSplitInvScale = 1.f / 2048.f;
// value of 4 consecutive elements of the split-fp16 residual: bits = (hi[0:2], hi[2:4], lo[0:2], lo[2:4])
__device__ __forceinline__ float4 split_decode(const float4& bits) {
  const uint32_t h0 = __float_as_uint(bits.x), h1 = __float_as_uint(bits.y);
  const uint32_t l0 = __float_as_uint(bits.z), l1 = __float_as_uint(bits.w);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h0));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h1));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&l0));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&l1));
  return make_float4(fmaf(c.x, kSplitInvScale, a.x), fmaf(c.y, kSplitInvScale, a.y), fmaf(d.x, kSplitInvScale, b.x),
                     fmaf(d.y, kSplitInvScale, b.y));
}
// hi plane = fp16(v) (saturating), lo plane = fp16((v - hi) * 2^11)
__device__ __forceinline__ void split_encode(const float4& o, uint2* hi, uint2* lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi->x) : "f"(o.y), "f"(o.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi->y) : "f"(o.w), "f"(o.z));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi->x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi->y));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo->x) : "f"((o.y - a.y) * kSplitScale), "f"((o.x - a.x) * kSplitScale));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo->y) : "f"((o.w - b.y) * kSplitScale), "f"((o.z - b.x) * kSplitScale));
}

template <int kEpiS>
__device__ __forceinline__ float epi_elem(const EpiConsts& c, float acc, float bias, float res, float accum,
                                          float hm = 0.5f) {
  constexpr int kEpi = kEpiS == EPI_RB_S ? EPI_RB : kEpiS == EPI_RB_ACC_S ? EPI_RB_ACC : kEpiS;
  float v = acc + bias;
  if (kEpi == EPI_BIAS) return v;
  if (kEpi == EPI_LRELU) return fmaxf(v, v * c.act_slope);  // 0 < slope <= 1
  if (kEpi == EPI_SILU) return fast_silu(v, hm);
  if (kEpi == EPI_RES) return v + res;
  if (kEpi == EPI_RB) {
    v += fminf(res, res * c.inv_slope);                     // inv_slope >= 1
    return fmaxf(v, v * c.act_slope);
  }
  if (kEpi == EPI_RB_ACC) {
    v += fminf(res, res * c.inv_slope);
    v = (v + accum) * c.out_scale;
    return fmaxf(v, v * c.act_slope);                       // slope 1 = no activation
  }
  const float rt = res >= 0.f ? res : res * c.inv_slope;
  v = fmaf(rt, c.pre_w, v);
  v += accum;
  v *= c.out_scale;
  if (kEpi == EPI_FULL_SILU) v = fast_silu(v);
  else v = v >= 0.f ? v : v * c.act_slope;
  return fmaf(rt, c.post_w, v);
}


// Per-warp epilogue state that does not change across tiles.
//
// Geometry: 16 epilogue warps, four per TMEM lane quadrant.  A warp works on units of 32 rows x 16 columns: thread
// (rr0 = lane / 4, cc = lane % 4) owns rows rr0 + 8 i (i = 0..3) and columns 4 cc .. 4 cc + 3 of the unit.  Half-width
// units halve the registers a thread needs (16 accumulators in flight, 4 float4 of residual / accumulate / transposed
// data instead of 8), which is what lets 18 warps fit the register file (112 registers each): the epilogue is a
// latency chain (TMEM -> SMEM transpose -> global load -> FMA -> global store) and with two warps per scheduler
// half of its time was dependency stalls (tools/trace_pair.py, profiles/README.md).
struct EpiWarp {
  EpiConsts ec;
  uint32_t stage;  // this warp's 2 KB transpose staging (shared::cta address)
  uint32_t bias_smem;  // staged bias vector (shared::cta address), 0 = read the bias from global memory
  int quad, grp, lane, rr0, cc;
  int mask_mode;
  bool has_res, has_acc;
  int dbg;  // probe switches (EngineParams::dbg): bit0 skip global stores, bit1 skip TMEM loads, bit2 skip SMEM transpose
};

__device__ __forceinline__ EpiWarp make_epi_warp(const Epilogue& e, uint32_t stage_base, int ew, int warp, int lane,
                                                 int dbg = 0, uint32_t bias_smem = 0u) {
  EpiWarp w;
  w.bias_smem = bias_smem;
  w.dbg = dbg;
  w.quad = warp & 3;  // TMEM lane quadrant this warp may access
  w.grp = ew >> 2;    // which of the kEpiGroups warps of the quadrant
  w.lane = lane;
  w.rr0 = lane >> 2;
  w.cc = lane & 3;
  w.stage = stage_base + ew * kEpiStageBytes;
  w.has_res = e.res != nullptr || e.res_hi != nullptr;
  w.has_acc = e.accum != nullptr;
  w.ec.inv_slope = e.res_inv_slope;
  w.ec.pre_w = (w.has_res && !e.res_after_act) ? 1.f : 0.f;
  w.ec.post_w = (w.has_res && e.res_after_act) ? 1.f : 0.f;
  w.ec.out_scale = e.out_scale;
  w.ec.act_slope = e.act == M2S_ACT_LRELU ? e.act_slope : 1.f;
  w.mask_mode = e.mask_mode;
  return w;
}

// row-loop variants: 0 = no image-border mask, 1 = image-border mask, 2 = decided at run time per row (one loop copy)
template <int kV> struct PitchTag { static constexpr int value = kV; };

// transpose staging: row r of the unit (16 floats = 64 bytes) lives at stage + 64 r, its 16-byte chunk c at
// chunk (c ^ ((r >> 1) & 3)): conflict-free for the row-per-lane writes and for the 4-lanes-per-row reads
__device__ __forceinline__ uint32_t epi_stage_addr(uint32_t stage, int row, int chunk) {
  return stage + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

// Epilogue of one accumulator tile for one warp: rows [q0, q0 + 128*msub) of batch item b, columns [n0, n0+n_tile).
// TMEM -> registers (row per thread) -> SMEM transpose (XOR-swizzled, conflict-free) -> coalesced global traffic:
// 4 lanes cover one 64-byte row segment, a warp instruction covers 8 rows (512 bytes per request).  `tmem_acc` already
// carries this warp's lane-quadrant offset.
// `q_end` (exclusive) bounds the rows this tile may write (the fused pair kernel keeps only part of a tile).
template <int kEpi>
__device__ __forceinline__ void epilogue_tile(const ConvProblem& p, const EpiWarp& ew_, uint32_t tmem_acc, int b, int q0,
                                              int n0, int msub, int n_tile, int q_end = 0x7fffffff) {
  constexpr int kR = 4;   // rows per thread: rr0 + 8 i
  const int row_end = min(p.l_out, q_end);
  const Epilogue& e = p.epi;
  const EpiConsts& ec = ew_.ec;
  const uint32_t stage = ew_.stage;
  const int quad = ew_.quad, grp = ew_.grp, lane = ew_.lane, rr0 = ew_.rr0, cc = ew_.cc;
  const int mask_mode = ew_.mask_mode;
  const bool has_res = ew_.has_res, has_acc = ew_.has_acc;
  const int nchunks = (n_tile + kEpiUnitCols - 1) / kEpiUnitCols;
  int len_rows = 0x7fffffff;
  if (mask_mode == M2S_MASK_LEN) len_rows = __ldg(e.lens + b) * e.len_scale;
  const size_t d_base = static_cast<size_t>(b) * p.d_batch_rows + p.d_row_offset;
  // The accumulator registers of the warp's next unit are requested from TMEM as soon as the current unit's have been
  // parked in the SMEM staging buffer: the TMEM read overlaps the arithmetic / global traffic of the current one.
  uint32_t r[16];
  auto issue_tmem_ld = [&](int sub_, int ci_) { tmem_ld16(tmem_acc + sub_ * n_tile + ci_ * kEpiUnitCols, r); };
  const bool pipelined = ew_.dbg == 0;
  // this warp's units: (sub-tile, 16-column chunk) pairs in row-major order, every kEpiGroups-th one, stepped without
  // divisions (the epilogue is half issue-bound: profiles/README.md)
  int sub = 0, ci = grp;
  while (ci >= nchunks) { ci -= nchunks; ++sub; }
  if (pipelined && sub < msub) issue_tmem_ld(sub, ci);
  const bool pitch_mask = mask_mode == M2S_MASK_PITCH;
  // image-border mask as one multiplier (1 or 0) per row of this thread, recomputed only when the 32-row slab changes
  // (the units of a slab share their rows): the per-row divmod + compare chain + selects were a third of the
  // instructions of the encoder's 3x3 convs, which are issue-bound in the epilogue (62 % issue-active, profiles/README.md)
  float rm[kR] = {1.f, 1.f, 1.f, 1.f};
  int rm_sub = -1;   // (dead code in the accumulate programs)
  for (int sub_n, ci_n; sub < msub; sub = sub_n, ci = ci_n) {
    sub_n = sub;
    ci_n = ci + kEpiGroups;
    while (ci_n >= nchunks) { ci_n -= nchunks; ++sub_n; }
    const bool has_next = sub_n < msub;
    const int c0 = ci * kEpiUnitCols;
    const int qw = q0 + sub * 128 + quad * 32;  // first row of this warp's 32-row slab
    if (!pipelined) {
      if (!(ew_.dbg & 2)) {
        issue_tmem_ld(sub, ci);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0x3f800000u + j + lane;
      }
    }
    // While the TMEM read is in flight: bias / residual / accumulate loads of this unit (the output may alias
    // them in place, so every load is issued before the first store).  Thread (rr0, cc) owns rows rr0 + 8i,
    // columns n .. n+3; all row predicates reduce to "8i + rr0 < bound".
    const int n = n0 + c0 + cc * 4;
    constexpr bool kSplit = kEpi == EPI_RB_S || kEpi == EPI_RB_ACC_S;  // residual = float(hi) + float(lo), fp16 planes
    constexpr bool kHasRes = kEpi == EPI_FULL || kEpi == EPI_FULL_SILU || kEpi == EPI_RES || kEpi == EPI_RB || kEpi == EPI_RB_ACC || kSplit;
    constexpr bool kHasAcc = kEpi == EPI_FULL || kEpi == EPI_FULL_SILU || kEpi == EPI_RB_ACC || kEpi == EPI_RB_ACC_S;
    // programs that can write the lo plane (d16_lo): the producers of a split stream -- the transposed-conv epilogue
    // (EPI_LRELU) and the split ResBlock programs (compile-time: the hot loop carries no pointer it does not use)
    constexpr bool kLoOut = kSplit || kEpi == EPI_LRELU;
    // ---- fast path: all 32 rows of the unit exist and survive the length mask (warp-uniform test): no row predicates on
    // the loads, row pointers stepped by warp-uniform strides; a lane whose 4 columns fall outside a narrow / ragged N tile
    // simply sits out (one predicate per thread); the image-border mask is evaluated per row.  A unit cut only by q_end
    // (the fused pair kernel keeps M1 - (k-1) rows of a tile) takes this path too, with predicated stores: its rows exist
    // in memory, and the warp that owns the cut unit of EVERY tile must not be slower than the others (it delayed the
    // T-tile barrier of the next tile by 2-8 k cycles: tools/trace_pair.py).
    // Measured alternatives that lost or tied (profiles/README.md): a row-per-thread epilogue straight from the TMEM
    // registers to global memory; a fragment-layout epilogue (tcgen05.ld.16x256b + lane-pair shuffles, no SMEM transpose);
    // prefetching the residual one unit / one tile ahead in registers or with cp.async.bulk.prefetch.L2.
    {
      int rows_valid_u = 32;
      if (mask_mode == M2S_MASK_LEN) rows_valid_u = len_rows - (qw + p.d_row_offset);
      const int keep = row_end - qw;
      const bool fast = p.l_out - qw >= 32 && rows_valid_u >= 32 && !ew_.dbg;
      if (fast) {
        const bool lane_ok = (c0 + cc * 4 < n_tile) && n < p.n;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!lane_ok) {
        } else if (ew_.bias_smem)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(bias4.x), "=f"(bias4.y), "=f"(bias4.z), "=f"(bias4.w)
                       : "r"(ew_.bias_smem + 4u * n));
        else if (e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
        const size_t row0 = d_base + qw + rr0;
        float4 res4[kR], acc4[kR];
        if (kHasRes) {
          if (kSplit && lane_ok) {
            // Lanes (cc even, cc + 1) own adjacent 4-column groups of the same rows: the even lane fetches 16 bytes
            // (8 columns) of the hi plane, the odd lane 16 bytes of the lo plane, and they swap halves by shuffle when
            // the row is used: 8-byte loads (twice the requests for the same bytes) made the split stream much SLOWER
            // than fp32 (tools/pair_bench.py).  The hi plane is the tile conv1 just read: an L2 hit.
            const __half* plane = static_cast<const __half*>((cc & 1) ? e.res_lo : e.res_hi);
            const uint4* rp = reinterpret_cast<const uint4*>(plane + row0 * e.res_ld + (n & ~7));
#pragma unroll
            for (int i = 0; i < kR; ++i) {
              const uint4 v = rp[static_cast<size_t>(i) * e.res_ld];
              res4[i] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
            }
          } else if (!kSplit && has_res && lane_ok) {
            if (e.res_half) {   // fp16 residual stream: 8 bytes per lane
              const uint2* rp = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(e.res) + row0 * e.res_ld + n);
#pragma unroll
              for (int i = 0; i < kR; ++i) {
                const uint2 u = rp[static_cast<size_t>(i) * 2 * e.res_ld];
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                res4[i] = make_float4(a.x, a.y, b.x, b.y);
              }
            } else {
              const float4* rp = reinterpret_cast<const float4*>(e.res + row0 * e.res_ld + n);
#pragma unroll
              for (int i = 0; i < kR; ++i) res4[i] = rp[static_cast<size_t>(i) * 2 * e.res_ld];
            }
          } else {
#pragma unroll
            for (int i = 0; i < kR; ++i) res4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (kHasAcc) {
          const float4* ap = reinterpret_cast<const float4*>(e.accum + row0 * e.accum_ld + n);
          if (has_acc && lane_ok) {
#pragma unroll
            for (int i = 0; i < kR; ++i) acc4[i] = ap[static_cast<size_t>(i) * 2 * e.accum_ld];
          } else {
#pragma unroll
            for (int i = 0; i < kR; ++i) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(epi_stage_addr(stage, lane, j)),
                       "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        __syncwarp();
        // (fast path implies pipelined)  Programs with an accumulate stream hold 16 more registers (acc4) through the
        // row loop: they request the next unit's accumulators after it instead -- 36 bytes of spills in that loop cost
        // 10 % of the launch, the exposed TMEM read a few dozen cycles per unit.
        if (!kHasAcc && has_next) issue_tmem_ld(sub_n, ci_n);
        float4* dp = reinterpret_cast<float4*>(p.d + row0 * p.d_ld + n);
        uint2* hp = reinterpret_cast<uint2*>(static_cast<__half*>(p.d16) + row0 * p.d_ld + n);
        uint2* lp = reinterpret_cast<uint2*>(static_cast<__half*>(p.d16_lo) + row0 * p.d_ld + n);
        const bool st32 = p.d != nullptr && lane_ok, st16 = p.d16 != nullptr && lane_ok;
        const bool st_lo = kLoOut && p.d16_lo != nullptr;
        float4 a8[kR];
#pragma unroll
        for (int i = 0; i < kR; ++i) {
          const int rr = i * 8 + rr0;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a8[i].x), "=f"(a8[i].y), "=f"(a8[i].z), "=f"(a8[i].w)
                       : "r"(epi_stage_addr(stage, rr, cc)));
        }
        // (programs with an accumulate stream are at the register cap and never run on images in this path: they keep
        // the plain per-row test below)
        if (!kHasAcc && pitch_mask && sub != rm_sub) {   // (warp-uniform)
          rm_sub = sub;
          const int drow = qw + rr0 + p.d_row_offset;
          int mi = e.pitch_magic ? static_cast<int>(__umulhi(static_cast<unsigned>(drow), e.pitch_magic)) : drow / e.pitch;
          int mj = drow - mi * e.pitch;
#pragma unroll
          for (int i = 0; i < kR; ++i) {
            rm[i] = (mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi) ? 1.f : 0.f;
            mj += 8;
            while (mj >= e.pitch) { mj -= e.pitch; ++mi; }
          }
        }
        // one row loop per mask kind (compile-time inside)
        auto rows = [&](auto pitch_tag) {
          constexpr int kPitchMode = decltype(pitch_tag)::value;
          const bool kPitch = kPitchMode == 2 ? pitch_mask : kPitchMode == 1;
          int mi = 0, mj = 0;
          if (kHasAcc && kPitch) {
            const int drow = qw + rr0 + p.d_row_offset;
            mi = drow / e.pitch;
            mj = drow - mi * e.pitch;
          }
#pragma unroll
          for (int i = 0; i < kR; ++i) {
            const int rr = i * 8 + rr0;
            const float4 a4 = a8[i];
            float4 r4 = kHasRes ? res4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (kSplit) {
              // even lane holds hi[own 4 | neighbour's 4], odd lane lo[neighbour's 4 | own 4]: swap the neighbour's half
              const bool odd = (cc & 1) != 0;
              const float g0 = __shfl_xor_sync(0xffffffffu, odd ? r4.x : r4.z, 1);
              const float g1 = __shfl_xor_sync(0xffffffffu, odd ? r4.y : r4.w, 1);
              r4 = split_decode(odd ? make_float4(g0, g1, r4.z, r4.w) : make_float4(r4.x, r4.y, g0, g1));
            }
            const float4 c4 = kHasAcc ? acc4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            // plain SiLU program: the mask rides on the 0.5 of h = v / 2; every other program multiplies its result
            const float hm = (kEpi == EPI_SILU && kPitch) ? 0.5f * rm[i] : 0.5f;
            o.x = epi_elem<kEpi>(ec, a4.x, bias4.x, r4.x, c4.x, hm);
            o.y = epi_elem<kEpi>(ec, a4.y, bias4.y, r4.y, c4.y, hm);
            o.z = epi_elem<kEpi>(ec, a4.z, bias4.z, r4.z, c4.z, hm);
            o.w = epi_elem<kEpi>(ec, a4.w, bias4.w, r4.w, c4.w, hm);
            if (!kHasAcc && kEpi != EPI_SILU && kPitch) { o.x *= rm[i]; o.y *= rm[i]; o.z *= rm[i]; o.w *= rm[i]; }
            if (kHasAcc && kPitch) {
              if (!(mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi)) o = make_float4(0.f, 0.f, 0.f, 0.f);
              mj += 8;
              while (mj >= e.pitch) { mj -= e.pitch; ++mi; }
            }
            const bool row_ok = rr < keep;   // always true unless q_end cuts this unit
            if (st32 && row_ok) dp[static_cast<size_t>(i) * 2 * p.d_ld] = o;
            if (st16 && row_ok) {  // fp16 copy: the tensor-core operand of the next conv (saturating conversion, never inf)
              uint2 pk, pl;
              if (st_lo) {  // + the lo plane: (hi, lo) together are the residual source of the next pair
                split_encode(o, &pk, &pl);
                lp[static_cast<size_t>(i) * 2 * p.d_ld] = pl;
              } else {
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o.y), "f"(o.x));
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o.w), "f"(o.z));
              }
              hp[static_cast<size_t>(i) * 2 * p.d_ld] = pk;
            }
          }
        };
        // (programs with an accumulate stream are at the register cap: ONE copy of the loop, or ptxas spills in it)
        if (kHasAcc) rows(PitchTag<2>{});
        else if (pitch_mask) rows(PitchTag<1>{});
        else rows(PitchTag<0>{});
        if (kHasAcc && has_next) issue_tmem_ld(sub_n, ci_n);
        __syncwarp();
        continue;
      }
    }
    // ---- general path: rows that do not exist or are masked (tile / utterance ends), probe switches ----
    const bool col_ok = (c0 + cc * 4 < n_tile) && n < p.n;
    const int rows_ok = col_ok ? min(32, row_end - qw) : 0;                 // rows that exist
    int rows_valid = 32;                                                    // rows that survive the mask
    if (mask_mode == M2S_MASK_LEN) rows_valid = len_rows - (qw + p.d_row_offset);
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    const size_t row0 = d_base + qw + rr0;
    float* dptr = p.d + row0 * p.d_ld + n;
    __half* hptr = static_cast<__half*>(p.d16) + row0 * p.d_ld + n;
    __half* lptr = static_cast<__half*>(p.d16_lo) + row0 * p.d_ld + n;
    const size_t d_step = static_cast<size_t>(8) * p.d_ld;
    float4 res4[kR], acc4[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      res4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (kSplit) {
      const __half* rh = static_cast<const __half*>(e.res_hi) + row0 * e.res_ld + n;
      const __half* rl = static_cast<const __half*>(e.res_lo) + row0 * e.res_ld + n;
      const size_t r_step = static_cast<size_t>(8) * e.res_ld;
#pragma unroll
      for (int i = 0; i < kR; ++i)
        if (i * 8 + rr0 < rows_ok) {
          const uint2 h = *reinterpret_cast<const uint2*>(rh + i * r_step), l = *reinterpret_cast<const uint2*>(rl + i * r_step);
          res4[i] = make_float4(__uint_as_float(h.x), __uint_as_float(h.y), __uint_as_float(l.x), __uint_as_float(l.y));
        }
#pragma unroll
      for (int i = 0; i < kR; ++i) res4[i] = split_decode(res4[i]);   // zero bits decode to zero
    } else if (kHasRes) {
      if (has_res) {
        const size_t r_step = static_cast<size_t>(8) * e.res_ld;
        if (e.res_half) {
          const __half* rptr = reinterpret_cast<const __half*>(e.res) + row0 * e.res_ld + n;
#pragma unroll
          for (int i = 0; i < kR; ++i)
            if (i * 8 + rr0 < rows_ok) {
              const uint2 u = *reinterpret_cast<const uint2*>(rptr + i * r_step);
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
              const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
              res4[i] = make_float4(a.x, a.y, b.x, b.y);
            }
        } else {
          const float* rptr = e.res + row0 * e.res_ld + n;
#pragma unroll
          for (int i = 0; i < kR; ++i)
            if (i * 8 + rr0 < rows_ok) res4[i] = *reinterpret_cast<const float4*>(rptr + i * r_step);
        }
      }
    }
    if (kHasAcc) {
      if (has_acc) {
        const float* aptr = e.accum + row0 * e.accum_ld + n;
        const size_t a_step = static_cast<size_t>(8) * e.accum_ld;
#pragma unroll
        for (int i = 0; i < kR; ++i)
          if (i * 8 + rr0 < rows_ok) acc4[i] = *reinterpret_cast<const float4*>(aptr + i * a_step);
      }
    }
    // image-border mask: (i, j) = divmod(row, pitch) once, then stepped by 8 rows
    int mi = 0, mj = 0;
    if (mask_mode == M2S_MASK_PITCH) {
      const int drow = qw + rr0 + p.d_row_offset;
      mi = drow / e.pitch;
      mj = drow - mi * e.pitch;
    }
    tmem_ld_wait();
    if (!(ew_.dbg & 4)) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(epi_stage_addr(stage, lane, j)),
                     "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                     : "memory");
      __syncwarp();
    }
    if (pipelined && has_next) issue_tmem_ld(sub_n, ci_n);
    float4 o[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const int rr = i * 8 + rr0;
      float4 a4;
      if (!(ew_.dbg & 4)) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w)
                     : "r"(epi_stage_addr(stage, rr, cc)));
      } else {
        a4 = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 4]), __uint_as_float(r[i + 8]), __uint_as_float(r[i + 12]));
      }
      o[i].x = epi_elem<kEpi>(ec, a4.x, bias4.x, res4[i].x, acc4[i].x);
      o[i].y = epi_elem<kEpi>(ec, a4.y, bias4.y, res4[i].y, acc4[i].y);
      o[i].z = epi_elem<kEpi>(ec, a4.z, bias4.z, res4[i].z, acc4[i].z);
      o[i].w = epi_elem<kEpi>(ec, a4.w, bias4.w, res4[i].w, acc4[i].w);
    }
    if (mask_mode == M2S_MASK_PITCH) {
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        const bool valid = mi >= e.i_lo && mi < e.i_hi && mj >= e.j_lo && mj < e.j_hi;
        if (!valid) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        mj += 8;
        while (mj >= e.pitch) { mj -= e.pitch; ++mi; }
      }
    } else if (mask_mode == M2S_MASK_LEN) {
#pragma unroll
      for (int i = 0; i < kR; ++i)
        if (i * 8 + rr0 >= rows_valid) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.d && !(ew_.dbg & 1)) {
#pragma unroll
      for (int i = 0; i < kR; ++i)
        if (i * 8 + rr0 < rows_ok) *reinterpret_cast<float4*>(dptr + i * d_step) = o[i];
    }
    if (p.d16 && !(ew_.dbg & 1)) {  // fp16 copy of the tile: the tensor-core operand of the next conv (saturating, never inf)
#pragma unroll
      for (int i = 0; i < kR; ++i)
        if (i * 8 + rr0 < rows_ok) {
          uint2 pk, pl;
          split_encode(o[i], &pk, &pl);
          *reinterpret_cast<uint2*>(hptr + i * d_step) = pk;
          if (kLoOut && p.d16_lo) *reinterpret_cast<uint2*>(lptr + i * d_step) = pl;
        }
    }
    __syncwarp();
  }
}

}  // namespace engine
}  // namespace m2s

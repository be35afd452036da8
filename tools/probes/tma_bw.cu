// Probe: per-SM load throughput of (a) 2-D TMA boxes of 128-byte rows out of a row-major matrix (the way mb_project /
// mb_expand_dw fetch their operands), (b) 2-D TMA boxes of 64-byte rows (SWIZZLE_64B), (c) cp.async.bulk of the same
// number of contiguous bytes (the way the conv engine fetches its pre-swizzled weight blocks).  All SMs active, one
// producer thread per CTA, a ring of 4 stages, nobody consumes.   nvcc -arch=sm_100a -o tools/probes/tma_bw tools/probes/tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// mode 0: TMA 2-D box (box_cols x box_rows); mode 1: bulk copies of `bytes`
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, const uint8_t* flat, int mode, int iters,
                                               int box_rows, int box_cols_elems, uint32_t bytes, int kblocks, long long rows_total) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 4u * 49152u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint32_t ph[4] = {0, 0, 0, 0};
  int issued = 0;
  const long long tiles = rows_total / box_rows;
  for (int i = 0; i < iters + 4; ++i) {
    const int s = i & 3;
    if (i >= 4) { mbar_wait(bars + 8u * s, ph[s]); ph[s] ^= 1; }
    if (i < iters) {
      const long long item = (static_cast<long long>(blockIdx.x) + static_cast<long long>(issued) * gridDim.x);
      const long long tile = (item / kblocks) % tiles;
      const int kb = static_cast<int>(item % kblocks);
      mbar_expect(bars + 8u * s, bytes);
      if (mode == 0) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(base + s * 49152u), "l"(&tm), "r"(bars + 8u * s), "r"(kb * box_cols_elems), "r"(static_cast<int>(tile * box_rows)) : "memory");
      } else {
        const uint8_t* src = flat + (static_cast<size_t>(tile) * kblocks + kb) * bytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(base + s * 49152u), "l"(src), "r"(bytes), "r"(bars + 8u * s) : "memory");
      }
      ++issued;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  const int C = 1248;                       // fp16 columns (stage 5's C_mid)
  const long long R = 131072;               // rows: 327 MB, beyond L2
  uint8_t* buf;
  CK(cudaMalloc(&buf, static_cast<size_t>(R) * C * 2));
  CK(cudaMemset(buf, 1, static_cast<size_t>(R) * C * 2));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct Case { const char* name; int mode, box_rows, box_cols; CUtensorMapSwizzle sw; };
  const Case cases[] = {{"tma 256 rows x 128 B (SWIZZLE_128B)", 0, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B},
                        {"tma 128 rows x 128 B (SWIZZLE_128B)", 0, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B},
                        {"tma 256 rows x  64 B (SWIZZLE_64B)", 0, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B},
                        {"bulk 32 KB contiguous", 1, 256, 64, CU_TENSOR_MAP_SWIZZLE_NONE},
                        {"bulk 16 KB contiguous", 1, 128, 64, CU_TENSOR_MAP_SWIZZLE_NONE}};
  for (const Case& c : cases) {
    CUtensorMap tm;
    cuuint64_t gd[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(R)};
    cuuint64_t gs[1] = {static_cast<cuuint64_t>(C) * 2};
    cuuint32_t bx[2] = {static_cast<cuuint32_t>(c.box_cols), static_cast<cuuint32_t>(c.box_rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      c.mode == 0 ? c.sw : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", static_cast<int>(cr)); return 1; }
    const uint32_t bytes = static_cast<uint32_t>(c.box_rows) * c.box_cols * 2;
    const int kblocks = C / c.box_cols;     // (19 full K blocks of 64, or 39 of 32)
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      probe<<<sms, 128, 4 * 49152 + 2048, 0>>>(tm, buf, c.mode, iters, c.box_rows, c.box_cols, bytes, kblocks, R);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1)
        printf("%-40s %8.1f GB/s aggregate, %6.1f B/clk/SM at 1.9 GHz, %5.2f us per %u-byte load\n", c.name,
               static_cast<double>(bytes) * iters * sms / (ms * 1e6), static_cast<double>(bytes) * iters / (ms * 1e-3 * 1.9e9),
               ms * 1e3 / iters, bytes);
    }
  }
  return 0;
}

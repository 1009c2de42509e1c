// Is TMA throughput per SM limited by the ISSUING THREAD (one box every ~N cycles) or by the TMA unit?
// W producer warps per CTA, each streams 16 KB boxes into its own 2 slots. nvcc -arch=sm_100a
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  uint32_t d = 0;
  while (!d) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(bar), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"((uint64_t)src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int BOX = 128 * 128, SLOTS = 4;   // per warp: 4 slots of 16 KB
template <int W>
__global__ void __launch_bounds__(W * 32, 1) k(const __grid_constant__ CUtensorMap map, int iters, int mode, const char* gbase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = (uint64_t*)(smem + W * SLOTS * BOX);
  const uint32_t base = smem_u32(smem), bb = smem_u32(bars);
  const int w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int s = 0; s < W * SLOTS; ++s) mbar_init(bb + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    const int row0 = (blockIdx.x * W + w) * 1024;
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t bar = bb + 8 * (w * SLOTS + s);
      if (i >= SLOTS) mbar_wait(bar, ph ^ 1);
      mbar_expect(bar, BOX);
      if (mode == 0) tma2d(base + (w * SLOTS + s) * BOX, &map, bar, 0, row0 + (i & 7) * 128);
      else if (mode == 1) bulk1d(base + (w * SLOTS + s) * BOX, gbase + (size_t)(row0 + (i & 7) * 128) * 128, BOX, bar);
      else {   // one tensor box (8 KB = 64 rows) + one bulk copy (8 KB)
        tma2d(base + (w * SLOTS + s) * BOX, &map, bar, 0, row0 + (i & 7) * 128);
      }
      if (++s == SLOTS) { s = 0; ph ^= 1; }
    }
    for (int j = 0; j < SLOTS; ++j) { mbar_wait(bb + 8 * (w * SLOTS + s), ph ^ 1); if (++s == SLOTS) { s = 0; ph ^= 1; } }
  }
}
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int W> void run(const CUtensorMap& map, int sms, int mode, const char* gbase) {
  const int smem = W * SLOTS * BOX + 256, iters = 4000;
  CK(cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0)); k<W><<<sms, W * 32, smem>>>(map, iters, mode, gbase); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("mode %d (%s) %d producer warps x %d slots: %.3f ms, %.0f ns per box per warp, %.0f ns per box per SM, %.2f TB/s\n", mode, mode == 0 ? "tensor box 128 rows x 128 B" : "bulk 1-D copy of 16 KB", W, SLOTS, ms,
                    ms * 1e6 / iters, ms * 1e6 / iters / W, (double)iters * W * BOX * sms / ms / 1e9);
  }
}
int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  PFN_enc enc = (PFN_enc)f;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const size_t rows = (size_t)1024 * 24 * sms;   // 465 MB (strided views need room; L2-resident working set stays 1024 rows x stride per CTA)
  void* d; CK(cudaMalloc(&d, rows * 128)); CK(cudaMemset(d, 0, rows * 128));
  CUtensorMap map; cuuint64_t gdim[2] = {64, rows}, gstr[1] = {128}; cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  for (int mode = 0; mode < 2; ++mode) { run<1>(map, sms, mode, (const char*)d); }
  for (int stride : {128, 512, 1280, 2816}) {
    // view the same buffer as [rows2, stride/2] bf16 and take the first 64 columns of each row
    const size_t rows2 = rows * 128 / stride;
    CUtensorMap m2; cuuint64_t gd[2] = {(cuuint64_t)stride / 2, rows2}, gs[1] = {(cuuint64_t)stride};
    if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc fail\n"); return 1; }
    printf("row stride %d B: ", stride);
    // each warp's region: 1024 rows of the strided view must stay inside the buffer: rows2 >= sms*1024?
    if (rows2 < (size_t)sms * 1024) { printf("skipped (buffer too small)\n"); continue; }
    run<1>(m2, sms, 0, (const char*)d);
  }
  return 0;
}

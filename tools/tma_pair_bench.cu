// Does the .cta_group::2 TMA form (completion signalled on the leader CTA's mbarrier) or a 3-D tensor map cost
// throughput? Clusters of 2; each CTA streams 16 KB boxes into its own 4 slots.
//   variant 0: plain TMA, own barrier          variant 1: 3-D map, own barrier
//   variant 2: cta_group::2 TMA, BOTH CTAs' boxes complete on the leader's barrier (leader waits, then remote-arrives
//              on the peer's "go" barrier so the peer may reuse its slot) -- the pair kernel's protocol
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
  uint32_t d = 0; long long t0 = clock64();
  while (!d) { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(bar), "r"(ph) : "memory");
    if (clock64() - t0 > 3000000000LL) { printf("timeout blk %d\n", blockIdx.x); __trap(); } }
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2d_cg2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
constexpr int BOX = 128 * 128, SLOTS = 4;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3, int iters, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = (uint64_t*)(smem + SLOTS * BOX);     // [0..SLOTS) full, [SLOTS..2SLOTS) go (peer may refill)
  const uint32_t base = smem_u32(smem), bb = smem_u32(bars);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { for (int s = 0; s < 2 * SLOTS; ++s) mbar_init(bb + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int row0 = blockIdx.x * 1024;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t full = bb + 8 * s, go = bb + 8 * (SLOTS + s);
      if (variant < 2) {
        if (i >= SLOTS) mbar_wait(full, ph ^ 1);
        mbar_expect(full, BOX);
        if (variant == 0) tma2d(base + s * BOX, &m2, full, 0, row0 + (i & 7) * 128);
        else tma3d(base + s * BOX, &m3, full, 0, (i & 7) * 128, blockIdx.x);
      } else {
        // leader: waits for the pair's 2 boxes, then tells the peer the slot is free again
        if (i >= SLOTS) {
          if (rank == 0) {
            mbar_wait(full, ph ^ 1);
            uint32_t remote; asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote) : "r"(go));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
          } else {
            mbar_wait(go, ph ^ 1);
          }
        }
        if (rank == 0) mbar_expect(full, 2 * BOX);
        tma2d_cg2(base + s * BOX, &m2, full & 0xFEFFFFFFu, 0, row0 + (i & 7) * 128);
      }
      if (++s == SLOTS) { s = 0; ph ^= 1; }
    }
    if (variant < 2) { for (int j = 0; j < SLOTS; ++j) { mbar_wait(bb + 8 * s, ph ^ 1); if (++s == SLOTS) { s = 0; ph ^= 1; } } }
    else if (rank == 0) { for (int j = 0; j < SLOTS; ++j) { mbar_wait(bb + 8 * s, ph ^ 1); if (++s == SLOTS) { s = 0; ph ^= 1; } } }
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  PFN_enc enc = (PFN_enc)f;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount / 2 * 2;
  const size_t rows = (size_t)1024 * sms;
  void* d; CK(cudaMalloc(&d, rows * 128)); CK(cudaMemset(d, 0, rows * 128));
  CUtensorMap m2, m3;
  { cuuint64_t gd[2] = {64, rows}, gs[1] = {128}; cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  { cuuint64_t gd[3] = {64, 1024, (cuuint64_t)sms}, gs[2] = {128, 128 * 1024}; cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    if (enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  const int smem = SLOTS * BOX + 256, iters = 4000;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* names[3] = {"plain 2-D TMA, own barrier", "3-D map, own barrier", "cta_group::2 TMA, pair completes on the leader's barrier"};
  for (int v = 0; v < 3; ++v) for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0)); k<<<sms, 64, smem>>>(m2, m3, iters, v); CK(cudaEventRecord(e1));
    cudaError_t err = cudaDeviceSynchronize(); if (err != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(err)); return 1; }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("variant %d (%s): %.3f ms, %.0f ns per box per SM, %.2f TB/s\n", v, names[v], ms, ms * 1e6 / iters, (double)iters * BOX * sms / ms / 1e9);
  }
  return 0;
}

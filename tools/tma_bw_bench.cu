// L2 -> shared-memory TMA bandwidth probe (development tool; decides how many bytes per tensor-core cycle a
// GEMM-shaped kernel may pull from L2 on this part).  nvcc -arch=sm_100a -O3 -o tma_bw_bench tma_bw_bench.cu
// Modes: 0 = every CTA streams the SAME L2-resident matrix (weights-like)
//        1 = every CTA streams its OWN L2-resident region (activation-like)
//        2 = clusters of C CTAs: rank 0 multicasts the same box to all C (mode 0 data)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  uint32_t d = 0;
  long long t0 = clock64();
  while (!d) {
#ifdef USE_TEST_WAIT
    asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(bar), "r"(ph) : "memory");
#else
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(bar), "r"(ph) : "memory");
#endif
    if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int STAGES = 6, BOX_ROWS = 256, BOX_BYTES = BOX_ROWS * 128;   // 32 KB boxes, 192 KB ring

template <int CLUSTER>
__global__ void __launch_bounds__(128, 1) bw_kernel(const __grid_constant__ CUtensorMap map, int mode, int rows_per_cta,
                                                     int iters, int box_rows, unsigned long long* cycles, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = (uint64_t*)(smem + STAGES * BOX_BYTES);
  const uint32_t base = smem_u32(smem), bb = smem_u32(bars);
  uint32_t rank = 0;
  if (CLUSTER > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bb + 8 * s, 1);
    if (stages > STAGES) stages = STAGES;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CLUSTER > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const int bytes = box_rows * 128;
  if (threadIdx.x == 0) {
    const int row0 = (mode == 1) ? blockIdx.x * rows_per_cta : 0;
    const int nbox = rows_per_cta / box_rows;
    long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(bb + 8 * s, ((i / stages) - 1) & 1);
      if (i < iters) {
        mbar_expect(bb + 8 * s, bytes);
        const int r = row0 + (i % nbox) * box_rows;
        if (mode == 2) {
          // every CTA expects the bytes; only rank 0 issues, multicast to the whole cluster.
          if (rank == 0) tma2d_mc(base + s * BOX_BYTES, &map, bb + 8 * s, 0, r, (uint16_t)((1u << CLUSTER) - 1));
        } else {
          tma2d(base + s * BOX_BYTES, &map, bb + 8 * s, 0, r);
        }
      }
      if (CLUSTER > 1 && mode == 2 && (i % STAGES) == STAGES - 1) {
        // keep the cluster in lockstep so a multicast never lands in a slot a peer has not re-armed
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (CLUSTER > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  PFN_enc enc = (PFN_enc)f;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int rows_per_cta = 4096;                       // 4096 rows x 128 B = 512 KB per CTA region
  const size_t rows = (size_t)rows_per_cta * sms;       // 74 MB total: L2 resident
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, rows * 128));
  CK(cudaMemset(d, 0, rows * 128));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, sms * 8));
  const int smem = STAGES * BOX_BYTES + 64;
  CK(cudaFuncSetAttribute(bw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(bw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(bw_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int box_rows : {256, 128, 64}) {
    CUtensorMap map;
    cuuint64_t gdim[2] = {64, rows}, gstr[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        bw_kernel<1><<<sms, 128, smem>>>(map, mode, rows_per_cta, iters, box_rows, cyc, (int)STAGES);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto c : h) avg += c;
        avg /= sms;
        const double bytes = (double)iters * box_rows * 128 * sms;
        if (rep) printf("box %3d rows  mode %d (%s): %.3f ms  %.2f TB/s  %.1f B/clk/SM (avg %.0f cycles, %.2f GHz)\n", box_rows, mode,
                        mode == 0 ? "same data for all CTAs" : "private region per CTA", ms, bytes / ms / 1e9,
                        (double)iters * box_rows * 128 / avg, avg, avg / ms / 1e6);
      }
    }
    for (int st : {1, 2, 3, 4, 6}) {
      CK(cudaEventRecord(e0));
      bw_kernel<1><<<sms, 128, smem>>>(map, 1, rows_per_cta, iters, box_rows, cyc, st);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("box %3d rows  private, %d in flight: %.3f ms  -> %.0f ns per box, latency estimate %.0f ns, %.2f TB/s\n", box_rows, st, ms,
             ms * 1e6 / iters, ms * 1e6 / iters * st, (double)iters * box_rows * 128 * sms / ms / 1e9);
    }
    // multicast, clusters of 2 and 4 (same data): bytes DELIVERED to shared memory per SM
    for (int cl : {2, 4}) {
      const int grid = sms / cl * cl;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cl == 2) CK(cudaLaunchKernelEx(&cfg, bw_kernel<2>, map, 2, rows_per_cta, iters, box_rows, cyc, (int)STAGES));
        else CK(cudaLaunchKernelEx(&cfg, bw_kernel<4>, map, 2, rows_per_cta, iters, box_rows, cyc, (int)STAGES));
        CK(cudaEventRecord(e1));
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("cluster %d: %s\n", cl, cudaGetErrorString(err)); return 1; }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)iters * box_rows * 128 * grid;
        if (rep) printf("box %3d rows  multicast cluster %d (grid %d): %.3f ms  %.2f TB/s delivered\n", box_rows, cl, grid, ms, bytes / ms / 1e9);
      }
    }
  }
  return 0;
}

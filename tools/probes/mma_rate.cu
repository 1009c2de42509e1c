// Probe: issue rate of back-to-back tcgen05.mma instructions (operands: zeros in shared memory, no loads), per kind and
// shape, one CTA per SM.  Answers "how many cycles does one M128 N256 kind::tf32 (K = 8) MMA take next to kind::f16 (K = 16)".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I text_to_speech_b200/csrc -I include -o /tmp/mma_rate tools/probes/mma_rate.cu -lcuda
#include <cstdio>
#include <vector>
#include "tc_tf32_kernels.cuh"
using namespace wg;

// mode 0: f16 N256 one accumulator; 1: tf32 N256 one accumulator; 2: tf32 N256 two accumulators alternating;
// 3: tf32 N128; 4: f16 N128; 5: tf32 SWIZZLE_64B descriptors N256; 6 / 7 / 8: CTA pair (cta_group::2, M = 256): f16 N256,
// tf32 N256, tf32 N128 (launched as clusters of 2; the leader issues)
template <bool pair>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode, int n_mma, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  const bool leader = !pair || cluster_ctarank() == 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (pair) { tmem2_alloc(smem_u32(&slot), 512); tmem2_relinquish(); }
    else { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && leader) {
    const uint64_t a = mode == 5 ? umma_desc_sw64(base) : umma_desc_sw128(base);
    const uint64_t b = mode == 5 ? umma_desc_sw64(base + 32768) : umma_desc_sw128(base + 32768);
    const int N = (mode == 3 || mode == 4 || mode == 8) ? 128 : 256;
    const uint32_t id16 = umma_idesc_bf16(pair ? 256 : 128, N), id32 = umma_idesc_tf32(pair ? 256 : 128, N);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tm + ((mode == 2 && (i & 1)) ? 256u : 0u);
        if (pair && mode == 6) umma2_bf16(d, a + 2 * (i & 3), b + 2 * (i & 3), id16, i > 1);
        else if (pair) umma2_tf32(d, a + 2 * (i & 3), b + 2 * (i & 3), id32, i > 1);
        else if (mode == 0 || mode == 4) umma_bf16(d, a + 2 * (i & 3), b + 2 * (i & 3), id16, i > 1);
        else umma_tf32(d, a + 2 * (i & 3), b + 2 * (i & 3), id32, i > 1);
      }
      if (pair) tc2_commit(smem_u32(&bar));
      else tc_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    long long tt = 0;   // the elected lane recorded t0
    for (int l = 0; l < 32; ++l) {
      const long long v = __shfl_sync(0xffffffffu, t0, l);
      if (v) tt = v;
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - tt);
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem2_dealloc(tm, 512);
    else tmem_dealloc(tm, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  unsigned long long* d;
  cudaMalloc(&d, sms * 8);
  const char* names[] = {"f16  M128 N256 K16", "tf32 M128 N256 K8", "tf32 M128 N256 K8, two accumulators", "tf32 M128 N128 K8",
                         "f16  M128 N128 K16", "tf32 M128 N256 K8 (SWIZZLE_64B)", "f16  pair M256 N256 K16",
                         "tf32 pair M256 N256 K8", "tf32 pair M256 N128 K8"};
  const int n = 2048;
  for (int grid : {2, sms & ~1}) {
    for (int mode = 0; mode < 9; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d, 0, sms * 8);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 96 * 1024;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = mode >= 6 ? 1 : 0;
        cudaError_t le = mode >= 6 ? cudaLaunchKernelEx(&cfg, mma_rate_kernel<true>, mode, n, d)
                                  : cudaLaunchKernelEx(&cfg, mma_rate_kernel<false>, mode, n, d);
        if (le != cudaSuccess) { printf("launch error (mode %d): %s\n", mode, cudaGetErrorString(le)); return 1; }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      std::vector<unsigned long long> h(grid);
      cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0;
      int cnt = 0;
      for (auto v : h) if (v) { s += (double)v; ++cnt; }
      printf("grid %3d  %-40s %7.1f cycles / MMA\n", grid, names[mode], s / (cnt ? cnt : 1) / n);
    }
  }
  return 0;
}

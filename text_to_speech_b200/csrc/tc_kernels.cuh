// tcgen05 / TMEM / TMA kernels of the WaveGlow engine (WG_MODE_BF16), sm_100a only.
//
//   tc_gemm_kernel      D[M,N] = A[M,K] @ W[N,K]^T + bias -- the polyphase ConvTranspose upsample
//                       (waveglow_arch.py:196-198, :245-253) and the stand-alone GEMM self-test.
//   tc_wn_layer_kernel  one fused WN layer (waveglow_arch.py:105-141 loop body):
//                         GEMM1  [128 x (3C+S)] @ [(3C+S) x 2C]  dilated k=3 conv (3 row-shifted A tiles,
//                                TMA zero fill = 'same' padding) + cond 1x1 as extra K
//                         gate   tanh * sigmoid on the fp32 accumulator (TMEM -> registers)
//                         fold   acc8 += acts @ (Wskip @ Wend)   (skip accumulation + end conv, fp32 FMA)
//                         GEMM2  acts[128 x C] (smem, written by the gate epilogue) @ Wres[C x C]
//                         res    h += GEMM2 + b   (fp32 master + bf16 shadow for the next layer's TMA)
//
// Operands are BF16, K-major, 128-byte swizzled in shared memory (TMA SWIZZLE_128B == UMMA
// LayoutType::SWIZZLE_128B); accumulators are fp32 in TMEM; one elected thread issues tcgen05.mma.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "simt_kernels.cuh"

namespace wg {

// ================================================================================================
// PTX wrappers
// ================================================================================================
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
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("wg_b200: mbarrier timeout block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Programmatic dependent launch: the trigger lets the next kernel of the stream start launching once every CTA of this grid
// has issued it; the wait blocks until the previous grid has COMPLETED and its memory is visible.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a fully converged warp. The producer / MMA issue loops keep the WHOLE warp in the loop and elect only around
// the asm: ptxas then keeps descriptors and coordinates in uniform registers; a loop entered by `if (lane == 0)` made every
// UTCHMMA / UTMALDG pay an ELECT + R2UR.BROADCAST waterfall (profiles/r01_probes.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO>>4 in [16,30) (unused for swizzled K-major), SBO>>4 in [32,46) = 8 rows * 128 B,
// version=1 in [46,48), layout_type=2 (SWIZZLE_128B) in [61,64). Tile bases are 1024-B aligned.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10),
// both K-major, N>>3 in [17,23), M>>4 in [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ================================================================================================
// Generic GEMM: D[M,N] = A[M,K] @ W[N,K]^T + bias[N].  Tile 128 x 256 x 64, 4-stage TMA ring.
// warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2..5: epilogue.
// ================================================================================================
constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 64, TG_STAGES = 4, TG_THREADS = 192;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 2, TG_B_BYTES = TG_BN * TG_BK * 2;
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_SMEM = TG_STAGES * TG_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

template <typename OutT>
__global__ void __launch_bounds__(TG_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const float* __restrict__ bias, OutT* __restrict__ D, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TG_STAGES * TG_STAGE_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S] accumulator full; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TG_STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TG_STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * TG_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TG_BN, m0 = blockIdx.y * TG_BM;
  const int num_kb = K / TG_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TG_STAGES;
        const uint32_t ph = (kb / TG_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), TG_STAGE_BYTES);
        const uint32_t a_dst = smem_base + s * TG_STAGE_BYTES;
        tma_load_2d(a_dst, &map_a, full_bar(s), kb * TG_BK, m0);
        tma_load_2d(a_dst + TG_A_BYTES, &map_w, full_bar(s), kb * TG_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TG_BM, TG_BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TG_STAGES;
        const uint32_t ph = (kb / TG_STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * TG_STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + TG_A_BYTES);
#pragma unroll
        for (int k = 0; k < TG_BK / 16; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        tc_commit(empty_bar(s));
      }
      tc_commit(acc_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int m = m0 + row;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int cb = 0; cb < TG_BN / 16; ++cb) {
      uint32_t r[16];
      tmem_ld16(taddr + cb * 16, r);
      tmem_ld_wait();
      if (m < M) {
        const int n = n0 + cb * 16;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + n + j) : 0.f);
        if constexpr (sizeof(OutT) == 4) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(D) + (size_t)m * N + n);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(D) + (size_t)m * N + n);
#pragma unroll
          for (int q = 0; q < 2; ++q)
            o[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                              pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ================================================================================================
// Fused WN layer (C = 256, S = 640).  One CTA per SM, persistent over 128-row tiles.
//   warp 0      TMA producer        warp 1      TMEM alloc + MMA issuer       warps 2..9  epilogue
//
// The residual stream h is kept as a bf16 pair (hi, lo) with hi = bf16(h), lo = bf16(h - hi): hi is the
// MMA operand of the next layer's dilated conv, hi + lo carries ~16 mantissa bits. The residual add
// itself runs on the tensor core:  D2 = acts @ Wres + hi @ I + lo @ I  (I = 64x64 identity B tile), so
// the epilogue never reads h back from global memory.
//
// TMEM: two 256-column fp32 regions R0/R1.  Tile with parity p:
//   GEMM1 chunk a (gate channels 0..127)   -> R[p]      K = 12 conv blocks of 64 + 10 (spect) or 5 (mel window) cond blocks;
//                                                       FIRST: 1 start-fold block + cond blocks
//   GEMM1 chunk b (gate channels 128..255) -> R[p^1]    (issued while the epilogue drains chunk a)
//   GEMM2 + residual                       -> R[p]      the half that depends only on chunk a's acts
//                                                       is issued before chunk b's epilogue finishes
// ================================================================================================
constexpr int WL_C = 256, WL_S = 640, WL_BM = 128, WL_BK = 64;
constexpr int WL_STAGES = 3;
constexpr int WL_A_BYTES = WL_BM * WL_BK * 2;          // 16 KB
constexpr int WL_B_BYTES = 256 * WL_BK * 2;            // 32 KB
constexpr int WL_STAGE_BYTES = WL_A_BYTES + WL_B_BYTES;
constexpr int WL_KB_CONV = 3 * WL_C / WL_BK;           // 12
constexpr int WL_KB1 = WL_KB_CONV + WL_S / WL_BK;      // 22
constexpr int WL_KB2 = WL_C / WL_BK;                   // 4
constexpr int WL_K1 = 3 * WL_C + WL_S;                 // 1408
constexpr int WL_ACTS_BYTES = WL_BM * WL_C * 2;        // 64 KB
constexpr int WL_EPI_WARPS = 8, WL_EPI_THREADS = WL_EPI_WARPS * 32;
constexpr int WL_THREADS = 64 + WL_EPI_THREADS;        // 320
constexpr int WL_OFF_ACTS = WL_STAGES * WL_STAGE_BYTES;
constexpr int WL_OFF_I64 = WL_OFF_ACTS + WL_ACTS_BYTES;   // 64x64 bf16 identity, K-major SWIZZLE_128B (8 KB)
constexpr int WL_OFF_B1 = WL_OFF_I64 + 64 * 128;
constexpr int WL_OFF_B2 = WL_OFF_B1 + 2 * WL_C * 4;
constexpr int WL_OFF_O8 = WL_OFF_B2 + WL_C * 4;           // [128][8] fp32: partial fold sums of the hf=1 warps
constexpr int WL_OFF_BARS = WL_OFF_O8 + WL_BM * 8 * 4;
constexpr int WL_NBARS = 2 * WL_STAGES + 3 + 2 + 4;
constexpr int WL_SMEM = WL_OFF_BARS + WL_NBARS * 8 + 16;
static_assert(WL_SMEM <= 232448, "shared memory budget");

// Row geometry. Position l of utterance b is stored at row ((b*R + r)*T + t) with l = R*t + r:
//   R = 1  : position-major (T = L rows per utterance), conditioning = upsampled spect [.., 640] (10 K-blocks)
//   R = 32 : phase-major. All 128 rows of a tile share the upsample phase r, so the
//            conditioning GEMM runs at its intrinsic rank: A = 4-frame mel window [.., 320] (5 K-blocks),
//            B = (Wup_r @ Wcond) folded per (layer, phase). A row shift by s positions is the tile of phase
//            (r+s) mod R moved by floor((r+s)/R) frames, still one contiguous TMA box. The host passes the batch
//            as ONE utterance (b = 0) of T = B*Tp frames -- all utterances of a phase in one row sequence, Tp =
//            frames + gap rows apart (RowGeom, simt_kernels.cuh); the gap rows are stored as zeros (Tp / Tv below),
//            which is the zero padding between utterances, and tiles span utterance boundaries.
struct WnLayerParams {
  int T, R, tiles_per_row, n_tiles;
  int L, tiles_per_b;   // legacy names used by the CTA-pair kernel (position-major only): L = T, tiles_per_b = tiles_per_row
  int n_cond_kb;        // K-blocks of the conditioning operand: 10 (spect) or 5 (mel window)
  int wc_col0;          // first K column of the conditioning weights inside map_wc
  int wc_row0;          // first N row of this layer's conditioning weights inside map_wc
  int wc_rstride;       // extra N rows per phase (0 or 512)
  int tile_order; // 1: tile index = (row tile, phase) with the phase fastest (phase-major); 0: phase blocks one after the other
  int Tp, Tv;     // gap layout (phase-major): row t of a phase block is valid iff (t % Tp) < Tv; Tp = 0: every row < T is valid
  const int* row_b;   // ragged batches (wg_infer_ragged): row t is valid iff row_b[t] >= 0 (overrides Tp / Tv); else null
  int layer;      // row block in the stacked W1 / W2 matrices
  int flow;       // row block in the stacked start-fold matrices W0 / H0 (FIRST variant)
  int dilation;
  const float* b1;   // [512] chunk-packed
  const float* b2;   // [256]
  __nv_bfloat16* hi_out;   // [B*L, 256] bf16(h) written for the next layer
  __nv_bfloat16* lo;       // [B*L, 256] bf16(h - hi), updated in place (rows are tile-private)
  float* acc8;       // [B*L, 8] folded skip/end accumulator (read-modify-write, one thread per row)
  unsigned long long* timing;   // optional [16] cycle counters (debug), may be null
  int pdl;     // launched with the programmatic-dependent-launch attribute: run the griddepcontrol instructions (small grids:
               // the next layer's CTAs start on idle SMs and run their prologue while this grid is still working)
  int flags;   // only read when compiled with -DWG_PROBES (result-breaking A/B probes, profiles/r01_probes.md):
               //   1 = skip every other W1 tile load, 2 = skip every other activation tile load
};

// Row validity of phase-block row t (gap rows between utterances hold zeros and never touch acc8).
__device__ __forceinline__ bool wn_row_valid(const WnLayerParams& p, int t) {
  if (t >= p.T) return false;
  if (p.row_b) return p.row_b[t] >= 0;
  return p.Tp == 0 || t % p.Tp < p.Tv;
}

struct WnLayerConst {   // kernel-parameter (constant bank) copy: every read is warp-uniform
  float wse[WL_C * 8];  // Wskip @ Wend, [256][8]
};

// Residual epilogue for 16 columns of one row: v = accumulator + bias; pass 0 stages hi = bf16(v), pass 1
// stages lo = bf16(v - hi) into the SWIZZLE_128B staging tile.
__device__ __forceinline__ void resid_step(const uint32_t (&r)[16], const float* bb, uint8_t* stg, int gi, int row,
                                           int pass, bool valid = true) {
  uint32_t w[8];
#pragma unroll
  for (int j2 = 0; j2 < 8; ++j2) {
    if (!valid) {      // gap row between two utterances: h stays exactly zero
      w[j2] = 0u;
      continue;
    }
    const float2 b2v = *reinterpret_cast<const float2*>(bb + 2 * j2);
    const float v0 = __uint_as_float(r[2 * j2]) + b2v.x;
    const float v1 = __uint_as_float(r[2 * j2 + 1]) + b2v.y;
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
    const float2 hf2 = __bfloat1622float2(h2);
    const uint32_t hi_bits = *reinterpret_cast<const uint32_t*>(&h2);
    const uint32_t lo_bits = pack_bf16x2(v0 - hf2.x, v1 - hf2.y);
    w[j2] = pass == 0 ? hi_bits : lo_bits;
  }
  uint8_t* dst = stg + (gi >> 2) * (128 * 64 * 2);
  const int c0 = (gi & 3) * 2;
  *reinterpret_cast<uint4*>(dst + ((c0 ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(dst + (((c0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
}

// Gate epilogue for 16 gate channels of one row: tanh * sigmoid on the fp32 accumulator (registers t/g),
// fold into the skip/end accumulator (weights from the kernel-parameter bank, warp-uniform address), and
// bf16 acts into the GEMM2 A tile (K-major SWIZZLE_128B: 16-byte chunk c of row r sits at c ^ (r & 7)).
template <bool LAST>
__device__ __forceinline__ void gate_step(const uint32_t (&t)[16], const uint32_t (&g)[16], const float* bT,
                                          const float* wse, uint8_t* kblk, int st, int row, float (&o8)[8]) {
  const float* bG = bT + 128;
  float a[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 bt4 = *reinterpret_cast<const float4*>(bT + 4 * j4);
    const float4 bg4 = *reinterpret_cast<const float4*>(bG + 4 * j4);
    const float btv[4] = {bt4.x, bt4.y, bt4.z, bt4.w}, bgv[4] = {bg4.x, bg4.y, bg4.z, bg4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * j4 + jj;
      const float xt = __uint_as_float(t[j]) + btv[jj];
      const float xg = __uint_as_float(g[j]) + bgv[jj];
      a[j] = tanh_fast(xt) * fmaf(0.5f, tanh_fast(0.5f * xg), 0.5f);
      const float4 w0 = *reinterpret_cast<const float4*>(wse + j * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wse + j * 8 + 4);
      o8[0] = fmaf(a[j], w0.x, o8[0]); o8[1] = fmaf(a[j], w0.y, o8[1]);
      o8[2] = fmaf(a[j], w0.z, o8[2]); o8[3] = fmaf(a[j], w0.w, o8[3]);
      o8[4] = fmaf(a[j], w1.x, o8[4]); o8[5] = fmaf(a[j], w1.y, o8[5]);
      o8[6] = fmaf(a[j], w1.z, o8[6]); o8[7] = fmaf(a[j], w1.w, o8[7]);
    }
  }
  if (!LAST) {
    const uint4 v0 = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
    const uint4 v1 = make_uint4(pack_bf16x2(a[8], a[9]), pack_bf16x2(a[10], a[11]), pack_bf16x2(a[12], a[13]), pack_bf16x2(a[14], a[15]));
    *reinterpret_cast<uint4*>(kblk + (((st * 2) ^ (row & 7)) << 4)) = v0;
    *reinterpret_cast<uint4*>(kblk + (((st * 2 + 1) ^ (row & 7)) << 4)) = v1;
  }
}

// Packed-math (f32x2: FADD2 / FMUL2 / FFMA2, sm_100) versions used by the single-CTA kernel. Two adjacent gate
// channels travel as one float2 through bias add, gate and fold; the fold keeps EVEN and ODD channels in the two
// lanes of eight float2 accumulators (summed once per tile), so no operand has to be duplicated:
//   o8p[c] += (a_j, a_j+1) * (wse[j][c], wse[j+1][c])      wse2 = [channel pair][8] float2 (host-permuted)
template <bool LAST>
__device__ __forceinline__ void gate_step2(const uint32_t (&t)[16], const uint32_t (&g)[16], const float* bT,
                                           const float2* wse2, uint8_t* kblk, int st, int row, float2 (&o8p)[8]) {
  const float* bG = bT + 128;
  const float2 half2 = make_float2(0.5f, 0.5f);
  uint32_t pk[8];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 bt4 = *reinterpret_cast<const float4*>(bT + 4 * j4);
    const float4 bg4 = *reinterpret_cast<const float4*>(bG + 4 * j4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j2 = 2 * j4 + h;   // channel pair (2 j2, 2 j2 + 1)
      const float2 xt = __fadd2_rn(make_float2(__uint_as_float(t[2 * j2]), __uint_as_float(t[2 * j2 + 1])),
                                   h == 0 ? make_float2(bt4.x, bt4.y) : make_float2(bt4.z, bt4.w));
      const float2 xg = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(g[2 * j2]), __uint_as_float(g[2 * j2 + 1])),
                                              h == 0 ? make_float2(bg4.x, bg4.y) : make_float2(bg4.z, bg4.w)), half2);
      const float2 th = make_float2(tanh_fast(xt.x), tanh_fast(xt.y));
      const float2 sg = __ffma2_rn(make_float2(tanh_fast(xg.x), tanh_fast(xg.y)), half2, half2);
      const float2 a = __fmul2_rn(th, sg);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 w = *reinterpret_cast<const float4*>(wse2 + j2 * 8 + 2 * c4);   // (w[j][2c4], w[j+1][2c4], w[j][2c4+1], w[j+1][2c4+1])
        o8p[2 * c4] = __ffma2_rn(a, make_float2(w.x, w.y), o8p[2 * c4]);
        o8p[2 * c4 + 1] = __ffma2_rn(a, make_float2(w.z, w.w), o8p[2 * c4 + 1]);
      }
      pk[j2] = pack_bf16x2(a.x, a.y);
    }
  }
  if (!LAST) {
    *reinterpret_cast<uint4*>(kblk + (((st * 2) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(kblk + (((st * 2 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// Residual epilogue, packed: PASS 0 stages hi = bf16(v) only, PASS 1 recomputes it and stages lo = bf16(v - hi).
template <int PASS>
__device__ __forceinline__ void resid_step2(const uint32_t (&r)[16], const float* bb, uint8_t* stg, int gi, int row, bool valid) {
  uint32_t w[8];
#pragma unroll
  for (int j2 = 0; j2 < 8; ++j2) {
    if (!valid) {      // gap row between two utterances: h stays exactly zero (the next layer's zero padding)
      w[j2] = 0u;
      continue;
    }
    const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[2 * j2]), __uint_as_float(r[2 * j2 + 1])),
                                *reinterpret_cast<const float2*>(bb + 2 * j2));
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v.x, v.y);
    if (PASS == 0) {
      w[j2] = *reinterpret_cast<const uint32_t*>(&h2);
    } else {
      const float2 d = __ffma2_rn(__bfloat1622float2(h2), make_float2(-1.f, -1.f), v);   // exact: v - hi
      w[j2] = pack_bf16x2(d.x, d.y);
    }
  }
  uint8_t* dst = stg + (gi >> 2) * (128 * 64 * 2);
  const int c0 = (gi & 3) * 2;
  *reinterpret_cast<uint4*>(dst + ((c0 ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(dst + (((c0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
}

// FIRST (layer 0 of a flow, phase-major only): the start conv is folded into the layer (simt_kernels.cuh,
// a0_build_kernel). The 12 conv K-blocks become ONE K-block  a0[128 x 64] @ W0[64 x 512]  (three K = 16 MMAs,
// one per tap) and the residual operand comes from the same tile,  a0 @ H0[64 x 256]  (one K = 16 MMA), instead
// of eight identity MMAs on (hi, lo): h0 never exists in HBM.
template <bool LAST, bool FIRST = false>
__global__ void __launch_bounds__(WL_THREADS, 1)
tc_wn_layer_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_ho,
                   const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_cond, const __grid_constant__ CUtensorMap map_w1,
                   const __grid_constant__ CUtensorMap map_wc,
                   const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_a0,
                   const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_h0, const WnLayerParams p,
                   const __grid_constant__ WnLayerConst cw) {
  static_assert(!(LAST && FIRST), "the start fold needs a residual layer");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + WL_OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + WL_OFF_B2);
  float* s_o8 = reinterpret_cast<float*>(smem + WL_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WL_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WL_NBARS);
  const uint32_t bar_base = smem_base + WL_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WL_STAGES + s); };
  auto dfull_bar = [&](int i) { return bar_base + 8u * (2 * WL_STAGES + i); };        // 0: chunk a, 1: chunk b, 2: GEMM2
  auto drained_bar = [&](int i) { return bar_base + 8u * (2 * WL_STAGES + 3 + i); };  // LAST: chunk a / b accumulators read out
  const uint32_t actsa_bar = bar_base + 8u * (2 * WL_STAGES + 5);                     // acts K-blocks 0,1 in smem, D1a read out
  const uint32_t acts_bar = bar_base + 8u * (2 * WL_STAGES + 6);                      // acts K-blocks 2,3 in smem, D1b read out
  const uint32_t epi2_bar = bar_base + 8u * (2 * WL_STAGES + 7);                      // GEMM2 accumulator read out
  const uint32_t acts2_bar = bar_base + 8u * (2 * WL_STAGES + 8);                     // acts K-block 2 in smem

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_base & 1023u) != 0u) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned bases
  if (p.pdl) pdl_trigger();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    prefetch_tmap(&map_ho);
    prefetch_tmap(&map_lo);
    prefetch_tmap(&map_cond);
    prefetch_tmap(&map_wc);
    prefetch_tmap(&map_w1);
    prefetch_tmap(&map_w2);
    if (FIRST) {
      prefetch_tmap(&map_a0);
      prefetch_tmap(&map_w0);
      prefetch_tmap(&map_h0);
    }
    for (int s = 0; s < WL_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(dfull_bar(i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(drained_bar(i), WL_EPI_THREADS);
    mbar_init(actsa_bar, WL_EPI_THREADS);
    mbar_init(acts_bar, WL_EPI_THREADS);
    mbar_init(epi2_bar, WL_EPI_THREADS);
    mbar_init(acts2_bar, WL_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * WL_C; i += WL_THREADS) s_b1[i] = p.b1[i];
  if (!LAST) {
    for (int i = threadIdx.x; i < WL_C; i += WL_THREADS) s_b2[i] = p.b2[i];
  }
  if (!LAST && !FIRST) {
    // identity B tile: element (n, k) of a [64 x 64] K-major SWIZZLE_128B tile sits at byte
    // n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2
    uint32_t* i64w = reinterpret_cast<uint32_t*>(smem + WL_OFF_I64);
    for (int i = threadIdx.x; i < 64 * 32; i += WL_THREADS) {
      const int n = i >> 5, w = i & 31;                 // 32-bit word w of row n (physical position)
      const int chunk_phys = w >> 2, chunk_log = chunk_phys ^ (n & 7);
      const int k0 = chunk_log * 8 + (w & 3) * 2;       // logical k of the low half-word
      uint32_t v = 0;
      if (k0 == n) v = 0x00003F80u;                      // bf16 1.0 in the low half
      if (k0 + 1 == n) v = 0x3F800000u;
      i64w[i] = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched only weights / kernel parameters; from here on the previous kernel's results are read
  if (p.pdl) pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const bool timing = p.timing != nullptr;
  const bool pm = p.R > 1;   // phase-major: maps are (channels, frames, phases, batch); else (channels, rows, batch, 1)
  constexpr int KB_CONV = FIRST ? 1 : WL_KB_CONV;   // conv K-blocks of GEMM1: 12, or the single start-fold block
  const int kb1 = KB_CONV + p.n_cond_kb;   // K-blocks of GEMM1 per chunk: 22 (spect), 17 (mel window), 6 (FIRST)
  // tile -> (utterance b, phase r, first row t0 of the 128-row tile inside the (b, r) row block)
  auto tile_coords = [&](int tile, int& b, int& r, int& t0) {
    if (p.tile_order) {
      // phase fastest: the CTAs running at the same time cover ALL phases of a few 128-row ranges, so the dilated-conv
      // taps (phases r +- d of the same rows) are L2 hits whatever the dilation
      r = tile % p.R;
      const int bt = tile / p.R;
      b = bt / p.tiles_per_row;
      t0 = (bt - b * p.tiles_per_row) * WL_BM;
      return;
    }
    const int tt = tile % p.tiles_per_row, br = tile / p.tiles_per_row;
    r = br % p.R;
    b = br / p.R;
    t0 = tt * WL_BM;
  };

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    {
      uint32_t it = 0;  // running stage counter
      long long t_wait = 0;
      auto acquire = [&](uint32_t bytes) -> uint32_t {
        const int s = it % WL_STAGES;
        const uint32_t ph = (it / WL_STAGES) & 1;
        long long t0 = 0;
        if (timing) t0 = clock64();
        mbar_wait(empty_bar(s), ph ^ 1);
        if (timing) t_wait += clock64() - t0;
        if (elect_one()) mbar_expect_tx(full_bar(s), bytes);
        __syncwarp();
        return static_cast<uint32_t>(s);
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int b, r, t0;
        tile_coords(tile, b, r, t0);
        for (int q = 0; q < 2; ++q) {
          for (int kb = 0; kb < kb1; ++kb, ++it) {
#ifdef WG_PROBES
            const bool skip_b = (p.flags & 1) && (kb & 1), skip_a = (p.flags & 2) && (kb & 1);
#else
            constexpr bool skip_b = false, skip_a = false;
#endif
            const uint32_t s = acquire((skip_a ? 0 : WL_A_BYTES) + (skip_b ? 0 : WL_B_BYTES));
            const uint32_t a_dst = smem_base + s * WL_STAGE_BYTES;
            if (elect_one()) {
              if (FIRST && kb == 0) {
                if (!skip_a) tma_load_4d(a_dst, &map_a0, full_bar(s), 0, t0, pm ? r : b, pm ? b : 0);
                if (!skip_b) tma_load_2d(a_dst + WL_A_BYTES, &map_w0, full_bar(s), 0, p.flow * 2 * WL_C + q * 256);
              } else if (kb < KB_CONV) {
                // tap shifted by sh positions: phase (r+sh) mod R, frames moved by floor((r+sh)/R)
                const int tap = kb >> 2, cblk = kb & 3;
                const int rs = r + (tap - 1) * p.dilation;
                const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
                if (!skip_a) tma_load_4d(a_dst, &map_h, full_bar(s), cblk * WL_BK, t0 + carry, pm ? rs - carry * p.R : b, pm ? b : 0);
                if (!skip_b) tma_load_2d(a_dst + WL_A_BYTES, &map_w1, full_bar(s), kb * WL_BK, p.layer * 2 * WL_C + q * 256);
              } else {
                const int kc = kb - KB_CONV;
                if (!skip_a) tma_load_4d(a_dst, &map_cond, full_bar(s), kc * WL_BK, t0, pm ? 0 : b, pm ? b : 0);
                if (!skip_b) tma_load_2d(a_dst + WL_A_BYTES, &map_wc, full_bar(s), p.wc_col0 + kc * WL_BK,
                                         p.wc_row0 + r * p.wc_rstride + q * 256);
              }
            }
            __syncwarp();
          }
        }
        if (FIRST) {
          // consumption order of the MMA warp: a0 / H0 (residual operand) | W2 blocks 0..3 (A = acts in smem)
          for (int step = 0; step < 5; ++step, ++it) {
            const uint32_t s = acquire(step == 0 ? WL_STAGE_BYTES : WL_B_BYTES);
            const uint32_t dst = smem_base + s * WL_STAGE_BYTES;
            if (elect_one()) {
              if (step == 0) {
                tma_load_4d(dst, &map_a0, full_bar(s), 0, t0, pm ? r : b, pm ? b : 0);
                tma_load_2d(dst + WL_A_BYTES, &map_h0, full_bar(s), 0, p.flow * WL_C);
              } else {
                tma_load_2d(dst + WL_A_BYTES, &map_w2, full_bar(s), (step - 1) * WL_BK, p.layer * WL_C);
              }
            }
            __syncwarp();
          }
        } else if (!LAST) {
          // consumption order of the MMA warp: W2/hi blocks 0,1 | lo blocks 0..3 | W2/hi blocks 2,3
          for (int step = 0; step < 6; ++step, ++it) {
            if (step == 2 || step == 3) {
              const uint32_t s = acquire(2 * WL_A_BYTES);
              const uint32_t dst = smem_base + s * WL_STAGE_BYTES;
              const int kb = (step - 2) * 2;
              if (elect_one()) {
                tma_load_4d(dst, &map_lo, full_bar(s), kb * WL_BK, t0, pm ? r : b, pm ? b : 0);
                tma_load_4d(dst + WL_A_BYTES, &map_lo, full_bar(s), (kb + 1) * WL_BK, t0, pm ? r : b, pm ? b : 0);
              }
            } else {
              const int kb = step < 2 ? step : step - 2;
              const uint32_t s = acquire(WL_STAGE_BYTES);
              const uint32_t dst = smem_base + s * WL_STAGE_BYTES;
              if (elect_one()) {
                tma_load_4d(dst, &map_h, full_bar(s), kb * WL_BK, t0, pm ? r : b, pm ? b : 0);
                tma_load_2d(dst + WL_A_BYTES, &map_w2, full_bar(s), kb * WL_BK, p.layer * WL_C);
              }
            }
            __syncwarp();
          }
        }
      }
      if (timing && lane == 0) atomicAdd(p.timing + 8, static_cast<unsigned long long>(t_wait));
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer =======================================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
      constexpr uint32_t idesc_id = umma_idesc_bf16(128, 64);
      const uint64_t idesc64 = umma_desc_sw128(smem_base + WL_OFF_I64);
      uint32_t it = 0;
      uint32_t n = 0;  // local tile counter
      long long t_full = 0, t_epi = 0, t_begin = 0;
      if (timing) t_begin = clock64();
      auto wait_full = [&]() -> uint32_t {
        const int s = it % WL_STAGES;
        const uint32_t ph = (it / WL_STAGES) & 1;
        long long t0 = 0;
        if (timing) t0 = clock64();
        mbar_wait(full_bar(s), ph);
        if (timing) t_full += clock64() - t0;
        tc_fence_after();
        return smem_base + s * WL_STAGE_BYTES;
      };
      auto wait_epi = [&](uint32_t bar, uint32_t ph) {
        long long t0 = 0;
        if (timing) t0 = clock64();
        mbar_wait(bar, ph);
        if (timing) t_epi += clock64() - t0;
        tc_fence_after();
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n) {
        const uint32_t par = LAST ? 0u : (n & 1u);
        const uint32_t prev_ph = (n - 1) & 1u;
        for (int q = 0; q < 2; ++q) {
          const uint32_t d_tmem = tmem_base + 256u * (q == 0 ? par : (par ^ 1u));
          if (n > 0) {
            if (LAST) wait_epi(drained_bar(q), prev_ph);   // this region still holds tile n-1's chunk q
            else if (q == 1) wait_epi(epi2_bar, prev_ph);  // R[p^1] held GEMM2 of tile n-1
          }
          for (int kb = 0; kb < kb1; ++kb, ++it) {
            const uint32_t a_addr = wait_full();
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + WL_A_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < WL_BK / 16; ++k) {
                if (FIRST && kb == 0 && k == 3) break;   // columns 48..63 of a0 are the residual operand (GEMM2)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              }
              tc_commit(empty_bar(it % WL_STAGES));
              if (kb == kb1 - 1) tc_commit(dfull_bar(q));
            }
            __syncwarp();
          }
        }
        if (FIRST) {
          const uint32_t d_tmem = tmem_base + 256u * par;
          wait_epi(actsa_bar, n & 1u);   // acts blocks 0,1 written, D1a (this region) read out
          for (int step = 0; step < 5; ++step, ++it) {
            if (step == 3) wait_epi(acts2_bar, n & 1u);  // acts block 2 written
            if (step == 4) wait_epi(acts_bar, n & 1u);   // acts block 3 written, D1b read out
            const uint32_t st_addr = wait_full();
            if (elect_one()) {
              if (step == 0) {
                // h0 = [a(l), 1] @ [Wstart; bstart]: K columns 48..63 of the a0 tile against H0
                umma_bf16(d_tmem, umma_desc_sw128(st_addr) + 6, umma_desc_sw128(st_addr + WL_A_BYTES) + 6, idesc, 0u);
              } else {
                const int kb = step - 1;
                const uint64_t adesc = umma_desc_sw128(smem_base + WL_OFF_ACTS + kb * WL_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(st_addr + WL_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
              }
              tc_commit(empty_bar(it % WL_STAGES));
              if (step == 4) tc_commit(dfull_bar(2));
            }
            __syncwarp();
          }
        } else if (!LAST) {
          const uint32_t d_tmem = tmem_base + 256u * par;
          wait_epi(actsa_bar, n & 1u);   // acts blocks 0,1 written, D1a (this region) read out
          for (int step = 0; step < 6; ++step, ++it) {
            if (step == 4) wait_epi(acts2_bar, n & 1u);  // acts block 2 written (all epilogue warps work on it first)
            if (step == 5) wait_epi(acts_bar, n & 1u);   // acts block 3 written, D1b read out
            const uint32_t st_addr = wait_full();
            if (elect_one()) {
              if (step == 2 || step == 3) {
                // lo blocks 2j, 2j+1 -> identity add into their 64-column slices
                const int kb = (step - 2) * 2;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const uint64_t adesc = umma_desc_sw128(st_addr + h2 * WL_A_BYTES);
#pragma unroll
                  for (int k = 0; k < WL_BK / 16; ++k)
                    umma_bf16(d_tmem + 64u * (kb + h2), adesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
                }
              } else {
                const int kb = step < 2 ? step : step - 2;
                const uint64_t adesc = umma_desc_sw128(smem_base + WL_OFF_ACTS + kb * WL_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(st_addr + WL_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (step | k) ? 1u : 0u);
                const uint64_t hdesc = umma_desc_sw128(st_addr);   // hi block kb (centre tap rows)
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma_bf16(d_tmem + 64u * kb, hdesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
              }
              tc_commit(empty_bar(it % WL_STAGES));
              if (step == 5) tc_commit(dfull_bar(2));
            }
            __syncwarp();
          }
        }
      }
      if (timing && lane == 0) {
        atomicAdd(p.timing + 0, static_cast<unsigned long long>(clock64() - t_begin));
        atomicAdd(p.timing + 1, static_cast<unsigned long long>(t_full));
        atomicAdd(p.timing + 2, static_cast<unsigned long long>(t_epi));
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int we = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int hf = we >> 2;             // which half of the columns this warp handles
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* acts = smem + WL_OFF_ACTS;
    const bool tmr = timing && we == 0 && lane == 0;
    long long t_w0 = 0, t_w1 = 0, t_w2 = 0, t_e1 = 0, t_e2 = 0;
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n) {
      int b, r, t0;
      tile_coords(tile, b, r, t0);
      const uint32_t par = LAST ? 0u : (n & 1u);
      const uint32_t ph = n & 1u;
      const bool valid = wn_row_valid(p, t0 + row);
      const size_t m = (static_cast<size_t>(b) * p.R + r) * p.T + t0 + row;
      // (even-channel, odd-channel) partial sums of the eight fold columns, one accumulator per 16-channel column class
      // this thread owns ((channel mod 64) / 16 = 2 hf and 2 hf + 1): the classes are summed separately and then pairwise,
      // ((S0 + S1) + (S2 + S3)), so that every layer-kernel variant (CTA pair, 8 or 16 epilogue warps) produces the same bits
      float2 o8p[8], o8q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8p[j] = o8q[j] = make_float2(0.f, 0.f);

      // ---- gate epilogue: chunk q holds gate channels [128 q, 128 q + 128) ---------------------
      // (loops deliberately NOT fully unrolled: the kernel must stay inside the instruction cache)
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        long long tw0 = 0;
        if (tmr) tw0 = clock64();
        mbar_wait(dfull_bar(q), ph);
        tc_fence_after();
        long long tw1 = 0;
        if (tmr) { tw1 = clock64(); (q == 0 ? t_w0 : t_w1) += tw1 - tw0; }
        // All eight warps work on the same 64-channel K-block (two blocks per chunk, one after the other):
        // thread = (row, hf) handles channels [32 hf, 32 hf + 32) of the block in two steps of 16. The first
        // block of chunk b is therefore complete half-way through this epilogue and GEMM2 can consume it
        // while the second is still being gated.
        const uint32_t taddr = tmem_base + lane_addr + 256u * (q == 0 ? par : (par ^ 1u)) + hf * 32;
        uint32_t t0r[16], g0r[16], t1r[16], g1r[16];
        tmem_ld16(taddr, t0r);
        tmem_ld16(taddr + 128, g0r);
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          uint8_t* kblk = acts + (q * 2 + blk) * WL_A_BYTES + row * 128;   // K-block (64 channels) row
          const int ch0 = blk * 64 + hf * 32;                                // first channel inside the chunk
          const float* bT0 = s_b1 + q * 256 + ch0;
          const float2* wse0 = reinterpret_cast<const float2*>(cw.wse) + (q * 128 + ch0) * 4;   // [channel pair][8]
          tmem_ld_wait();
          tmem_ld16(taddr + blk * 64 + 16, t1r);
          tmem_ld16(taddr + 128 + blk * 64 + 16, g1r);
          gate_step2<LAST>(t0r, g0r, bT0, wse0, kblk, hf * 2, row, o8p);
          tmem_ld_wait();
          if (blk == 0) {
            tmem_ld16(taddr + 64, t0r);
            tmem_ld16(taddr + 128 + 64, g0r);
          }
          gate_step2<LAST>(t1r, g1r, bT0 + 16, wse0 + 64, kblk, hf * 2 + 1, row, o8q);
          if (!LAST && q == 1 && blk == 0) {
            fence_proxy_async_smem();
            mbar_arrive(acts2_bar);
          }
        }
        tc_fence_before();
        if (LAST) {
          mbar_arrive(drained_bar(q));
        } else {
          fence_proxy_async_smem();   // acts (generic-proxy writes) -> visible to the MMA (async proxy)
          mbar_arrive(q == 0 ? actsa_bar : acts_bar);
        }
        if (tmr) t_e1 += clock64() - tw1;
      }
      // fold accumulator: even + odd channels, then the two column halves of a row, in a fixed order (bit-reproducible)
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o8[j] = o8p[j].x + o8p[j].y;
        o8[j] += o8q[j].x + o8q[j].y;
      }
      if (hf == 1) {
        *reinterpret_cast<float4*>(s_o8 + row * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(s_o8 + row * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(WL_EPI_THREADS) : "memory");
      if (hf == 0 && valid) {
        const float4 p0 = *reinterpret_cast<const float4*>(s_o8 + row * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(s_o8 + row * 8 + 4);
        float4* o = reinterpret_cast<float4*>(p.acc8 + m * 8);
        float4 a0 = o[0], a1 = o[1];
        a0.x += o8[0] + p0.x; a0.y += o8[1] + p0.y; a0.z += o8[2] + p0.z; a0.w += o8[3] + p0.w;
        a1.x += o8[4] + p1.x; a1.y += o8[5] + p1.y; a1.z += o8[6] + p1.z; a1.w += o8[7] + p1.w;
        o[0] = a0; o[1] = a1;
      }

      // ---- residual epilogue: h = GEMM2 + hi + lo (already in the accumulator) + b2 -> (hi, lo) ----
      // Staged through the (now free) acts tile in the TMA SWIZZLE_128B layout and written with TMA
      // stores (rows beyond L are clipped by the tensor map). Two passes over the accumulator: hi, then lo.
      if (!LAST) {
        long long tw0 = 0;
        if (tmr) tw0 = clock64();
        mbar_wait(dfull_bar(2), ph);
        tc_fence_after();
        long long tw1 = 0;
        if (tmr) { tw1 = clock64(); t_w2 += tw1 - tw0; }
        const uint32_t taddr = tmem_base + lane_addr + 256u * par + hf * 128;
        uint8_t* stg = acts + (hf * 2) * WL_A_BYTES + row * 128;      // this half's two 64-column blocks
        const uint32_t stg_addr = smem_base + WL_OFF_ACTS + (hf * 2) * WL_A_BYTES;
        const bool issuer = (we == hf * 4) && lane == 0;
        auto resid_pass = [&](auto pass_tag) {
          constexpr int pass = decltype(pass_tag)::value;
          if (pass == 1) {
            if (issuer) bulk_wait_read0();   // the hi store has finished reading the staging tile
            if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
            else asm volatile("bar.sync 4, 128;" ::: "memory");
          }
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr, r0);
#pragma unroll 1
          for (int gp = 0; gp < 4; ++gp) {          // two groups of 16 columns per iteration
            tmem_ld_wait();
            tmem_ld16(taddr + (2 * gp + 1) * 16, r1);
            resid_step2<pass>(r0, s_b2 + hf * 128 + (2 * gp) * 16, stg, 2 * gp, row, valid);
            tmem_ld_wait();
            if (gp < 3) tmem_ld16(taddr + (2 * gp + 2) * 16, r0);
            else if (pass == 1) {
              tc_fence_before();
              mbar_arrive(epi2_bar);   // all TMEM reads of this tile are done
            }
            resid_step2<pass>(r1, s_b2 + hf * 128 + (2 * gp + 1) * 16, stg, 2 * gp + 1, row, valid);
          }
          fence_proxy_async_smem();
          if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          if (issuer) {
            const CUtensorMap* om = pass == 0 ? &map_ho : &map_lo;
            tma_store_4d(om, stg_addr, (hf * 2) * WL_BK, t0, pm ? r : b, pm ? b : 0);
            tma_store_4d(om, stg_addr + WL_A_BYTES, (hf * 2 + 1) * WL_BK, t0, pm ? r : b, pm ? b : 0);
            bulk_commit();
          }
        };
        resid_pass(std::integral_constant<int, 0>{});
        resid_pass(std::integral_constant<int, 1>{});
        if (issuer) bulk_wait_read0();   // staging tile may be overwritten by the next gate epilogue
        if (tmr) t_e2 += clock64() - tw1;
      }
      // s_o8 and the staging tile are reused by the next tile: no epilogue warp may run ahead
      asm volatile("bar.sync 1, %0;" ::"n"(WL_EPI_THREADS) : "memory");
    }
    if (!LAST && lane == 0 && (we == 0 || we == 4)) bulk_wait0();   // all TMA stores of this CTA have landed
    if (tmr) {
      atomicAdd(p.timing + 3, static_cast<unsigned long long>(t_w0));
      atomicAdd(p.timing + 4, static_cast<unsigned long long>(t_w1));
      atomicAdd(p.timing + 5, static_cast<unsigned long long>(t_w2));
      atomicAdd(p.timing + 6, static_cast<unsigned long long>(t_e1));
      atomicAdd(p.timing + 7, static_cast<unsigned long long>(t_e2));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// Small helper kernels
// ================================================================================================
// A operand of the polyphase upsample GEMM / the phase-major conditioning: row (b, t) of ONE phase block (RowGeom with
// R = 1: utterances Tp rows apart, or the ragged tables) holds aup[row, j*n_mel + i] = bf16(mel[b, t-j, i]), 0 for t < j
// and for gap rows. mel is the caller's padded [B, T, n_mel].
__global__ void upsample_im2col_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ aup, const RowGeom geo,
                                       int n_mel, int Kup) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(geo.rows_per_phase()) * Kup;
  if (idx >= total) return;
  const int kk = static_cast<int>(idx % Kup);
  const int row = static_cast<int>(idx / Kup);
  int b, t;
  const bool valid = geo.decode_row(row, b, t);
  const int j = kk / n_mel, i = kk - j * n_mel;
  float v = 0.f;
  if (valid && j < 4 && t - j >= 0) v = mel[(static_cast<size_t>(b) * geo.T + t - j) * n_mel + i];
  aup[idx] = __float2bfloat16_rn(v);
}

// weight prep: out[n*K + k] = bf16(in[k*N + n])  (fp32 [K,N] -> bf16 K-major [N,K])
__global__ void transpose_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int K, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int n = idx / K, k = idx - n * K;
  out[idx] = __float2bfloat16_rn(in[static_cast<size_t>(k) * N + n]);
}

// weight prep of the folded conditioning weights: in [(r, k), n] fp32 (one GEMM over all phases) ->
// out[(r*N + n)*K + k] = bf16(in[(r*K + k)*N + n])   (per phase a K-major [N, K] matrix)
__global__ void fold_store_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int K, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(R) * K * N) return;
  const int k = static_cast<int>(idx % K);
  const size_t rn = idx / K;
  const int n = static_cast<int>(rn % N), r = static_cast<int>(rn / N);
  out[idx] = __float2bfloat16_rn(in[(static_cast<size_t>(r) * K + k) * N + n]);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = __bfloat162float(in[idx]);
}

// ================================================================================================
// Host side: tensor maps + launches
// ================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled& encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  return fn;
}

inline void tc_init() {
  if (encode_fn()) return;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qres;
  WG_CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
  if (!f || qres != cudaDriverEntryPointSuccess) fail(WG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  encode_fn() = reinterpret_cast<PFN_encodeTiled>(f);
  WG_CK(cudaFuncSetAttribute(tc_gemm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
  WG_CK(cudaFuncSetAttribute(tc_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL_SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL_SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_layer_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL_SMEM));
}

// bf16 tensor map, innermost dim contiguous, SWIZZLE_128B, box inner = 64 elements (128 B).
inline void make_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(WG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

inline void make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows}, str[1] = {cols * 2};
  const uint32_t box[2] = {64, box_rows};
  make_map(m, ptr, 2, dims, str, box);
}

inline void make_map_3d(CUtensorMap* m, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const uint64_t dims[3] = {cols, rows, batch}, str[2] = {cols * 2, rows * cols * 2};
  const uint32_t box[3] = {64, box_rows, 1};
  make_map(m, ptr, 3, dims, str, box);
}

inline void make_map_4d(CUtensorMap* m, const void* ptr, uint64_t batch, uint64_t phases, uint64_t rows, uint64_t cols,
                        uint32_t box_rows) {
  const uint64_t dims[4] = {cols, rows, phases, batch};
  const uint64_t str[3] = {cols * 2, rows * cols * 2, phases * rows * cols * 2};
  const uint32_t box[4] = {64, box_rows, 1, 1};
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 3; ++i) gstr[i] = str[i];
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(WG_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
}

struct TcPlan {
  // position-major maps used by the upsample GEMM and the CTA-pair kernel
  CUtensorMap m_aup, m_wup, m_w1, m_w2, m_h16[2], m_lo, m_spect;
  // (phase, frame) geometry of the single-CTA layer kernel: 4-D maps (channels, rows, phases, batch)
  CUtensorMap m4_h[2], m4_lo, m4_cond, m_wc;
  // start-conv fold (FIRST layer variant, phase-major only): a0 [M, 64], W0 [n_flows*2C, 64], H0 [n_flows*C, 64]
  CUtensorMap m4_a0{}, m_w0{}, m_h0{};
  bool fold0 = false;
  bool pdl = true;          // WG_PDL=0 switches programmatic dependent launch of the layer kernels off (A/B)
  int n_layers = 0;
  int tile_order = 1;
  int Breal = 0, Treal = 0, Tp = 0;   // gap layout: B utterances of T frames, Tp rows apart inside a phase block (pm only)
  RowGeom geo1{};                     // one phase block (R = 1) in the caller's utterance geometry: the mel-window rows
  bool pm = false;           // phase-major layout (R = 32) with the rank-320 conditioning
  int R = 1, Trows = 0, tiles_per_row = 0;
  int n_cond_kb = 0, wc_col0 = 0, wc_rows_per_layer = 0, wc_rstride = 0;
  int sm_count = 0, B = 0, T = 0, L = 0, C = 0, S = 0, Kup = 0, n_mel = 0, NupN = 0;
  int tiles_per_b = 0, n_tiles = 0;
  __nv_bfloat16 *aup16 = nullptr, *spect16 = nullptr, *h16[2] = {nullptr, nullptr}, *hlo = nullptr;
};

// `pm`: phase-major layout. `V` = folded conditioning weights [n_layers_total * R * 2C, Kup] (pm only).
inline void tc_prepare(TcPlan& pl, int sm_count, int B, int T, int L, int C, int S, int Kup, int n_mel,
                       int n_layers_total, const __nv_bfloat16* Wup16, int NupN, const __nv_bfloat16* W1,
                       const __nv_bfloat16* W2, __nv_bfloat16* aup16, __nv_bfloat16* spect16, __nv_bfloat16* h16a,
                       __nv_bfloat16* h16b, __nv_bfloat16* hlo, bool pm, int R, const __nv_bfloat16* V,
                       __nv_bfloat16* a0 = nullptr, const __nv_bfloat16* W0 = nullptr, const __nv_bfloat16* H0 = nullptr,
                       int n_flows = 0, int gap = 0, const RowGeom* ragged = nullptr) {
  if ((C != 256 && C != 512) || S != WL_S) fail(WG_ERR_UNSUPPORTED, "tensor path is built for C in {256, 512}, S=640 (got C=%d, S=%d)", C, S);
  pl.sm_count = sm_count; pl.B = B; pl.T = T; pl.L = L; pl.C = C; pl.S = S; pl.Kup = Kup; pl.n_mel = n_mel; pl.NupN = NupN;
  pl.aup16 = aup16; pl.spect16 = spect16; pl.h16[0] = h16a; pl.h16[1] = h16b; pl.hlo = hlo;
  pl.pm = pm;
  pl.R = pm ? R : 1;
  // Phase-major: ONE row sequence per phase holding all utterances Tp = T + gap rows apart (RowGeom, simt_kernels.cuh);
  // the kernels see a single "utterance" of B*Tp frames whose gap rows are kept zero, and tiles may span utterances.
  pl.Breal = B; pl.Treal = T; pl.Tp = pm ? T + gap : 0;
  pl.Trows = pm ? B * pl.Tp : L;
  pl.geo1 = RowGeom{1, T, pm ? pl.Tp : T, B};
  if (ragged) {   // per-utterance lengths: phase-major only, rows per phase block from the tables
    if (!pm) fail(WG_ERR_INVALID, "ragged batches need the phase-major layout");
    pl.Trows = ragged->rpp;
    pl.geo1 = *ragged;
    pl.geo1.R = 1;
  }
  const int Bk = pm ? 1 : B;     // batch extent the layer kernels iterate over
  pl.tiles_per_row = (pl.Trows + WL_BM - 1) / WL_BM;
  pl.n_tiles = pl.tiles_per_row * pl.R * Bk;
  pl.tiles_per_b = pl.tiles_per_row;
  make_map_2d(&pl.m_w1, W1, (uint64_t)n_layers_total * 2 * C, 3 * C + S, 256);
  make_map_2d(&pl.m_w2, W2, (uint64_t)n_layers_total * C, C, 256);
  // phase-major: (channels, frames, phases, batch); position-major: (channels, rows, batch, 1)
  const uint64_t d3 = 1, d2 = pm ? (uint64_t)R : (uint64_t)B;
  make_map_4d(&pl.m4_h[0], h16a, d3, d2, pl.Trows, C, WL_BM);
  make_map_4d(&pl.m4_h[1], h16b, d3, d2, pl.Trows, C, WL_BM);
  make_map_4d(&pl.m4_lo, hlo, d3, d2, pl.Trows, C, WL_BM);
  if (pm) {
    if (Kup % WL_BK) fail(WG_ERR_UNSUPPORTED, "mel window K (%d) must be a multiple of 64", Kup);
    make_map_4d(&pl.m4_cond, aup16, 1, 1, pl.Trows, Kup, WL_BM);
    make_map_2d(&pl.m_wc, V, (uint64_t)n_layers_total * R * 2 * C, Kup, 256);
    pl.n_cond_kb = Kup / WL_BK; pl.wc_col0 = 0; pl.wc_rows_per_layer = R * 2 * C; pl.wc_rstride = 2 * C;
    pl.fold0 = a0 && W0 && H0 && n_flows > 0;
    if (pl.fold0) {
      pl.n_layers = n_layers_total / n_flows;
      make_map_4d(&pl.m4_a0, a0, 1, R, pl.Trows, WL_BK, WL_BM);
      make_map_2d(&pl.m_w0, W0, (uint64_t)n_flows * 2 * C, WL_BK, 256);
      make_map_2d(&pl.m_h0, H0, (uint64_t)n_flows * C, WL_BK, 256);
    }
  } else {
    make_map_2d(&pl.m_aup, aup16, (uint64_t)B * T, Kup, TG_BM);
    make_map_2d(&pl.m_wup, Wup16, NupN, Kup, TG_BN);
    make_map_4d(&pl.m4_cond, spect16, 1, B, L, S, WL_BM);
    make_map_2d(&pl.m_wc, W1, (uint64_t)n_layers_total * 2 * C, 3 * C + S, 256);
    pl.n_cond_kb = S / WL_BK; pl.wc_col0 = 3 * C; pl.wc_rows_per_layer = 2 * C; pl.wc_rstride = 0;
    // 3-D maps for the CTA-pair kernel
    make_map_3d(&pl.m_h16[0], h16a, B, L, C, WL_BM);
    make_map_3d(&pl.m_h16[1], h16b, B, L, C, WL_BM);
    make_map_3d(&pl.m_lo, hlo, B, L, C, WL_BM);
    make_map_3d(&pl.m_spect, spect16, B, L, S, WL_BM);
  }
}

// A operand of the conditioning: 4-frame mel window (always), plus the polyphase upsample GEMM when the
// position-major path needs the materialised spect.
inline int tc_upsample(const TcPlan& pl, const float* mel, const float* bup, cudaStream_t st) {
  const size_t total = (size_t)pl.geo1.rows_per_phase() * pl.Kup;
  upsample_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mel, pl.aup16, pl.geo1, pl.n_mel, pl.Kup);
  WG_CK(cudaGetLastError());
  if (pl.pm) return 1;
  const int M = pl.B * pl.T;
  dim3 grid(pl.NupN / TG_BN, (M + TG_BM - 1) / TG_BM);
  tc_gemm_kernel<__nv_bfloat16><<<grid, TG_THREADS, TG_SMEM, st>>>(pl.m_aup, pl.m_wup, bup, pl.spect16, M, pl.NupN, pl.Kup);
  WG_CK(cudaGetLastError());
  return 2;
}

inline void tc_fill_params(const TcPlan& pl, WnLayerParams& p, int layer, int dilation, int hcur, float* acc8,
                           const float* b1, const float* b2, unsigned long long* timing, int flags) {
  p.T = pl.Trows; p.R = pl.R; p.tiles_per_row = pl.tiles_per_row; p.n_tiles = pl.n_tiles;
  p.Tp = pl.pm ? pl.Tp : 0; p.Tv = pl.Treal;
  p.row_b = pl.geo1.row_b;
  p.tile_order = pl.pm && pl.tile_order ? 1 : 0;
  p.L = pl.Trows; p.tiles_per_b = pl.tiles_per_row;
  p.n_cond_kb = pl.n_cond_kb; p.wc_col0 = pl.wc_col0; p.wc_row0 = layer * pl.wc_rows_per_layer; p.wc_rstride = pl.wc_rstride;
  p.layer = layer; p.dilation = dilation;
  p.flow = pl.n_layers > 0 ? layer / pl.n_layers : 0;
  p.b1 = b1; p.b2 = b2; p.hi_out = pl.h16[hcur ^ 1]; p.lo = pl.hlo; p.acc8 = acc8; p.timing = timing; p.flags = flags;
}

template <typename... KArgs, typename... Args>
inline void tc_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, const Args&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  WG_CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}

inline int tc_wn_layer(const TcPlan& pl, int layer, int dilation, bool last, int hcur, float* acc8, const float* b1,
                       const float* b2, const float* wse_host, unsigned long long* timing, int flags, cudaStream_t st,
                       bool first = false) {
  WnLayerParams p{};
  tc_fill_params(pl, p, layer, dilation, hcur, acc8, b1, b2, timing, flags);
  WnLayerConst cw;   // wse_host: packed-fold layout [channel pair][column][even, odd] (LayerW::wse_p)
  std::memcpy(cw.wse, wse_host, sizeof cw.wse);
  const int grid = pl.n_tiles < pl.sm_count ? pl.n_tiles : pl.sm_count;
  if (first && (!pl.fold0 || last || dilation != 1)) fail(WG_ERR_INVALID, "start fold requested for a layer it does not apply to");
  // programmatic dependent launch only where it can pay: a grid that leaves SMs idle (single-utterance calls)
  p.pdl = pl.pdl && grid < pl.sm_count ? 1 : 0;
  if (last)
    tc_launch(tc_wn_layer_kernel<true, false>, grid, WL_THREADS, WL_SMEM, st, p.pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pl.m_w1, pl.m_wc, pl.m_w2, pl.m4_a0, pl.m_w0, pl.m_h0, p, cw);
  else if (first)
    tc_launch(tc_wn_layer_kernel<false, true>, grid, WL_THREADS, WL_SMEM, st, p.pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pl.m_w1, pl.m_wc, pl.m_w2, pl.m4_a0, pl.m_w0, pl.m_h0, p, cw);
  else
    tc_launch(tc_wn_layer_kernel<false, false>, grid, WL_THREADS, WL_SMEM, st, p.pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pl.m_w1, pl.m_wc, pl.m_w2, pl.m4_a0, pl.m_w0, pl.m_h0, p, cw);
  WG_CK(cudaGetLastError());
  return 1;
}

// debug: h = hi + lo as float32, rows re-ordered from the internal order (RowGeom) to position-major (b, R*t + r)
__global__ void hilo_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                   float* __restrict__ out, const RowGeom geo, int C) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(geo.rows()) * C) return;
  const int m = static_cast<int>(idx / C);
  const int c = static_cast<int>(idx - static_cast<size_t>(m) * C);
  int b, r, t;
  if (!geo.decode(m, b, r, t)) return;
  out[geo.position_major(b, r, t) * C + c] = __bfloat162float(hi[idx]) + __bfloat162float(lo[idx]);
}

// debug: float rows [rows, W] re-ordered the same way; `parts` partial buffers `stride` floats apart are summed
__global__ void unpermute_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const RowGeom geo, int W,
                                      int parts = 1, size_t stride = 0) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(geo.rows()) * W) return;
  const int m = static_cast<int>(idx / W);
  const int c = static_cast<int>(idx - static_cast<size_t>(m) * W);
  int b, r, t;
  if (!geo.decode(m, b, r, t)) return;
  float v = in[idx];
  for (int pp = 1; pp < parts; ++pp) v += in[pp * stride + idx];
  out[geo.position_major(b, r, t) * W + c] = v;
}

// debug (tf32x3 mode): h = hi + lo, both fp32
__global__ void hilo_f32_to_f32_kernel(const float* __restrict__ hi, const float* __restrict__ lo, float* __restrict__ out,
                                       const RowGeom geo, int C) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(geo.rows()) * C) return;
  const int m = static_cast<int>(idx / C);
  const int c = static_cast<int>(idx - static_cast<size_t>(m) * C);
  int b, r, t;
  if (!geo.decode(m, b, r, t)) return;
  out[geo.position_major(b, r, t) * C + c] = hi[idx] + lo[idx];
}

inline void tc_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t st) {
  bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
}

inline void tc_debug_gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* D, int M, int N,
                          int K, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0 || N % TG_BN || K % TG_BK)
    fail(WG_ERR_INVALID, "debug GEMM needs N %% 256 == 0 and K %% 64 == 0 (got M=%d N=%d K=%d)", M, N, K);
  CUtensorMap ma, mw;
  make_map_2d(&ma, A, M, K, TG_BM);
  make_map_2d(&mw, W, N, K, TG_BN);
  dim3 grid(N / TG_BN, (M + TG_BM - 1) / TG_BM);
  tc_gemm_kernel<float><<<grid, TG_THREADS, TG_SMEM, st>>>(ma, mw, bias, D, M, N, K);
  WG_CK(cudaGetLastError());
}

}  // namespace wg

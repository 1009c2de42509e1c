// Fused WN layer (C = 256) on a CTA PAIR -- tcgen05 cta_group::2 -- for the gapped phase-major layout.
//
// Same computation, same bits per row as tc_wn_layer_kernel (tc_kernels.cuh); what changes is who holds the operands.
// Two SMs of one TPC form a cluster and issue ONE 256-row MMA per instruction: each CTA owns one 128-row tile (its own
// A tiles, its own TMEM accumulators, its own epilogue) of the SAME upsample phase -- so both use the same weights --
// and the B operand (weights) is split in halves across the two CTAs' shared memory and read by both tensor cores.
// Per CTA a pipeline stage is 16 KB of A + 16 KB of B instead of 16 + 32 KB: a tile needs 1.25 MB through the SM's TMA
// port instead of 1.9 MB (the single-CTA kernel is bounded by that port: 67 B/clk/SM, profiles/r01_probes.md), and the
// same shared memory holds a 4-deep ring instead of 3.
//
// Roles per CTA: warp 0 TMA producer (own A tile + own half of B; the bytes of BOTH CTAs complete on the LEADER's
// full barrier through the .cta_group::2 TMA form), warp 1 TMEM allocator and -- in the leader CTA (cluster rank 0)
// only -- the MMA issuer; warps 2..9 epilogue. Epilogue -> MMA barriers live in the leader and count the epilogue
// threads of both CTAs (remote arrive); MMA -> epilogue and MMA -> producer barriers are signalled in both CTAs by
// multicast tcgen05.commit.
#pragma once
#include "tc_kernels.cuh"

namespace wg {

// ---- cluster / cta_group::2 PTX wrappers ----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs are done
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of CTA 1 -> same offset in CTA 0

// ---- geometry -------------------------------------------------------------------------------------
constexpr int WP_A_BYTES = WL_BM * WL_BK * 2;          // 16 KB: own 128 rows x 64 K
constexpr int WP_B_BYTES = 128 * WL_BK * 2;            // 16 KB: own half (128 of 256 N rows) x 64 K
constexpr int WP_STAGE_BYTES = WP_A_BYTES + WP_B_BYTES;
// The LAST layer of a flow has no GEMM2 / residual, hence no acts tile, identity tile or staging: its ring can be 6 deep.
// EW = epilogue warps (8 or 16): NQ = EW / 4 warps share a TMEM lane quarter and split the accumulator columns.
template <bool LAST, int EW = 8>
struct WpGeom {
  static constexpr int NQ = EW / 4;
  static constexpr int EPI_THREADS = EW * 32, THREADS = 64 + EPI_THREADS;
  static constexpr int STAGES = LAST ? 6 : 4;
  static constexpr int OFF_ACTS = STAGES * WP_STAGE_BYTES;
  static constexpr int OFF_I64 = OFF_ACTS + (LAST ? 0 : WL_ACTS_BYTES);   // own half (32 N rows) of the 64x64 identity
  static constexpr int OFF_B1 = OFF_I64 + (LAST ? 0 : 32 * 128);
  static constexpr int OFF_B2 = OFF_B1 + 2 * WL_C * 4;
  static constexpr int OFF_O8 = OFF_B2 + WL_C * 4;
  static constexpr int OFF_BARS = OFF_O8 + (NQ - 1) * WL_BM * 8 * 4;   // fold partials of column groups 1 .. NQ-1
  static constexpr int NBARS = 2 * STAGES + 3 + 2 + 4;
  static constexpr int SMEM = OFF_BARS + NBARS * 8 + 16;
  static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <bool LAST, bool FIRST = false, int EW = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EW * 32, 1)
tc_wn_pair_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_ho,
                  const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_cond,
                  const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_wc,
                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_a0,
                  const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_h0, const WnLayerParams p,
                  const __grid_constant__ WnLayerConst cw) {
  static_assert(!(LAST && FIRST), "the start fold needs a residual layer");
  using G = WpGeom<LAST, EW>;
  constexpr int STAGES = G::STAGES, NQ = G::NQ, EPI_THREADS = G::EPI_THREADS, THREADS = G::THREADS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + G::OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + G::OFF_B2);
  float* s_o8 = reinterpret_cast<float*>(smem + G::OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G::NBARS);
  const uint32_t bar_base = smem_base + G::OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                           // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };               // both CTAs (multicast commit)
  auto dfull_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + i); };           // both CTAs (multicast commit)
  auto drained_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 3 + i); };     // leader only, 2 x epilogue threads
  const uint32_t actsa_bar = bar_base + 8u * (2 * STAGES + 5);                        // leader only
  const uint32_t acts_bar = bar_base + 8u * (2 * STAGES + 6);                         // leader only
  const uint32_t epi2_bar = bar_base + 8u * (2 * STAGES + 7);                         // leader only
  const uint32_t acts2_bar = bar_base + 8u * (2 * STAGES + 8);                        // leader only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((smem_base & 1023u) != 0u) __trap();
  if (p.pdl) pdl_trigger();

  // pair-tile pt -> phase r (fastest, as in the single-CTA kernel's tile order) and the pair's row range; this CTA's
  // tile is the rank-th 128-row tile of it. With an odd tile count per phase block the last pair has a GHOST tile
  // (t0 >= T): every TMA box of it is out of bounds (zeros in, nothing out) and none of its rows is valid.
  const int pairs_per_row = (p.tiles_per_row + 1) / 2;
  const int n_pair_tiles = pairs_per_row * p.R;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  auto tile_coords = [&](int pt, int& r, int& t0) {
    r = pt % p.R;
    t0 = (2 * (pt / p.R) + static_cast<int>(rank)) * WL_BM;
  };

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
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(dfull_bar(i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(drained_bar(i), 2 * EPI_THREADS);
    mbar_init(actsa_bar, 2 * EPI_THREADS);
    mbar_init(acts_bar, 2 * EPI_THREADS);
    mbar_init(epi2_bar, 2 * EPI_THREADS);
    mbar_init(acts2_bar, 2 * EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_slot), 512);
    tmem2_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * WL_C; i += THREADS) s_b1[i] = p.b1[i];
  if (!LAST) {
    for (int i = threadIdx.x; i < WL_C; i += THREADS) s_b2[i] = p.b2[i];
  }
  if (!LAST && !FIRST) {
    // own half of the identity B tile: local row nl is N row n = 32 * rank + nl; element (n, k) = [n == k]
    // (K-major SWIZZLE_128B: 16-byte chunk c of row nl sits at c ^ (nl & 7))
    uint32_t* i64w = reinterpret_cast<uint32_t*>(smem + G::OFF_I64);
    for (int i = threadIdx.x; i < 32 * 32; i += THREADS) {
      const int nl = i >> 5, w = i & 31;
      const int n = 32 * static_cast<int>(rank) + nl;
      const int chunk_log = (w >> 2) ^ (nl & 7);
      const int k0 = chunk_log * 8 + (w & 3) * 2;
      uint32_t v = 0;
      if (k0 == n) v = 0x00003F80u;
      if (k0 + 1 == n) v = 0x3F800000u;
      i64w[i] = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers / TMEM / constant tiles exist before anyone signals across
  tc_fence_after();
  if (p.pdl) pdl_wait();   // from here on the previous kernel's results are read
  const uint32_t tmem_base = *tmem_slot;
  const bool timing = p.timing != nullptr;   // WG_LAYER_TIMING=1: in-kernel cycle counters (same slots as the single-CTA kernel)
  constexpr int KB_CONV = FIRST ? 1 : WL_KB_CONV;
  const int kb1 = KB_CONV + p.n_cond_kb;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ============================
    uint32_t it = 0;
    // `pair_bytes`: what BOTH CTAs load into this stage (the leader's barrier counts them all)
    long long t_wait = 0;
    auto acquire = [&](uint32_t pair_bytes) -> uint32_t {
      const int s = it % STAGES;
      long long tq = 0;
      if (timing) tq = clock64();
      mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
      if (timing) t_wait += clock64() - tq;
      if (leader && elect_one()) mbar_expect_tx(full_bar(s), pair_bytes);
      __syncwarp();
      return static_cast<uint32_t>(s);
    };
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs) {
      int r, t0;
      tile_coords(pt, r, t0);
      for (int q = 0; q < 2; ++q) {
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const uint32_t s = acquire(2 * WP_STAGE_BYTES);
          const uint32_t fb = full_bar(s) & kPeerBitMask;
          const uint32_t a_dst = smem_base + s * WP_STAGE_BYTES;
          if (elect_one()) {
            if (FIRST && kb == 0) {
              tma2_load_4d(a_dst, &map_a0, fb, 0, t0, r, 0);
              tma2_load_2d(a_dst + WP_A_BYTES, &map_w0, fb, 0, p.flow * 2 * WL_C + q * 256 + rank * 128);
            } else if (kb < KB_CONV) {
              const int tap = kb >> 2, cblk = kb & 3;
              const int rs = r + (tap - 1) * p.dilation;
              const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
              tma2_load_4d(a_dst, &map_h, fb, cblk * WL_BK, t0 + carry, rs - carry * p.R, 0);
              tma2_load_2d(a_dst + WP_A_BYTES, &map_w1, fb, kb * WL_BK, p.layer * 2 * WL_C + q * 256 + rank * 128);
            } else {
              const int kc = kb - KB_CONV;
              tma2_load_4d(a_dst, &map_cond, fb, kc * WL_BK, t0, 0, 0);
              tma2_load_2d(a_dst + WP_A_BYTES, &map_wc, fb, p.wc_col0 + kc * WL_BK,
                           p.wc_row0 + r * p.wc_rstride + q * 256 + rank * 128);
            }
          }
          __syncwarp();
        }
      }
      if (FIRST) {
        // consumption order of the MMA warp: a0 / H0 (residual operand) | W2 blocks 0..3 (A = acts in smem)
        for (int step = 0; step < 5; ++step, ++it) {
          const uint32_t s = acquire(step == 0 ? 2 * WP_STAGE_BYTES : 2 * WP_B_BYTES);
          const uint32_t fb = full_bar(s) & kPeerBitMask;
          const uint32_t dst = smem_base + s * WP_STAGE_BYTES;
          if (elect_one()) {
            if (step == 0) {
              tma2_load_4d(dst, &map_a0, fb, 0, t0, r, 0);
              tma2_load_2d(dst + WP_A_BYTES, &map_h0, fb, 0, p.flow * WL_C + rank * 128);
            } else {
              tma2_load_2d(dst + WP_A_BYTES, &map_w2, fb, (step - 1) * WL_BK, p.layer * WL_C + rank * 128);
            }
          }
          __syncwarp();
        }
      } else if (!LAST) {
        // consumption order of the MMA warp: W2/hi blocks 0,1 | lo blocks 0..3 | W2/hi blocks 2,3
        for (int step = 0; step < 6; ++step, ++it) {
          const uint32_t s = acquire(2 * WP_STAGE_BYTES);
          const uint32_t fb = full_bar(s) & kPeerBitMask;
          const uint32_t dst = smem_base + s * WP_STAGE_BYTES;
          if (elect_one()) {
            if (step == 2 || step == 3) {
              const int kb = (step - 2) * 2;
              tma2_load_4d(dst, &map_lo, fb, kb * WL_BK, t0, r, 0);
              tma2_load_4d(dst + WP_A_BYTES, &map_lo, fb, (kb + 1) * WL_BK, t0, r, 0);
            } else {
              const int kb = step < 2 ? step : step - 2;
              tma2_load_4d(dst, &map_h, fb, kb * WL_BK, t0, r, 0);
              tma2_load_2d(dst + WP_A_BYTES, &map_w2, fb, kb * WL_BK, p.layer * WL_C + rank * 128);
            }
          }
          __syncwarp();
        }
      }
    }
    if (timing && leader && lane == 0) atomicAdd(p.timing + 8, static_cast<unsigned long long>(t_wait));
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
      constexpr uint32_t idesc_id = umma_idesc_bf16(256, 64);
      const uint64_t idesc64 = umma_desc_sw128(smem_base + G::OFF_I64);
      uint32_t it = 0, n = 0;
      long long t_full = 0, t_epi = 0, t_begin = 0;
      if (timing) t_begin = clock64();
      auto wait_full = [&]() -> uint32_t {
        const int s = it % STAGES;
        long long tq = 0;
        if (timing) tq = clock64();
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        if (timing) t_full += clock64() - tq;
        tc_fence_after();
        return smem_base + s * WP_STAGE_BYTES;
      };
      auto wait_epi = [&](uint32_t bar, uint32_t ph) {
        long long tq = 0;
        if (timing) tq = clock64();
        mbar_wait(bar, ph);
        if (timing) t_epi += clock64() - tq;
        tc_fence_after();
      };
      for (int pt = pair; pt < n_pair_tiles; pt += n_pairs, ++n) {
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
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + WP_A_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < WL_BK / 16; ++k) {
                if (FIRST && kb == 0 && k == 3) break;   // columns 48..63 of a0 are the residual operand (GEMM2)
                umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              }
              tc2_commit(empty_bar(it % STAGES));
              if (kb == kb1 - 1) tc2_commit(dfull_bar(q));
            }
            __syncwarp();
          }
        }
        if (FIRST) {
          const uint32_t d_tmem = tmem_base + 256u * par;
          wait_epi(actsa_bar, n & 1u);   // acts blocks 0,1 written in both CTAs, D1a (this region) read out
          for (int step = 0; step < 5; ++step, ++it) {
            if (step == 3) wait_epi(acts2_bar, n & 1u);
            if (step == 4) wait_epi(acts_bar, n & 1u);
            const uint32_t st_addr = wait_full();
            if (elect_one()) {
              if (step == 0) {
                umma2_bf16(d_tmem, umma_desc_sw128(st_addr) + 6, umma_desc_sw128(st_addr + WP_A_BYTES) + 6, idesc, 0u);
              } else {
                const int kb = step - 1;
                const uint64_t adesc = umma_desc_sw128(smem_base + G::OFF_ACTS + kb * WL_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(st_addr + WP_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k) umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
              }
              tc2_commit(empty_bar(it % STAGES));
              if (step == 4) tc2_commit(dfull_bar(2));
            }
            __syncwarp();
          }
        } else if (!LAST) {
          const uint32_t d_tmem = tmem_base + 256u * par;
          wait_epi(actsa_bar, n & 1u);
          for (int step = 0; step < 6; ++step, ++it) {
            if (step == 4) wait_epi(acts2_bar, n & 1u);
            if (step == 5) wait_epi(acts_bar, n & 1u);
            const uint32_t st_addr = wait_full();
            if (elect_one()) {
              if (step == 2 || step == 3) {
                // lo blocks 2j, 2j+1 -> identity add into their 64-column slices
                const int kb = (step - 2) * 2;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const uint64_t adesc = umma_desc_sw128(st_addr + h2 * WP_A_BYTES);
#pragma unroll
                  for (int k = 0; k < WL_BK / 16; ++k)
                    umma2_bf16(d_tmem + 64u * (kb + h2), adesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
                }
              } else {
                const int kb = step < 2 ? step : step - 2;
                const uint64_t adesc = umma_desc_sw128(smem_base + G::OFF_ACTS + kb * WL_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(st_addr + WP_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (step | k) ? 1u : 0u);
                const uint64_t hdesc = umma_desc_sw128(st_addr);   // hi block kb (centre tap rows)
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma2_bf16(d_tmem + 64u * kb, hdesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
              }
              tc2_commit(empty_bar(it % STAGES));
              if (step == 5) tc2_commit(dfull_bar(2));
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
    // ======================================= epilogue (both CTAs) ==============================
    // NQ warps share a TMEM lane quarter; column group cq handles, of every 64-channel block, the 64 / NQ channels
    // [cq * 64 / NQ, ...) in STEPS steps of 16 ("column class" = step index st = (channel mod 64) / 16).
    constexpr int STEPS = 4 / NQ;
    const int we = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int cq = we >> 2;             // column group of this warp
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* acts = smem + G::OFF_ACTS;
    // epilogue -> MMA barriers live in the leader CTA
    const uint32_t r_drained0 = mapa_u32(drained_bar(0), 0), r_drained1 = mapa_u32(drained_bar(1), 0);
    const uint32_t r_actsa = mapa_u32(actsa_bar, 0), r_acts = mapa_u32(acts_bar, 0);
    const uint32_t r_epi2 = mapa_u32(epi2_bar, 0), r_acts2 = mapa_u32(acts2_bar, 0);
    const bool tmr = timing && leader && we == 0 && lane == 0;
    long long t_w0 = 0, t_w1 = 0, t_w2 = 0, t_e1 = 0, t_e2 = 0;
    uint32_t n = 0;
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs, ++n) {
      int r, t0;
      tile_coords(pt, r, t0);
      const uint32_t par = LAST ? 0u : (n & 1u);
      const uint32_t ph = n & 1u;
      const bool valid = wn_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      // one fold accumulator per column class this thread owns (summed per class, then pairwise: the same order
      // whatever NQ is, so every layer-kernel variant produces the same bits)
      float2 o8p[STEPS][8];
#pragma unroll
      for (int i = 0; i < STEPS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) o8p[i][j] = make_float2(0.f, 0.f);

#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        long long tw0 = 0, tw1 = 0;
        if (tmr) tw0 = clock64();
        mbar_wait(dfull_bar(q), ph);
        tc_fence_after();
        if (tmr) { tw1 = clock64(); (q == 0 ? t_w0 : t_w1) += tw1 - tw0; }
        const uint32_t taddr = tmem_base + lane_addr + 256u * (q == 0 ? par : (par ^ 1u)) + cq * (64 / NQ);
        if constexpr (NQ == 2) {
          uint32_t t0r[16], g0r[16], t1r[16], g1r[16];
          tmem_ld16(taddr, t0r);
          tmem_ld16(taddr + 128, g0r);
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {
            uint8_t* kblk = acts + (q * 2 + blk) * WL_A_BYTES + row * 128;
            const int ch0 = blk * 64 + cq * 32;
            const float* bT0 = s_b1 + q * 256 + ch0;
            const float2* wse0 = reinterpret_cast<const float2*>(cw.wse) + (q * 128 + ch0) * 4;
            tmem_ld_wait();
            tmem_ld16(taddr + blk * 64 + 16, t1r);
            tmem_ld16(taddr + 128 + blk * 64 + 16, g1r);
            gate_step2<LAST>(t0r, g0r, bT0, wse0, kblk, cq * 2, row, o8p[0]);
            tmem_ld_wait();
            if (blk == 0) {
              tmem_ld16(taddr + 64, t0r);
              tmem_ld16(taddr + 128 + 64, g0r);
            }
            gate_step2<LAST>(t1r, g1r, bT0 + 16, wse0 + 64, kblk, cq * 2 + 1, row, o8p[STEPS - 1]);
            if (!LAST && q == 1 && blk == 0) {
              fence_proxy_async_smem();
              mbar_arrive_cluster(r_acts2);
            }
          }
        } else {
          uint32_t tr[16], gr[16];
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {
            uint8_t* kblk = acts + (q * 2 + blk) * WL_A_BYTES + row * 128;
            const int ch0 = blk * 64 + cq * 16;
            tmem_ld16(taddr + blk * 64, tr);
            tmem_ld16(taddr + 128 + blk * 64, gr);
            tmem_ld_wait();
            gate_step2<LAST>(tr, gr, s_b1 + q * 256 + ch0, reinterpret_cast<const float2*>(cw.wse) + (q * 128 + ch0) * 4, kblk, cq,
                             row, o8p[0]);
            if (!LAST && q == 1 && blk == 0) {
              fence_proxy_async_smem();
              mbar_arrive_cluster(r_acts2);
            }
          }
        }
        tc_fence_before();
        if (LAST) {
          mbar_arrive_cluster(q == 0 ? r_drained0 : r_drained1);
        } else {
          fence_proxy_async_smem();   // acts (generic-proxy writes) -> visible to the MMA (async proxy)
          mbar_arrive_cluster(q == 0 ? r_actsa : r_acts);
        }
        if (tmr) t_e1 += clock64() - tw1;
      }
      // fold: per class even + odd channel lanes, then the classes pairwise ((S0 + S1) + (S2 + S3)), in a fixed order
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o8[j] = o8p[0][j].x + o8p[0][j].y;
        if (STEPS == 2) o8[j] += o8p[STEPS - 1][j].x + o8p[STEPS - 1][j].y;
      }
      if (cq > 0) {
        float* dst = s_o8 + ((cq - 1) * WL_BM + row) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
      if (cq == 0 && valid) {
        float tot[8];
        if constexpr (NQ == 2) {
          const float* p1 = s_o8 + row * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) tot[j] = o8[j] + p1[j];
        } else {
          const float* p1 = s_o8 + row * 8;
          const float* p2 = s_o8 + (WL_BM + row) * 8;
          const float* p3 = s_o8 + (2 * WL_BM + row) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) tot[j] = (o8[j] + p1[j]) + (p2[j] + p3[j]);
        }
        float4* o = reinterpret_cast<float4*>(p.acc8 + m * 8);
        float4 a0 = o[0], a1 = o[1];
        a0.x += tot[0]; a0.y += tot[1]; a0.z += tot[2]; a0.w += tot[3];
        a1.x += tot[4]; a1.y += tot[5]; a1.z += tot[6]; a1.w += tot[7];
        o[0] = a0; o[1] = a1;
      }

      // ---- residual epilogue: column group cq owns 256 / NQ accumulator columns = 4 / NQ blocks of 64 ----
      if (!LAST) {
        constexpr int NBLK = 4 / NQ, NG = 16 / NQ;     // 64-column blocks and 16-column groups per column group
        long long tw0 = 0, tw1 = 0;
        if (tmr) tw0 = clock64();
        mbar_wait(dfull_bar(2), ph);
        tc_fence_after();
        if (tmr) { tw1 = clock64(); t_w2 += tw1 - tw0; }
        const uint32_t taddr = tmem_base + lane_addr + 256u * par + cq * (256 / NQ);
        uint8_t* stg = acts + (cq * NBLK) * WL_A_BYTES + row * 128;
        const uint32_t stg_addr = smem_base + G::OFF_ACTS + (cq * NBLK) * WL_A_BYTES;
        const bool issuer = (we == cq * 4) && lane == 0;
        auto group_sync = [&]() {      // the 128 threads of this column group
          if (cq == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else if (cq == 1) asm volatile("bar.sync 4, 128;" ::: "memory");
          else if (cq == 2) asm volatile("bar.sync 5, 128;" ::: "memory");
          else asm volatile("bar.sync 6, 128;" ::: "memory");
        };
        auto resid_pass = [&](auto pass_tag) {
          constexpr int pass = decltype(pass_tag)::value;
          if (pass == 1) {
            if (issuer) bulk_wait_read0();   // the hi store has finished reading the staging tile
            group_sync();
          }
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr, r0);
#pragma unroll 1
          for (int gp = 0; gp < NG / 2; ++gp) {          // two groups of 16 columns per iteration
            tmem_ld_wait();
            tmem_ld16(taddr + (2 * gp + 1) * 16, r1);
            resid_step2<pass>(r0, s_b2 + cq * (256 / NQ) + (2 * gp) * 16, stg, 2 * gp, row, valid);
            tmem_ld_wait();
            if (gp < NG / 2 - 1) tmem_ld16(taddr + (2 * gp + 2) * 16, r0);
            else if (pass == 1) {
              tc_fence_before();
              mbar_arrive_cluster(r_epi2);   // all TMEM reads of this tile are done
            }
            resid_step2<pass>(r1, s_b2 + cq * (256 / NQ) + (2 * gp + 1) * 16, stg, 2 * gp + 1, row, valid);
          }
          fence_proxy_async_smem();
          group_sync();
          if (issuer) {
            const CUtensorMap* om = pass == 0 ? &map_ho : &map_lo;
#pragma unroll
            for (int bk = 0; bk < NBLK; ++bk)
              tma_store_4d(om, stg_addr + bk * WL_A_BYTES, (cq * NBLK + bk) * WL_BK, t0, r, 0);
            bulk_commit();
          }
        };
        resid_pass(std::integral_constant<int, 0>{});
        resid_pass(std::integral_constant<int, 1>{});
        if (issuer) bulk_wait_read0();   // staging tile may be overwritten by the next gate epilogue
        if (tmr) t_e2 += clock64() - tw1;
      }
      // s_o8 and the staging tile are reused by the next tile: no epilogue warp may run ahead
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
    if (!LAST && lane == 0 && (we & 3) == 0) bulk_wait0();   // all TMA stores of this CTA have landed
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
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, 512);
  }
}

// Returns how many CTA pairs (clusters of 2) of the layer kernel can be resident at once on this device. A pair needs
// both SMs of one TPC; a part whose disabled SMs are spread over TPCs has fewer complete TPCs than sm_count / 2, and a
// persistent grid larger than that would run its last clusters as a second round.
template <int EW>
inline void tc_pair_set_attrs() {
  WG_CK(cudaFuncSetAttribute(tc_wn_pair_kernel<false, false, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WpGeom<false, EW>::SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_pair_kernel<true, false, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WpGeom<true, EW>::SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_pair_kernel<false, true, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WpGeom<false, EW>::SMEM));
}
// The 16-epilogue-warp instantiation (EW = 16, four warps per TMEM lane quarter) is a measured probe: bit-identical, no
// faster (K2 55.3 ms vs 54.6-54.8 ms with 8 warps on the same box) -- the step is power-capped, not epilogue-bound. It is
// compiled only into a -DWG_PROBES build (WG_PAIR_EPI=16 selects it there).
inline int tc_pair_init() {
  tc_pair_set_attrs<8>();
#ifdef WG_PROBES
  tc_pair_set_attrs<16>();
#endif
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.gridDim = dim3(2 * 1024); cfg.blockDim = dim3(WpGeom<false, 8>::THREADS); cfg.dynamicSmemBytes = WpGeom<false, 8>::SMEM;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tc_wn_pair_kernel<false, false, 8>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

// 128-row boxes (one CTA's half of a 256-row B chunk) over the stacked weight matrices
struct TcPairMaps {
  CUtensorMap m_w1, m_wc, m_w2, m_w0, m_h0;
  int max_pairs = 0;     // resident clusters of 2 on this device (tc_pair_init)
  bool ready = false;
};

inline void tc_pair_prepare(TcPairMaps& pm, const TcPlan& pl, int n_layers_total, int n_flows, int R, const __nv_bfloat16* W1,
                            const __nv_bfloat16* W2, const __nv_bfloat16* V, const __nv_bfloat16* W0, const __nv_bfloat16* H0,
                            int max_pairs) {
  const int C = pl.C;
  pm.max_pairs = max_pairs;
  make_map_2d(&pm.m_w1, W1, (uint64_t)n_layers_total * 2 * C, 3 * C + pl.S, 128);
  make_map_2d(&pm.m_w2, W2, (uint64_t)n_layers_total * C, C, 128);
  make_map_2d(&pm.m_wc, V, (uint64_t)n_layers_total * R * 2 * C, pl.Kup, 128);
  if (pl.fold0) {
    make_map_2d(&pm.m_w0, W0, (uint64_t)n_flows * 2 * C, WL_BK, 128);
    make_map_2d(&pm.m_h0, H0, (uint64_t)n_flows * C, WL_BK, 128);
  } else {
    pm.m_w0 = pm.m_w1;
    pm.m_h0 = pm.m_w1;
  }
  pm.ready = true;
}

template <int EW>
inline void tc_pair_launch(const TcPlan& pl, const TcPairMaps& pm, const WnLayerParams& p, const WnLayerConst& cw, int grid, bool last,
                           bool first, int hcur, cudaStream_t st) {
  const bool pdl = p.pdl != 0;
  if (last)
    tc_launch(tc_wn_pair_kernel<true, false, EW>, grid, WpGeom<true, EW>::THREADS, WpGeom<true, EW>::SMEM, st, pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pm.m_w1, pm.m_wc, pm.m_w2, pl.m4_a0, pm.m_w0, pm.m_h0, p, cw);
  else if (first)
    tc_launch(tc_wn_pair_kernel<false, true, EW>, grid, WpGeom<false, EW>::THREADS, WpGeom<false, EW>::SMEM, st, pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pm.m_w1, pm.m_wc, pm.m_w2, pl.m4_a0, pm.m_w0, pm.m_h0, p, cw);
  else
    tc_launch(tc_wn_pair_kernel<false, false, EW>, grid, WpGeom<false, EW>::THREADS, WpGeom<false, EW>::SMEM, st, pdl, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m4_cond, pm.m_w1, pm.m_wc, pm.m_w2, pl.m4_a0, pm.m_w0, pm.m_h0, p, cw);
}

inline int tc_wn_layer_pair(const TcPlan& pl, const TcPairMaps& pm, int layer, int dilation, bool last, int hcur,
                            float* acc8, const float* b1, const float* b2, const float* wse_host, cudaStream_t st,
                            bool first = false, unsigned long long* timing = nullptr, int epi_warps = 8) {
  if (!pl.pm || pl.C != 256) fail(WG_ERR_UNSUPPORTED, "the CTA-pair kernel is built for the phase-major layout, C = 256");
  WnLayerParams p{};
  tc_fill_params(pl, p, layer, dilation, hcur, acc8, b1, b2, timing, 0);
  WnLayerConst cw;
  std::memcpy(cw.wse, wse_host, sizeof cw.wse);
  const int max_pairs = pm.max_pairs;
  const int need_pairs = ((pl.tiles_per_row + 1) / 2) * pl.R;
  const int grid = 2 * (need_pairs < max_pairs ? need_pairs : max_pairs);
  if (first && (!pl.fold0 || last || dilation != 1)) fail(WG_ERR_INVALID, "start fold requested for a layer it does not apply to");
  p.pdl = pl.pdl && grid < pl.sm_count ? 1 : 0;
#ifdef WG_PROBES
  if (epi_warps == 16) {
    tc_pair_launch<16>(pl, pm, p, cw, grid, last, first, hcur, st);
    WG_CK(cudaGetLastError());
    return 1;
  }
#endif
  (void)epi_warps;
  tc_pair_launch<8>(pl, pm, p, cw, grid, last, first, hcur, st);
  WG_CK(cudaGetLastError());
  return 1;
}

}  // namespace wg

// Fused WN layer on a CTA PAIR (tcgen05 cta_group::2): the same computation as tc_wn_layer_kernel, but two
// SMs of one TPC form a cluster and run ONE 256-row MMA: each CTA owns a 128-row tile (its own A tiles, its
// own TMEM accumulators, its own epilogue), the B operand (weights) is split in halves across the two CTAs'
// shared memory and read by both tensor cores. Per CTA a pipeline stage is 16 KB of A + 16 KB of B instead of
// 16 + 32 KB, so the same shared memory holds 4 stages instead of 3 and the bytes that must be in flight to
// cover the L2 latency drop from 96 to 64 per tensor-core cycle -- the stage ring, not the tensor pipe, was
// what bounded the single-CTA kernel (profiles/r01_v3_*).
//
// Roles per CTA: warp 0 TMA producer (own A tile + own half of B; completion is signalled on the LEADER's
// full barrier), warp 1 TMEM allocator, and in the leader CTA (cluster rank 0) the MMA issuer; warps 2..9
// epilogue. Epilogue->MMA barriers live in the leader and count the epilogue threads of BOTH CTAs;
// MMA->epilogue and MMA->producer barriers are signalled in both CTAs by multicast tcgen05.commit.
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
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
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

// (elect_one() lives in tc_kernels.cuh: keeping the whole warp in the producer / MMA loops and electing only
// around the asm lets ptxas keep descriptors and coordinates in UNIFORM registers; a loop entered by
// `if (lane == 0)` made every UTCHMMA / UTMALDG pay an ELECT + R2UR.BROADCAST waterfall.)

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of CTA 1 -> same offset in CTA 0

// ---- geometry -------------------------------------------------------------------------------------
constexpr int WP_A_BYTES = WL_BM * WL_BK * 2;          // 16 KB: own 128 rows x 64 K
constexpr int WP_B_BYTES = 128 * WL_BK * 2;            // 16 KB: own half (128 of 256 N rows) x 64 K
constexpr int WP_STAGE_BYTES = WP_A_BYTES + WP_B_BYTES;
// The LAST layer of a flow has no GEMM2 / residual, hence no acts tile, identity tile or staging: its ring
// can be 6 deep. The other layers keep a 64 KB acts tile and get 4 stages.
template <bool LAST>
struct WpGeom {
  static constexpr int STAGES = LAST ? 6 : 4;
  static constexpr int OFF_ACTS = STAGES * WP_STAGE_BYTES;
  static constexpr int OFF_I64 = OFF_ACTS + (LAST ? 0 : WL_ACTS_BYTES);   // own half (32 N rows) of the 64x64 identity
  static constexpr int OFF_B1 = OFF_I64 + (LAST ? 0 : 32 * 128);
  static constexpr int OFF_B2 = OFF_B1 + 2 * WL_C * 4;
  static constexpr int OFF_O8 = OFF_B2 + WL_C * 4;
  static constexpr int OFF_BARS = OFF_O8 + WL_BM * 8 * 4;
  static constexpr int NBARS = 2 * STAGES + 3 + 2 + 3;
  static constexpr int SMEM = OFF_BARS + NBARS * 8 + 16 + 2 * STAGES * 8 + 64 * 8;   // + debug timestamps + wait histogram
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <bool LAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WL_THREADS, 1)
tc_wn_pair_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_ho,
                  const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_spect,
                  const __grid_constant__ CUtensorMap map_w1h, const __grid_constant__ CUtensorMap map_w2h,
                  const WnLayerParams p, const __grid_constant__ WnLayerConst cw) {
  using G = WpGeom<LAST>;
  constexpr int WP_STAGES = G::STAGES, WP_OFF_ACTS = G::OFF_ACTS, WP_OFF_I64 = G::OFF_I64, WP_OFF_B1 = G::OFF_B1,
                WP_OFF_B2 = G::OFF_B2, WP_OFF_O8 = G::OFF_O8, WP_OFF_BARS = G::OFF_BARS, WP_NBARS = G::NBARS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + WP_OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + WP_OFF_B2);
  float* s_o8 = reinterpret_cast<float*>(smem + WP_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WP_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WP_NBARS);
  volatile long long* ts_issue = reinterpret_cast<volatile long long*>(smem + WP_OFF_BARS + WP_NBARS * 8 + 16);
  volatile long long* ts_commit = ts_issue + WP_STAGES;
  volatile long long* wait_pos = ts_commit + WP_STAGES;     // [64] MMA wait-for-data cycles by stage position in the tile
  const uint32_t bar_base = smem_base + WP_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                            // leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (WP_STAGES + s); };             // both CTAs
  auto dfull_bar = [&](int i) { return bar_base + 8u * (2 * WP_STAGES + i); };         // both CTAs
  auto drained_bar = [&](int i) { return bar_base + 8u * (2 * WP_STAGES + 3 + i); };   // leader only
  const uint32_t actsa_bar = bar_base + 8u * (2 * WP_STAGES + 5);                      // leader only
  const uint32_t acts_bar = bar_base + 8u * (2 * WP_STAGES + 6);                       // leader only
  const uint32_t epi2_bar = bar_base + 8u * (2 * WP_STAGES + 7);                       // leader only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst_req = (p.flags >> 8) & 15;                                   // probe: use fewer ring slots
  const uint32_t nst = (nst_req > 0 && nst_req < WP_STAGES) ? nst_req : WP_STAGES;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((smem_base & 1023u) != 0u) __trap();

  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  // tile of this CTA in pair-iteration i: 2 * (pair + i * n_pairs) + rank; may be a ghost tile (>= n_tiles)
  auto tile_coords = [&](int i, int& b, int& l0, bool& ghost) {
    const int t = 2 * (pair + i * n_pairs) + static_cast<int>(rank);
    ghost = t >= p.n_tiles;
    b = ghost ? p.n_tiles / p.tiles_per_b : t / p.tiles_per_b;   // ghost: batch index == B -> every TMA box is out of bounds
    l0 = ghost ? 0 : (t - b * p.tiles_per_b) * WL_BM;
  };
  const int npt = (p.n_tiles + 1) / 2;                                        // pair-tiles in this launch
  const int n_iter = pair < npt ? (npt - pair + n_pairs - 1) / n_pairs : 0;  // identical for both CTAs of a pair

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    prefetch_tmap(&map_ho);
    prefetch_tmap(&map_lo);
    prefetch_tmap(&map_spect);
    prefetch_tmap(&map_w1h);
    prefetch_tmap(&map_w2h);
    for (int s = 0; s < WP_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(dfull_bar(i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(drained_bar(i), 2 * WL_EPI_THREADS);
    mbar_init(actsa_bar, 2 * WL_EPI_THREADS);
    mbar_init(acts_bar, 2 * WL_EPI_THREADS);
    mbar_init(epi2_bar, 2 * WL_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_slot), 512);
    tmem2_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * WL_C; i += WL_THREADS) s_b1[i] = p.b1[i];
  if (!LAST) {
    for (int i = threadIdx.x; i < WL_C; i += WL_THREADS) s_b2[i] = p.b2[i];
    // own half of the identity B tile: local row nl is N row n = 32 * rank + nl; element (n, k) = [n == k]
    uint32_t* i64w = reinterpret_cast<uint32_t*>(smem + WP_OFF_I64);
    for (int i = threadIdx.x; i < 32 * 32; i += WL_THREADS) {
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
  const uint32_t tmem_base = *tmem_slot;
  const bool timing = p.timing != nullptr;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ============================
    {
      uint32_t it = 0;
      long long t_wait = 0, t_c2p = 0, n_c2p = 0;
      auto acquire = [&](uint32_t pair_bytes) -> uint32_t {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        long long t0 = 0;
        if (timing) t0 = clock64();
        mbar_wait(empty_bar(s), ph ^ 1);
        if (timing) {
          const long long now = clock64();
          t_wait += now - t0;
          if (leader && it >= nst) { t_c2p += now - ts_commit[s]; ++n_c2p; }
          if (leader && lane == 0) ts_issue[s] = now;
        }
        if (leader && elect_one()) mbar_expect_tx(full_bar(s), pair_bytes);   // bytes of BOTH CTAs land on the leader's barrier
        __syncwarp();
        return static_cast<uint32_t>(s);
      };
      for (int i = 0; i < n_iter && !(p.flags & 16); ++i) {
        int b, l0;
        bool ghost;
        tile_coords(i, b, l0, ghost);
        // While the second chunk streams (its activation tiles are L2 hits: the first chunk just read them),
        // pull the NEXT tile's activation rows from HBM into L2, one box per stage, so the first chunk of the
        // next tile does not pay DRAM latency through a 4-deep ring.
        int nb, nl0;
        bool nghost;
        tile_coords(i + 1, nb, nl0, nghost);
        const bool pf_on = (p.flags & 8) && i + 1 < n_iter && !nghost;
        for (int q = 0; q < 2; ++q) {
          for (int kb = 0; kb < WL_KB1; ++kb, ++it) {
            const bool skip_b = p.flags & 1, skip_a = p.flags & 2;     // probes (wrong results)
            const uint32_t s = acquire(2 * ((skip_a ? 0 : WP_A_BYTES) + (skip_b ? 0 : WP_B_BYTES)));
            const uint32_t fb = full_bar(s) & kPeerBitMask;
            const uint32_t a_dst = smem_base + s * WP_STAGE_BYTES;
            if (elect_one()) {
              if (kb < WL_KB_CONV) {
                const int tap = kb >> 2, cblk = kb & 3;
                if (!skip_a) tma2_load_3d(a_dst, &map_h, fb, cblk * WL_BK, l0 + (tap - 1) * p.dilation, b);
                if (pf_on && q == 1 && tap == 1) tma_prefetch_3d(&map_h, cblk * WL_BK, nl0, nb);   // centre rows cover most of the halo
              } else {
                if (!skip_a) tma2_load_3d(a_dst, &map_spect, fb, (kb - WL_KB_CONV) * WL_BK, l0, b);
                if (pf_on && q == 1) tma_prefetch_3d(&map_spect, (kb - WL_KB_CONV) * WL_BK, nl0, nb);
              }
              if (!skip_b) tma2_load_2d(a_dst + WP_A_BYTES, &map_w1h, fb, kb * WL_BK, p.layer * 2 * WL_C + q * 256 + rank * 128);
            }
            __syncwarp();
          }
        }
        if (!LAST) {
          // consumption order of the MMA warp: W2/hi blocks 0,1 | lo blocks 0..3 | W2/hi blocks 2,3
          for (int step = 0; step < 6; ++step, ++it) {
            const uint32_t s = acquire(2 * WP_STAGE_BYTES);
            const uint32_t fb = full_bar(s) & kPeerBitMask;
            const uint32_t dst = smem_base + s * WP_STAGE_BYTES;
            if (elect_one()) {
              if (step == 2 || step == 3) {
                const int kb = (step - 2) * 2;
                tma2_load_3d(dst, &map_lo, fb, kb * WL_BK, l0, b);
                tma2_load_3d(dst + WP_A_BYTES, &map_lo, fb, (kb + 1) * WL_BK, l0, b);
              } else {
                const int kb = step < 2 ? step : step - 2;
                tma2_load_3d(dst, &map_h, fb, kb * WL_BK, l0, b);
                tma2_load_2d(dst + WP_A_BYTES, &map_w2h, fb, kb * WL_BK, p.layer * WL_C + rank * 128);
              }
            }
            __syncwarp();
          }
        }
      }
      if (timing && leader && lane == 0) {
        atomicAdd(p.timing + 8, static_cast<unsigned long long>(t_wait));
        atomicAdd(p.timing + 9, static_cast<unsigned long long>(t_c2p));    // sum (producer wake - commit issue)
        atomicAdd(p.timing + 10, static_cast<unsigned long long>(n_c2p));
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
      constexpr uint32_t idesc_id = umma_idesc_bf16(256, 64);
      const uint64_t idesc64 = umma_desc_sw128(smem_base + WP_OFF_I64);
      uint32_t it = 0;
      long long t_full = 0, t_epi = 0, t_begin = 0, t_i2f = 0, n_i2f = 0, t_i2f_w = 0, n_i2f_w = 0;
      long long t_issue_mma = 0, t_issue_commit = 0, t_syncwarp = 0;
      if (timing) {
        t_begin = clock64();
        if (lane == 0) for (int i = 0; i < 64; ++i) wait_pos[i] = 0;
        __syncwarp();
      }
      uint32_t pos = 0;   // stage position inside the current tile
      auto wait_full = [&]() -> uint32_t {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        long long t0 = 0;
        if (timing) t0 = clock64();
        if (!(p.flags & 16)) mbar_wait(full_bar(s), ph);   // flag 16: tensor-pipe-only probe (no TMA)
        if (timing) {
          const long long now = clock64();
          t_full += now - t0;
          t_i2f += now - ts_issue[s]; ++n_i2f;                       // TMA issue -> data seen by the MMA thread
          if (lane == 0) wait_pos[pos & 63] += now - t0;
          ++pos;
          if (now - t0 > 64) { t_i2f_w += now - ts_issue[s]; ++n_i2f_w; }   // ... only when the MMA really waited
        }
        tc_fence_after();
        return smem_base + s * WP_STAGE_BYTES;
      };
      auto wait_epi = [&](uint32_t bar, uint32_t ph) {
        long long t0 = 0;
        if (timing) t0 = clock64();
        mbar_wait(bar, ph);
        if (timing) t_epi += clock64() - t0;
        tc_fence_after();
      };
      for (int n = 0; n < n_iter; ++n) {
        const uint32_t par = LAST ? 0u : (static_cast<uint32_t>(n) & 1u);
        const uint32_t prev_ph = static_cast<uint32_t>(n - 1) & 1u;
        pos = 0;
        for (int q = 0; q < 2; ++q) {
          const uint32_t d_tmem = tmem_base + 256u * (q == 0 ? par : (par ^ 1u));
          if (n > 0) {
            if (LAST) wait_epi(drained_bar(q), prev_ph);
            else if (q == 1) wait_epi(epi2_bar, prev_ph);
          }
          for (int kb = 0; kb < WL_KB1; ++kb, ++it) {
            const uint32_t a_addr = wait_full();
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + WP_A_BYTES);
            if (elect_one()) {
              long long c0 = 0, c1 = 0;
              if (timing) c0 = clock64();
              if (!(p.flags & 64)) {   // flag 64: feed-only probe (no MMA issued, stages are released at once)
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              }
              if (timing) { c1 = clock64(); ts_commit[it % nst] = c1; t_issue_mma += c1 - c0; }
              tc2_commit(empty_bar(it % nst));
              if (kb == WL_KB1 - 1) tc2_commit(dfull_bar(q));
              if (timing) t_issue_commit += clock64() - c1;
            }
            if (timing) { const long long c2 = clock64(); __syncwarp(); t_syncwarp += clock64() - c2; } else
            __syncwarp();
          }
        }
        if (!LAST) {
          const uint32_t d_tmem = tmem_base + 256u * par;
          wait_epi(actsa_bar, static_cast<uint32_t>(n) & 1u);
          for (int step = 0; step < 6; ++step, ++it) {
            if (step == 4) wait_epi(acts_bar, static_cast<uint32_t>(n) & 1u);
            const uint32_t st_addr = wait_full();
            if (elect_one()) {
              if (step == 2 || step == 3) {
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
                const uint64_t adesc = umma_desc_sw128(smem_base + WP_OFF_ACTS + kb * WL_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(st_addr + WP_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (step | k) ? 1u : 0u);
                const uint64_t hdesc = umma_desc_sw128(st_addr);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma2_bf16(d_tmem + 64u * kb, hdesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
              }
              if (timing) ts_commit[it % nst] = clock64();
              tc2_commit(empty_bar(it % nst));
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
        atomicAdd(p.timing + 11, static_cast<unsigned long long>(t_i2f));
        atomicAdd(p.timing + 12, static_cast<unsigned long long>(n_i2f));
        atomicAdd(p.timing + 13, static_cast<unsigned long long>(t_i2f_w));
        atomicAdd(p.timing + 14, static_cast<unsigned long long>(n_i2f_w));
        for (int i = 0; i < 64; ++i) atomicAdd(p.timing + 16 + i, static_cast<unsigned long long>(wait_pos[i]));
      }
      if (timing) {   // elected-lane counters: reduce over the warp (only the elected lane accumulated)
        for (int o = 16; o > 0; o >>= 1) {
          t_issue_mma += __shfl_xor_sync(0xffffffffu, t_issue_mma, o);
          t_issue_commit += __shfl_xor_sync(0xffffffffu, t_issue_commit, o);
        }
        if (lane == 0) {
          atomicAdd(p.timing + 80, static_cast<unsigned long long>(t_issue_mma));
          atomicAdd(p.timing + 81, static_cast<unsigned long long>(t_issue_commit));
          atomicAdd(p.timing + 82, static_cast<unsigned long long>(t_syncwarp));
        }
      }
    }
  } else {
    // ======================================= epilogue (both CTAs) ==============================
    const int we = warp - 2;
    const int quarter = warp & 3;
    const int hf = we >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* acts = smem + WP_OFF_ACTS;
    // epilogue -> MMA barriers live in the leader CTA
    const uint32_t r_drained0 = mapa_u32(drained_bar(0), 0), r_drained1 = mapa_u32(drained_bar(1), 0);
    const uint32_t r_actsa = mapa_u32(actsa_bar, 0), r_acts = mapa_u32(acts_bar, 0), r_epi2 = mapa_u32(epi2_bar, 0);
    const bool tmr = timing && we == 0 && lane == 0;
    long long t_w0 = 0, t_w1 = 0, t_w2 = 0, t_e1 = 0, t_e2 = 0;
    for (int n = 0; n < n_iter; ++n) {
      int b, l0;
      bool ghost;
      tile_coords(n, b, l0, ghost);
      const uint32_t par = LAST ? 0u : (static_cast<uint32_t>(n) & 1u);
      const uint32_t ph = static_cast<uint32_t>(n) & 1u;
      const bool valid = !ghost && (l0 + row) < p.L;
      const size_t m = static_cast<size_t>(b) * p.L + l0 + row;
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = 0.f;

#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        long long t0 = 0;
        if (tmr) t0 = clock64();
        mbar_wait(dfull_bar(q), ph);
        tc_fence_after();
        long long t1 = 0;
        if (tmr) { t1 = clock64(); (q == 0 ? t_w0 : t_w1) += t1 - t0; }
        const uint32_t taddr = tmem_base + lane_addr + 256u * (q == 0 ? par : (par ^ 1u)) + hf * 64;
        uint8_t* kblk = acts + (q * 2 + hf) * WL_A_BYTES + row * 128;
        const float* bT0 = s_b1 + q * 256 + hf * 64;
        const float* wse0 = cw.wse + (q * 128 + hf * 64) * 8;
        uint32_t t0r[16], g0r[16], t1r[16], g1r[16];
        tmem_ld16(taddr, t0r);
        tmem_ld16(taddr + 128, g0r);
#pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {
          const int st = 2 * sp;
          tmem_ld_wait();
          tmem_ld16(taddr + (st + 1) * 16, t1r);
          tmem_ld16(taddr + 128 + (st + 1) * 16, g1r);
          if (!(p.flags & 32)) gate_step<LAST>(t0r, g0r, bT0 + st * 16, wse0 + st * 128, kblk, st, row, o8);
          tmem_ld_wait();
          if (sp == 0) {
            tmem_ld16(taddr + (st + 2) * 16, t0r);
            tmem_ld16(taddr + 128 + (st + 2) * 16, g0r);
          }
          if (!(p.flags & 32)) gate_step<LAST>(t1r, g1r, bT0 + (st + 1) * 16, wse0 + (st + 1) * 128, kblk, st + 1, row, o8);
        }
        tc_fence_before();
        if (LAST) {
          mbar_arrive_cluster(q == 0 ? r_drained0 : r_drained1);
        } else {
          fence_proxy_async_smem();
          mbar_arrive_cluster(q == 0 ? r_actsa : r_acts);
        }
        if (tmr) t_e1 += clock64() - t1;
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

      if (!LAST) {
        long long t0 = 0;
        if (tmr) t0 = clock64();
        mbar_wait(dfull_bar(2), ph);
        tc_fence_after();
        long long t1 = 0;
        if (tmr) { t1 = clock64(); t_w2 += t1 - t0; }
        const uint32_t taddr = tmem_base + lane_addr + 256u * par + hf * 128;
        uint8_t* stg = acts + (hf * 2) * WL_A_BYTES + row * 128;
        const uint32_t stg_addr = smem_base + WP_OFF_ACTS + (hf * 2) * WL_A_BYTES;
        const bool issuer = (we == hf * 4) && lane == 0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 1) {
            if (issuer) bulk_wait_read0();
            if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
            else asm volatile("bar.sync 4, 128;" ::: "memory");
          }
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr, r0);
#pragma unroll 1
          for (int gp = 0; gp < 4; ++gp) {
            tmem_ld_wait();
            tmem_ld16(taddr + (2 * gp + 1) * 16, r1);
            if (!(p.flags & 32)) resid_step(r0, s_b2 + hf * 128 + (2 * gp) * 16, stg, 2 * gp, row, pass);
            tmem_ld_wait();
            if (gp < 3) tmem_ld16(taddr + (2 * gp + 2) * 16, r0);
            else if (pass == 1) {
              tc_fence_before();
              mbar_arrive_cluster(r_epi2);
            }
            if (!(p.flags & 32)) resid_step(r1, s_b2 + hf * 128 + (2 * gp + 1) * 16, stg, 2 * gp + 1, row, pass);
          }
          fence_proxy_async_smem();
          if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          if (issuer) {
            const CUtensorMap* om = pass == 0 ? &map_ho : &map_lo;
            tma_store_3d(om, stg_addr, (hf * 2) * WL_BK, l0, b);
            tma_store_3d(om, stg_addr + WL_A_BYTES, (hf * 2 + 1) * WL_BK, l0, b);
            bulk_commit();
          }
        }
        if (issuer) bulk_wait_read0();
        if (tmr) t_e2 += clock64() - t1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(WL_EPI_THREADS) : "memory");
    }
    if (!LAST && lane == 0 && (we == 0 || we == 4)) bulk_wait0();
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

inline void tc_pair_init() {
  WG_CK(cudaFuncSetAttribute(tc_wn_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WpGeom<false>::SMEM));
  WG_CK(cudaFuncSetAttribute(tc_wn_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WpGeom<true>::SMEM));
}

struct TcPairMaps {
  CUtensorMap m_w1h, m_w2h;   // 128-row boxes over the stacked W1 / W2 matrices
};

inline void tc_pair_prepare(TcPairMaps& pm, int n_layers_total, int C, int S, const __nv_bfloat16* W1,
                            const __nv_bfloat16* W2) {
  make_map_2d(&pm.m_w1h, W1, (uint64_t)n_layers_total * 2 * C, 3 * C + S, 128);
  make_map_2d(&pm.m_w2h, W2, (uint64_t)n_layers_total * C, C, 128);
}

inline int tc_wn_layer_pair(const TcPlan& pl, const TcPairMaps& pm, int layer, int dilation, bool last, int hcur,
                            float* acc8, const float* b1, const float* b2, const float* wse_host,
                            unsigned long long* timing, int flags, cudaStream_t st) {
  if (pl.pm) fail(WG_ERR_UNSUPPORTED, "the CTA-pair kernel only supports the position-major layout");
  WnLayerParams p{};
  tc_fill_params(pl, p, layer, dilation, hcur, acc8, b1, b2, timing, flags);
  WnLayerConst cw;
  std::memcpy(cw.wse, wse_host, sizeof cw.wse);
  const int max_pairs = pl.sm_count / 2;
  const int need_pairs = (pl.n_tiles + 1) / 2;
  const int grid = 2 * (need_pairs < max_pairs ? need_pairs : max_pairs);
  if (last)
    tc_wn_pair_kernel<true><<<grid, WL_THREADS, WpGeom<true>::SMEM, st>>>(pl.m_h16[hcur], pl.m_h16[hcur ^ 1], pl.m_lo, pl.m_spect,
                                                                pm.m_w1h, pm.m_w2h, p, cw);
  else
    tc_wn_pair_kernel<false><<<grid, WL_THREADS, WpGeom<false>::SMEM, st>>>(pl.m_h16[hcur], pl.m_h16[hcur ^ 1], pl.m_lo, pl.m_spect,
                                                                 pm.m_w1h, pm.m_w2h, p, cw);
  WG_CK(cudaGetLastError());
  return 1;
}

}  // namespace wg

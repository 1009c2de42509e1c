// WaveGlow-512 (the reference's default width, BASELINE.json configs[2]) on tcgen05.
//
// With C = 512 the acts tile of one 128-row tile is 128 KB, so GEMM1+gate and GEMM2+residual no longer fit
// into one CTA next to the TMA ring. The layer is two kernels that exchange the bf16 acts through HBM
// (+2 KB/row, still ~700 FLOP/B):
//   tc512_gate_kernel<LAST>  GEMM1 [128 x (3*512 + cond)] . [.. x 1024] in four 256-column chunks (gate
//                            channels 128 q .. 128 q + 127), gate epilogue, skip/end fold, acts -> HBM (TMA store)
//   tc512_res_kernel         GEMM2 acts . Wres + hi . I + lo . I in two 256-column chunks, residual epilogue
//                            (hi, lo) -> HBM (TMA store)
// Same row geometry (position- or phase-major), PTX wrappers, gate_step / resid_step as tc_kernels.cuh.
#pragma once
#include "tc_kernels.cuh"
#include "tc_pair_kernels.cuh"   // cluster / cta_group::2 PTX wrappers

namespace wg {

constexpr int W5_C = 512;
constexpr int W5_KB_CONV = 3 * W5_C / WL_BK;            // 24
constexpr int W5_STAGES = 3;
constexpr int W5_OFF_STG = W5_STAGES * WL_STAGE_BYTES;  // 144 KB ring, then 64 KB of staging tiles
constexpr int W5_STG_BYTES = 4 * WL_A_BYTES;            // 64 KB
constexpr int W5_OFF_I64 = W5_OFF_STG + W5_STG_BYTES;
constexpr int W5_OFF_B = W5_OFF_I64 + 64 * 128;         // gate: b1 [1024] f32; res: b2 [512] f32
constexpr int W5_OFF_O8 = W5_OFF_B + 4 * W5_C * 2;
constexpr int W5_OFF_BARS = W5_OFF_O8 + WL_BM * 8 * 4;
constexpr int W5_NBARS = 2 * W5_STAGES + 2 + 2;         // full, empty, dfull[2], drained[2]
constexpr int W5_SMEM = W5_OFF_BARS + W5_NBARS * 8 + 16;
static_assert(W5_SMEM <= 232448, "shared memory budget");

struct Wn512Const {        // kernel-parameter bank
  float wse[W5_C * 8];     // Wskip @ Wend, [256 channel pairs][8][even, odd]
};

// ------------------------------------------------------------------------------------------------
// Gate kernel: chunk q of tile n is accumulated into TMEM region (q & 1); the epilogue drains a region while
// the MMA fills the other one. Region x is filled for the f-th time after its (f-1)-th drain.
// ------------------------------------------------------------------------------------------------
// FIRST (layer 0 of a flow, phase-major): start-conv fold as in tc_wn_layer_kernel -- the 24 conv K-blocks are one
// block a0[128 x 64] @ W0[64 x 1024] (three K = 16 MMAs per chunk).
template <bool LAST, bool FIRST = false>
__global__ void __launch_bounds__(WL_THREADS, 1)
tc512_gate_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_cond,
                  const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_wc,
                  const __grid_constant__ CUtensorMap map_acts, const __grid_constant__ CUtensorMap map_a0,
                  const __grid_constant__ CUtensorMap map_w0, const WnLayerParams p,
                  const __grid_constant__ Wn512Const cw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + W5_OFF_B);
  float* s_o8 = reinterpret_cast<float*>(smem + W5_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W5_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + W5_NBARS);
  const uint32_t bar_base = smem_base + W5_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (W5_STAGES + s); };
  auto dfull_bar = [&](int x) { return bar_base + 8u * (2 * W5_STAGES + x); };
  auto drained_bar = [&](int x) { return bar_base + 8u * (2 * W5_STAGES + 2 + x); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_base & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    prefetch_tmap(&map_cond);
    prefetch_tmap(&map_w1);
    prefetch_tmap(&map_wc);
    prefetch_tmap(&map_acts);
    for (int s = 0; s < W5_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(dfull_bar(x), 1);
      mbar_init(drained_bar(x), WL_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * W5_C; i += WL_THREADS) s_b1[i] = p.b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool pm = p.R > 1;
  constexpr int KB_CONV = FIRST ? 1 : W5_KB_CONV;
  const int kb1 = KB_CONV + p.n_cond_kb;
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
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int b, r, t0;
      tile_coords(tile, b, r, t0);
      for (int q = 0; q < 4; ++q) {
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const int s = it % W5_STAGES;
          mbar_wait(empty_bar(s), ((it / W5_STAGES) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), WL_STAGE_BYTES);
            const uint32_t a_dst = smem_base + s * WL_STAGE_BYTES;
            if (FIRST && kb == 0) {
              tma_load_4d(a_dst, &map_a0, full_bar(s), 0, t0, pm ? r : b, pm ? b : 0);
              tma_load_2d(a_dst + WL_A_BYTES, &map_w0, full_bar(s), 0, p.flow * 2 * W5_C + q * 256);
            } else if (kb < KB_CONV) {
              const int tap = kb >> 3, cblk = kb & 7;
              const int rs = r + (tap - 1) * p.dilation;
              const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
              tma_load_4d(a_dst, &map_h, full_bar(s), cblk * WL_BK, t0 + carry, pm ? rs - carry * p.R : b, pm ? b : 0);
              tma_load_2d(a_dst + WL_A_BYTES, &map_w1, full_bar(s), kb * WL_BK, p.layer * 2 * W5_C + q * 256);
            } else {
              const int kc = kb - KB_CONV;
              tma_load_4d(a_dst, &map_cond, full_bar(s), kc * WL_BK, t0, pm ? 0 : b, pm ? b : 0);
              tma_load_2d(a_dst + WL_A_BYTES, &map_wc, full_bar(s), p.wc_col0 + kc * WL_BK,
                          p.wc_row0 + r * p.wc_rstride + q * 256);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    uint32_t it = 0, fills = 0;   // fills = chunks issued so far (region = fills & 1)
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int q = 0; q < 4; ++q, ++fills) {
        const uint32_t x = fills & 1u, f = fills >> 1;      // f-th fill of region x
        if (f > 0) {
          mbar_wait(drained_bar(x), (f - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + 256u * x;
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const int s = it % W5_STAGES;
          mbar_wait(full_bar(s), (it / W5_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * WL_STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + WL_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < WL_BK / 16; ++k) {
              if (FIRST && kb == 0 && k == 3) break;   // columns 48..63 of a0: the residual operand (tc512_res_kernel)
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
            tc_commit(empty_bar(s));
            if (kb == kb1 - 1) tc_commit(dfull_bar(x));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int we = warp - 2;
    const int quarter = warp & 3;
    const int hf = we >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* stg_all = smem + W5_OFF_STG;
    const bool issuer = we == 0 && lane == 0;
    uint32_t fills = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int b, r, t0;
      tile_coords(tile, b, r, t0);
      const bool valid = wn_row_valid(p, t0 + row);
      const size_t m = (static_cast<size_t>(b) * p.R + r) * p.T + t0 + row;
      float2 o8p[8];   // (even-channel, odd-channel) partial sums of the fold columns (gate_step2)
#pragma unroll
      for (int j = 0; j < 8; ++j) o8p[j] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++fills) {
        const uint32_t x = fills & 1u, f = fills >> 1;
        mbar_wait(dfull_bar(x), f & 1u);
        tc_fence_after();
        // staging tile (q & 1): two 64-channel blocks of this chunk's acts; its previous TMA store (chunk q-2)
        // must have finished reading shared memory
        uint8_t* stg = stg_all + (q & 1) * (2 * WL_A_BYTES);
        if (!LAST) {
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 3, %0;" ::"n"(WL_EPI_THREADS) : "memory");
        }
        const uint32_t taddr = tmem_base + lane_addr + 256u * x + hf * 32;
        uint32_t t0r[16], g0r[16], t1r[16], g1r[16];
        tmem_ld16(taddr, t0r);
        tmem_ld16(taddr + 128, g0r);
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          uint8_t* kblk = stg + blk * WL_A_BYTES + row * 128;
          const int ch0 = blk * 64 + hf * 32;
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
          gate_step2<LAST>(t1r, g1r, bT0 + 16, wse0 + 64, kblk, hf * 2 + 1, row, o8p);
        }
        tc_fence_before();
        mbar_arrive(drained_bar(x));
        if (!LAST) {
          fence_proxy_async_smem();
          asm volatile("bar.sync 3, %0;" ::"n"(WL_EPI_THREADS) : "memory");
          if (issuer) {
            const uint32_t src = smem_base + W5_OFF_STG + (q & 1) * (2 * WL_A_BYTES);
            tma_store_4d(&map_acts, src, (q * 2) * WL_BK, t0, pm ? r : b, pm ? b : 0);
            tma_store_4d(&map_acts, src + WL_A_BYTES, (q * 2 + 1) * WL_BK, t0, pm ? r : b, pm ? b : 0);
            bulk_commit();
          }
        }
      }
      // fold accumulator (fixed combination order -> bit-reproducible)
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = o8p[j].x + o8p[j].y;
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
      asm volatile("bar.sync 1, %0;" ::"n"(WL_EPI_THREADS) : "memory");
    }
    if (!LAST && issuer) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Gate kernel on a CTA PAIR (tcgen05 cta_group::2), phase-major layout only: the same computation and the same bits as
// tc512_gate_kernel; each CTA owns one 128-row tile of the same phase, the weights are split in halves across the two
// CTAs' shared memory (protocol as in tc_pair_kernels.cuh: both CTAs' TMA bytes complete on the leader's full barrier,
// multicast tcgen05.commit frees stages / publishes accumulators in both CTAs, the epilogues' "drained" arrivals go to
// the leader). A stage is 16 KB A + 16 KB B: the gate kernel streams 5.6 MB per tile through its SM's TMA port (39 GB
// per K3 launch = 76 % of the measured streaming ceiling) -- 3.7 MB here -- and the ring is 4 deep instead of 3.
// ------------------------------------------------------------------------------------------------
constexpr int W5P_STAGES = 4;
constexpr int W5P_OFF_STG = W5P_STAGES * WP_STAGE_BYTES;     // 128 KB ring, then 64 KB of staging tiles
constexpr int W5P_OFF_B = W5P_OFF_STG + W5_STG_BYTES;
constexpr int W5P_OFF_O8 = W5P_OFF_B + 4 * W5_C * 2;
constexpr int W5P_OFF_BARS = W5P_OFF_O8 + WL_BM * 8 * 4;
constexpr int W5P_NBARS = 2 * W5P_STAGES + 2 + 2;
constexpr int W5P_SMEM = W5P_OFF_BARS + W5P_NBARS * 8 + 16;
static_assert(W5P_SMEM <= 232448, "shared memory budget");

template <bool LAST, bool FIRST = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WL_THREADS, 1)
tc512_gate_pair_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_cond,
                       const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_wc,
                       const __grid_constant__ CUtensorMap map_acts, const __grid_constant__ CUtensorMap map_a0,
                       const __grid_constant__ CUtensorMap map_w0, const WnLayerParams p,
                       const __grid_constant__ Wn512Const cw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + W5P_OFF_B);
  float* s_o8 = reinterpret_cast<float*>(smem + W5P_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W5P_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + W5P_NBARS);
  const uint32_t bar_base = smem_base + W5P_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                           // leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (W5P_STAGES + s); };           // both CTAs
  auto dfull_bar = [&](int x) { return bar_base + 8u * (2 * W5P_STAGES + x); };       // both CTAs
  auto drained_bar = [&](int x) { return bar_base + 8u * (2 * W5P_STAGES + 2 + x); }; // leader only, 2 x epilogue threads

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((smem_base & 1023u) != 0u) __trap();
  const int pairs_per_row = (p.tiles_per_row + 1) / 2;
  const int n_pair_tiles = pairs_per_row * p.R;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  auto tile_coords = [&](int pt, int& r, int& t0) {      // phase fastest; the rank-th tile of the pair's 256-row range
    r = pt % p.R;
    t0 = (2 * (pt / p.R) + static_cast<int>(rank)) * WL_BM;
  };
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    prefetch_tmap(&map_cond);
    prefetch_tmap(&map_w1);
    prefetch_tmap(&map_wc);
    prefetch_tmap(&map_acts);
    for (int s = 0; s < W5P_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(dfull_bar(x), 1);
      mbar_init(drained_bar(x), 2 * WL_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_slot), 512);
    tmem2_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * W5_C; i += WL_THREADS) s_b1[i] = p.b1[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int KB_CONV = FIRST ? 1 : W5_KB_CONV;
  const int kb1 = KB_CONV + p.n_cond_kb;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ============================
    uint32_t it = 0;
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs) {
      int r, t0;
      tile_coords(pt, r, t0);
      for (int q = 0; q < 4; ++q) {
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const int s = it % W5P_STAGES;
          mbar_wait(empty_bar(s), ((it / W5P_STAGES) & 1) ^ 1);
          if (leader && elect_one()) mbar_expect_tx(full_bar(s), 2 * WP_STAGE_BYTES);
          __syncwarp();
          if (elect_one()) {
            const uint32_t fb = full_bar(s) & kPeerBitMask;
            const uint32_t a_dst = smem_base + s * WP_STAGE_BYTES;
            if (FIRST && kb == 0) {
              tma2_load_4d(a_dst, &map_a0, fb, 0, t0, r, 0);
              tma2_load_2d(a_dst + WP_A_BYTES, &map_w0, fb, 0, p.flow * 2 * W5_C + q * 256 + rank * 128);
            } else if (kb < KB_CONV) {
              const int tap = kb >> 3, cblk = kb & 7;
              const int rs = r + (tap - 1) * p.dilation;
              const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
              tma2_load_4d(a_dst, &map_h, fb, cblk * WL_BK, t0 + carry, rs - carry * p.R, 0);
              tma2_load_2d(a_dst + WP_A_BYTES, &map_w1, fb, kb * WL_BK, p.layer * 2 * W5_C + q * 256 + rank * 128);
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
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
      uint32_t it = 0, fills = 0;   // fills = chunks issued so far (region = fills & 1)
      for (int pt = pair; pt < n_pair_tiles; pt += n_pairs) {
        for (int q = 0; q < 4; ++q, ++fills) {
          const uint32_t x = fills & 1u, f = fills >> 1;      // f-th fill of region x
          if (f > 0) {
            mbar_wait(drained_bar(x), (f - 1) & 1u);
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + 256u * x;
          for (int kb = 0; kb < kb1; ++kb, ++it) {
            const int s = it % W5P_STAGES;
            mbar_wait(full_bar(s), (it / W5P_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_base + s * WP_STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + WP_A_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < WL_BK / 16; ++k) {
                if (FIRST && kb == 0 && k == 3) break;   // columns 48..63 of a0: the residual operand (tc512_res_kernel)
                umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              }
              tc2_commit(empty_bar(s));
              if (kb == kb1 - 1) tc2_commit(dfull_bar(x));
            }
            __syncwarp();
          }
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
    uint8_t* stg_all = smem + W5P_OFF_STG;
    const bool issuer = we == 0 && lane == 0;
    const uint32_t r_drained0 = mapa_u32(drained_bar(0), 0), r_drained1 = mapa_u32(drained_bar(1), 0);
    uint32_t fills = 0;
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs) {
      int r, t0;
      tile_coords(pt, r, t0);
      const bool valid = wn_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      float2 o8p[8];   // (even-channel, odd-channel) partial sums of the fold columns (gate_step2)
#pragma unroll
      for (int j = 0; j < 8; ++j) o8p[j] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++fills) {
        const uint32_t x = fills & 1u, f = fills >> 1;
        mbar_wait(dfull_bar(x), f & 1u);
        tc_fence_after();
        uint8_t* stg = stg_all + (q & 1) * (2 * WL_A_BYTES);
        if (!LAST) {
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 3, %0;" ::"n"(WL_EPI_THREADS) : "memory");
        }
        const uint32_t taddr = tmem_base + lane_addr + 256u * x + hf * 32;
        uint32_t t0r[16], g0r[16], t1r[16], g1r[16];
        tmem_ld16(taddr, t0r);
        tmem_ld16(taddr + 128, g0r);
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          uint8_t* kblk = stg + blk * WL_A_BYTES + row * 128;
          const int ch0 = blk * 64 + hf * 32;
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
          gate_step2<LAST>(t1r, g1r, bT0 + 16, wse0 + 64, kblk, hf * 2 + 1, row, o8p);
        }
        tc_fence_before();
        mbar_arrive_cluster(x == 0 ? r_drained0 : r_drained1);
        if (!LAST) {
          fence_proxy_async_smem();
          asm volatile("bar.sync 3, %0;" ::"n"(WL_EPI_THREADS) : "memory");
          if (issuer) {
            const uint32_t src = smem_base + W5P_OFF_STG + (q & 1) * (2 * WL_A_BYTES);
            tma_store_4d(&map_acts, src, (q * 2) * WL_BK, t0, r, 0);
            tma_store_4d(&map_acts, src + WL_A_BYTES, (q * 2 + 1) * WL_BK, t0, r, 0);
            bulk_commit();
          }
        }
      }
      // fold accumulator (fixed combination order -> bit-reproducible, the same as tc512_gate_kernel)
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = o8p[j].x + o8p[j].y;
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
      asm volatile("bar.sync 1, %0;" ::"n"(WL_EPI_THREADS) : "memory");
    }
    if (!LAST && issuer) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Residual kernel: D2 chunk nn (output channels 256 nn .. 256 nn + 255) -> TMEM region nn.
// Stages per chunk: 8 x [acts block kb | W2 rows of the chunk, K-block kb], then 2 x [hi blocks] and 2 x [lo blocks]
// of the chunk's four 64-channel blocks (identity adds).
// ------------------------------------------------------------------------------------------------
// FIRST: the residual operand h0 = [a(l), 1] @ [Wstart; bstart] comes from the a0 tile (one K = 16 MMA per chunk)
// instead of four identity stages on (hi, lo).
template <bool FIRST = false>
__global__ void __launch_bounds__(WL_THREADS, 1)
tc512_res_kernel(const __grid_constant__ CUtensorMap map_acts, const __grid_constant__ CUtensorMap map_h,
                 const __grid_constant__ CUtensorMap map_ho, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_a0,
                 const __grid_constant__ CUtensorMap map_h0, const WnLayerParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b2 = reinterpret_cast<float*>(smem + W5_OFF_B);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W5_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + W5_NBARS);
  const uint32_t bar_base = smem_base + W5_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (W5_STAGES + s); };
  auto dfull_bar = [&](int x) { return bar_base + 8u * (2 * W5_STAGES + x); };
  auto drained_bar = [&](int x) { return bar_base + 8u * (2 * W5_STAGES + 2 + x); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_base & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_acts);
    prefetch_tmap(&map_h);
    prefetch_tmap(&map_ho);
    prefetch_tmap(&map_lo);
    prefetch_tmap(&map_w2);
    for (int s = 0; s < W5_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(dfull_bar(x), 1);
      mbar_init(drained_bar(x), WL_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < W5_C; i += WL_THREADS) s_b2[i] = p.b2[i];
  {
    uint32_t* i64w = reinterpret_cast<uint32_t*>(smem + W5_OFF_I64);
    for (int i = threadIdx.x; i < 64 * 32; i += WL_THREADS) {
      const int n = i >> 5, w = i & 31;
      const int chunk_log = (w >> 2) ^ (n & 7);
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool pm = p.R > 1;
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
  constexpr int KB2 = W5_C / WL_BK;   // 8
  constexpr int NSTEP = FIRST ? KB2 + 1 : KB2 + 4;

  if (warp == 0) {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int b, r, t0;
      tile_coords(tile, b, r, t0);
      const int c2 = pm ? r : b, c3 = pm ? b : 0;
      for (int nn = 0; nn < 2; ++nn) {
        for (int step = 0; step < NSTEP; ++step, ++it) {
          const int s = it % W5_STAGES;
          mbar_wait(empty_bar(s), ((it / W5_STAGES) & 1) ^ 1);
          if (elect_one()) {
            const uint32_t dst = smem_base + s * WL_STAGE_BYTES;
            if (FIRST && step == KB2) {
              mbar_expect_tx(full_bar(s), WL_STAGE_BYTES);
              tma_load_4d(dst, &map_a0, full_bar(s), 0, t0, c2, c3);
              tma_load_2d(dst + WL_A_BYTES, &map_h0, full_bar(s), 0, p.flow * W5_C + nn * 256);
            } else if (step < KB2) {
              mbar_expect_tx(full_bar(s), WL_STAGE_BYTES);
              tma_load_4d(dst, &map_acts, full_bar(s), step * WL_BK, t0, c2, c3);
              tma_load_2d(dst + WL_A_BYTES, &map_w2, full_bar(s), step * WL_BK, p.layer * W5_C + nn * 256);
            } else {
              // identity stages: two 64-channel blocks of hi (steps KB2, KB2+1) or lo (KB2+2, KB2+3)
              const int j = step - KB2;
              const CUtensorMap* mp = j < 2 ? &map_h : &map_lo;
              const int blk = nn * 4 + (j & 1) * 2;
              mbar_expect_tx(full_bar(s), 2 * WL_A_BYTES);
              tma_load_4d(dst, mp, full_bar(s), blk * WL_BK, t0, c2, c3);
              tma_load_4d(dst + WL_A_BYTES, mp, full_bar(s), (blk + 1) * WL_BK, t0, c2, c3);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    constexpr uint32_t idesc_id = umma_idesc_bf16(128, 64);
    const uint64_t idesc64 = umma_desc_sw128(smem_base + W5_OFF_I64);
    uint32_t it = 0, n = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n) {
      for (int nn = 0; nn < 2; ++nn) {
        if (n > 0) {
          mbar_wait(drained_bar(nn), (n - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + 256u * nn;
        for (int step = 0; step < NSTEP; ++step, ++it) {
          const int s = it % W5_STAGES;
          mbar_wait(full_bar(s), (it / W5_STAGES) & 1);
          tc_fence_after();
          const uint32_t st_addr = smem_base + s * WL_STAGE_BYTES;
          if (elect_one()) {
            if (FIRST && step == KB2) {
              umma_bf16(d_tmem, umma_desc_sw128(st_addr) + 6, umma_desc_sw128(st_addr + WL_A_BYTES) + 6, idesc, 1u);
            } else if (step < KB2) {
              const uint64_t adesc = umma_desc_sw128(st_addr), bdesc = umma_desc_sw128(st_addr + WL_A_BYTES);
#pragma unroll
              for (int k = 0; k < WL_BK / 16; ++k)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (step | k) ? 1u : 0u);
            } else {
              const int j = step - KB2;
              const int blk_local = (j & 1) * 2;     // column block inside this 256-column chunk
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const uint64_t adesc = umma_desc_sw128(st_addr + h2 * WL_A_BYTES);
#pragma unroll
                for (int k = 0; k < WL_BK / 16; ++k)
                  umma_bf16(d_tmem + 64u * (blk_local + h2), adesc + 2 * k, idesc64 + 2 * k, idesc_id, 1u);
              }
            }
            tc_commit(empty_bar(s));
            if (step == NSTEP - 1) tc_commit(dfull_bar(nn));
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int we = warp - 2;
    const int quarter = warp & 3;
    const int hf = we >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* stg = smem + W5_OFF_STG + (hf * 2) * WL_A_BYTES + row * 128;
    const uint32_t stg_addr = smem_base + W5_OFF_STG + (hf * 2) * WL_A_BYTES;
    const bool issuer = (we == hf * 4) && lane == 0;
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n) {
      int b, r, t0;
      tile_coords(tile, b, r, t0);
      const int c2 = pm ? r : b, c3 = pm ? b : 0;
      const bool valid = wn_row_valid(p, t0 + row);   // gap rows are stored as zeros
#pragma unroll 1
      for (int nn = 0; nn < 2; ++nn) {
        mbar_wait(dfull_bar(nn), n & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + 256u * nn + hf * 128;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          // the staging half is reused: the previous store from it must have finished reading shared memory
          if (issuer) bulk_wait_read0();
          if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr, r0);
#pragma unroll 1
          for (int gp = 0; gp < 4; ++gp) {
            tmem_ld_wait();
            tmem_ld16(taddr + (2 * gp + 1) * 16, r1);
            resid_step(r0, s_b2 + nn * 256 + hf * 128 + (2 * gp) * 16, stg, 2 * gp, row, pass, valid);
            tmem_ld_wait();
            if (gp < 3) tmem_ld16(taddr + (2 * gp + 2) * 16, r0);
            else if (pass == 1) {
              tc_fence_before();
              mbar_arrive(drained_bar(nn));
            }
            resid_step(r1, s_b2 + nn * 256 + hf * 128 + (2 * gp + 1) * 16, stg, 2 * gp + 1, row, pass, valid);
          }
          fence_proxy_async_smem();
          if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          if (issuer) {
            const CUtensorMap* om = pass == 0 ? &map_ho : &map_lo;
            tma_store_4d(om, stg_addr, (nn * 4 + hf * 2) * WL_BK, t0, c2, c3);
            tma_store_4d(om, stg_addr + WL_A_BYTES, (nn * 4 + hf * 2 + 1) * WL_BK, t0, c2, c3);
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0 && (we == 0 || we == 4)) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline void tc512_init() {
  WG_CK(cudaFuncSetAttribute(tc512_gate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_gate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_gate_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_res_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_res_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_gate_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5P_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_gate_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5P_SMEM));
  WG_CK(cudaFuncSetAttribute(tc512_gate_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5P_SMEM));
}

// 128-row weight boxes for the CTA-pair gate kernel (each CTA loads its half of a 256-row B chunk)
struct Tc512PairMaps {
  CUtensorMap m_w1, m_wc, m_w0;
  int max_pairs = 0;
  bool ready = false;
};

inline void tc512_pair_prepare(Tc512PairMaps& pm, const TcPlan& pl, int n_layers_total, int n_flows, int R, const __nv_bfloat16* W1,
                               const __nv_bfloat16* V, const __nv_bfloat16* W0, int max_pairs) {
  const int C = pl.C;
  make_map_2d(&pm.m_w1, W1, (uint64_t)n_layers_total * 2 * C, 3 * C + pl.S, 128);
  make_map_2d(&pm.m_wc, V, (uint64_t)n_layers_total * R * 2 * C, pl.Kup, 128);
  if (pl.fold0) make_map_2d(&pm.m_w0, W0, (uint64_t)n_flows * 2 * C, WL_BK, 128);
  else pm.m_w0 = pm.m_w1;
  pm.max_pairs = max_pairs;
  pm.ready = true;
}

// one WN layer of WaveGlow-512: gate kernel (on CTA pairs when `pairs` is ready), then -- unless it is the last layer of the
// flow -- the residual kernel
inline int tc512_wn_layer(const TcPlan& pl, const CUtensorMap& m_acts, int layer, int dilation, bool last, int hcur,
                          float* acc8, const float* b1, const float* b2, const float* wse_host, cudaStream_t st,
                          bool first = false, const Tc512PairMaps* pairs = nullptr) {
  WnLayerParams p{};
  tc_fill_params(pl, p, layer, dilation, hcur, acc8, b1, b2, nullptr, 0);
  Wn512Const cw;
  std::memcpy(cw.wse, wse_host, sizeof cw.wse);   // wse_host: packed-fold layout (LayerW::wse_p)
  const int grid = pl.n_tiles < pl.sm_count ? pl.n_tiles : pl.sm_count;
  if (first && (!pl.fold0 || last || dilation != 1)) fail(WG_ERR_INVALID, "start fold requested for a layer it does not apply to");
  const bool pair = pairs && pairs->ready && pl.pm;
  int grid2 = 0;
  if (pair) {
    const int need = ((pl.tiles_per_row + 1) / 2) * pl.R;
    grid2 = 2 * (need < pairs->max_pairs ? need : pairs->max_pairs);
  }
  if (last) {
    if (pair) tc512_gate_pair_kernel<true><<<grid2, WL_THREADS, W5P_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pairs->m_w1, pairs->m_wc, m_acts, pl.m4_a0, pairs->m_w0, p, cw);
    else tc512_gate_kernel<true><<<grid, WL_THREADS, W5_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pl.m_w1, pl.m_wc, m_acts, pl.m4_a0, pl.m_w0, p, cw);
    WG_CK(cudaGetLastError());
    return 1;
  }
  if (first) {
    if (pair) tc512_gate_pair_kernel<false, true><<<grid2, WL_THREADS, W5P_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pairs->m_w1, pairs->m_wc, m_acts, pl.m4_a0, pairs->m_w0, p, cw);
    else tc512_gate_kernel<false, true><<<grid, WL_THREADS, W5_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pl.m_w1, pl.m_wc, m_acts, pl.m4_a0, pl.m_w0, p, cw);
    WG_CK(cudaGetLastError());
    tc512_res_kernel<true><<<grid, WL_THREADS, W5_SMEM, st>>>(m_acts, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m_w2, pl.m4_a0, pl.m_h0, p);
    WG_CK(cudaGetLastError());
    return 2;
  }
  if (pair) tc512_gate_pair_kernel<false><<<grid2, WL_THREADS, W5P_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pairs->m_w1, pairs->m_wc, m_acts, pl.m4_a0, pairs->m_w0, p, cw);
  else tc512_gate_kernel<false><<<grid, WL_THREADS, W5_SMEM, st>>>(pl.m4_h[hcur], pl.m4_cond, pl.m_w1, pl.m_wc, m_acts, pl.m4_a0, pl.m_w0, p, cw);
  WG_CK(cudaGetLastError());
  tc512_res_kernel<false><<<grid, WL_THREADS, W5_SMEM, st>>>(m_acts, pl.m4_h[hcur], pl.m4_h[hcur ^ 1], pl.m4_lo, pl.m_w2, pl.m4_a0, pl.m_h0, p);
  WG_CK(cudaGetLastError());
  return 2;
}

}  // namespace wg

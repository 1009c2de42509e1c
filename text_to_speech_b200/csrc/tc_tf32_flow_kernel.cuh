// WG_MODE_TF32X3, single short utterance: ALL WN layers of one flow in ONE persistent launch.
//
// When the whole batch is one wave (every CTA pair owns exactly one (tile pair, chunk) item -- K1, 1 x 200 frames, is 64
// pairs on 74), the per-layer kernels of tc_tf32_kernels.cuh pay a kernel boundary twice per layer (drain, launch,
// prologue, ring fill: ~7 us each against ~50 us of work). This kernel keeps the CTAs resident for the n_layers layers
// of a flow and replaces the kernel boundaries by a grid barrier in HBM (arrive counter + generation word):
//
//   layer l:   [wait: h of layer l-1 complete everywhere]  gate GEMM (taps of h, mel window) -> gate epilogue -> acts
//              [grid barrier]                               residual GEMM (acts) -> residual epilogue -> h of layer l
//              [grid barrier]
//
// Same items, same MMAs in the same order, same epilogue arithmetic as tf32_gate_kernel / tf32_res_kernel (CTA pairs,
// 32-float K-blocks, 16 epilogue warps, staged TMA stores): the waveform is bit-identical to theirs.
// It is launched COOPERATIVELY (all CTAs co-resident or the launch fails), and every spin has a time-out trap.
#pragma once
#include "tc_tf32_kernels.cuh"

namespace wg {

constexpr int T3F_MAX_LAYERS = 16;
constexpr int T3F_SYNC_WORDS = 2 + 2 * 128;          // grid barrier + one barrier per tile pair (at most one CTA pair each: <= 74)
constexpr int T3F_EW = 16, T3F_EPI_THREADS = T3F_EW * 32, T3F_THREADS = 64 + T3F_EPI_THREADS;
using T3FG = T3G<32, true>;                      // gate ring: 3 stages of 64 KB
using T3FR = T3RG<true>;                         // residual ring: 4 stages of 48 KB (the same shared memory)
constexpr int T3F_RING_BYTES = 192 * 1024;
static_assert(T3FG::STAGES * T3FG::STAGE_BYTES <= T3F_RING_BYTES && T3FR::STAGES * T3FR::STAGE_BYTES <= T3F_RING_BYTES, "ring");
constexpr int T3F_OFF_B1 = T3F_RING_BYTES;                       // [256] gate bias of this CTA's chunk
constexpr int T3F_OFF_B2 = T3F_OFF_B1 + 256 * 4;                 // [128] residual bias of this CTA's chunk
constexpr int T3F_OFF_O8 = T3F_OFF_B2 + T3R_BN * 4;              // fold partials of column groups 1..3
constexpr int T3F_OFF_BARS = T3F_OFF_O8 + 3 * T3_BM * 8 * 4;
constexpr int T3F_NBARS = 2 * T3FG::STAGES + 2 * T3FR::STAGES + 2;
constexpr int T3F_SMEM = T3F_OFF_BARS + T3F_NBARS * 8 + 16;      // + TMEM slot, 2 generation words
static_assert(T3F_SMEM <= 232448, "shared memory budget");

struct Tf32FlowMaps {
  // "?h": fp32 words holding tf32(x); "?b": the bf16 companion [.., 2K] = bf16(hi) | bf16(lo)  (tc_tf32_kernels.cuh)
  CUtensorMap hh[2], hb[2];        // residual stream, ping-pong: 128-row boxes (gate A operand, taps)
  CUtensorMap ch, cb;              // mel window
  CUtensorMap w1h, w1b, vh, vb;    // gate weights, 128-row boxes (one CTA's half of a chunk)
  CUtensorMap ah, ab;              // acts (residual A operand)
  CUtensorMap w2h, w2b;            // residual weights, 64-row boxes
  CUtensorMap sah, sab;            // 32 x 32 store boxes: acts
  CUtensorMap shh[2], shl[2], shb[2];   // 32 x 32 store boxes: residual stream (hi, exact lo, bf16 companion)
};

struct Tf32FlowParams {
  Tf32Params base;                 // geometry, acts / acc8 pointers (layer, dilation, bias, wse, h_* unused)
  int n_layers, layer0, hcur0;     // layers of this flow, global index of its first layer, ping-pong index of its input
  int dilation[T3F_MAX_LAYERS];
  const float* b1[T3F_MAX_LAYERS];
  const float* b2[T3F_MAX_LAYERS];
  const float* wse[T3F_MAX_LAYERS];
  const float* h_hi[2];            // residual stream, generic reads of the residual epilogue
  const float* h_lo[2];
  unsigned int* sync;              // grid barrier: [0] arrivals, [1] generation; tile-pair barriers: [2 + 2 i], [3 + 2 i]
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// One arrival per CTA (called by one thread once the CTA's stores are complete); the last arriver re-arms the counter
// and publishes the next generation.
__device__ __forceinline__ void t3f_grid_arrive(unsigned int* sync, unsigned int n_ctas) {
  __threadfence();
  const unsigned int prev = atomicAdd(sync, 1u);
  if (prev == n_ctas - 1u) {
    atomicExch(sync, 0u);
    __threadfence();
    atomicAdd(sync + 1, 1u);
  }
}
// Spin until `target` barriers have completed (generation counter, wrap-safe); traps instead of hanging the device.
__device__ __forceinline__ void t3f_grid_wait(const unsigned int* sync, unsigned int target) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_gpu(sync + 1) - target) < 0) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
  fence_proxy_async_all();   // what other SMs wrote is read through TMA (async proxy) from here on
}

__global__ void __launch_bounds__(T3F_THREADS, 1)
tf32_flow_kernel(const __grid_constant__ Tf32FlowMaps maps, const __grid_constant__ Tf32FlowParams fp) {
  constexpr int GS = T3FG::STAGES, GSB = T3FG::STAGE_BYTES, GA = T3FG::A_BYTES, GB = T3FG::B_BYTES;
  constexpr int RS = T3FR::STAGES, RSB = T3FR::STAGE_BYTES, RB = T3FR::B_BYTES;
  constexpr int BK = 32, NCG = T3F_EW / 4, CH = 128 / NCG;   // 32 channels / columns per epilogue thread
  static_assert(CH == 32, "one 128-byte row per thread");
  const Tf32Params& p = fp.base;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + T3F_OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + T3F_OFF_B2);
  float* s_o8 = reinterpret_cast<float*>(smem + T3F_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T3F_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T3F_NBARS);
  const uint32_t bar_base = smem_base + T3F_OFF_BARS;
  auto gfull = [&](int s) { return bar_base + 8u * s; };                       // leader only
  auto gempty = [&](int s) { return bar_base + 8u * (GS + s); };               // both CTAs (multicast commit)
  auto rfull = [&](int s) { return bar_base + 8u * (2 * GS + s); };
  auto rempty = [&](int s) { return bar_base + 8u * (2 * GS + RS + s); };
  const uint32_t gacc_bar = bar_base + 8u * (2 * GS + 2 * RS);                 // gate accumulator complete (both CTAs)
  const uint32_t racc_bar = gacc_bar + 8u;                                     // residual accumulator complete
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((smem_base & 1023u) != 0u) __trap();
  // WG_LAYER_TIMING=1 (debug): cycle counters, slots 48.. of Tf32Params::timing (see tools/layer_timing_tf32.py)
  unsigned long long* const tm = fp.base.timing;
  const long long t_entry = tm ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(&maps);
    for (int i = 0; i < static_cast<int>(sizeof(Tf32FlowMaps) / sizeof(CUtensorMap)); ++i) prefetch_tmap(m + i);
    for (int s = 0; s < GS; ++s) { mbar_init(gfull(s), 1); mbar_init(gempty(s), 1); }
    for (int s = 0; s < RS; ++s) { mbar_init(rfull(s), 1); mbar_init(rempty(s), 1); }
    mbar_init(gacc_bar, 1);
    mbar_init(racc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_slot), 512);
    tmem2_relinquish();
  }
  // the generation the barrier stands at when this launch starts (no CTA arrives before the cluster sync below)
  // Two barriers: the GRID one (words 0, 1; h of a layer complete everywhere -- the taps of the next layer read other tiles
  // and other phases) and one per TILE PAIR (words 2 + 2 * tile pair ..; acts complete -- the residual GEMM reads all C
  // channels of its OWN rows, written by the CTA pairs of that tile pair's chunks only).
  volatile uint32_t* s_gen0 = tmem_slot + 1;
  unsigned int* const tsync = fp.sync + 2 + 2 * ((blockIdx.x >> 1) / (2 * p.C / T3G_BN));
  if (threadIdx.x == 0) {
    s_gen0[0] = ld_acquire_gpu(fp.sync + 1);
    s_gen0[1] = ld_acquire_gpu(tsync + 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const unsigned int gen0 = s_gen0[0], tgen0 = s_gen0[1];
  const unsigned int n_ctas = gridDim.x;

  // this CTA's item, the same in every layer: pair item -> (chunk q, phase r, first row t0 of this CTA's tile)
  const int n_chunks = 2 * p.C / T3G_BN;       // = C / 128: gate chunks == residual chunks
  int q, r, t0;
  {
    const int item = blockIdx.x >> 1;
    q = item % n_chunks;
    const int tile = item / n_chunks;
    r = tile % p.R;
    t0 = (2 * (tile / p.R) + static_cast<int>(rank)) * T3_BM;
  }
  const int kb_conv = 3 * p.C / BK, kb_cond = fp.base.kb_cond, kb1 = kb_conv + kb_cond, kb2 = p.C / BK;
  const int cblks = p.C / BK;

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    uint32_t itg = 0, itr = 0;
    long long t_gw = 0, t_gw1 = 0;
    for (int l = 0; l < fp.n_layers; ++l) {
      const bool last = l == fp.n_layers - 1;
      const int hcur = fp.hcur0 ^ (l & 1);
      const int layer = fp.layer0 + l, dil = fp.dilation[l];
      if (l > 0) {
        const long long tq = tm ? clock64() : 0;
        if (lane == 0) t3f_grid_wait(fp.sync, gen0 + l);   // every CTA's h of layer l-1 is in memory
        __syncwarp();
        if (tm) t_gw += clock64() - tq;
      }
      const int bq = q * T3G_BN + static_cast<int>(rank) * (T3G_BN / 2);
      for (int kb = 0; kb < kb1; ++kb, ++itg) {
        const int s = itg % GS;
        mbar_wait(gempty(s), ((itg / GS) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(gfull(s), 2 * GSB);
          const uint32_t fb = gfull(s) & kPeerBitMask;
          const uint32_t a_hi = smem_base + s * GSB, a_b = a_hi + GA, b_hi = a_b + GA, b_b = b_hi + GB;
          if (kb < kb_conv) {
            const int tap = kb / cblks, cblk = kb - tap * cblks;
            const int rs = r + (tap - 1) * dil;
            const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
            tma2_load_4d(a_hi, &maps.hh[hcur], fb, cblk * BK, t0 + carry, rs - carry * p.R, 0);
            t3_load_b_4d<true>(a_b, &maps.hb[hcur], gfull(s), cblk * BK, t0 + carry, rs - carry * p.R, 0);
            tma2_load_2d(b_hi, &maps.w1h, fb, kb * BK, layer * 2 * p.C + bq);
            t3_load_b_2d<true>(b_b, &maps.w1b, gfull(s), kb * BK, layer * 2 * p.C + bq);
          } else {
            const int kc = kb - kb_conv;
            const int vrow = layer * p.R * 2 * p.C + r * 2 * p.C + bq;
            tma2_load_4d(a_hi, &maps.ch, fb, kc * BK, t0, 0, 0);
            t3_load_b_4d<true>(a_b, &maps.cb, gfull(s), kc * BK, t0, 0, 0);
            tma2_load_2d(b_hi, &maps.vh, fb, kc * BK, vrow);
            t3_load_b_2d<true>(b_b, &maps.vb, gfull(s), kc * BK, vrow);
          }
        }
        __syncwarp();
      }
      if (last) break;
      const long long tq1 = tm ? clock64() : 0;
      if (lane == 0) t3f_grid_wait(tsync, tgen0 + l + 1u);   // the acts of this tile pair (all chunks) are in memory
      __syncwarp();
      if (tm) t_gw1 += clock64() - tq1;
      const int b2q = q * T3R_BN + static_cast<int>(rank) * (T3R_BN / 2);
      for (int kb = 0; kb < kb2; ++kb, ++itr) {
        const int s = itr % RS;
        mbar_wait(rempty(s), ((itr / RS) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(rfull(s), 2 * RSB);
          const uint32_t fb = rfull(s) & kPeerBitMask;
          const uint32_t a_hi = smem_base + s * RSB, a_b = a_hi + T3_A_BYTES, b_hi = a_b + T3_A_BYTES, b_b = b_hi + RB;
          tma2_load_4d(a_hi, &maps.ah, fb, kb * BK, t0, r, 0);
          t3_load_b_4d<true>(a_b, &maps.ab, rfull(s), kb * BK, t0, r, 0);
          tma2_load_2d(b_hi, &maps.w2h, fb, kb * BK, layer * p.C + b2q);
          t3_load_b_2d<true>(b_b, &maps.w2b, rfull(s), kb * BK, layer * p.C + b2q);
        }
        __syncwarp();
      }
    }
    if (tm && lane == 0) {
      atomicAdd(tm + 48, static_cast<unsigned long long>(t_gw));
      atomicAdd(tm + 57, static_cast<unsigned long long>(t_gw1));
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (the leader CTA of the pair) =======================
    if (leader) {
      constexpr uint32_t idesc_g = umma_idesc_tf32(2 * T3_BM, T3G_BN), idesc_r = umma_idesc_tf32(2 * T3_BM, T3R_BN);
      constexpr uint32_t idesc16_g = umma_idesc_bf16(2 * T3_BM, T3G_BN), idesc16_r = umma_idesc_bf16(2 * T3_BM, T3R_BN);
      const uint32_t d_g = tmem_base, d_r = tmem_base + 256u;
      uint32_t itg = 0, itr = 0;
      long long t_gf = 0, t_rf = 0;
      for (int l = 0; l < fp.n_layers; ++l) {
        const bool last = l == fp.n_layers - 1;
        for (int kb = 0; kb < kb1; ++kb, ++itg) {
          const int s = itg % GS;
          const long long tq = tm ? clock64() : 0;
          mbar_wait(gfull(s), (itg / GS) & 1);
          if (tm) t_gf += clock64() - tq;
          tc_fence_after();
          const uint32_t base = smem_base + s * GSB;
          if (elect_one()) {
            t3_mma_stage<true>(d_g, base, GA, base + 2 * GA, GB, idesc_g, idesc16_g, kb == 0);
            tc2_commit(gempty(s));
            if (kb == kb1 - 1) tc2_commit(gacc_bar);
          }
          __syncwarp();
        }
        if (last) break;
        for (int kb = 0; kb < kb2; ++kb, ++itr) {
          const int s = itr % RS;
          const long long tq = tm ? clock64() : 0;
          mbar_wait(rfull(s), (itr / RS) & 1);
          if (tm) t_rf += clock64() - tq;
          tc_fence_after();
          const uint32_t base = smem_base + s * RSB;
          if (elect_one()) {
            t3_mma_stage<true>(d_r, base, T3_A_BYTES, base + 2 * T3_A_BYTES, RB, idesc_r, idesc16_r, kb == 0);
            tc2_commit(rempty(s));
            if (kb == kb2 - 1) tc2_commit(racc_bar);
          }
          __syncwarp();
        }
      }
      if (tm && lane == 0) {
        atomicAdd(tm + 49, static_cast<unsigned long long>(t_gf));
        atomicAdd(tm + 50, static_cast<unsigned long long>(t_rf));
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int we = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int cg = we >> 2;             // column group: 32 of the chunk's 128 gate channels / residual columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    // staging tiles of this warp: gate (acts) hi 4 KB | companion 4 KB;  residual (h) hi | lo | companion, 4 KB each
    const uint32_t stg = smem_base + static_cast<uint32_t>(we) * 12288u;
    const bool valid = t3_row_valid(p, t0 + row);
    const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
    const int tid = threadIdx.x - 64;
    for (int l = 0; l < fp.n_layers; ++l) {
      const bool last = l == fp.n_layers - 1;
      const int hcur = fp.hcur0 ^ (l & 1);
      // ---- biases of this CTA's chunks (the previous layer's readers passed bar.sync 3 / the fold's bar.sync 2)
      for (int i = tid; i < 256; i += T3F_EPI_THREADS) s_b1[i] = fp.b1[l][q * 256 + i];
      if (!last)
        for (int i = tid; i < T3R_BN; i += T3F_EPI_THREADS) s_b2[i] = fp.b2[l][q * T3R_BN + i];
      asm volatile("bar.sync 1, %0;" ::"n"(T3F_EPI_THREADS) : "memory");
      // ---- gate epilogue (tf32_gate_kernel, 16 warps, staged stores)
      const long long te0 = tm ? clock64() : 0;
      mbar_wait(gacc_bar, l & 1);
      const long long te1 = tm ? clock64() : 0;
      tc_fence_after();
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = 0.f;
      {
        const uint32_t taddr = tmem_base + lane_addr + cg * CH;
#pragma unroll 1
        for (int g2 = 0; g2 < 2; ++g2) {       // 16 gate channels per step
          uint32_t tr[16], gr[16];
          tmem_ld16(taddr + g2 * 16, tr);
          tmem_ld16(taddr + 128 + g2 * 16, gr);
          tmem_ld_wait();
          const int ch0 = cg * CH + g2 * 16;
          const float* wse = fp.wse[l] + static_cast<size_t>(q * 128 + ch0) * 8;
          float a[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xt = __uint_as_float(tr[j]) + s_b1[ch0 + j];
            const float xg = __uint_as_float(gr[j]) + s_b1[128 + ch0 + j];
            a[j] = t3_gate_act(xt, xg);                              // waveglow_arch.py:19-24
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wse + j * 8));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wse + j * 8 + 4));
            o8[0] = fmaf(a[j], w0.x, o8[0]); o8[1] = fmaf(a[j], w0.y, o8[1]);
            o8[2] = fmaf(a[j], w0.z, o8[2]); o8[3] = fmaf(a[j], w0.w, o8[3]);
            o8[4] = fmaf(a[j], w1.x, o8[4]); o8[5] = fmaf(a[j], w1.y, o8[5]);
            o8[6] = fmaf(a[j], w1.z, o8[6]); o8[7] = fmaf(a[j], w1.w, o8[7]);
          }
          if (!last) {
            T3Split16 sp;
            t3_split16(a, valid, sp);
            t3_stage_f32(stg, lane, g2, sp.hi);
            t3_stage_b16(stg + 4096u, lane, g2, sp.hb, false);
            t3_stage_b16(stg + 4096u, lane, g2, sp.lb, true);
          }
        }
      }
      if (!last) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&maps.sah, stg, q * 128 + cg * CH, t0 + quarter * 32, r, 0);
          tma_store_4d(&maps.sab, stg + 4096u, 2 * (q * 128 + cg * CH), t0 + quarter * 32, r, 0);
          bulk_commit();
        }
      }
      tc_fence_before();
      // skip/end fold: (P0 + P1) + (P2 + P3), one partial per column group, into this chunk's own accumulator
      if (cg > 0) {
        float* d = s_o8 + (static_cast<size_t>(cg - 1) * T3_BM + row) * 8;
        *reinterpret_cast<float4*>(d) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(T3F_EPI_THREADS) : "memory");
      if (cg == 0 && valid) {
        float tot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          tot[j] = (o8[j] + s_o8[row * 8 + j]) + (s_o8[(T3_BM + row) * 8 + j] + s_o8[(2 * T3_BM + row) * 8 + j]);
        float4* o = reinterpret_cast<float4*>(p.acc8 + q * p.acc8_stride + m * 8);
        float4 a0 = o[0], a1 = o[1];
        a0.x += tot[0]; a0.y += tot[1]; a0.z += tot[2]; a0.w += tot[3];
        a1.x += tot[4]; a1.y += tot[5]; a1.z += tot[6]; a1.w += tot[7];
        o[0] = a0; o[1] = a1;
      }
      if (tm && tid == 0) {
        atomicAdd(tm + 51, static_cast<unsigned long long>(te1 - te0));
        atomicAdd(tm + 52, static_cast<unsigned long long>(clock64() - te1));
      }
      if (last) break;
      // ---- acts complete -> grid barrier (the residual GEMM of every CTA reads all C channels of its rows)
      const long long te2 = tm ? clock64() : 0;
      if (lane == 0) bulk_wait0();
      asm volatile("bar.sync 3, %0;" ::"n"(T3F_EPI_THREADS) : "memory");
      if (tid == 0) t3f_grid_arrive(tsync, 2u * static_cast<unsigned int>(n_chunks));
      const long long te3 = tm ? clock64() : 0;
      // ---- residual epilogue (tf32_res_kernel): the old value of h is fetched while the residual GEMM runs
      const size_t off0 = m * p.C + q * T3R_BN + cg * CH;
      // (L2 loads: this buffer was read two layers ago and rewritten since by TMA stores, which do not update L1)
      float old[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) old[j] = 0.f;
      if (valid) {
        const float* hh = fp.h_hi[hcur] + off0;
        const float* hl = fp.h_lo[hcur] + off0;
#pragma unroll
        for (int v = 0; v < CH / 4; ++v) {
          const float4 oh = __ldcg(reinterpret_cast<const float4*>(hh + 4 * v));
          const float4 ol = __ldcg(reinterpret_cast<const float4*>(hl + 4 * v));
          old[4 * v] = oh.x + ol.x; old[4 * v + 1] = oh.y + ol.y; old[4 * v + 2] = oh.z + ol.z; old[4 * v + 3] = oh.w + ol.w;
        }
      }
      mbar_wait(racc_bar, l & 1);
      const long long te4 = tm ? clock64() : 0;
      tc_fence_after();
      {
        const uint32_t taddr = tmem_base + lane_addr + 256u + cg * CH;
#pragma unroll
        for (int g = 0; g < CH / 16; ++g) {
          uint32_t rr[16];
          tmem_ld16(taddr + g * 16, rr);
          tmem_ld_wait();
          // h_new = (acts @ Wres + b) + h   (waveglow_arch.py:131-133); gap rows: the next layer's zero padding
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = (__uint_as_float(rr[j]) + s_b2[cg * CH + g * 16 + j]) + old[g * 16 + j];
          T3Split16 sp;
          t3_split16(x, valid, sp);
          t3_stage_f32(stg, lane, g, sp.hi);
          t3_stage_f32(stg + 4096u, lane, g, sp.lo);
          t3_stage_b16(stg + 8192u, lane, g, sp.hb, false);
          t3_stage_b16(stg + 8192u, lane, g, sp.lb, true);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&maps.shh[hcur ^ 1], stg, q * T3R_BN + cg * CH, t0 + quarter * 32, r, 0);
        tma_store_4d(&maps.shl[hcur ^ 1], stg + 4096u, q * T3R_BN + cg * CH, t0 + quarter * 32, r, 0);
        tma_store_4d(&maps.shb[hcur ^ 1], stg + 8192u, 2 * (q * T3R_BN + cg * CH), t0 + quarter * 32, r, 0);
        bulk_commit();
        bulk_wait0();
      }
      tc_fence_before();
      asm volatile("bar.sync 3, %0;" ::"n"(T3F_EPI_THREADS) : "memory");
      if (tid == 0) t3f_grid_arrive(fp.sync, n_ctas);
      if (tm && tid == 0) {
        atomicAdd(tm + 53, static_cast<unsigned long long>(te3 - te2));       // store completion + arrive (acts)
        atomicAdd(tm + 54, static_cast<unsigned long long>(te4 - te3));       // arrive -> residual accumulator complete
        atomicAdd(tm + 55, static_cast<unsigned long long>(clock64() - te4)); // residual epilogue incl. store completion + arrive
      }
    }
    if (tm && tid == 0) {
      atomicAdd(tm + 56, static_cast<unsigned long long>(clock64() - t_entry));
      atomicAdd(tm + 58, 1ull);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, 512);
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
struct Tf32FlowState {
  unsigned int* sync = nullptr;   // device: [arrivals, generation]
  int max_pairs = 0;              // co-resident CTA pairs of tf32_flow_kernel (0: unavailable)
  bool cooperative = true;        // launch attribute: all CTAs co-resident or the launch fails (WG_TF32_COOP=0: plain launch)
  bool refuse = false;            // WG_TF32_FLOW=2 (tests): behave as if the device refused the cooperative launch
};

inline int tf32_flow_init() {
  WG_CK(cudaFuncSetAttribute(tf32_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3F_SMEM));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.gridDim = dim3(2 * 1024); cfg.blockDim = dim3(T3F_THREADS); cfg.dynamicSmemBytes = T3F_SMEM;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tf32_flow_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

// Can this plan run a flow as one launch?  Every pair item needs its own CTA pair, all co-resident.
inline bool tf32_flow_fits(const Tf32Plan& pl, int max_pairs, int n_layers) {
  if (!pl.pair || max_pairs < 1 || n_layers > T3F_MAX_LAYERS) return false;
  const int n_chunks = 2 * pl.base.C / T3G_BN;
  const int items = ((pl.base.tiles_per_row + 1) / 2) * pl.base.R * n_chunks;
  return items <= max_pairs;
}

// All `n_layers` layers of one flow (dilation 2^i). Returns the launch count (1), or 0 when the device refuses the
// cooperative launch (not all CTAs can be co-resident right now): the caller falls back to the per-layer kernels.
inline int tf32_wn_flow(const Tf32Plan& pl, const Tf32FlowState& fs, int layer0, int n_layers, int hcur0, const float* const* b1,
                        const float* const* b2, const float* const* wse, cudaStream_t st, unsigned long long* timing = nullptr) {
  if (fs.refuse) return 0;
  Tf32FlowMaps m;
  for (int i = 0; i < 2; ++i) {
    m.hh[i] = pl.m_h_hi[i]; m.hb[i] = pl.m_h_b[i];
    m.shh[i] = pl.s_h_hi[i]; m.shl[i] = pl.s_h_lo[i]; m.shb[i] = pl.s_h_b[i];
  }
  m.ch = pl.m_c_hi; m.cb = pl.m_c_b;
  m.w1h = pl.p_w1h; m.w1b = pl.p_w1b; m.vh = pl.p_vh; m.vb = pl.p_vb;
  m.ah = pl.m_a_hi; m.ab = pl.m_a_b; m.w2h = pl.p_w2h; m.w2b = pl.p_w2b;
  m.sah = pl.s_a_hi; m.sab = pl.s_a_b;
  Tf32FlowParams fp{};
  fp.base = pl.base;
  fp.base.kb_cond = pl.Kup / 32;
  fp.base.timing = timing;
  fp.n_layers = n_layers; fp.layer0 = layer0; fp.hcur0 = hcur0;
  for (int i = 0; i < n_layers; ++i) {
    fp.dilation[i] = 1 << i;
    fp.b1[i] = b1[i]; fp.b2[i] = b2[i]; fp.wse[i] = wse[i];
  }
  for (int i = 0; i < 2; ++i) { fp.h_hi[i] = pl.h_hi[i]; fp.h_lo[i] = pl.h_lo[i]; }
  fp.sync = fs.sync;
  const int n_chunks = 2 * pl.base.C / T3G_BN;
  const int items = ((pl.base.tiles_per_row + 1) / 2) * pl.base.R * n_chunks;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * items);
  cfg.blockDim = dim3(T3F_THREADS);
  cfg.dynamicSmemBytes = T3F_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = fs.cooperative ? 2 : 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, tf32_flow_kernel, m, fp);
  if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
    cudaGetLastError();
    return 0;
  }
  WG_CK(le);
  return 1;
}

}  // namespace wg

// WG_MODE_TF32X3 -- the fp32-grade WN layer on the 5th-gen tensor cores (tcgen05.mma kind::tf32), sm_100a only.
//
// Reference arithmetic: architectures/waveglow_arch.py:19-24 (gate), :105-141 (WN layer) in fp32. The tensor cores
// read TF32 (10 explicit mantissa bits), so every fp32 operand x travels as the pair
//     x_hi = tf32(x)  (cvt.rna)          x_lo = x - x_hi   (exact in fp32; the MMA reads its top 11 bits)
// and every product is issued three times:  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (the dropped a_lo*b_lo term is
// <= 2^-22 relative), accumulated in fp32 in TMEM: "3xTF32" (BASELINE.json north_star). The gate uses the accurate
// tanhf / expf. Everything else -- phase-major gapped row layout (RowGeom), the conditioning folded to the 4-frame mel
// window (V = Wup_r @ Wcond, rank 320, exact in fp32), the skip->end fold into [C, 8] -- is the algebra of the BF16
// path (tc_kernels.cuh), so ragged batches (wg_infer_ragged) work the same way.
//
// Two kernels per layer (the fp32 acts tile, 128 x C x 2 x 4 B, does not fit beside the operand ring):
//   tf32_gate_kernel   item = (128-row tile, 256-column chunk q of the gate pre-activation = 128 tanh + 128 sigmoid
//                      channels).  GEMM1 [128 x (3C + 320)] @ [(3C + 320) x 256], K-blocks of 16 floats (64-byte
//                      swizzled rows), 4-stage TMA ring of {A_hi, A_lo, B_hi, B_lo} = 48 KB, 6 MMAs per stage
//                      (WG_TF32_BK=32: K-blocks of 32 floats, 2 stages of 96 KB -- measured 7.5 % slower at K1);
//                      epilogue: gate, acts -> (hi, lo) in HBM, skip/end fold into this chunk's OWN partial
//                      accumulator acc8[q] (no cross-CTA race: the flow boundary sums the partials in a fixed order).
//   tf32_res_kernel    item = (128-row tile, 128-column chunk of the residual half).  GEMM2 acts[128 x C] @ Wres,
//                      3-stage ring of 64 KB; epilogue: h = acc + b + (h_hi + h_lo) -> (hi, lo) for the next layer.
// Both are persistent (static schedule) with a double-buffered TMEM accumulator so the epilogue of one item overlaps
// the MMAs of the next. K1 (1 x 200 frames) is 32 phases x 2 tiles = 64 tiles -> 128 gate items: one item per CTA, one wave.
#pragma once
#include "tc_kernels.cuh"

namespace wg {

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor, kind::tf32: D = f32 (bit 4), A = B = TF32 (format 2 in bits [7,10) and [10,13)), K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__host__ __device__ inline float tf32_rna_host(float x) {   // same rounding on the host (weights are split at wg_create)
#ifdef __CUDA_ARCH__
  return tf32_rna(x);
#else
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
  float y;
  std::memcpy(&y, &u, 4);
  return y;
#endif
}

constexpr int T3_BM = 128, T3_BK = 32;                 // 32 floats = one 128-byte swizzle row
constexpr int T3_A_BYTES = T3_BM * T3_BK * 4;          // 16 KB
constexpr int T3_EPI_WARPS = 8, T3_EPI_THREADS = T3_EPI_WARPS * 32, T3_THREADS = 64 + T3_EPI_THREADS;
// gate kernel: N = 256 per item. K-block width BK floats per stage: 32 (one 128-byte swizzle row, 2 stages of 96 KB) or
// 16 (64-byte swizzle rows, 4 stages of 48 KB): the same 192 KB ring in finer slices keeps more bytes in flight while the
// MMA works on a stage -- the kernel is bound by the latency of its operand feed (8 B per operand element for the fp32 pairs).
constexpr int T3G_BN = 256;
template <int BK>
struct T3G {
  static_assert(BK == 32 || BK == 16, "K-block of 32 floats (SWIZZLE_128B) or 16 floats (SWIZZLE_64B)");
  static constexpr int A_BYTES = T3_BM * BK * 4, B_BYTES = T3G_BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;                // 96 KB / 48 KB
  static constexpr int STAGES = BK == 32 ? 2 : 4;
  static constexpr int OFF_B1 = STAGES * STAGE_BYTES;
  static constexpr int OFF_O8 = OFF_B1 + 256 * 4;
  static constexpr int OFF_BARS = OFF_O8 + T3_BM * 8 * 4;
  static constexpr int NBARS = 2 * STAGES + 4;
  static constexpr int SMEM = OFF_BARS + NBARS * 8 + 16;
  static_assert(SMEM <= 232448, "shared memory budget");
};
// K-major SWIZZLE_64B shared-memory matrix descriptor: 8-row groups of 64-byte rows (SBO = 512 B), layout type 4
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
template <int BK>
__device__ __forceinline__ uint64_t t3_desc(uint32_t smem_addr) {
  return BK == 32 ? umma_desc_sw128(smem_addr) : umma_desc_sw64(smem_addr);
}
// residual kernel: N = 128 per item
constexpr int T3R_BN = 128, T3R_B_BYTES = T3R_BN * T3_BK * 4;                 // 16 KB
constexpr int T3R_STAGES = 3, T3R_STAGE_BYTES = 2 * T3_A_BYTES + 2 * T3R_B_BYTES;   // 64 KB
constexpr int T3R_OFF_B2 = T3R_STAGES * T3R_STAGE_BYTES;
constexpr int T3R_OFF_BARS = T3R_OFF_B2 + T3R_BN * 4;
constexpr int T3R_NBARS = 2 * T3R_STAGES + 4;
constexpr int T3R_SMEM = T3R_OFF_BARS + T3R_NBARS * 8 + 16;
static_assert(T3R_SMEM <= 232448, "shared memory budget");

struct Tf32Params {
  int T, R, tiles_per_row, n_tiles;   // phase-block rows, phases, 128-row tiles per phase block, tiles in all
  int n_chunks;        // items per tile: gate 2C/256, residual C/128
  int C;
  int kb_conv;         // gate: 3C/32 conv K-blocks;  residual: C/32
  int kb_cond;         // gate: K-blocks of the mel window (320/32 = 10)
  int wc_row0, wc_rstride;   // first row of this layer's folded conditioning weights in V, rows per phase (2C)
  int Tp, Tv;
  const int* row_b;    // ragged validity table (see WnLayerParams)
  int layer, dilation;
  const float* bias;   // gate: [2C] chunk-packed (cond bias folded in);  residual: [C]
  const float* wse;    // gate: [C, 8] = Wskip @ Wend (fp32)
  float* acts_hi;      // gate out / residual A operand: [rows, C] internal row order
  float* acts_lo;
  float* acc8;         // gate: partial fold accumulators [n_chunks][rows][8]
  size_t acc8_stride;  // floats between two partials
  const float* h_hi;   // residual epilogue: the residual stream of this layer (read), [rows, C]
  const float* h_lo;
  float* ho_hi;        // residual epilogue: the next layer's stream (written; gap rows as zeros)
  float* ho_lo;
};

__device__ __forceinline__ bool t3_row_valid(const Tf32Params& p, int t) {
  if (t >= p.T) return false;
  if (p.row_b) return p.row_b[t] >= 0;
  return p.Tp == 0 || t % p.Tp < p.Tv;
}

// item -> (tile, chunk); tile -> (phase r, first row t0): phase fastest, as in the BF16 kernel's tile order
__device__ __forceinline__ void t3_item_coords(const Tf32Params& p, int item, int& q, int& r, int& t0) {
  q = item % p.n_chunks;
  const int tile = item / p.n_chunks;
  r = tile % p.R;
  t0 = (tile / p.R) * T3_BM;
}

template <bool LAST, int BK = 32>
__global__ void __launch_bounds__(T3_THREADS, 1)
tf32_gate_kernel(const __grid_constant__ CUtensorMap map_hh, const __grid_constant__ CUtensorMap map_hl,
                 const __grid_constant__ CUtensorMap map_ch, const __grid_constant__ CUtensorMap map_cl,
                 const __grid_constant__ CUtensorMap map_w1h, const __grid_constant__ CUtensorMap map_w1l,
                 const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl,
                 const Tf32Params p) {
  using G = T3G<BK>;
  constexpr int T3G_STAGES = G::STAGES, T3G_STAGE_BYTES = G::STAGE_BYTES, T3G_B_BYTES = G::B_BYTES, T3G_A_BYTES = G::A_BYTES;
  constexpr int T3G_OFF_B1 = G::OFF_B1, T3G_OFF_O8 = G::OFF_O8, T3G_OFF_BARS = G::OFF_BARS, T3G_NBARS = G::NBARS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b1 = reinterpret_cast<float*>(smem + T3G_OFF_B1);
  float* s_o8 = reinterpret_cast<float*>(smem + T3G_OFF_O8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T3G_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T3G_NBARS);
  const uint32_t bar_base = smem_base + T3G_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (T3G_STAGES + s); };
  auto accfull_bar = [&](int s) { return bar_base + 8u * (2 * T3G_STAGES + s); };
  auto accempty_bar = [&](int s) { return bar_base + 8u * (2 * T3G_STAGES + 2 + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_base & 1023u) != 0u) __trap();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_hh); prefetch_tmap(&map_hl); prefetch_tmap(&map_ch); prefetch_tmap(&map_cl);
    prefetch_tmap(&map_w1h); prefetch_tmap(&map_w1l); prefetch_tmap(&map_vh); prefetch_tmap(&map_vl);
    for (int s = 0; s < T3G_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), T3_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_tiles * p.n_chunks;
  const int kb1 = p.kb_conv + p.kb_cond;
  const int cblks = p.C / BK;   // K-blocks per tap

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int q, r, t0;
      t3_item_coords(p, item, q, r, t0);
      for (int kb = 0; kb < kb1; ++kb, ++it) {
        const int s = it % T3G_STAGES;
        mbar_wait(empty_bar(s), ((it / T3G_STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), T3G_STAGE_BYTES);
          const uint32_t a_hi = smem_base + s * T3G_STAGE_BYTES, a_lo = a_hi + T3G_A_BYTES;
          const uint32_t b_hi = a_lo + T3G_A_BYTES, b_lo = b_hi + T3G_B_BYTES;
          if (kb < p.kb_conv) {
            // tap shifted by (tap-1)*dilation positions: phase (r+sh) mod R, frames moved by floor((r+sh)/R)
            const int tap = kb / cblks, cblk = kb - tap * cblks;
            const int rs = r + (tap - 1) * p.dilation;
            const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
            tma_load_4d(a_hi, &map_hh, full_bar(s), cblk * BK, t0 + carry, rs - carry * p.R, 0);
            tma_load_4d(a_lo, &map_hl, full_bar(s), cblk * BK, t0 + carry, rs - carry * p.R, 0);
            tma_load_2d(b_hi, &map_w1h, full_bar(s), kb * BK, p.layer * 2 * p.C + q * T3G_BN);
            tma_load_2d(b_lo, &map_w1l, full_bar(s), kb * BK, p.layer * 2 * p.C + q * T3G_BN);
          } else {
            const int kc = kb - p.kb_conv;
            tma_load_4d(a_hi, &map_ch, full_bar(s), kc * BK, t0, 0, 0);
            tma_load_4d(a_lo, &map_cl, full_bar(s), kc * BK, t0, 0, 0);
            tma_load_2d(b_hi, &map_vh, full_bar(s), kc * BK, p.wc_row0 + r * p.wc_rstride + q * T3G_BN);
            tma_load_2d(b_lo, &map_vl, full_bar(s), kc * BK, p.wc_row0 + r * p.wc_rstride + q * T3G_BN);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_tf32(T3_BM, T3G_BN);
    uint32_t it = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t as = n & 1u;
      mbar_wait(accempty_bar(as), ((n >> 1) & 1u) ^ 1u);   // the epilogue has read this accumulator out (2 items ago)
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 256u * as;
      for (int kb = 0; kb < kb1; ++kb, ++it) {
        const int s = it % T3G_STAGES;
        mbar_wait(full_bar(s), (it / T3G_STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_base + s * T3G_STAGE_BYTES;
        const uint64_t ahi = t3_desc<BK>(base), alo = t3_desc<BK>(base + T3G_A_BYTES);
        const uint64_t bhi = t3_desc<BK>(base + 2 * T3G_A_BYTES), blo = t3_desc<BK>(base + 2 * T3G_A_BYTES + T3G_B_BYTES);
        if (elect_one()) {
          // small cross terms first, then the main product (K = 8 floats = 32 bytes per instruction: +2 in the descriptor)
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) umma_tf32(d_tmem, alo + 2 * k, bhi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) umma_tf32(d_tmem, ahi + 2 * k, blo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) umma_tf32(d_tmem, ahi + 2 * k, bhi + 2 * k, idesc, 1u);
          tc_commit(empty_bar(s));
          if (kb == kb1 - 1) tc_commit(accfull_bar(as));
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int we = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int hf = we >> 2;             // which half of the 128 gate channels this warp handles
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int q, r, t0;
      t3_item_coords(p, item, q, r, t0);
      const uint32_t as = n & 1u;
      const bool valid = t3_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      // this chunk's bias: [tanh 128 | sigmoid 128]
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");   // previous item's s_b1 / s_o8 readers are done
      for (int i = threadIdx.x - 64; i < 256; i += T3_EPI_THREADS) s_b1[i] = p.bias[q * 256 + i];
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      mbar_wait(accfull_bar(as), (n >> 1) & 1u);
      tc_fence_after();
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = 0.f;
      const uint32_t taddr = tmem_base + lane_addr + 256u * as + hf * 64;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {       // 16 gate channels per step
        uint32_t tr[16], gr[16];
        tmem_ld16(taddr + g * 16, tr);
        tmem_ld16(taddr + 128 + g * 16, gr);
        tmem_ld_wait();
        const int ch0 = hf * 64 + g * 16;                 // first channel inside the chunk
        const float* wse = p.wse + static_cast<size_t>(q * 128 + ch0) * 8;
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float xt = __uint_as_float(tr[j]) + s_b1[ch0 + j];
          const float xg = __uint_as_float(gr[j]) + s_b1[128 + ch0 + j];
          a[j] = tanhf(xt) * (1.0f / (1.0f + expf(-xg)));          // waveglow_arch.py:19-24
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wse + j * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wse + j * 8 + 4));
          o8[0] = fmaf(a[j], w0.x, o8[0]); o8[1] = fmaf(a[j], w0.y, o8[1]);
          o8[2] = fmaf(a[j], w0.z, o8[2]); o8[3] = fmaf(a[j], w0.w, o8[3]);
          o8[4] = fmaf(a[j], w1.x, o8[4]); o8[5] = fmaf(a[j], w1.y, o8[5]);
          o8[6] = fmaf(a[j], w1.z, o8[6]); o8[7] = fmaf(a[j], w1.w, o8[7]);
        }
        if (!LAST && (t0 + row) < p.T) {
          float hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            hi[j] = valid ? tf32_rna(a[j]) : 0.f;
            lo[j] = valid ? a[j] - hi[j] : 0.f;
          }
          float4* dh = reinterpret_cast<float4*>(p.acts_hi + m * p.C + q * 128 + ch0);
          float4* dl = reinterpret_cast<float4*>(p.acts_lo + m * p.C + q * 128 + ch0);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            dh[v] = make_float4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
            dl[v] = make_float4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(accempty_bar(as));
      // fold: the two column halves of a row are summed in a fixed order, then added to this chunk's own partial
      if (hf == 1) {
        *reinterpret_cast<float4*>(s_o8 + row * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(s_o8 + row * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      if (hf == 0 && valid) {
        const float4 p0 = *reinterpret_cast<const float4*>(s_o8 + row * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(s_o8 + row * 8 + 4);
        float4* o = reinterpret_cast<float4*>(p.acc8 + q * p.acc8_stride + m * 8);
        float4 a0 = o[0], a1 = o[1];
        a0.x += o8[0] + p0.x; a0.y += o8[1] + p0.y; a0.z += o8[2] + p0.z; a0.w += o8[3] + p0.w;
        a1.x += o8[4] + p1.x; a1.y += o8[5] + p1.y; a1.z += o8[6] + p1.z; a1.w += o8[7] + p1.w;
        o[0] = a0; o[1] = a1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(T3_THREADS, 1)
tf32_res_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_w2h, const __grid_constant__ CUtensorMap map_w2l,
                const Tf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* s_b2 = reinterpret_cast<float*>(smem + T3R_OFF_B2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T3R_OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T3R_NBARS);
  const uint32_t bar_base = smem_base + T3R_OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (T3R_STAGES + s); };
  auto accfull_bar = [&](int s) { return bar_base + 8u * (2 * T3R_STAGES + s); };
  auto accempty_bar = [&](int s) { return bar_base + 8u * (2 * T3R_STAGES + 2 + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_base & 1023u) != 0u) __trap();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_ah); prefetch_tmap(&map_al); prefetch_tmap(&map_w2h); prefetch_tmap(&map_w2l);
    for (int s = 0; s < T3R_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), T3_EPI_THREADS);
    }
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
  const int n_items = p.n_tiles * p.n_chunks;
  const int kb2 = p.kb_conv;

  if (warp == 0) {
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int q, r, t0;
      t3_item_coords(p, item, q, r, t0);
      for (int kb = 0; kb < kb2; ++kb, ++it) {
        const int s = it % T3R_STAGES;
        mbar_wait(empty_bar(s), ((it / T3R_STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), T3R_STAGE_BYTES);
          const uint32_t a_hi = smem_base + s * T3R_STAGE_BYTES, a_lo = a_hi + T3_A_BYTES;
          const uint32_t b_hi = a_lo + T3_A_BYTES, b_lo = b_hi + T3R_B_BYTES;
          tma_load_4d(a_hi, &map_ah, full_bar(s), kb * T3_BK, t0, r, 0);
          tma_load_4d(a_lo, &map_al, full_bar(s), kb * T3_BK, t0, r, 0);
          tma_load_2d(b_hi, &map_w2h, full_bar(s), kb * T3_BK, p.layer * p.C + q * T3R_BN);
          tma_load_2d(b_lo, &map_w2l, full_bar(s), kb * T3_BK, p.layer * p.C + q * T3R_BN);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_tf32(T3_BM, T3R_BN);
    uint32_t it = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t as = n & 1u;
      mbar_wait(accempty_bar(as), ((n >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 128u * as;
      for (int kb = 0; kb < kb2; ++kb, ++it) {
        const int s = it % T3R_STAGES;
        mbar_wait(full_bar(s), (it / T3R_STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_base + s * T3R_STAGE_BYTES;
        const uint64_t ahi = umma_desc_sw128(base), alo = umma_desc_sw128(base + T3_A_BYTES);
        const uint64_t bhi = umma_desc_sw128(base + 2 * T3_A_BYTES), blo = umma_desc_sw128(base + 2 * T3_A_BYTES + T3R_B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < T3_BK / 8; ++k) umma_tf32(d_tmem, alo + 2 * k, bhi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < T3_BK / 8; ++k) umma_tf32(d_tmem, ahi + 2 * k, blo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < T3_BK / 8; ++k) umma_tf32(d_tmem, ahi + 2 * k, bhi + 2 * k, idesc, 1u);
          tc_commit(empty_bar(s));
          if (kb == kb2 - 1) tc_commit(accfull_bar(as));
        }
        __syncwarp();
      }
    }
  } else {
    const int we = warp - 2;
    const int quarter = warp & 3;
    const int hf = we >> 2;             // which 64 of the item's 128 columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int q, r, t0;
      t3_item_coords(p, item, q, r, t0);
      const uint32_t as = n & 1u;
      const bool in_range = (t0 + row) < p.T;
      const bool valid = t3_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      for (int i = threadIdx.x - 64; i < T3R_BN; i += T3_EPI_THREADS) s_b2[i] = p.bias[q * T3R_BN + i];
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      mbar_wait(accfull_bar(as), (n >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + 128u * as + hf * 64;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t rr[16];
        tmem_ld16(taddr + g * 16, rr);
        tmem_ld_wait();
        if (in_range) {
          const int c0 = q * T3R_BN + hf * 64 + g * 16;
          const size_t off = m * p.C + c0;
          float hi[16], lo[16];
          if (valid) {
            // h_new = (acts @ Wres + b) + h   with h = h_hi + h_lo exactly (waveglow_arch.py:131-133)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const float4 oh = *reinterpret_cast<const float4*>(p.h_hi + off + 4 * v);
              const float4 ol = *reinterpret_cast<const float4*>(p.h_lo + off + 4 * v);
              const float old[4] = {oh.x + ol.x, oh.y + ol.y, oh.z + ol.z, oh.w + ol.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = (__uint_as_float(rr[4 * v + j]) + s_b2[hf * 64 + g * 16 + 4 * v + j]) + old[j];
                hi[4 * v + j] = tf32_rna(x);
                lo[4 * v + j] = x - hi[4 * v + j];
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) hi[j] = lo[j] = 0.f;     // gap row: the next layer's zero padding
          }
          float4* dh = reinterpret_cast<float4*>(p.ho_hi + off);
          float4* dl = reinterpret_cast<float4*>(p.ho_lo + off);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            dh[v] = make_float4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
            dl[v] = make_float4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(accempty_bar(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- helper kernels ---------------------------------------------------------------------------------------------
// mel window (A operand of the conditioning) as an fp32 (hi, lo) pair; rows follow one phase block (RowGeom, R = 1)
__global__ void tf32_im2col_kernel(const float* __restrict__ mel, float* __restrict__ a_hi, float* __restrict__ a_lo,
                                   const RowGeom geo, int n_mel, int Kup) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(geo.rows_per_phase()) * Kup) return;
  const int kk = static_cast<int>(idx % Kup);
  const int row = static_cast<int>(idx / Kup);
  int b, t;
  const bool valid = geo.decode_row(row, b, t);
  const int j = kk / n_mel, i = kk - j * n_mel;
  float v = 0.f;
  if (valid && j < 4 && t - j >= 0) v = mel[(static_cast<size_t>(b) * geo.T + t - j) * n_mel + i];
  const float hi = tf32_rna(v);
  a_hi[idx] = hi;
  a_lo[idx] = v - hi;
}

// folded conditioning weights: in [(r, k), n] fp32 -> (hi, lo)[(r*N + n)*K + k]
__global__ void fold_store_tf32_kernel(const float* __restrict__ in, float* __restrict__ o_hi, float* __restrict__ o_lo,
                                       int R, int K, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(R) * K * N) return;
  const int k = static_cast<int>(idx % K);
  const size_t rn = idx / K;
  const int n = static_cast<int>(rn % N), r = static_cast<int>(rn / N);
  const float v = in[(static_cast<size_t>(r) * K + k) * N + n];
  const float hi = tf32_rna(v);
  o_hi[idx] = hi;
  o_lo[idx] = v - hi;
}

// ---- host side --------------------------------------------------------------------------------------------------
inline void make_map_f32(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
  const CUtensorMapSwizzle swz = box[0] == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;   // 64- or 128-byte rows
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  uint64_t stride = 4;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(WG_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
}
inline void make_map_f32_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int bk = T3_BK) {
  const uint64_t dims[2] = {cols, rows};
  const uint32_t box[2] = {(uint32_t)bk, box_rows};
  make_map_f32(m, ptr, 2, dims, box);
}
inline void make_map_f32_4d(CUtensorMap* m, const void* ptr, uint64_t phases, uint64_t rows, uint64_t cols, int bk = T3_BK) {
  const uint64_t dims[4] = {cols, rows, phases, 1};
  const uint32_t box[4] = {(uint32_t)bk, T3_BM, 1, 1};
  make_map_f32(m, ptr, 4, dims, box);
}

inline void tf32_init() {
  WG_CK(cudaFuncSetAttribute(tf32_gate_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3G<32>::SMEM));
  WG_CK(cudaFuncSetAttribute(tf32_gate_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3G<32>::SMEM));
  WG_CK(cudaFuncSetAttribute(tf32_gate_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3G<16>::SMEM));
  WG_CK(cudaFuncSetAttribute(tf32_gate_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3G<16>::SMEM));
  WG_CK(cudaFuncSetAttribute(tf32_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3R_SMEM));
}

struct Tf32Plan {
  CUtensorMap m_h_hi[2], m_h_lo[2], m_c_hi, m_c_lo, m_w1h, m_w1l, m_vh, m_vl, m_a_hi, m_a_lo, m_w2h, m_w2l;
  Tf32Params base{};
  RowGeom geo1{};
  int sm_count = 0, n_mel = 0, Kup = 0;
  int gate_bk = 16;      // K-block width of the gate kernel (WG_TF32_BK=32 selects the 2-stage SWIZZLE_128B variant)
  float *h_hi[2] = {nullptr, nullptr}, *h_lo[2] = {nullptr, nullptr}, *aup_hi = nullptr, *aup_lo = nullptr;
};

struct Tf32Weights {   // device pointers, stacked over all layers
  const float *W1h, *W1l;   // [n_layers_total * 2C, 3C]  chunk-packed rows, K-major
  const float *Vh, *Vl;     // [n_layers_total * R * 2C, Kup]
  const float *W2h, *W2l;   // [n_layers_total * C, C]
};

// rows1 = rows of one phase block (B*(T+gap), or the ragged total); geo1 = that block's utterance geometry (R = 1)
inline void tf32_prepare(Tf32Plan& pl, int sm_count, int C, int R, int Kup, int n_mel, int n_layers_total, int rows1,
                         const RowGeom& geo1, int Tp, int Tv, const Tf32Weights& w, float* h_hi0, float* h_hi1,
                         float* h_lo0, float* h_lo1, float* aup_hi, float* aup_lo, float* acts_hi, float* acts_lo,
                         float* acc8, size_t acc8_stride, int gate_bk = 16) {
  if (C % 128 || Kup % T3_BK) fail(WG_ERR_UNSUPPORTED, "tf32x3 path needs n_channels %% 128 == 0 (got %d)", C);
  pl.gate_bk = gate_bk == 32 ? 32 : 16;
  const int gbk = pl.gate_bk;
  pl.sm_count = sm_count; pl.n_mel = n_mel; pl.Kup = Kup; pl.geo1 = geo1;
  pl.h_hi[0] = h_hi0; pl.h_hi[1] = h_hi1; pl.h_lo[0] = h_lo0; pl.h_lo[1] = h_lo1; pl.aup_hi = aup_hi; pl.aup_lo = aup_lo;
  Tf32Params& p = pl.base;
  p.T = rows1; p.R = R; p.tiles_per_row = (rows1 + T3_BM - 1) / T3_BM; p.n_tiles = p.tiles_per_row * R;
  p.C = C; p.Tp = Tp; p.Tv = Tv; p.row_b = geo1.row_b;
  p.acts_hi = acts_hi; p.acts_lo = acts_lo; p.acc8 = acc8; p.acc8_stride = acc8_stride;
  for (int i = 0; i < 2; ++i) {
    make_map_f32_4d(&pl.m_h_hi[i], pl.h_hi[i], R, rows1, C, gbk);
    make_map_f32_4d(&pl.m_h_lo[i], pl.h_lo[i], R, rows1, C, gbk);
  }
  make_map_f32_4d(&pl.m_c_hi, aup_hi, 1, rows1, Kup, gbk);
  make_map_f32_4d(&pl.m_c_lo, aup_lo, 1, rows1, Kup, gbk);
  make_map_f32_4d(&pl.m_a_hi, acts_hi, R, rows1, C);
  make_map_f32_4d(&pl.m_a_lo, acts_lo, R, rows1, C);
  make_map_f32_2d(&pl.m_w1h, w.W1h, (uint64_t)n_layers_total * 2 * C, 3 * C, T3G_BN, gbk);
  make_map_f32_2d(&pl.m_w1l, w.W1l, (uint64_t)n_layers_total * 2 * C, 3 * C, T3G_BN, gbk);
  make_map_f32_2d(&pl.m_vh, w.Vh, (uint64_t)n_layers_total * R * 2 * C, Kup, T3G_BN, gbk);
  make_map_f32_2d(&pl.m_vl, w.Vl, (uint64_t)n_layers_total * R * 2 * C, Kup, T3G_BN, gbk);
  make_map_f32_2d(&pl.m_w2h, w.W2h, (uint64_t)n_layers_total * C, C, T3R_BN);
  make_map_f32_2d(&pl.m_w2l, w.W2l, (uint64_t)n_layers_total * C, C, T3R_BN);
}

inline int tf32_upsample(const Tf32Plan& pl, const float* mel, cudaStream_t st) {
  const size_t total = (size_t)pl.geo1.rows_per_phase() * pl.Kup;
  tf32_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mel, pl.aup_hi, pl.aup_lo, pl.geo1, pl.n_mel, pl.Kup);
  WG_CK(cudaGetLastError());
  return 1;
}

// One WN layer: gate kernel (+ residual kernel unless it is the flow's last layer). Returns the launch count.
inline int tf32_wn_layer(const Tf32Plan& pl, int layer, int dilation, bool last, int hcur, const float* b1, const float* b2,
                         const float* wse, cudaStream_t st) {
  Tf32Params g = pl.base;
  g.layer = layer; g.dilation = dilation;
  g.n_chunks = 2 * g.C / T3G_BN; g.kb_conv = 3 * g.C / pl.gate_bk; g.kb_cond = pl.Kup / pl.gate_bk;
  g.wc_row0 = layer * g.R * 2 * g.C; g.wc_rstride = 2 * g.C;
  g.bias = b1; g.wse = wse;
  const int items_g = g.n_tiles * g.n_chunks;
  const int grid_g = items_g < pl.sm_count ? items_g : pl.sm_count;
#define WG_T3G_LAUNCH(LASTV, BKV)                                                                                          \
  tf32_gate_kernel<LASTV, BKV><<<grid_g, T3_THREADS, T3G<BKV>::SMEM, st>>>(pl.m_h_hi[hcur], pl.m_h_lo[hcur], pl.m_c_hi, pl.m_c_lo, \
                                                                           pl.m_w1h, pl.m_w1l, pl.m_vh, pl.m_vl, g)
  if (pl.gate_bk == 32) {
    if (last) WG_T3G_LAUNCH(true, 32);
    else WG_T3G_LAUNCH(false, 32);
  } else {
    if (last) WG_T3G_LAUNCH(true, 16);
    else WG_T3G_LAUNCH(false, 16);
  }
#undef WG_T3G_LAUNCH
  WG_CK(cudaGetLastError());
  if (last) return 1;
  Tf32Params r = pl.base;
  r.layer = layer; r.dilation = dilation;
  r.n_chunks = r.C / T3R_BN; r.kb_conv = r.C / T3_BK; r.kb_cond = 0;
  r.bias = b2;
  r.h_hi = pl.h_hi[hcur]; r.h_lo = pl.h_lo[hcur]; r.ho_hi = pl.h_hi[hcur ^ 1]; r.ho_lo = pl.h_lo[hcur ^ 1];
  const int items_r = r.n_tiles * r.n_chunks;
  const int grid_r = items_r < pl.sm_count ? items_r : pl.sm_count;
  tf32_res_kernel<<<grid_r, T3_THREADS, T3R_SMEM, st>>>(pl.m_a_hi, pl.m_a_lo, pl.m_w2h, pl.m_w2l, r);
  WG_CK(cudaGetLastError());
  return 2;
}

}  // namespace wg

// WG_MODE_TF32X3 -- the fp32-grade WN layer on the 5th-gen tensor cores (tcgen05.mma), sm_100a only.
//
// Reference arithmetic: architectures/waveglow_arch.py:19-24 (gate), :105-141 (WN layer) in fp32. The tensor cores
// read at most TF32 (10 explicit mantissa bits), so every fp32 operand x is split as
//     x_hi = tf32(x)  (cvt.rna, kept in fp32 words)          x_lo = x - x_hi   (exact in fp32)
// and every multiply is issued as three products, small terms first, into one fp32 TMEM accumulator:
//     a_lo*b_hi + a_hi*b_lo   kind::f16 MMAs on a BF16 companion [.., 2K] of every operand (per K-block of 32: hi x 32 | lo x 32)
//                             (the cross terms are 2^-11 of the product: 8 mantissa bits carry them to ~2e-7)
//     a_hi*b_hi               kind::tf32 MMAs
// (the dropped a_lo*b_lo term is <= 2^-22 relative): the "fp32/3xTF32 mode" of BASELINE.json's north_star. The gate is
// evaluated in fp32 to ~2e-7 absolute (t3_gate_act). Everything else -- phase-major gapped row layout (RowGeom), the
// conditioning folded to the 4-frame mel window (V = Wup_r @ Wcond, rank 320, exact in fp32), the skip->end fold into
// [C, 8] -- is the algebra of the BF16 path (tc_kernels.cuh), so ragged batches (wg_infer_ragged) work the same way.
//
// Two kernels per layer (the fp32 acts tile does not fit beside the operand ring):
//   tf32_gate_kernel   item = (128-row tile, 256-column chunk q of the gate pre-activation = 128 tanh + 128 sigmoid
//                      channels).  GEMM1 [128 x (3C + 320)] @ [(3C + 320) x 256] in K-blocks of 32; one pipeline stage =
//                      [A_hi fp32 | A_b bf16][B_hi | B_b], 128-byte rows (t3_mma_stage);
//                      epilogue: gate, acts -> split -> HBM, skip/end fold into this chunk's OWN partial accumulator
//                      acc8[q] (no cross-CTA race: the flow boundary sums the partials in a fixed order).
//   tf32_res_kernel    item = (128-row tile, 128-column chunk of the residual half).  GEMM2 acts[128 x C] @ Wres;
//                      epilogue: h = acc + b + (h_hi + h_lo) -> split for the next layer (h keeps its exact fp32 x_lo).
// Both are persistent (static schedule) with a double-buffered TMEM accumulator so the epilogue of one item overlaps
// the MMAs of the next, and exist for single CTAs and for CTA pairs (PAIR: cta_group::2, tc_pair_kernels.cuh's
// protocol), with 8 or 16 epilogue warps (EW). All variants, and tf32_flow_kernel (tc_tf32_flow_kernel.cuh: one
// persistent launch per flow for a single utterance), produce the same bits.
#pragma once
#include "tc_pair_kernels.cuh"

namespace wg {

// cta_group::2 form: ONE instruction drives both tensor cores of a CTA pair (M = 256: 128 rows per CTA; each CTA's shared
// memory holds its own A rows and its own HALF of the B rows -- see tc_pair_kernels.cuh for the protocol)
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
// Gate activation tanh(x) * sigmoid(y) (waveglow_arch.py:19-24) with ONE reciprocal:
//   u = e^(-2|x|), v = e^(-y):   tanh(x) = sign(x) (1 - u) / (1 + u),  sigmoid(y) = 1 / (1 + v)
//   => a = sign(x) (1 - u) / ((1 + u)(1 + v))
// ex2.approx (<= 2 ulp) for both exponentials, rcp.approx + one Newton step (< 1 ulp) for the reciprocal: absolute error
// <= ~2e-7 on a value in (-1, 1) -- the cancellation in 1 - u near x = 0 costs relative, not absolute, accuracy -- which
// is 1/100 of this mode's measured waveform error. y is clamped at -87 so that (1 + u)(1 + v) stays finite (a -> 0).
// ~15 instructions instead of ~55 for tanhf + expf + IEEE division: the epilogue of a single-wave call is not hidden.
__device__ __forceinline__ float t3_gate_act(float x, float y) {
  float u, v, r;
  const float tu = -2.8853900817779268f * fabsf(x);                // -2 log2(e) |x|
  const float tv = -1.4426950408889634f * fmaxf(y, -87.0f);        // -log2(e) y
  asm("ex2.approx.f32 %0, %1;" : "=f"(u) : "f"(tu));
  asm("ex2.approx.f32 %0, %1;" : "=f"(v) : "f"(tv));
  const float v1 = 1.0f + v;
  const float d = fmaf(u, v1, v1);                                 // (1 + u)(1 + v)
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = r * fmaf(-d, r, 2.0f);
  return copysignf((1.0f - u) * r, x);
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
// Epilogue warps EW (template parameter, 8 or 16): NCG = EW / 4 warps share a TMEM lane quarter and split the accumulator
// columns. A single utterance runs ONE item per CTA, so the epilogue is not hidden behind the next item's MMAs and its
// latency (2 warps per scheduler with EW = 8) is paid per layer: 16 warps there (K1 7.8 -> 7.1 ms when it was introduced);
// with several items per CTA the epilogue is hidden and 8 warps are faster (8 x 860: 130 vs 145 ms). Same bits either way:
// the skip/end fold has one canonical order.
constexpr int t3_threads(int ew) { return 64 + ew * 32; }
// gate kernel: N = 256 per item, K-blocks of BK = 32 (a 128-byte fp32 row / a 64-byte bf16 row): 8 B per operand element
// (4 of hi + 2 + 2 of the bf16 companion), a 192 KB ring of 2 stages (single CTA) or 3 stages (CTA pair).
constexpr int T3G_BN = 256;
template <int BK, bool PAIR = false>
struct T3G {
  static_assert(BK == 32, "K-block of 32: 128-byte fp32 rows (SWIZZLE_128B), 64-byte bf16 rows (SWIZZLE_64B)");
  // PAIR: this CTA holds its own 128 A rows and 128 of the chunk's 256 B rows (the other half sits in the peer CTA)
  static constexpr int A_BYTES = T3_BM * BK * 4, B_BYTES = (PAIR ? T3G_BN / 2 : T3G_BN) * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;                // 96 KB;  pair: 64 KB  (hi + companion, A and B)
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;                    // 2;  pair: 3
  static constexpr int OFF_B1 = STAGES * STAGE_BYTES;
  static constexpr int OFF_O8 = OFF_B1 + 256 * 4;
  static constexpr int OFF_BARS = OFF_O8 + 3 * T3_BM * 8 * 4;   // fold partials of column groups 1 .. NCG-1 (NCG <= 4)
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
// residual kernel: N = 128 per item
constexpr int T3R_BN = 128;
template <bool PAIR = false>
struct T3RG {
  static constexpr int B_BYTES = (PAIR ? T3R_BN / 2 : T3R_BN) * T3_BK * 4;     // 16 KB;  pair: 8 KB (64 of the 128 B rows)
  static constexpr int STAGE_BYTES = 2 * T3_A_BYTES + 2 * B_BYTES;             // 64 KB;  pair: 48 KB
  static constexpr int STAGES = PAIR ? 4 : 3;
  static constexpr int OFF_B2 = STAGES * STAGE_BYTES;
  static constexpr int OFF_BARS = OFF_B2 + T3R_BN * 4;
  static constexpr int NBARS = 2 * STAGES + 4;
  static constexpr int SMEM = OFF_BARS + NBARS * 8 + 16;
  static_assert(SMEM <= 232448, "shared memory budget");
};

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
  __nv_bfloat16* acts_b;   // [rows, 2C]: bf16(acts_hi) | bf16(acts - acts_hi), the operands of the cross-term MMAs
  float* acc8;         // gate: partial fold accumulators [n_chunks][rows][8]
  size_t acc8_stride;  // floats between two partials
  const float* h_hi;   // residual epilogue: the residual stream of this layer (read), [rows, C]
  const float* h_lo;
  float* ho_hi;        // residual epilogue: the next layer's stream (written; gap rows as zeros)
  float* ho_lo;
  __nv_bfloat16* ho_b;     // [rows, 2C]: bf16(hi) | bf16(lo) of the next layer's stream
  // WG_LAYER_TIMING=1 (debug): cycle counters, summed over CTAs. Gate kernel slots 16..: [16] MMA warp first wait -> last
  // accumulator complete, [17] MMA waiting for TMA data, [18] MMA waiting for the epilogue, [19] epilogue waiting for an
  // accumulator, [20] epilogue work, [21] kernel entry -> MMA loop entry, [22] producer waiting for a free stage,
  // [23] MMA-issuing CTAs; residual kernel: the same at 32..
  unsigned long long* timing;
};

__device__ __forceinline__ bool t3_row_valid(const Tf32Params& p, int t) {
  if (t >= p.T) return false;
  if (p.row_b) return p.row_b[t] >= 0;
  return p.Tp == 0 || t % p.Tp < p.Tv;
}

// item -> (tile, chunk); tile -> (phase r, first row t0): phase fastest, as in the BF16 kernel's tile order.
// PAIR: an item belongs to a CTA pair and covers two adjacent 128-row tiles of one phase, this CTA's being the rank-th;
// with an odd tile count per phase block the last pair has a GHOST tile (t0 >= T: zeros in, nothing out).
template <bool PAIR>
__device__ __forceinline__ void t3_item_coords(const Tf32Params& p, int item, uint32_t rank, int& q, int& r, int& t0) {
  q = item % p.n_chunks;
  const int tile = item / p.n_chunks;
  r = tile % p.R;
  t0 = PAIR ? (2 * (tile / p.R) + static_cast<int>(rank)) * T3_BM : (tile / p.R) * T3_BM;
}
template <bool PAIR>
__device__ __forceinline__ int t3_n_items(const Tf32Params& p) {
  return (PAIR ? ((p.tiles_per_row + 1) / 2) * p.R : p.n_tiles) * p.n_chunks;
}
// TMA loads whose bytes complete on `bar`: the CTA's own barrier, or (PAIR) the LEADER CTA's barrier at the same offset
template <bool PAIR>
__device__ __forceinline__ void t3_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  if (PAIR) tma2_load_2d(dst, m, bar & kPeerBitMask, c0, c1);
  else tma_load_2d(dst, m, bar, c0, c1);
}
template <bool PAIR>
__device__ __forceinline__ void t3_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  if (PAIR) tma2_load_4d(dst, m, bar & kPeerBitMask, c0, c1, c2, c3);
  else tma_load_4d(dst, m, bar, c0, c1, c2, c3);
}
template <bool PAIR>
__device__ __forceinline__ void t3_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma2_tf32(d, a, b, idesc, acc);
  else umma_tf32(d, a, b, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void t3_mma16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma2_bf16(d, a, b, idesc, acc);
  else umma_bf16(d, a, b, idesc, acc);
}
// The bf16 companion of an operand: [.., 2K], per K-block of 32 one 128-byte row segment of 64 bf16 --
//     A operands (h, acts, mel window):  [ bf16(hi) x 32 | bf16(lo) x 32 ]
//     B operands (W1, V, W2):            [ bf16(lo) x 32 | bf16(hi) x 32 ]
// so that ONE K = 64 sweep of BF16 MMAs over a companion tile pair gives both cross terms, a_hi*b_lo + a_lo*b_hi, and a
// K-block is one TMA box of 128-byte rows (two boxes of 64-byte rows cost the SM's TMA port twice the row requests).
template <bool PAIR>
__device__ __forceinline__ void t3_load_b_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  t3_load_2d<PAIR>(dst, m, bar, 2 * c0, c1);
}
template <bool PAIR>
__device__ __forceinline__ void t3_load_b_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  t3_load_4d<PAIR>(dst, m, bar, 2 * c0, c1, c2, c3);
}
// One pipeline stage = one K-block of 32: [A_hi fp32 | A_b bf16][B_hi | B_b], all 128-byte rows (SWIZZLE_128B). Small cross
// terms first (4 BF16 MMAs, K = 16 each, over the 64 companion columns), then the main product (4 TF32 MMAs, K = 8 each),
// all into the same fp32 accumulator. Call from ONE elected thread.
template <bool PAIR>
__device__ __forceinline__ void t3_mma_stage(uint32_t d_tmem, uint32_t a_base, uint32_t a_bytes, uint32_t b_base, uint32_t b_bytes,
                                             uint32_t idesc32, uint32_t idesc16, bool first) {
  const uint64_t ahi = umma_desc_sw128(a_base), ab = umma_desc_sw128(a_base + a_bytes);
  const uint64_t bhi = umma_desc_sw128(b_base), bb = umma_desc_sw128(b_base + b_bytes);
#pragma unroll
  for (int k = 0; k < 4; ++k) t3_mma16<PAIR>(d_tmem, ab + 2 * k, bb + 2 * k, idesc16, (first && k == 0) ? 0u : 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) t3_mma<PAIR>(d_tmem, ahi + 2 * k, bhi + 2 * k, idesc32, 1u);
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Operand split of 16 consecutive values: hi = tf32(x) (fp32 words), and the BF16 companions of hi and of lo = x - hi
struct T3Split16 {
  float hi[16];
  float lo[16];
  uint32_t hb[8], lb[8];
};
__device__ __forceinline__ void t3_split16(const float* x, bool valid, T3Split16& o) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    o.hi[j] = valid ? tf32_rna(x[j]) : 0.f;
    o.lo[j] = valid ? x[j] - o.hi[j] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o.hb[j] = pack_bf16x2(o.hi[2 * j], o.hi[2 * j + 1]);
    o.lb[j] = pack_bf16x2(o.lo[2 * j], o.lo[2 * j + 1]);
  }
}
// staged tiles of one warp (32 rows x 32 values), 128-byte rows, SWIZZLE_128B: an fp32 tile, and the companion tile
// [hb x 32 | lb x 32].  `j4` = which group of 16 values of the thread's 32 (0 / 1).
__device__ __forceinline__ void t3_stage_f32(uint32_t tile, int lane, int j4, const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t o = static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>((j4 * 4 + c) ^ (lane & 7)) << 4);
    st_shared_f4(tile + o, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  }
}
__device__ __forceinline__ void t3_stage_b16(uint32_t tile, int lane, int j4, const uint32_t* w, bool is_lo) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t o = static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(((is_lo ? 4 : 0) + j4 * 2 + c) ^ (lane & 7)) << 4);
    st_shared_u4(tile + o, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  }
}
// companion of 16 channels starting at channel ch (a multiple of 16), row base `rowp` = &b[row * 2C]: direct stores
__device__ __forceinline__ void t3_store_b16(__nv_bfloat16* rowp, int ch, const uint32_t* hb, const uint32_t* lb) {
  uint4* dh = reinterpret_cast<uint4*>(rowp + 64 * (ch >> 5) + (ch & 16));
  uint4* dl = reinterpret_cast<uint4*>(rowp + 64 * (ch >> 5) + 32 + (ch & 16));
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    dh[v] = make_uint4(hb[4 * v], hb[4 * v + 1], hb[4 * v + 2], hb[4 * v + 3]);
    dl[v] = make_uint4(lb[4 * v], lb[4 * v + 1], lb[4 * v + 2], lb[4 * v + 3]);
  }
}
template <bool PAIR>
__device__ __forceinline__ void t3_commit(uint32_t bar) {   // PAIR: arrives at this offset in BOTH CTAs
  if (PAIR) tc2_commit(bar);
  else tc_commit(bar);
}

template <bool LAST, int BK = 32, bool PAIR = false, int EW = 8>
__global__ void __launch_bounds__(t3_threads(EW), 1)
tf32_gate_kernel(const __grid_constant__ CUtensorMap map_hh, const __grid_constant__ CUtensorMap map_hb,
                 const __grid_constant__ CUtensorMap map_ch, const __grid_constant__ CUtensorMap map_cb,
                 const __grid_constant__ CUtensorMap map_w1h, const __grid_constant__ CUtensorMap map_w1b,
                 const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vb,
                 const __grid_constant__ CUtensorMap smap_hi, const __grid_constant__ CUtensorMap smap_b,
                 const Tf32Params p) {
  static_assert(BK == 32, "one 128-byte fp32 row / one 64-byte bf16 row per K-block");
  using G = T3G<BK, PAIR>;
  static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
  constexpr int T3_EPI_THREADS = EW * 32, T3_NCG = EW / 4;
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
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;     // PAIR: full / accempty barriers are used in the leader only, the MMA warp too
  const int item0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, item_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  if ((smem_base & 1023u) != 0u) __trap();
  const bool timing = p.timing != nullptr;
  const long long t_entry = timing ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_hh); prefetch_tmap(&map_hb); prefetch_tmap(&map_ch); prefetch_tmap(&map_cb);
    prefetch_tmap(&map_w1h); prefetch_tmap(&map_w1b); prefetch_tmap(&map_vh); prefetch_tmap(&map_vb);
    for (int s = 0; s < T3G_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), (PAIR ? 2 : 1) * T3_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem2_alloc(smem_u32(tmem_slot), 512);
      tmem2_relinquish();
    } else {
      tmem_alloc(smem_u32(tmem_slot), 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers and TMEM exist before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = t3_n_items<PAIR>(p);
  const int kb1 = p.kb_conv + p.kb_cond;
  const int cblks = p.C / BK;   // K-blocks per tap

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    uint32_t it = 0;
    long long t_wait = 0;
    for (int item = item0; item < n_items; item += item_step) {
      int q, r, t0;
      t3_item_coords<PAIR>(p, item, rank, q, r, t0);
      for (int kb = 0; kb < kb1; ++kb, ++it) {
        const int s = it % T3G_STAGES;
        const long long tq = timing ? clock64() : 0;
        mbar_wait(empty_bar(s), ((it / T3G_STAGES) & 1) ^ 1);
        if (timing) t_wait += clock64() - tq;
        if (elect_one()) {
          if (leader) mbar_expect_tx(full_bar(s), (PAIR ? 2 : 1) * T3G_STAGE_BYTES);   // PAIR: the bytes of both CTAs
          const int bq = q * T3G_BN + (PAIR ? static_cast<int>(rank) * (T3G_BN / 2) : 0);   // first B row this CTA loads
          // stage = [A_hi | A_b][B_hi | B_b]  (t3_mma_stage)
          const uint32_t a_hi = smem_base + s * T3G_STAGE_BYTES, a_b = a_hi + T3G_A_BYTES;
          const uint32_t b_hi = a_b + T3G_A_BYTES, b_b = b_hi + T3G_B_BYTES;
          if (kb < p.kb_conv) {
            // tap shifted by (tap-1)*dilation positions: phase (r+sh) mod R, frames moved by floor((r+sh)/R)
            const int tap = kb / cblks, cblk = kb - tap * cblks;
            const int rs = r + (tap - 1) * p.dilation;
            const int carry = (rs >= 0) ? rs / p.R : -((-rs + p.R - 1) / p.R);
            t3_load_4d<PAIR>(a_hi, &map_hh, full_bar(s), cblk * BK, t0 + carry, rs - carry * p.R, 0);
            t3_load_b_4d<PAIR>(a_b, &map_hb, full_bar(s), cblk * BK, t0 + carry, rs - carry * p.R, 0);
            t3_load_2d<PAIR>(b_hi, &map_w1h, full_bar(s), kb * BK, p.layer * 2 * p.C + bq);
            t3_load_b_2d<PAIR>(b_b, &map_w1b, full_bar(s), kb * BK, p.layer * 2 * p.C + bq);
          } else {
            const int kc = kb - p.kb_conv;
            t3_load_4d<PAIR>(a_hi, &map_ch, full_bar(s), kc * BK, t0, 0, 0);
            t3_load_b_4d<PAIR>(a_b, &map_cb, full_bar(s), kc * BK, t0, 0, 0);
            t3_load_2d<PAIR>(b_hi, &map_vh, full_bar(s), kc * BK, p.wc_row0 + r * p.wc_rstride + bq);
            t3_load_b_2d<PAIR>(b_b, &map_vb, full_bar(s), kc * BK, p.wc_row0 + r * p.wc_rstride + bq);
          }
        }
        __syncwarp();
      }
    }
    if (timing && lane == 0) atomicAdd(p.timing + 22, static_cast<unsigned long long>(t_wait));
  } else if (warp == 1) {
    // =========================== MMA issuer (PAIR: the leader CTA only) ========================
    constexpr uint32_t idesc = umma_idesc_tf32(PAIR ? 2 * T3_BM : T3_BM, T3G_BN);
    constexpr uint32_t idesc16 = umma_idesc_bf16(PAIR ? 2 * T3_BM : T3_BM, T3G_BN);
    uint32_t it = 0, n = 0;
    long long t_full = 0, t_epi = 0;
    const long long t_begin = timing ? clock64() : 0;
    for (int item = item0; leader && item < n_items; item += item_step, ++n) {
      const uint32_t as = n & 1u;
      long long tq = timing ? clock64() : 0;
      mbar_wait(accempty_bar(as), ((n >> 1) & 1u) ^ 1u);
      if (timing) t_epi += clock64() - tq;   // the epilogue has read this accumulator out (2 items ago)
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 256u * as;
      for (int kb = 0; kb < kb1; ++kb, ++it) {
        const int s = it % T3G_STAGES;
        tq = timing ? clock64() : 0;
        mbar_wait(full_bar(s), (it / T3G_STAGES) & 1);
        if (timing) t_full += clock64() - tq;
        tc_fence_after();
        const uint32_t base = smem_base + s * T3G_STAGE_BYTES;
        if (elect_one()) {
          t3_mma_stage<PAIR>(d_tmem, base, T3G_A_BYTES, base + 2 * T3G_A_BYTES, T3G_B_BYTES, idesc, idesc16, kb == 0);
          t3_commit<PAIR>(empty_bar(s));
          if (kb == kb1 - 1) t3_commit<PAIR>(accfull_bar(as));
        }
        __syncwarp();
      }
    }
    if (timing && leader && n > 0) {
      // the last accumulator is complete when its accfull barrier flips (this CTA's copy; waiting does not consume it)
      mbar_wait(accfull_bar((n - 1) & 1u), ((n - 1) >> 1) & 1u);
      if (lane == 0) {
        atomicAdd(p.timing + 16, static_cast<unsigned long long>(clock64() - t_begin));
        atomicAdd(p.timing + 17, static_cast<unsigned long long>(t_full));
        atomicAdd(p.timing + 18, static_cast<unsigned long long>(t_epi));
        atomicAdd(p.timing + 21, static_cast<unsigned long long>(t_begin - t_entry));
        atomicAdd(p.timing + 23, 1ull);
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int we = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int cg = we >> 2;             // column group: CH of the chunk's 128 gate channels
    constexpr int CH = 128 / T3_NCG, NP = CH / 32;   // channels per thread; 32-channel fold partials per thread
    // One item per CTA (the 16-warp case): once the accumulator is complete the operand ring is idle, so each warp
    // stages its 32 x 32 (hi, lo) output tiles there (128-byte swizzled rows) and writes them with two TMA stores
    // instead of 16 row-strided 16-byte stores per thread (32 half-filled sectors per instruction).
    const bool stage_out = !LAST && EW == 16 && n_items <= item_step;
    const uint32_t stg_hi = smem_base + static_cast<uint32_t>(we) * 8192u, stg_b = stg_hi + 4096u;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t n = 0;
    for (int item = item0; item < n_items; item += item_step, ++n) {
      int q, r, t0;
      t3_item_coords<PAIR>(p, item, rank, q, r, t0);
      const uint32_t as = n & 1u;
      const bool valid = t3_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      // this chunk's bias: [tanh 128 | sigmoid 128]
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");   // previous item's s_b1 / s_o8 readers are done
      for (int i = threadIdx.x - 64; i < 256; i += T3_EPI_THREADS) s_b1[i] = p.bias[q * 256 + i];
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      const long long tq = timing ? clock64() : 0;
      mbar_wait(accfull_bar(as), (n >> 1) & 1u);
      const long long tw = timing ? clock64() : 0;
      tc_fence_after();
      // skip/end fold: one partial sum per 32 channels, P0..P3, combined as (P0 + P1) + (P2 + P3) whatever the warp count
      float o8[NP][8];
#pragma unroll
      for (int pi = 0; pi < NP; ++pi)
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[pi][j] = 0.f;
      const uint32_t taddr = tmem_base + lane_addr + 256u * as + cg * CH;
#pragma unroll
      for (int pi = 0; pi < NP; ++pi) {
#pragma unroll 1
        for (int g2 = 0; g2 < 2; ++g2) {       // 16 gate channels per step
          const int g = pi * 2 + g2;
          uint32_t tr[16], gr[16];
          tmem_ld16(taddr + g * 16, tr);
          tmem_ld16(taddr + 128 + g * 16, gr);
          tmem_ld_wait();
          const int ch0 = cg * CH + g * 16;                 // first channel inside the chunk
          const float* wse = p.wse + static_cast<size_t>(q * 128 + ch0) * 8;
          float a[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xt = __uint_as_float(tr[j]) + s_b1[ch0 + j];
            const float xg = __uint_as_float(gr[j]) + s_b1[128 + ch0 + j];
            a[j] = t3_gate_act(xt, xg);                              // waveglow_arch.py:19-24
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wse + j * 8));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wse + j * 8 + 4));
            o8[pi][0] = fmaf(a[j], w0.x, o8[pi][0]); o8[pi][1] = fmaf(a[j], w0.y, o8[pi][1]);
            o8[pi][2] = fmaf(a[j], w0.z, o8[pi][2]); o8[pi][3] = fmaf(a[j], w0.w, o8[pi][3]);
            o8[pi][4] = fmaf(a[j], w1.x, o8[pi][4]); o8[pi][5] = fmaf(a[j], w1.y, o8[pi][5]);
            o8[pi][6] = fmaf(a[j], w1.z, o8[pi][6]); o8[pi][7] = fmaf(a[j], w1.w, o8[pi][7]);
          }
          if (!LAST && (stage_out || (t0 + row) < p.T)) {
            T3Split16 sp;
            t3_split16(a, valid, sp);
            if (stage_out) {
              t3_stage_f32(stg_hi, lane, g2, sp.hi);
              t3_stage_b16(stg_b, lane, g2, sp.hb, false);
              t3_stage_b16(stg_b, lane, g2, sp.lb, true);
            } else {
              float4* dh = reinterpret_cast<float4*>(p.acts_hi + m * p.C + q * 128 + ch0);
#pragma unroll
              for (int v = 0; v < 4; ++v) dh[v] = make_float4(sp.hi[4 * v], sp.hi[4 * v + 1], sp.hi[4 * v + 2], sp.hi[4 * v + 3]);
              t3_store_b16(p.acts_b + m * 2 * p.C, q * 128 + ch0, sp.hb, sp.lb);
            }
          }
        }
      }
      if (stage_out) {       // rows beyond the phase block are clipped by the TMA store
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&smap_hi, stg_hi, q * 128 + cg * CH, t0 + quarter * 32, r, 0);
          tma_store_4d(&smap_b, stg_b, 2 * (q * 128 + cg * CH), t0 + quarter * 32, r, 0);
          bulk_commit();
        }
      }
      tc_fence_before();
      if (timing && threadIdx.x == 64) {
        atomicAdd(p.timing + 19, static_cast<unsigned long long>(tw - tq));
        atomicAdd(p.timing + 20, static_cast<unsigned long long>(clock64() - tw));
      }
      if (PAIR) mbar_arrive_cluster(mapa_u32(accempty_bar(as), 0));   // the leader counts the epilogue threads of both CTAs
      else mbar_arrive(accempty_bar(as));
      float mine[8];      // this thread's partials: P[cg] (16 warps) or P[2cg] + P[2cg+1] (8 warps)
#pragma unroll
      for (int j = 0; j < 8; ++j) mine[j] = NP == 2 ? o8[0][j] + o8[NP - 1][j] : o8[0][j];
      if (cg > 0) {
        float* d = s_o8 + (static_cast<size_t>(cg - 1) * T3_BM + row) * 8;
        *reinterpret_cast<float4*>(d) = make_float4(mine[0], mine[1], mine[2], mine[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(mine[4], mine[5], mine[6], mine[7]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      if (cg == 0 && valid) {
        float tot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s0 = s_o8[row * 8 + j];
          if (T3_NCG == 2) tot[j] = mine[j] + s0;
          else tot[j] = (mine[j] + s0) + (s_o8[(T3_BM + row) * 8 + j] + s_o8[(2 * T3_BM + row) * 8 + j]);
        }
        float4* o = reinterpret_cast<float4*>(p.acc8 + q * p.acc8_stride + m * 8);
        float4 a0 = o[0], a1 = o[1];
        a0.x += tot[0]; a0.y += tot[1]; a0.z += tot[2]; a0.w += tot[3];
        a1.x += tot[4]; a1.y += tot[5]; a1.z += tot[6]; a1.w += tot[7];
        o[0] = a0; o[1] = a1;
      }
    }
    if (stage_out && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem2_dealloc(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <bool PAIR = false, int EW = 8>
__global__ void __launch_bounds__(t3_threads(EW), 1)
tf32_res_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_ab,
                const __grid_constant__ CUtensorMap map_w2h, const __grid_constant__ CUtensorMap map_w2b,
                const __grid_constant__ CUtensorMap smap_hi, const __grid_constant__ CUtensorMap smap_lo,
                const __grid_constant__ CUtensorMap smap_b, const Tf32Params p) {
  using G = T3RG<PAIR>;
  static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
  constexpr int T3_EPI_THREADS = EW * 32, T3_NCG = EW / 4;
  constexpr int T3R_STAGES = G::STAGES, T3R_STAGE_BYTES = G::STAGE_BYTES, T3R_B_BYTES = G::B_BYTES;
  constexpr int T3R_OFF_B2 = G::OFF_B2, T3R_OFF_BARS = G::OFF_BARS, T3R_NBARS = G::NBARS;
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
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int item0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, item_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  if ((smem_base & 1023u) != 0u) __trap();
  const bool timing = p.timing != nullptr;
  const long long t_entry = timing ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_ah); prefetch_tmap(&map_ab); prefetch_tmap(&map_w2h); prefetch_tmap(&map_w2b);
    for (int s = 0; s < T3R_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), (PAIR ? 2 : 1) * T3_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem2_alloc(smem_u32(tmem_slot), 256);
      tmem2_relinquish();
    } else {
      tmem_alloc(smem_u32(tmem_slot), 256);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = t3_n_items<PAIR>(p);
  const int kb2 = p.kb_conv;

  if (warp == 0) {
    uint32_t it = 0;
    long long t_wait = 0;
    for (int item = item0; item < n_items; item += item_step) {
      int q, r, t0;
      t3_item_coords<PAIR>(p, item, rank, q, r, t0);
      for (int kb = 0; kb < kb2; ++kb, ++it) {
        const int s = it % T3R_STAGES;
        const long long tq = timing ? clock64() : 0;
        mbar_wait(empty_bar(s), ((it / T3R_STAGES) & 1) ^ 1);
        if (timing) t_wait += clock64() - tq;
        if (elect_one()) {
          if (leader) mbar_expect_tx(full_bar(s), (PAIR ? 2 : 1) * T3R_STAGE_BYTES);
          const int bq = q * T3R_BN + (PAIR ? static_cast<int>(rank) * (T3R_BN / 2) : 0);
          const uint32_t a_hi = smem_base + s * T3R_STAGE_BYTES, a_b = a_hi + T3_A_BYTES;
          const uint32_t b_hi = a_b + T3_A_BYTES, b_b = b_hi + T3R_B_BYTES;
          t3_load_4d<PAIR>(a_hi, &map_ah, full_bar(s), kb * T3_BK, t0, r, 0);
          t3_load_b_4d<PAIR>(a_b, &map_ab, full_bar(s), kb * T3_BK, t0, r, 0);
          t3_load_2d<PAIR>(b_hi, &map_w2h, full_bar(s), kb * T3_BK, p.layer * p.C + bq);
          t3_load_b_2d<PAIR>(b_b, &map_w2b, full_bar(s), kb * T3_BK, p.layer * p.C + bq);
        }
        __syncwarp();
      }
    }
    if (timing && lane == 0) atomicAdd(p.timing + 38, static_cast<unsigned long long>(t_wait));
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_tf32(PAIR ? 2 * T3_BM : T3_BM, T3R_BN);
    constexpr uint32_t idesc16 = umma_idesc_bf16(PAIR ? 2 * T3_BM : T3_BM, T3R_BN);
    uint32_t it = 0, n = 0;
    long long t_full = 0, t_epi = 0;
    const long long t_begin = timing ? clock64() : 0;
    for (int item = item0; leader && item < n_items; item += item_step, ++n) {
      const uint32_t as = n & 1u;
      long long tq = timing ? clock64() : 0;
      mbar_wait(accempty_bar(as), ((n >> 1) & 1u) ^ 1u);
      if (timing) t_epi += clock64() - tq;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 128u * as;
      for (int kb = 0; kb < kb2; ++kb, ++it) {
        const int s = it % T3R_STAGES;
        tq = timing ? clock64() : 0;
        mbar_wait(full_bar(s), (it / T3R_STAGES) & 1);
        if (timing) t_full += clock64() - tq;
        tc_fence_after();
        const uint32_t base = smem_base + s * T3R_STAGE_BYTES;
        if (elect_one()) {
          t3_mma_stage<PAIR>(d_tmem, base, T3_A_BYTES, base + 2 * T3_A_BYTES, T3R_B_BYTES, idesc, idesc16, kb == 0);
          t3_commit<PAIR>(empty_bar(s));
          if (kb == kb2 - 1) t3_commit<PAIR>(accfull_bar(as));
        }
        __syncwarp();
      }
    }
    if (timing && leader && n > 0) {
      // the last accumulator is complete when its accfull barrier flips (this CTA's copy; waiting does not consume it)
      mbar_wait(accfull_bar((n - 1) & 1u), ((n - 1) >> 1) & 1u);
      if (lane == 0) {
        atomicAdd(p.timing + 32, static_cast<unsigned long long>(clock64() - t_begin));
        atomicAdd(p.timing + 33, static_cast<unsigned long long>(t_full));
        atomicAdd(p.timing + 34, static_cast<unsigned long long>(t_epi));
        atomicAdd(p.timing + 37, static_cast<unsigned long long>(t_begin - t_entry));
        atomicAdd(p.timing + 39, 1ull);
      }
    }
  } else {
    const int we = warp - 2;
    const int quarter = warp & 3;
    const int cg = we >> 2;             // column group: CH of the item's 128 columns
    constexpr int CH = T3R_BN / T3_NCG;
    // With 32 columns per thread the residual stream's old value is fetched BEFORE the accumulator is waited for
    // (it does not depend on this layer's MMAs): the loads' latency hides behind the MMAs instead of following them.
    constexpr bool PREFETCH = CH <= 32;
    const bool stage_out = EW == 16 && n_items <= item_step;   // see tf32_gate_kernel
    // per warp: h_hi and the exact h_lo (fp32: the next residual add reads h = hi + lo) and the BF16 MMA companions
    const uint32_t stg_hi = smem_base + static_cast<uint32_t>(we) * 12288u, stg_lo = stg_hi + 4096u, stg_b = stg_hi + 8192u;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t n = 0;
    for (int item = item0; item < n_items; item += item_step, ++n) {
      int q, r, t0;
      t3_item_coords<PAIR>(p, item, rank, q, r, t0);
      const uint32_t as = n & 1u;
      const bool in_range = (t0 + row) < p.T;
      const bool valid = t3_row_valid(p, t0 + row);
      const size_t m = static_cast<size_t>(r) * p.T + t0 + row;
      const size_t off0 = m * p.C + q * T3R_BN + cg * CH;
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      for (int i = threadIdx.x - 64; i < T3R_BN; i += T3_EPI_THREADS) s_b2[i] = p.bias[q * T3R_BN + i];
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      // h = h_hi + h_lo exactly
      auto load_old = [&](size_t off, float* old) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 oh = *reinterpret_cast<const float4*>(p.h_hi + off + 4 * v);
          const float4 ol = *reinterpret_cast<const float4*>(p.h_lo + off + 4 * v);
          old[4 * v] = oh.x + ol.x; old[4 * v + 1] = oh.y + ol.y; old[4 * v + 2] = oh.z + ol.z; old[4 * v + 3] = oh.w + ol.w;
        }
      };
      float pre[PREFETCH ? CH : 16];
      if (PREFETCH && valid) {
#pragma unroll
        for (int g = 0; g < CH / 16; ++g) load_old(off0 + g * 16, pre + g * 16);
      }
      const long long tq = timing ? clock64() : 0;
      mbar_wait(accfull_bar(as), (n >> 1) & 1u);
      const long long tw = timing ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + 128u * as + cg * CH;
      auto finish = [&](int g, const uint32_t* rr, const float* old) {
        if (!in_range && !stage_out) return;
        const size_t off = off0 + g * 16;
        float x[16];
        // h_new = (acts @ Wres + b) + h   (waveglow_arch.py:131-133); a gap row is the next layer's zero padding
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (__uint_as_float(rr[j]) + s_b2[cg * CH + g * 16 + j]) + old[j];
        T3Split16 sp;
        t3_split16(x, valid, sp);
        if (stage_out) {
          t3_stage_f32(stg_hi, lane, g, sp.hi);
          t3_stage_f32(stg_lo, lane, g, sp.lo);
          t3_stage_b16(stg_b, lane, g, sp.hb, false);
          t3_stage_b16(stg_b, lane, g, sp.lb, true);
          return;
        }
        float4* dh = reinterpret_cast<float4*>(p.ho_hi + off);
        float4* dl = reinterpret_cast<float4*>(p.ho_lo + off);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          dh[v] = make_float4(sp.hi[4 * v], sp.hi[4 * v + 1], sp.hi[4 * v + 2], sp.hi[4 * v + 3]);
          dl[v] = make_float4(sp.lo[4 * v], sp.lo[4 * v + 1], sp.lo[4 * v + 2], sp.lo[4 * v + 3]);
        }
        t3_store_b16(p.ho_b + m * 2 * p.C, static_cast<int>(off - m * p.C), sp.hb, sp.lb);
      };
      if (PREFETCH) {
#pragma unroll
        for (int g = 0; g < CH / 16; ++g) {
          uint32_t rr[16];
          tmem_ld16(taddr + g * 16, rr);
          tmem_ld_wait();
          finish(g, rr, pre + g * 16);
        }
      } else {
#pragma unroll 1
        for (int g = 0; g < CH / 16; ++g) {
          uint32_t rr[16];
          tmem_ld16(taddr + g * 16, rr);
          tmem_ld_wait();
          float old[16];
          if (in_range && valid) load_old(off0 + g * 16, old);
          finish(g, rr, old);
        }
      }
      if (stage_out) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&smap_hi, stg_hi, q * T3R_BN + cg * CH, t0 + quarter * 32, r, 0);
          tma_store_4d(&smap_lo, stg_lo, q * T3R_BN + cg * CH, t0 + quarter * 32, r, 0);
          tma_store_4d(&smap_b, stg_b, 2 * (q * T3R_BN + cg * CH), t0 + quarter * 32, r, 0);
          bulk_commit();
        }
      }
      tc_fence_before();
      if (timing && threadIdx.x == 64) {
        atomicAdd(p.timing + 35, static_cast<unsigned long long>(tw - tq));
        atomicAdd(p.timing + 36, static_cast<unsigned long long>(clock64() - tw));
      }
      if (PAIR) mbar_arrive_cluster(mapa_u32(accempty_bar(as), 0));
      else mbar_arrive(accempty_bar(as));
    }
    if (stage_out && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem2_dealloc(tmem_base, 256);
    else tmem_dealloc(tmem_base, 256);
  }
}

// ---- helper kernels ---------------------------------------------------------------------------------------------
// mel window (A operand of the conditioning) as an fp32 (hi, lo) pair; rows follow one phase block (RowGeom, R = 1)
__global__ void tf32_im2col_kernel(const float* __restrict__ mel, float* __restrict__ a_hi, __nv_bfloat16* __restrict__ a_b,
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
  __nv_bfloat16* bp = a_b + static_cast<size_t>(row) * 2 * Kup + 64 * (kk >> 5) + (kk & 31);   // A operand: [hb | lb] per K-block
  bp[0] = __float2bfloat16_rn(hi);
  bp[32] = __float2bfloat16_rn(v - hi);
}

// folded conditioning weights: in [(r, k), n] fp32 -> hi[(r*N + n)*K + k] (fp32) and the B-operand companion b[(r*N + n)*2K + ..]
__global__ void fold_store_tf32_kernel(const float* __restrict__ in, float* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_b,
                                       int R, int K, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(R) * K * N) return;
  const int k = static_cast<int>(idx % K);
  const size_t rn = idx / K;
  const int n = static_cast<int>(rn % N), r = static_cast<int>(rn / N);
  const float v = in[(static_cast<size_t>(r) * K + k) * N + n];
  const float hi = tf32_rna(v);
  o_hi[idx] = hi;
  __nv_bfloat16* bp = o_b + rn * 2 * K + 64 * (k >> 5) + (k & 31);   // B operand: [lb | hb] per K-block
  bp[0] = __float2bfloat16_rn(v - hi);
  bp[32] = __float2bfloat16_rn(hi);
}

// ---- host side --------------------------------------------------------------------------------------------------
inline void make_map_f32(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;   // boxes of 32 floats = 128-byte rows
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
// bf16 companion arrays [.., rows, 2K]: boxes of 64 elements = one K-block (128-byte rows, SWIZZLE_128B)
inline void make_map_b16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(WG_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 companion) failed with CUresult %d", (int)r);
}
inline void make_map_b16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t K, uint32_t box_rows) {
  const uint64_t dims[2] = {2 * K, rows};
  const uint32_t box[2] = {64, box_rows};
  make_map_b16(m, ptr, 2, dims, box);
}
inline void make_map_b16_4d(CUtensorMap* m, const void* ptr, uint64_t phases, uint64_t rows, uint64_t K, uint32_t box_rows = T3_BM) {
  const uint64_t dims[4] = {2 * K, rows, phases, 1};
  const uint32_t box[4] = {64, box_rows, 1, 1};
  make_map_b16(m, ptr, 4, dims, box);
}
inline void make_map_f32_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const int bk = T3_BK;
  const uint64_t dims[2] = {cols, rows};
  const uint32_t box[2] = {(uint32_t)bk, box_rows};
  make_map_f32(m, ptr, 2, dims, box);
}
inline void make_map_f32_4d(CUtensorMap* m, const void* ptr, uint64_t phases, uint64_t rows, uint64_t cols) {
  const int bk = T3_BK;
  const uint64_t dims[4] = {cols, rows, phases, 1};
  const uint32_t box[4] = {(uint32_t)bk, T3_BM, 1, 1};
  make_map_f32(m, ptr, 4, dims, box);
}

// launch with an optional cluster of 2 (the CTA-pair variants)
template <typename... KArgs, typename... Args>
inline void t3_launch(void (*kernel)(KArgs...), int grid, int ew, size_t smem, cudaStream_t st, bool pair, const Args&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(t3_threads(ew));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pair ? 1 : 0;
  WG_CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}

template <typename K>
inline int t3_max_pairs(K kernel, int ew, size_t smem) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.gridDim = dim3(2 * 1024); cfg.blockDim = dim3(t3_threads(ew)); cfg.dynamicSmemBytes = smem;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

// Returns how many CTA pairs of the tf32 kernels can be resident at once (0: pairs unavailable); see tc_pair_init().
inline int tf32_init() {
#define WG_T3_ATTR(K, BYTES) WG_CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES))
#define WG_T3_ATTRS(EWV)                                                                   \
  WG_T3_ATTR((tf32_gate_kernel<false, 32, false, EWV>), (T3G<32, false>::SMEM));           \
  WG_T3_ATTR((tf32_gate_kernel<true, 32, false, EWV>), (T3G<32, false>::SMEM));            \
  WG_T3_ATTR((tf32_gate_kernel<false, 32, true, EWV>), (T3G<32, true>::SMEM));             \
  WG_T3_ATTR((tf32_gate_kernel<true, 32, true, EWV>), (T3G<32, true>::SMEM));              \
  WG_T3_ATTR((tf32_res_kernel<false, EWV>), T3RG<false>::SMEM);                            \
  WG_T3_ATTR((tf32_res_kernel<true, EWV>), T3RG<true>::SMEM)
  WG_T3_ATTRS(8);
  WG_T3_ATTRS(16);
#undef WG_T3_ATTRS
#undef WG_T3_ATTR
  const int a = t3_max_pairs(tf32_gate_kernel<false, 32, true, 16>, 16, T3G<32, true>::SMEM);
  const int b = t3_max_pairs(tf32_res_kernel<true, 16>, 16, T3RG<true>::SMEM);
  return a < b ? a : b;
}

struct Tf32Plan {
  // "*_hi": fp32 words holding tf32(x) (TF32 MMAs); "*_b": the bf16 companion [.., 2K] = bf16(hi) | bf16(lo) (cross terms)
  CUtensorMap m_h_hi[2], m_h_b[2], m_c_hi, m_c_b, m_w1h, m_w1b, m_vh, m_vb, m_a_hi, m_a_b, m_w2h, m_w2b;
  CUtensorMap p_w1h, p_w1b, p_vh, p_vb, p_w2h, p_w2b;   // CTA-pair variants: boxes of HALF a chunk's B rows
  CUtensorMap s_a_hi, s_a_b, s_h_hi[2], s_h_lo[2], s_h_b[2];   // 32 x 32 boxes for the staged TMA stores of the epilogues
  int max_pairs = 0;     // resident CTA pairs (tf32_init); 0 = single-CTA kernels only
  bool pair = false;     // this plan runs the CTA-pair kernels
  int epi_warps = 0;     // 0 = by shape (16 when every CTA runs one item, else 8); 8 / 16 force (WG_TF32_EPI, A/B)
  Tf32Params base{};
  RowGeom geo1{};
  int sm_count = 0, n_mel = 0, Kup = 0;
  float *h_hi[2] = {nullptr, nullptr}, *h_lo[2] = {nullptr, nullptr}, *aup_hi = nullptr;
  __nv_bfloat16 *h_b[2] = {nullptr, nullptr}, *aup_b = nullptr;
};

struct Tf32Weights {   // device pointers, stacked over all layers
  const float* W1h;  const __nv_bfloat16* W1b;   // [n_layers_total * 2C, 3C] (+ [.., 6C])  chunk-packed rows, K-major
  const float* Vh;   const __nv_bfloat16* Vb;    // [n_layers_total * R * 2C, Kup]
  const float* W2h;  const __nv_bfloat16* W2b;   // [n_layers_total * C, C]
};

struct Tf32Buffers {   // per-call scratch (engine workspace)
  float *h_hi[2], *h_lo[2];        // residual stream: tf32(h) and the exact remainder (the residual add reads h = hi + lo)
  __nv_bfloat16* h_b[2];           // its bf16 companion
  float* aup_hi; __nv_bfloat16* aup_b;       // mel window (conditioning A operand)
  float* acts_hi; __nv_bfloat16* acts_b;     // gate output / residual A operand
  float* acc8; size_t acc8_stride;           // skip/end fold partials
};

// rows1 = rows of one phase block (B*(T+gap), or the ragged total); geo1 = that block's utterance geometry (R = 1)
inline void tf32_prepare(Tf32Plan& pl, int sm_count, int C, int R, int Kup, int n_mel, int n_layers_total, int rows1,
                         const RowGeom& geo1, int Tp, int Tv, const Tf32Weights& w, const Tf32Buffers& b,
                         int max_pairs = 0, int pair_policy = -1, int epi_warps = 0) {
  if (C % 128 || Kup % T3_BK) fail(WG_ERR_UNSUPPORTED, "tf32x3 path needs n_channels %% 128 == 0 (got %d)", C);
  pl.epi_warps = epi_warps;
  pl.sm_count = sm_count; pl.n_mel = n_mel; pl.Kup = Kup; pl.geo1 = geo1;
  for (int i = 0; i < 2; ++i) { pl.h_hi[i] = b.h_hi[i]; pl.h_lo[i] = b.h_lo[i]; pl.h_b[i] = b.h_b[i]; }
  pl.aup_hi = b.aup_hi; pl.aup_b = b.aup_b;
  Tf32Params& p = pl.base;
  p.T = rows1; p.R = R; p.tiles_per_row = (rows1 + T3_BM - 1) / T3_BM; p.n_tiles = p.tiles_per_row * R;
  p.C = C; p.Tp = Tp; p.Tv = Tv; p.row_b = geo1.row_b;
  p.acts_hi = b.acts_hi; p.acts_b = b.acts_b; p.acc8 = b.acc8; p.acc8_stride = b.acc8_stride;
  const uint64_t LT = (uint64_t)n_layers_total;
  for (int i = 0; i < 2; ++i) {
    make_map_f32_4d(&pl.m_h_hi[i], pl.h_hi[i], R, rows1, C);
    make_map_b16_4d(&pl.m_h_b[i], pl.h_b[i], R, rows1, C);
  }
  make_map_f32_4d(&pl.m_c_hi, b.aup_hi, 1, rows1, Kup);
  make_map_b16_4d(&pl.m_c_b, b.aup_b, 1, rows1, Kup);
  make_map_f32_4d(&pl.m_a_hi, b.acts_hi, R, rows1, C);
  make_map_b16_4d(&pl.m_a_b, b.acts_b, R, rows1, C);
  make_map_f32_2d(&pl.m_w1h, w.W1h, LT * 2 * C, 3 * C, T3G_BN);
  make_map_b16_2d(&pl.m_w1b, w.W1b, LT * 2 * C, 3 * C, T3G_BN);
  make_map_f32_2d(&pl.m_vh, w.Vh, LT * R * 2 * C, Kup, T3G_BN);
  make_map_b16_2d(&pl.m_vb, w.Vb, LT * R * 2 * C, Kup, T3G_BN);
  make_map_f32_2d(&pl.m_w2h, w.W2h, LT * C, C, T3R_BN);
  make_map_b16_2d(&pl.m_w2b, w.W2b, LT * C, C, T3R_BN);
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)rows1, (uint64_t)R, 1};
    const uint32_t box[4] = {32, 32, 1, 1};
    make_map_f32(&pl.s_a_hi, b.acts_hi, 4, dims, box);
    make_map_b16_4d(&pl.s_a_b, b.acts_b, R, rows1, C, 32);
    for (int i = 0; i < 2; ++i) {
      make_map_f32(&pl.s_h_hi[i], pl.h_hi[i], 4, dims, box);
      make_map_f32(&pl.s_h_lo[i], pl.h_lo[i], 4, dims, box);
      make_map_b16_4d(&pl.s_h_b[i], pl.h_b[i], R, rows1, C, 32);
    }
  }
  make_map_f32_2d(&pl.p_w1h, w.W1h, LT * 2 * C, 3 * C, T3G_BN / 2);
  make_map_b16_2d(&pl.p_w1b, w.W1b, LT * 2 * C, 3 * C, T3G_BN / 2);
  make_map_f32_2d(&pl.p_vh, w.Vh, LT * R * 2 * C, Kup, T3G_BN / 2);
  make_map_b16_2d(&pl.p_vb, w.Vb, LT * R * 2 * C, Kup, T3G_BN / 2);
  make_map_f32_2d(&pl.p_w2h, w.W2h, LT * C, C, T3R_BN / 2);
  make_map_b16_2d(&pl.p_w2b, w.W2b, LT * C, C, T3R_BN / 2);
  // CTA pairs halve the B bytes each SM pulls in and reads per MMA and issue faster MMAs (tools/probes/mma_rate.cu);
  // a pair needs two tiles of one phase, so an odd tile count per phase block costs a ghost tile. pair_policy: 1 / 0
  // force, -1 = by wave count: (pair items / resident pairs) waves at ~0.7 of a single-CTA wave (measured).
  pl.max_pairs = max_pairs;
  const int n_chunks = 2 * C / T3G_BN;
  const long items1 = (long)p.n_tiles * n_chunks, items2 = (long)((p.tiles_per_row + 1) / 2) * R * n_chunks;
  const long waves1 = (items1 + sm_count - 1) / sm_count;
  const long waves2 = max_pairs > 0 ? (items2 + max_pairs - 1) / max_pairs : 0;
  pl.pair = max_pairs > 0 && (pair_policy == 1 || (pair_policy < 0 && 0.7 * (double)waves2 < (double)waves1));
}

inline int tf32_upsample(const Tf32Plan& pl, const float* mel, cudaStream_t st) {
  const size_t total = (size_t)pl.geo1.rows_per_phase() * pl.Kup;
  tf32_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mel, pl.aup_hi, pl.aup_b, pl.geo1, pl.n_mel, pl.Kup);
  WG_CK(cudaGetLastError());
  return 1;
}

// One WN layer: gate kernel (+ residual kernel unless it is the flow's last layer). Returns the launch count.
inline int tf32_wn_layer(const Tf32Plan& pl, int layer, int dilation, bool last, int hcur, const float* b1, const float* b2,
                         const float* wse, cudaStream_t st, unsigned long long* timing = nullptr) {
  Tf32Params g = pl.base;
  g.timing = timing;
  g.layer = layer; g.dilation = dilation;
  g.n_chunks = 2 * g.C / T3G_BN; g.kb_conv = 3 * g.C / T3_BK; g.kb_cond = pl.Kup / T3_BK;
  g.wc_row0 = layer * g.R * 2 * g.C; g.wc_rstride = 2 * g.C;
  g.bias = b1; g.wse = wse;
  const bool pair = pl.pair;
  auto grid_for = [&](const Tf32Params& q) {
    if (pair) {
      const int items = ((q.tiles_per_row + 1) / 2) * q.R * q.n_chunks;
      return 2 * (items < pl.max_pairs ? items : pl.max_pairs);
    }
    const int items = q.n_tiles * q.n_chunks;
    return items < pl.sm_count ? items : pl.sm_count;
  };
  // 16 epilogue warps when every CTA runs ONE item (nothing to hide the epilogue behind), else 8
  auto one_item = [&](const Tf32Params& q) {
    return pair ? ((q.tiles_per_row + 1) / 2) * q.R * q.n_chunks <= pl.max_pairs : q.n_tiles * q.n_chunks <= pl.sm_count;
  };
  const int grid_g = grid_for(g);
  const bool wide_g = pl.epi_warps == 16 || (pl.epi_warps == 0 && one_item(g));
#define WG_T3G_LAUNCH2(LASTV, EWV)                                                                                    \
  do {                                                                                                                \
    if (pair)                                                                                                         \
      t3_launch(tf32_gate_kernel<LASTV, 32, true, EWV>, grid_g, EWV, T3G<32, true>::SMEM, st, true, pl.m_h_hi[hcur],  \
                pl.m_h_b[hcur], pl.m_c_hi, pl.m_c_b, pl.p_w1h, pl.p_w1b, pl.p_vh, pl.p_vb, pl.s_a_hi, pl.s_a_b, g);   \
    else                                                                                                              \
      t3_launch(tf32_gate_kernel<LASTV, 32, false, EWV>, grid_g, EWV, T3G<32, false>::SMEM, st, false,                \
                pl.m_h_hi[hcur], pl.m_h_b[hcur], pl.m_c_hi, pl.m_c_b, pl.m_w1h, pl.m_w1b, pl.m_vh, pl.m_vb,           \
                pl.s_a_hi, pl.s_a_b, g);                                                                              \
  } while (0)
#define WG_T3G_LAUNCH(LASTV)              \
  do {                                    \
    if (wide_g) WG_T3G_LAUNCH2(LASTV, 16); \
    else WG_T3G_LAUNCH2(LASTV, 8);         \
  } while (0)
  if (last) WG_T3G_LAUNCH(true);
  else WG_T3G_LAUNCH(false);
#undef WG_T3G_LAUNCH
#undef WG_T3G_LAUNCH2
  WG_CK(cudaGetLastError());
  if (last) return 1;
  Tf32Params r = pl.base;
  r.timing = timing;
  r.layer = layer; r.dilation = dilation;
  r.n_chunks = r.C / T3R_BN; r.kb_conv = r.C / T3_BK; r.kb_cond = 0;
  r.bias = b2;
  r.h_hi = pl.h_hi[hcur]; r.h_lo = pl.h_lo[hcur]; r.ho_hi = pl.h_hi[hcur ^ 1]; r.ho_lo = pl.h_lo[hcur ^ 1];
  r.ho_b = pl.h_b[hcur ^ 1];
  const int grid_r = grid_for(r);
  const bool wide_r = pl.epi_warps == 16 || (pl.epi_warps == 0 && one_item(r));
#define WG_T3R_LAUNCH(EWV)                                                                                                      \
  do {                                                                                                                          \
    if (pair) t3_launch(tf32_res_kernel<true, EWV>, grid_r, EWV, T3RG<true>::SMEM, st, true, pl.m_a_hi, pl.m_a_b, pl.p_w2h, pl.p_w2b, pl.s_h_hi[hcur ^ 1], pl.s_h_lo[hcur ^ 1], pl.s_h_b[hcur ^ 1], r);   \
    else t3_launch(tf32_res_kernel<false, EWV>, grid_r, EWV, T3RG<false>::SMEM, st, false, pl.m_a_hi, pl.m_a_b, pl.m_w2h, pl.m_w2b, pl.s_h_hi[hcur ^ 1], pl.s_h_lo[hcur ^ 1], pl.s_h_b[hcur ^ 1], r);     \
  } while (0)
  if (wide_r) WG_T3R_LAUNCH(16);
  else WG_T3R_LAUNCH(8);
#undef WG_T3R_LAUNCH
  WG_CK(cudaGetLastError());
  return 2;
}

}  // namespace wg

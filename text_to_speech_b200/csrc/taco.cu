// Tacotron2 decoder loop of libwg_b200.so (C ABI: include/wg_taco_b200.h).
//
// One decoder frame (reference architectures/tacotron2_arch.py:422-486, :640-691) is EIGHT kernels,
// captured `graph_chunk` frames at a time into a CUDA graph:
//   lstm_mma_kernel     attention LSTM: gates = [prenet, context, h_a] . W (K = 1792) + LSTMCell update
//                       (split-bf16 mma.sync tensor-core path; dense16<32,LSTM> is the fp32 FFMA alternative)
//   dense16<8,QUERY>    query projection of the new h_a (1024 -> 128)
//   energy_kernel       location sensitive attention, 16 text positions per CTA: location conv + dense, energies
//   context_kernel      masked softmax, attention state update, context (4 CTAs per batch row)
//   lstm_mma_kernel     decoder LSTM: gates = [h_a, context, h_d] . W (K = 2560)
//   dense16<8,FRAME>    [h_d, context] -> frame + stop probability, finished/length bookkeeping
//   dense16<8,PRENET>x2 prenet (dense-relu-dropout) of the new frame for the next step
// The dense kernels are weight-streaming: 18 M fp32 parameters (72 MB) are read once per frame through a
// cp.async ring, each CTA owning a slice of output columns for all 16 batch rows of a group, so the state
// vectors are kept TRANSPOSED ([feature][16 rows]) and every weight element is used 16 times from a
// register (packed fp32x2 FMAs). State ping-pongs between two buffers by frame parity; nothing goes back
// to the host inside the loop (except the finished flags once per chunk when early stopping is requested).
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/wg_taco_b200.h"
#include "common.cuh"

namespace {

using wg::fail;
#define CK WG_CK

constexpr int kRows = 16;        // batch rows per group (the register tile of the LSTM kernels)
constexpr int kP = 256;          // prenet width
constexpr int kAD = 128;         // attention_dim
constexpr int kNF = 32;          // attention_filters
constexpr int kUnitsPerCta = 8;

// Per-decode values the kernels read from device memory, so that a captured graph does not bake the
// caller's pointers.
struct TacoIo {
  const float* memory;   // [B, S, E]
  const float* pm;       // [B, S, 128] processed memory
  const int* text_len;   // [B]
  float* outputs;        // [B, max_len, n_mel]
  float* stops;          // [B, max_len]
  float* attn;           // [B, max_len, S] or null
  int* lengths;          // [B]
  int* finished;         // [B]
  unsigned long long seed;
  int max_len;
  int t_base;
  int deterministic;
  int B;
};

struct Dims {
  int NM, E, A, D, KS, S;
  int KA, KD, KO;      // P + E + A, A + E + D, D + E
  float drop_rate;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ unsigned hash4(unsigned long long seed, unsigned a, unsigned b, unsigned c) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (1ull + a) + 0xBF58476D1CE4E5B9ull * b + 0x94D049BB133111EBull * c;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<unsigned>(z >> 32);
}

// ---------------------------------------------------------------------------------------------------
// dense16_kernel: out[16 rows, COLS columns of this CTA] = X^T . W + bias for one group of 16 batch rows,
// then an epilogue. X is [K][16] (transposed state), the CTA's weight slice is packed contiguously as
// [K][COLS]; both are streamed through a 4-stage cp.async ring in chunks of 128 k-rows (24 KB per stage
// for COLS = 32: ~72 KB in flight per SM, enough to cover the L2 latency). A lane owns a 4-column x 4-row
// register tile (packed fp32x2 FMAs); the 16 warps take different k-rows and are summed through shared memory.
constexpr int kChunk = 128;
constexpr int kStages = 4;
constexpr int kDenseThreads = 512;    // 1024 threads (4 k-rows per warp) measured 12 % slower
constexpr int kWarps = kDenseThreads / 32;
constexpr int kRpw = kChunk / kWarps;   // k-rows of a chunk owned by one warp (8)

enum { EPI_LSTM = 0, EPI_QUERY = 1, EPI_FRAME = 2, EPI_PRENET = 3 };

struct DenseArgs {
  const TacoIo* io;
  const float* Wp;    // [n_cta][K][COLS]
  const float* bp;    // [n_cta][COLS]
  const float* X;     // [K][16], group 0
  size_t x_gs;
  int K, step, n_cols;   // n_cols: valid output columns over all CTAs
  float* c;           // LSTM: cell state [U][16]
  size_t c_gs;
  float* o1;          // LSTM: new h (rows of the next input vector); others: transposed output [n_cols][16]
  size_t o1_gs;
  float* o2;          // LSTM: optional second copy of h
  size_t o2_gs;
  float* o3;          // LSTM: optional third copy of h
  size_t o3_gs;
  int layer;          // PRENET: dropout stream
  float drop_rate;
  int pdl;            // execute the griddepcontrol instructions (launched with the PDL attribute)
};

// Programmatic dependent launch: every kernel of the frame chain lets its successor start launching once its own
// main loop is done (pdl_trigger, just before the reduction / epilogue), and only waits for its predecessor's
// results (pdl_wait) after it has queued the loads that do not depend on them (its weights), so launch latency
// and the first weight fetch overlap the predecessor's tail. (Triggering at kernel start was measured slower
// than no PDL at all.)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int COLS, typename WT = float>
constexpr size_t dense_smem() {
  return kStages * (kChunk * COLS * sizeof(WT) + kChunk * kRows * sizeof(float)) +
         sizeof(float) * (kWarps * (32 / COLS) * kRows * COLS + COLS * kRows);
}

// WT = float, or __nv_bfloat16 for weights STORED in bf16 (arithmetic stays fp32): half the bytes per frame, and the
// two LSTM matrices (36 MB instead of 72 MB) then stay resident in L2 from one frame to the next.
template <int COLS, int EPI, typename WT = float>
__global__ void __launch_bounds__(kDenseThreads) dense16_kernel(DenseArgs a) {
  constexpr bool kBf16 = sizeof(WT) == 2;
  constexpr int WF = kChunk * COLS, XF = kChunk * kRows;   // elements per stage
  constexpr int WV = 16 / sizeof(WT);    // weight elements per 16-byte cp.async
  constexpr int LPR = COLS / 4;          // lanes across the columns of one k-row
  constexpr int RPW = 32 / COLS;         // k-rows one warp instruction covers
  constexpr int NPART = kWarps * RPW;    // partial sums per output element
  extern __shared__ __align__(16) float sm[];
  WT* ws = reinterpret_cast<WT*>(sm);
  float* xs = reinterpret_cast<float*>(ws + kStages * WF);
  float* red = xs + kStages * XF;
  float* zs = red + NPART * kRows * COLS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, g = blockIdx.y;
  const WT* Wg = reinterpret_cast<const WT*>(a.Wp) + static_cast<size_t>(cta) * a.K * COLS;
  const float* Xg = a.X + g * a.x_gs;
  const int n_chunks = (a.K + kChunk - 1) / kChunk;

  // Every warp loads exactly the 8 k-rows of a chunk it will consume (1 KB of W + 512 B of X for COLS = 32,
  // both contiguous), so the ring needs no block-wide barrier: cp.async.wait_group + __syncwarp is enough.
  auto issue_w = [&](int c) {
    if (c < n_chunks) {
      const int st = c % kStages, k0 = c * kChunk + warp * kRpw;
      for (int f = lane; f < kRpw * COLS / WV; f += 32) {
        const bool ok = k0 + f / (COLS / WV) < a.K;
        cp_async16(ws + st * WF + warp * kRpw * COLS + f * WV, ok ? Wg + static_cast<size_t>(k0) * COLS + f * WV : Wg, ok ? 16 : 0);
      }
    }
  };
  auto issue_x = [&](int c) {
    if (c < n_chunks) {
      const int st = c % kStages, k0 = c * kChunk + warp * kRpw;
      if (lane < kRpw * 4) {
        const bool ok = k0 + lane / 4 < a.K;
        cp_async16(xs + st * XF + warp * kRpw * kRows + lane * 4, ok ? Xg + static_cast<size_t>(k0) * kRows + lane * 4 : Xg, ok ? 16 : 0);
      }
    }
    cp_async_commit();   // always: keeps the group count in step with the chunk index
  };

  // weights first: they do not depend on the previous kernel
#pragma unroll
  for (int c = 0; c < kStages - 1; ++c) issue_w(c);
  if (a.pdl) pdl_wait();
  const TacoIo& io = *a.io;
  const int t = io.t_base + a.step;
  if (t >= io.max_len) {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    return;
  }
#pragma unroll
  for (int c = 0; c < kStages - 1; ++c) issue_x(c);   // group c = {X of chunk c} (+ all prologue weights in group 0)

  const int c4 = lane % LPR, r4 = (lane / LPR) & 3, ks = lane / COLS;
  float2 acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<kStages - 2>();
    __syncwarp();
    const WT* wst = ws + (c % kStages) * WF;
    const float4* x4 = reinterpret_cast<const float4*>(xs + (c % kStages) * XF);
#pragma unroll
    for (int kk = 0; kk < kRpw / RPW; ++kk) {
      const int row = warp * kRpw + ks * (kRpw / RPW) + kk;
      float4 wv;
      if (kBf16) {
        const uint2 u = reinterpret_cast<const uint2*>(wst)[row * LPR + c4];      // 4 bf16 -> fp32 is a shift
        wv = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                         __uint_as_float(u.y & 0xffff0000u));
      } else {
        wv = reinterpret_cast<const float4*>(wst)[row * LPR + c4];
      }
      const float4 xv = x4[row * 4 + r4];
      const float2 xa = make_float2(xv.x, xv.y), xb = make_float2(xv.z, xv.w);
      acc[0][0] = __ffma2_rn(make_float2(wv.x, wv.x), xa, acc[0][0]); acc[0][1] = __ffma2_rn(make_float2(wv.x, wv.x), xb, acc[0][1]);
      acc[1][0] = __ffma2_rn(make_float2(wv.y, wv.y), xa, acc[1][0]); acc[1][1] = __ffma2_rn(make_float2(wv.y, wv.y), xb, acc[1][1]);
      acc[2][0] = __ffma2_rn(make_float2(wv.z, wv.z), xa, acc[2][0]); acc[2][1] = __ffma2_rn(make_float2(wv.z, wv.z), xb, acc[2][1]);
      acc[3][0] = __ffma2_rn(make_float2(wv.w, wv.w), xa, acc[3][0]); acc[3][1] = __ffma2_rn(make_float2(wv.w, wv.w), xb, acc[3][1]);
    }
    __syncwarp();               // the warp is done reading stage (c-1) % kStages: refill it
    issue_w(c + kStages - 1);
    issue_x(c + kStages - 1);
  }
  if (a.pdl) pdl_trigger();     // late trigger: the successor's launch and weight prefetch overlap only our reduction / epilogue
  {
    const int part = warp * RPW + ks;
    float* r = red + (part * kRows + 4 * r4) * COLS + 4 * c4;
    *reinterpret_cast<float4*>(r) = make_float4(acc[0][0].x, acc[1][0].x, acc[2][0].x, acc[3][0].x);
    *reinterpret_cast<float4*>(r + COLS) = make_float4(acc[0][0].y, acc[1][0].y, acc[2][0].y, acc[3][0].y);
    *reinterpret_cast<float4*>(r + 2 * COLS) = make_float4(acc[0][1].x, acc[1][1].x, acc[2][1].x, acc[3][1].x);
    *reinterpret_cast<float4*>(r + 3 * COLS) = make_float4(acc[0][1].y, acc[1][1].y, acc[2][1].y, acc[3][1].y);
  }
  __syncthreads();
  if (tid < COLS * kRows) {
    const int col = tid % COLS, row = tid / COLS;
    float z = a.bp[cta * COLS + col];
#pragma unroll 8
    for (int q = 0; q < NPART; ++q) z += red[(q * kRows + row) * COLS + col];
    const int gcol = cta * COLS + col, br = g * kRows + row;
    if (EPI == EPI_LSTM) {
      zs[col * kRows + row] = z;
    } else if (EPI == EPI_QUERY) {
      if (gcol < a.n_cols) a.o1[g * a.o1_gs + gcol * kRows + row] = z;
    } else if (EPI == EPI_PRENET) {
      if (gcol < a.n_cols) {
        z = fmaxf(z, 0.0f);
        if (!io.deterministic) {
          const unsigned thresh = static_cast<unsigned>(fminf(a.drop_rate, 0.999999f) * 4294967296.0f);
          z = hash4(io.seed, t, br * 2 + a.layer, gcol) >= thresh ? z * (1.0f / (1.0f - a.drop_rate)) : 0.0f;
        }
        a.o1[g * a.o1_gs + gcol * kRows + row] = z;
      }
    } else if (EPI == EPI_FRAME) {
      if (gcol < a.n_cols - 1) {
        a.o1[g * a.o1_gs + gcol * kRows + row] = z;    // the frame, transposed, feeds the prenet
        if (br < io.B) io.outputs[(static_cast<size_t>(br) * io.max_len + t) * (a.n_cols - 1) + gcol] = z;
      } else if (gcol == a.n_cols - 1 && br < io.B) {
        const float stop = sigmoidf_(z);
        io.stops[static_cast<size_t>(br) * io.max_len + t] = stop;
        const int fin = io.finished[br] | (stop > 0.5f ? 1 : 0);     // tacotron2_arch.py:671-672
        io.lengths[br] += fin ? 0 : 1;
        io.finished[br] = fin;
      }
    }
  }
  if (EPI == EPI_LSTM) {
    __syncthreads();
    if (tid < kUnitsPerCta * kRows) {
      const int du = tid >> 4, b = tid & 15;
      const float zi = zs[du * kRows + b], zf = zs[(8 + du) * kRows + b], zc = zs[(16 + du) * kRows + b], zo = zs[(24 + du) * kRows + b];
      const int u = cta * kUnitsPerCta + du;
      float* cp = a.c + g * a.c_gs + u * kRows + b;
      const float c_new = sigmoidf_(zf) * *cp + sigmoidf_(zi) * tanhf(zc);
      const float h = sigmoidf_(zo) * tanhf(c_new);
      *cp = c_new;
      a.o1[g * a.o1_gs + u * kRows + b] = h;
      if (a.o2) a.o2[g * a.o2_gs + u * kRows + b] = h;
      if (a.o3) a.o3[g * a.o3_gs + u * kRows + b] = h;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// lstm_mma_kernel: the LSTM gate GEMM on the (legacy-interface) tensor cores. The batch group is exactly one
// mma.sync M = 16 tile, so tcgen05 (M >= 64) would idle three quarters of the array; m16n8k16 BF16 fits.
// fp32 accuracy is kept by SPLITTING both operands into bf16 hi + lo (x = hi + lo to 2^-17) and issuing three
// products per tile (hi.hi + hi.lo + lo.hi, fp32 accumulate; the dropped lo.lo term is 2^-16 relative): the
// weights are pre-split and pre-swizzled into mma B-fragment order at create time (same bytes as fp32), the
// state is split on the fly. Each of the 16 warps owns 16 k-rows of a 256-row chunk and streams exactly those
// through a 3-stage cp.async ring (no block barrier in the loop); partial tiles meet in shared memory.
constexpr int kMmaChunk = 256;
constexpr int kMmaStages = 3;
constexpr int kXPitch = 20;     // floats per k-row of the staged state: 16 rows + 4 pad -> conflict-free A-fragment loads
constexpr size_t kMmaSmem = kMmaStages * (16 * 2048 + kMmaChunk * kXPitch * sizeof(float)) + sizeof(float) * (16 * kRows * 32 + 32 * kRows);

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_bf16(float x0, float x1, unsigned& hi, unsigned& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);      // .x (low half) = x0: the smaller k index
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __bfloat162float(h.x), x1 - __bfloat162float(h.y));
  hi = *reinterpret_cast<const unsigned*>(&h);
  lo = *reinterpret_cast<const unsigned*>(&l);
}

__global__ void __launch_bounds__(kDenseThreads) lstm_mma_kernel(DenseArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  unsigned char* ws = smraw;                                                      // stages x 16 warps x 2 KB
  float* xs = reinterpret_cast<float*>(ws + kMmaStages * 16 * 2048);              // stages x 256 x 20
  float* red = xs + kMmaStages * kMmaChunk * kXPitch;                             // 16 x 16 x 32
  float* zs = red + 16 * kRows * 32;                                              // 32 x 16
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int cta = blockIdx.x, grp = blockIdx.y;
  const int n_chunks = (a.K + kMmaChunk - 1) / kMmaChunk, n_kb = n_chunks * 16;
  const unsigned char* Wg = reinterpret_cast<const unsigned char*>(a.Wp) + static_cast<size_t>(cta) * n_kb * 2048;
  const float* Xg = a.X + grp * a.x_gs;

  auto issue_w = [&](int c) {
    if (c < n_chunks) {
      const unsigned char* src = Wg + static_cast<size_t>(c * 16 + warp) * 2048;
      unsigned char* dst = ws + ((c % kMmaStages) * 16 + warp) * 2048;
#pragma unroll
      for (int u = 0; u < 4; ++u) cp_async16(dst + (lane + 32 * u) * 16, src + (lane + 32 * u) * 16, 16);
    }
  };
  auto issue_x = [&](int c) {
    if (c < n_chunks) {
      const int k0 = c * kMmaChunk + warp * 16;
      float* dst = xs + ((c % kMmaStages) * kMmaChunk + warp * 16) * kXPitch;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int unit = lane + 32 * u, row = unit >> 2, part = unit & 3;
        const bool ok = k0 + row < a.K;
        cp_async16(dst + row * kXPitch + part * 4, ok ? Xg + static_cast<size_t>(k0 + row) * kRows + part * 4 : Xg, ok ? 16 : 0);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int c = 0; c < kMmaStages - 1; ++c) issue_w(c);
  if (a.pdl) pdl_wait();
  const TacoIo& io = *a.io;
  if (io.t_base + a.step >= io.max_len) {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    return;
  }
#pragma unroll
  for (int c = 0; c < kMmaStages - 1; ++c) issue_x(c);

  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<kMmaStages - 2>();
    __syncwarp();
    const float* xw = xs + ((c % kMmaStages) * kMmaChunk + warp * 16) * kXPitch;
    const uint2* wf = reinterpret_cast<const uint2*>(ws + ((c % kMmaStages) * 16 + warp) * 2048);
    unsigned ahi[4], alo[4];
    // A fragment (16 batch rows x 16 k): a0 = (row g, k 2tg..), a1 = (row g+8, same k), a2 / a3 = k + 8
    split_bf16(xw[(2 * tg) * kXPitch + g], xw[(2 * tg + 1) * kXPitch + g], ahi[0], alo[0]);
    split_bf16(xw[(2 * tg) * kXPitch + g + 8], xw[(2 * tg + 1) * kXPitch + g + 8], ahi[1], alo[1]);
    split_bf16(xw[(2 * tg + 8) * kXPitch + g], xw[(2 * tg + 9) * kXPitch + g], ahi[2], alo[2]);
    split_bf16(xw[(2 * tg + 8) * kXPitch + g + 8], xw[(2 * tg + 9) * kXPitch + g + 8], ahi[3], alo[3]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint2 bh = wf[nt * 32 + lane], bl = wf[(4 + nt) * 32 + lane];
      mma_bf16_16816(acc[nt], ahi, bh.x, bh.y);
      mma_bf16_16816(acc[nt], ahi, bl.x, bl.y);
      mma_bf16_16816(acc[nt], alo, bh.x, bh.y);
    }
    __syncwarp();
    issue_w(c + kMmaStages - 1);
    issue_x(c + kMmaStages - 1);
  }
  if (a.pdl) pdl_trigger();
  // C fragment: c0,c1 = (row g, cols 2tg, 2tg+1), c2,c3 = (row g+8, same cols) of n-tile nt (= LSTM gate nt)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float* r = red + (warp * kRows + g) * 32 + nt * 8 + 2 * tg;
    *reinterpret_cast<float2*>(r) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(r + 8 * 32) = make_float2(acc[nt][2], acc[nt][3]);
  }
  __syncthreads();
  {
    const int col = tid & 31, row = tid >> 5;
    float z = a.bp[cta * 32 + col];
#pragma unroll
    for (int q = 0; q < 16; ++q) z += red[(q * kRows + row) * 32 + col];
    zs[col * kRows + row] = z;
  }
  __syncthreads();
  if (tid < kUnitsPerCta * kRows) {
    const int du = tid >> 4, b = tid & 15;
    const float zi = zs[du * kRows + b], zf = zs[(8 + du) * kRows + b], zc = zs[(16 + du) * kRows + b], zo = zs[(24 + du) * kRows + b];
    const int u = cta * kUnitsPerCta + du;
    float* cp = a.c + grp * a.c_gs + u * kRows + b;
    const float c_new = sigmoidf_(zf) * *cp + sigmoidf_(zi) * tanhf(zc);
    const float h = sigmoidf_(zo) * tanhf(c_new);
    *cp = c_new;
    a.o1[grp * a.o1_gs + u * kRows + b] = h;
    if (a.o2) a.o2[grp * a.o2_gs + u * kRows + b] = h;
    if (a.o3) a.o3[grp * a.o3_gs + u * kRows + b] = h;
  }
}

// ---------------------------------------------------------------------------------------------------
// Location sensitive attention (location_sensitive_attention.py:104-186) in two kernels so that one batch
// row is spread over several SMs: energy_kernel computes the raw energies of 16 text positions per CTA (one
// warp per position: location conv, location dense with float4 weights per lane, tanh, dot with v);
// context_kernel does the masked softmax of a row (redundantly in each of its 4 CTAs, S is small), updates
// the attention state and reduces its quarter of the context vector.
struct AttnArgs {
  const TacoIo* io;
  const float* qT;     // [G][128][16] projected query of this frame
  const float* Wc;     // [32, 2, KS]  (repacked location conv)
  const float* Wld;    // [32, 128]
  const float* v;      // [128]
  float* energy;       // [G*16, S] raw energies of this frame
  float* ctx1;         // XD[p] rows A.. (decoder LSTM input of this frame), group 0
  float* ctx2;         // XA[p^1] rows P.. (attention LSTM input of the next frame), group 0
  float* ctx3;         // XO rows D.. (frame projection input), group 0
  float* aw;           // [G*16, S]
  float* awc;          // [G*16, S]
  size_t xd_gs, xa_gs, xo_gs;
  Dims d;
  int step;
  int pdl;
};

constexpr int kEnergyThreads = 512;
constexpr int kPosPerCta = kEnergyThreads / 32;   // 16 text positions per CTA, one warp each
constexpr int kCtxThreads = 256;
constexpr int kCtxSplit = 4;                      // CTAs per batch row in context_kernel

__device__ __forceinline__ float tanh_exp(float x) {   // 1 - 2/(e^2x + 1): abs error ~1e-7, exact at +-inf
  return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

__global__ void __launch_bounds__(kEnergyThreads) energy_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const Dims& d = a.d;
  const int S = d.S, KS = d.KS, half = KS / 2, pitch = kPosPerCta + KS - 1;
  float* wld = sm;                        // 32 x 128 (16-byte aligned)
  float* q = wld + kNF * kAD;             // 128
  float* vv = q + kAD;                    // 128
  float* wc = vv + kAD;                   // 32 x 2 x KS
  float* ac = wc + kNF * 2 * KS;          // 2 x pitch
  float* fs = ac + 2 * pitch;             // 16 x 33
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int br = blockIdx.y, g = br / kRows, b = br % kRows, s0 = blockIdx.x * kPosPerCta;
  for (int i = tid; i < kNF * kAD / 4; i += kEnergyThreads)
    reinterpret_cast<float4*>(wld)[i] = __ldg(reinterpret_cast<const float4*>(a.Wld) + i);
  for (int i = tid; i < kNF * 2 * KS; i += kEnergyThreads) wc[i] = __ldg(a.Wc + i);
  if (tid < kAD) vv[tid] = __ldg(a.v + tid);
  if (a.pdl) pdl_wait();     // everything above is weights; everything below depends on the previous kernels
  const TacoIo& io = *a.io;
  const int t = io.t_base + a.step;
  if (t >= io.max_len) return;
  const int len = io.text_len[br];
  if (tid < kAD) q[tid] = a.qT[(static_cast<size_t>(g) * kAD + tid) * kRows + b];
  const float* awr = a.aw + static_cast<size_t>(br) * S;
  const float* awcr = a.awc + static_cast<size_t>(br) * S;
  for (int i = tid; i < pitch; i += kEnergyThreads) {
    const int s = s0 + i - half;
    const bool in = s >= 0 && s < S;
    ac[i] = in ? awr[s] : 0.0f;
    ac[pitch + i] = in ? awcr[s] : 0.0f;
  }
  __syncthreads();
  {  // location conv ('same', no bias): one (position, filter) output per thread
    const int sl = tid / kNF, f = tid % kNF;
    const float* w0 = wc + (f * 2) * KS;
    float acc0 = 0.0f, acc1 = 0.0f;
    for (int j = 0; j < KS; ++j) {
      acc0 = fmaf(ac[sl + j], w0[j], acc0);
      acc1 = fmaf(ac[pitch + sl + j], w0[KS + j], acc1);
    }
    fs[sl * (kNF + 1) + f] = acc0 + acc1;
  }
  __syncthreads();
  if (a.pdl) pdl_trigger();
  const int s = s0 + warp;
  if (s < S) {
    float part = 0.0f;
    if (s < len) {
      const float* pm = io.pm + (static_cast<size_t>(br) * S + s) * kAD;
      const float4 pmv = __ldg(reinterpret_cast<const float4*>(pm) + lane);
      float4 loc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 8
      for (int f = 0; f < kNF; ++f) {
        const float fv = fs[warp * (kNF + 1) + f];
        const float4 w = reinterpret_cast<const float4*>(wld + f * kAD)[lane];
        loc.x = fmaf(fv, w.x, loc.x); loc.y = fmaf(fv, w.y, loc.y); loc.z = fmaf(fv, w.z, loc.z); loc.w = fmaf(fv, w.w, loc.w);
      }
      const float4 qv = reinterpret_cast<const float4*>(q)[lane], v4 = reinterpret_cast<const float4*>(vv)[lane];
      part = v4.x * tanh_exp(qv.x + pmv.x + loc.x);
      part = fmaf(v4.y, tanh_exp(qv.y + pmv.y + loc.y), part);
      part = fmaf(v4.z, tanh_exp(qv.z + pmv.z + loc.z), part);
      part = fmaf(v4.w, tanh_exp(qv.w + pmv.w + loc.w), part);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) a.energy[static_cast<size_t>(br) * S + s] = s < len ? part : -INFINITY;
  }
}

__global__ void __launch_bounds__(kCtxThreads) context_kernel(AttnArgs a) {
  if (a.pdl) pdl_wait();
  const TacoIo& io = *a.io;
  const int t = io.t_base + a.step;
  if (t >= io.max_len) return;
  extern __shared__ __align__(16) float sm[];
  const Dims& d = a.d;
  const int S = d.S;
  float* e = sm;              // S
  float* red = e + S;         // 32
  float* part = red + 32;     // kCtxThreads
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int br = blockIdx.y, g = br / kRows, b = br % kRows, cb = blockIdx.x;
  const int len = io.text_len[br];
  const float* er = a.energy + static_cast<size_t>(br) * S;
  float m = -INFINITY;
  for (int s = tid; s < S; s += kCtxThreads) {
    const float x = er[s];
    e[s] = x;
    m = fmaxf(m, x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < kCtxThreads / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.0f;
  for (int s = tid; s < S; s += kCtxThreads) {
    const float p = s < len ? expf(e[s] - m) : 0.0f;
    e[s] = p;
    sum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int w = 0; w < kCtxThreads / 32; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  float* awr = a.aw + static_cast<size_t>(br) * S;
  float* awcr = a.awc + static_cast<size_t>(br) * S;
  float* attn = io.attn ? io.attn + (static_cast<size_t>(br) * io.max_len + t) * S : nullptr;
  for (int s = tid; s < S; s += kCtxThreads) {
    const float p = e[s] * inv;
    e[s] = p;
    if (cb == 0) {            // one CTA of the row owns the attention state
      awr[s] = p;
      awcr[s] += p;
      if (attn) attn[s] = p;
    }
  }
  __syncthreads();
  if (a.pdl) pdl_trigger();
  // this CTA's quarter of the context vector: two threads per column (even / odd positions)
  const int cols = d.E / kCtxSplit, x0 = cb * cols;
  const float* mem = io.memory + static_cast<size_t>(br) * S * d.E;
  for (int xb = 0; xb < cols; xb += kCtxThreads / 2) {
    const int xl = xb + (tid & (kCtxThreads / 2 - 1)), hf = tid / (kCtxThreads / 2);
    float acc0 = 0.0f, acc1 = 0.0f;
    if (xl < cols) {
      const float* mp = mem + x0 + xl;
      int s = hf;
      for (; s + 2 < len; s += 4) {
        const float m0 = __ldg(mp + static_cast<size_t>(s) * d.E), m1 = __ldg(mp + static_cast<size_t>(s + 2) * d.E);
        acc0 = fmaf(e[s], m0, acc0);
        acc1 = fmaf(e[s + 2], m1, acc1);
      }
      for (; s < len; s += 2) acc0 = fmaf(e[s], __ldg(mp + static_cast<size_t>(s) * d.E), acc0);
    }
    part[tid] = acc0 + acc1;
    __syncthreads();
    if (hf == 0 && xl < cols) {
      const float acc = part[tid] + part[tid + kCtxThreads / 2];
      const int x = x0 + xl;
      a.ctx1[g * a.xd_gs + x * kRows + b] = acc;
      a.ctx2[g * a.xa_gs + x * kRows + b] = acc;
      a.ctx3[g * a.xo_gs + x * kRows + b] = acc;
    }
    __syncthreads();
  }
}

__global__ void processed_memory_kernel(const float* memory, const float* Wm, float* pm, int E) {
  // pm[row, d] = sum_e memory[row, e] * Wm[e, d]   (LocationSensitiveAttention.process_memory, :96-102)
  const size_t row = blockIdx.x;
  const int dd = threadIdx.x;
  const float* m = memory + row * E;
  float s = 0.0f;
  for (int e = 0; e < E; ++e) s = fmaf(__ldg(m + e), __ldg(Wm + e * kAD + dd), s);
  pm[row * kAD + dd] = s;
}

__global__ void advance_kernel(TacoIo* io, int frames) { io->t_base += frames; }

}  // namespace

struct wg_taco_engine {
  wg_taco_config cfg{};
  Dims d{};
  int device = 0;
  int graph_chunk = 32;
  bool lstm_bf16 = false;  // LSTM weights stored as bf16 (wg_taco_config.lstm_weight_dtype == 1)
  int lstm_mode = 0;       // wg_taco_config.lstm_weight_dtype
  bool use_pdl = true;     // programmatic dependent launch along the frame chain (WG_TACO_PDL=0: plain stream order).
                           // Triggering at kernel START measured 6 % slower than no PDL; triggering after the main
                           // loop (the successor overlaps only our reduction / epilogue) measures 9 % faster.
  // weights
  float *WpA = nullptr, *bpA = nullptr, *WpD = nullptr, *bpD = nullptr;       // LSTMs, 32 columns per CTA
  float *WpQ = nullptr, *bpQ = nullptr;                                         // query projection, 8 columns per CTA
  float *WpO = nullptr, *bpO = nullptr;                                         // frame projection | stop gate, 8 per CTA
  float *WpP0 = nullptr, *bpP0 = nullptr, *WpP1 = nullptr, *bpP1 = nullptr;   // prenet layers, 8 per CTA
  float *Wm = nullptr, *Wc = nullptr, *Wld = nullptr, *v = nullptr;
  // state (grown on demand)
  int cap_groups = 0, cap_S = 0;
  size_t cap_pm = 0;
  float *XA = nullptr, *XD = nullptr, *XO = nullptr, *ca = nullptr, *cd = nullptr, *aw = nullptr, *awc = nullptr, *pm = nullptr;
  float *qT = nullptr, *frameT = nullptr, *x1T = nullptr, *energy = nullptr;
  int *finished = nullptr, *text_len = nullptr;
  TacoIo* io = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  std::map<std::tuple<int, int, int>, cudaGraphExec_t> graphs;   // (B, S, chunk)
  std::map<std::tuple<int, int, int>, unsigned long long> graph_used;   // last use (LRU: at most TACO_MAX_GRAPHS are kept)
  unsigned long long graph_clock = 0;
  std::string err;
};

namespace {

std::mutex g_taco_err_mu;
std::string g_taco_create_err;

constexpr size_t TACO_MAX_GRAPHS = 16;   // every new (B, S, chunk) captures a graph: a long-running process must not grow for ever

void drop_graphs(wg_taco_engine* e) {
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
  e->graph_used.clear();
}

void destroy_taco(wg_taco_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  drop_graphs(e);
  for (float* p : {e->WpA, e->bpA, e->WpD, e->bpD, e->WpQ, e->bpQ, e->WpO, e->bpO, e->WpP0, e->bpP0, e->WpP1, e->bpP1,
                   e->Wm, e->Wc, e->Wld, e->v, e->XA, e->XD, e->XO, e->ca, e->cd, e->aw, e->awc, e->pm, e->qT, e->frameT,
                   e->x1T, e->energy})
    cudaFree(p);
  cudaFree(e->finished); cudaFree(e->text_len); cudaFree(e->io);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

template <typename F>
int taco_guarded(wg_taco_engine* h, F&& f) {
  try {
    f();
    return WG_OK;
  } catch (const wg::Fail& x) {
    if (h) h->err = x.msg;
    else {
      std::lock_guard<std::mutex> g(g_taco_err_mu);
      g_taco_create_err = x.msg;
    }
    return x.code;
  } catch (const std::exception& x) {
    if (h) h->err = x.what();
    return WG_ERR_INVALID;
  }
}

float* upload(const std::vector<float>& v) {
  float* p = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&p), v.size() * sizeof(float)));
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return p;
}

const wg_tensor& find(const wg_tensor* ts, int n, const char* name, std::initializer_list<int64_t> shape) {
  for (int i = 0; i < n; ++i) {
    if (ts[i].name && std::strcmp(ts[i].name, name) == 0) {
      const wg_tensor& t = ts[i];
      if (!t.data) fail(WG_ERR_WEIGHTS, "wg_taco_create: tensor %s has no data", name);
      bool ok = t.ndim == static_cast<int>(shape.size());
      int j = 0;
      for (int64_t s : shape) ok = ok && t.shape[j++] == s;
      if (!ok) fail(WG_ERR_WEIGHTS, "wg_taco_create: tensor %s has the wrong shape", name);
      size_t numel = 1;
      for (int64_t s : shape) numel *= static_cast<size_t>(s);
      for (size_t q = 0; q < numel; ++q)
        if (!std::isfinite(t.data[q])) fail(WG_ERR_WEIGHTS, "wg_taco_create: tensor %s holds a non-finite value", name);
      return t;
    }
  }
  fail(WG_ERR_WEIGHTS, "wg_taco_create: tensor %s is missing", name);
}

// [K_in, 4U] kernel + [U, 4U] recurrent kernel -> per-CTA packed [U/8][K][32], bias -> [U/8][32]
void pack_lstm(const wg_tensor* ts, int n, const std::string& prefix, int n_in, int U, int mode, float** Wp, float** bp) {
  const wg_tensor& k = find(ts, n, (prefix + "/kernel").c_str(), {n_in, 4 * U});
  const wg_tensor& r = find(ts, n, (prefix + "/recurrent_kernel").c_str(), {U, 4 * U});
  const wg_tensor& b = find(ts, n, (prefix + "/bias").c_str(), {4 * U});
  const int K = n_in + U, n_cta = U / kUnitsPerCta;
  std::vector<float> w(static_cast<size_t>(n_cta) * K * 32), bb(static_cast<size_t>(n_cta) * 32);
  for (int c = 0; c < n_cta; ++c)
    for (int g = 0; g < 4; ++g)
      for (int du = 0; du < kUnitsPerCta; ++du) {
        const int colsrc = g * U + c * kUnitsPerCta + du, j = g * kUnitsPerCta + du;
        bb[c * 32 + j] = b.data[colsrc];
        for (int kk = 0; kk < K; ++kk) {
          const float val = kk < n_in ? k.data[static_cast<size_t>(kk) * 4 * U + colsrc]
                                      : r.data[static_cast<size_t>(kk - n_in) * 4 * U + colsrc];
          w[(static_cast<size_t>(c) * K + kk) * 32 + j] = val;
        }
      }
  if (mode == 2) {
    // split-bf16 mma B fragments: [cta][k16 block][hi | lo][n-tile = gate][lane][4 bf16], K padded to 256 with zeros
    auto to_bf16 = [](float x) {
      uint32_t u;
      std::memcpy(&u, &x, 4);
      return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    };
    auto from_bf16 = [](uint16_t h) {
      uint32_t u = static_cast<uint32_t>(h) << 16;
      float x;
      std::memcpy(&x, &u, 4);
      return x;
    };
    const int n_kb = (K + 255) / 256 * 16;
    std::vector<uint16_t> frag(static_cast<size_t>(n_cta) * n_kb * 1024, 0);
    for (int c = 0; c < n_cta; ++c)
      for (int kb = 0; kb < n_kb; ++kb)
        for (int nt = 0; nt < 4; ++nt)
          for (int lane = 0; lane < 32; ++lane) {
            const int gq = lane >> 2, tq = lane & 3, j = nt * 8 + gq;
            const int ks[4] = {kb * 16 + 2 * tq, kb * 16 + 2 * tq + 1, kb * 16 + 2 * tq + 8, kb * 16 + 2 * tq + 9};
            for (int e = 0; e < 4; ++e) {
              const float val = ks[e] < K ? w[(static_cast<size_t>(c) * K + ks[e]) * 32 + j] : 0.0f;
              const uint16_t hi = to_bf16(val), lo = to_bf16(val - from_bf16(hi));
              const size_t base = (static_cast<size_t>(c) * n_kb + kb) * 1024;
              frag[base + (static_cast<size_t>(nt) * 32 + lane) * 4 + e] = hi;
              frag[base + 512 + (static_cast<size_t>(nt) * 32 + lane) * 4 + e] = lo;
            }
          }
    std::vector<float> as_words(frag.size() / 2);
    std::memcpy(as_words.data(), frag.data(), frag.size() * 2);
    *Wp = upload(as_words);
  } else if (mode == 1) {
    // round to nearest even into bf16, two per 32-bit word (the kernel reads them back with shifts)
    std::vector<float> packed(w.size() / 2);
    for (size_t i = 0; i < w.size(); i += 2) {
      uint32_t lo, hi, out;
      std::memcpy(&lo, &w[i], 4);
      std::memcpy(&hi, &w[i + 1], 4);
      lo = (lo + 0x7fffu + ((lo >> 16) & 1u)) >> 16;
      hi = (hi + 0x7fffu + ((hi >> 16) & 1u)) >> 16;
      out = lo | (hi << 16);
      std::memcpy(&packed[i / 2], &out, 4);
    }
    *Wp = upload(packed);
  } else {
    *Wp = upload(w);
  }
  *bp = upload(bb);
}

// [K, N] row-major (optionally two matrices side by side: N = n_a + n_b) -> [ceil(N/cols)][K][cols], zero padded
void pack_dense(const float* wa, int n_a, const float* wb, int n_b, const float* ba, const float* bb_, int K, int cols,
                float** Wp, float** bp) {
  const int N = n_a + n_b, n_cta = (N + cols - 1) / cols;
  std::vector<float> w(static_cast<size_t>(n_cta) * K * cols, 0.0f), b(static_cast<size_t>(n_cta) * cols, 0.0f);
  for (int n = 0; n < N; ++n) {
    const int c = n / cols, j = n % cols;
    if (n < n_a ? ba != nullptr : bb_ != nullptr) b[c * cols + j] = n < n_a ? ba[n] : bb_[n - n_a];
    for (int k = 0; k < K; ++k)
      w[(static_cast<size_t>(c) * K + k) * cols + j] = n < n_a ? wa[static_cast<size_t>(k) * n_a + n] : wb[static_cast<size_t>(k) * n_b + (n - n_a)];
  }
  *Wp = upload(w);
  *bp = upload(b);
}

void build_taco(wg_taco_engine* e, const wg_taco_config* cfg, const wg_tensor* ts, int n, int device) {
  e->cfg = *cfg;
  e->device = device;
  const wg_taco_config& c = *cfg;
  if (c.prenet_dim != kP || c.attention_dim != kAD || c.attention_filters != kNF)
    fail(WG_ERR_UNSUPPORTED, "wg_taco_create: prenet_dim/attention_dim/attention_filters must be %d/%d/%d", kP, kAD, kNF);
  if (c.n_mel_channels < 1 || c.n_mel_channels > 128) fail(WG_ERR_INVALID, "wg_taco_create: n_mel_channels %d", c.n_mel_channels);
  if (c.embedding_dim < 4 || c.embedding_dim % 4) fail(WG_ERR_INVALID, "wg_taco_create: embedding_dim %d", c.embedding_dim);
  if (c.attention_rnn_dim < 8 || c.attention_rnn_dim % 8 || c.decoder_rnn_dim < 8 || c.decoder_rnn_dim % 8)
    fail(WG_ERR_INVALID, "wg_taco_create: rnn dims must be multiples of 8");
  if (c.attention_kernel_size < 1 || c.attention_kernel_size > 63 || c.attention_kernel_size % 2 == 0)
    fail(WG_ERR_INVALID, "wg_taco_create: attention_kernel_size %d must be odd and <= 63", c.attention_kernel_size);
  if (!(c.prenet_drop_rate >= 0.0f && c.prenet_drop_rate < 1.0f)) fail(WG_ERR_INVALID, "wg_taco_create: prenet_drop_rate");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    fail(WG_ERR_CUDA, "wg_taco_create: no CUDA device (the B200 Tacotron2 decoder has no CPU fallback)");
  if (device < 0 || device >= n_dev) fail(WG_ERR_INVALID, "wg_taco_create: device %d of %d", device, n_dev);
  wg::DeviceGuard dev_guard(device);
  Dims& d = e->d;
  d.NM = c.n_mel_channels; d.E = c.embedding_dim; d.A = c.attention_rnn_dim; d.D = c.decoder_rnn_dim;
  d.KS = c.attention_kernel_size; d.S = 0;
  d.KA = kP + d.E + d.A; d.KD = d.A + d.E + d.D; d.KO = d.D + d.E;
  d.drop_rate = c.prenet_drop_rate;

  if (c.lstm_weight_dtype < 0 || c.lstm_weight_dtype > 2)
    fail(WG_ERR_INVALID, "wg_taco_create: lstm_weight_dtype %d (0 = fp32, 1 = bf16, 2 = split-bf16 mma)", c.lstm_weight_dtype);
  e->lstm_mode = c.lstm_weight_dtype;
  e->lstm_bf16 = c.lstm_weight_dtype == 1;
  pack_lstm(ts, n, "decoder/attention_rnn", kP + d.E, d.A, e->lstm_mode, &e->WpA, &e->bpA);
  pack_lstm(ts, n, "decoder/decoder_rnn/cell_0", d.A + d.E, d.D, e->lstm_mode, &e->WpD, &e->bpD);
  CK(cudaFuncSetAttribute(lstm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMmaSmem)));
  auto plain = [&](const char* name, std::initializer_list<int64_t> shape) {
    const wg_tensor& t = find(ts, n, name, shape);
    size_t numel = 1;
    for (int64_t s : shape) numel *= static_cast<size_t>(s);
    return upload(std::vector<float>(t.data, t.data + numel));
  };
  e->Wm = plain("decoder/lsa/memory_layer/kernel", {d.E, kAD});
  e->v = plain("decoder/lsa/value_layer/kernel", {kAD, 1});
  e->Wld = plain("decoder/lsa/location_dense/kernel", {kNF, kAD});
  {
    const wg_tensor& t = find(ts, n, "decoder/lsa/location_conv/kernel", {d.KS, 2, kNF});   // Keras [k, in, out]
    std::vector<float> w(static_cast<size_t>(kNF) * 2 * d.KS);
    for (int j = 0; j < d.KS; ++j)
      for (int ch = 0; ch < 2; ++ch)
        for (int f = 0; f < kNF; ++f) w[(f * 2 + ch) * d.KS + j] = t.data[(j * 2 + ch) * kNF + f];
    e->Wc = upload(w);
  }
  pack_dense(find(ts, n, "decoder/lsa/query_layer/kernel", {d.A, kAD}).data, kAD, nullptr, 0, nullptr, nullptr, d.A, 8,
             &e->WpQ, &e->bpQ);
  pack_dense(find(ts, n, "decoder/linear_projection/kernel", {d.KO, d.NM}).data, d.NM,
             find(ts, n, "decoder/gate_output/kernel", {d.KO, 1}).data, 1,
             find(ts, n, "decoder/linear_projection/bias", {d.NM}).data, find(ts, n, "decoder/gate_output/bias", {1}).data,
             d.KO, 8, &e->WpO, &e->bpO);
  pack_dense(find(ts, n, "decoder/prenet/layer_0/kernel", {d.NM, kP}).data, kP, nullptr, 0, nullptr, nullptr, d.NM, 8,
             &e->WpP0, &e->bpP0);
  pack_dense(find(ts, n, "decoder/prenet/layer_1/kernel", {kP, kP}).data, kP, nullptr, 0, nullptr, nullptr, kP, 8,
             &e->WpP1, &e->bpP1);
  CK(cudaFuncSetAttribute(dense16_kernel<32, EPI_LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dense_smem<32>())));
  CK(cudaFuncSetAttribute(dense16_kernel<32, EPI_LSTM, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(dense_smem<32, __nv_bfloat16>())));
  CK(cudaFuncSetAttribute(dense16_kernel<8, EPI_QUERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dense_smem<8>())));
  CK(cudaFuncSetAttribute(dense16_kernel<8, EPI_FRAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dense_smem<8>())));
  CK(cudaFuncSetAttribute(dense16_kernel<8, EPI_PRENET>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dense_smem<8>())));
  if (const char* v = std::getenv("WG_TACO_PDL")) e->use_pdl = std::atoi(v) != 0;
  CK(cudaMalloc(reinterpret_cast<void**>(&e->io), sizeof(TacoIo)));
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
}

size_t energy_smem(const Dims& d) {
  return sizeof(float) * (static_cast<size_t>(kNF) * kAD + 2 * kAD + kNF * 2 * d.KS + 2 * (kPosPerCta + d.KS - 1) +
                          kPosPerCta * (kNF + 1));
}
size_t context_smem(int S) { return sizeof(float) * (static_cast<size_t>(S) + 32 + kCtxThreads); }

void ensure_state(wg_taco_engine* e, int G, int S, size_t pm_floats) {
  const Dims& d = e->d;
  if (G > e->cap_groups || S > e->cap_S) {
    drop_graphs(e);
    for (float* p : {e->XA, e->XD, e->XO, e->ca, e->cd, e->aw, e->awc, e->qT, e->frameT, e->x1T, e->energy}) cudaFree(p);
    cudaFree(e->finished); cudaFree(e->text_len);
    e->XA = e->XD = e->XO = e->ca = e->cd = e->aw = e->awc = e->qT = e->frameT = e->x1T = e->energy = nullptr;
    e->finished = e->text_len = nullptr;
    const int g = std::max(G, e->cap_groups), s = std::max(S, e->cap_S);
    e->cap_groups = e->cap_S = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&e->XA), sizeof(float) * g * 2 * d.KA * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->XD), sizeof(float) * g * 2 * d.KD * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->XO), sizeof(float) * g * d.KO * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->qT), sizeof(float) * g * kAD * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->frameT), sizeof(float) * g * d.NM * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->x1T), sizeof(float) * g * kP * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->ca), sizeof(float) * g * d.A * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->cd), sizeof(float) * g * d.D * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->aw), sizeof(float) * g * kRows * s));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->awc), sizeof(float) * g * kRows * s));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->energy), sizeof(float) * g * kRows * s));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->finished), sizeof(int) * g * kRows));
    CK(cudaMalloc(reinterpret_cast<void**>(&e->text_len), sizeof(int) * g * kRows));
    e->cap_groups = g;
    e->cap_S = s;
  }
  if (pm_floats > e->cap_pm) {
    cudaFree(e->pm);
    e->pm = nullptr;
    e->cap_pm = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&e->pm), sizeof(float) * pm_floats));
    e->cap_pm = pm_floats;
  }
}

template <typename Args>
void launch_chain(void (*kernel)(Args), dim3 grid, int block, size_t smem, cudaStream_t st, const Args& args, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, kernel, args));
}

// enqueues `frames` decoder frames (chunk-local indices 0..frames-1) on st: 7 kernels per frame
void enqueue_frames(wg_taco_engine* e, int B, int S, int frames, cudaStream_t st) {
  Dims d = e->d;
  d.S = S;
  const int G = (B + kRows - 1) / kRows;
  const size_t xa_gs = static_cast<size_t>(2) * d.KA * kRows, xd_gs = static_cast<size_t>(2) * d.KD * kRows;
  const size_t xo_gs = static_cast<size_t>(d.KO) * kRows;
  const bool pdl = e->use_pdl;
  const int pdl_instr = pdl ? 1 : 0;
  auto dense = [&](const float* Wp, const float* bp, const float* X, size_t x_gs, int K, int n_cols, int step) {
    DenseArgs a{};
    a.io = e->io; a.Wp = Wp; a.bp = bp; a.X = X; a.x_gs = x_gs; a.K = K; a.n_cols = n_cols; a.step = step;
    a.drop_rate = d.drop_rate;
    a.pdl = pdl_instr;
    return a;
  };
  for (int i = 0; i < frames; ++i) {
    const int p = i & 1;   // chunks start on even frames, so chunk-local parity == frame parity
    float* XAp = e->XA + static_cast<size_t>(p) * d.KA * kRows;
    float* XAq = e->XA + static_cast<size_t>(p ^ 1) * d.KA * kRows;
    float* XDp = e->XD + static_cast<size_t>(p) * d.KD * kRows;
    float* XDq = e->XD + static_cast<size_t>(p ^ 1) * d.KD * kRows;
    // attention LSTM on [prenet, context, h_a]: h_a -> next frame's XA, this frame's XD
    DenseArgs la = dense(e->WpA, e->bpA, XAp, xa_gs, d.KA, 4 * d.A, i);
    la.c = e->ca; la.c_gs = static_cast<size_t>(d.A) * kRows;
    la.o1 = XAq + static_cast<size_t>(kP + d.E) * kRows; la.o1_gs = xa_gs;
    la.o2 = XDp; la.o2_gs = xd_gs;
    if (e->lstm_mode == 2)
      launch_chain(lstm_mma_kernel, dim3(d.A / kUnitsPerCta, G), kDenseThreads, kMmaSmem, st, la, pdl);
    else if (e->lstm_bf16)
      launch_chain(dense16_kernel<32, EPI_LSTM, __nv_bfloat16>, dim3(d.A / kUnitsPerCta, G), kDenseThreads,
                   dense_smem<32, __nv_bfloat16>(), st, la, pdl);
    else
      launch_chain(dense16_kernel<32, EPI_LSTM>, dim3(d.A / kUnitsPerCta, G), kDenseThreads, dense_smem<32>(), st, la, pdl);
    // query projection of the new h_a
    DenseArgs qa = dense(e->WpQ, e->bpQ, XDp, xd_gs, d.A, kAD, i);
    qa.o1 = e->qT; qa.o1_gs = static_cast<size_t>(kAD) * kRows;
    launch_chain(dense16_kernel<8, EPI_QUERY>, dim3(kAD / 8, G), kDenseThreads, dense_smem<8>(), st, qa, pdl);
    AttnArgs aa{};
    aa.io = e->io; aa.qT = e->qT; aa.Wc = e->Wc; aa.Wld = e->Wld; aa.v = e->v;
    aa.ctx1 = XDp + static_cast<size_t>(d.A) * kRows; aa.ctx2 = XAq + static_cast<size_t>(kP) * kRows;
    aa.ctx3 = e->XO + static_cast<size_t>(d.D) * kRows;
    aa.aw = e->aw; aa.awc = e->awc; aa.energy = e->energy;
    aa.xd_gs = xd_gs; aa.xa_gs = xa_gs; aa.xo_gs = xo_gs; aa.d = d; aa.step = i; aa.pdl = pdl_instr;
    launch_chain(energy_kernel, dim3((S + kPosPerCta - 1) / kPosPerCta, B), kEnergyThreads, energy_smem(d), st, aa, pdl);
    launch_chain(context_kernel, dim3(kCtxSplit, B), kCtxThreads, context_smem(S), st, aa, pdl);
    // decoder LSTM on [h_a, context, h_d]: h_d -> next frame's XD and the frame projection input
    DenseArgs ld = dense(e->WpD, e->bpD, XDp, xd_gs, d.KD, 4 * d.D, i);
    ld.c = e->cd; ld.c_gs = static_cast<size_t>(d.D) * kRows;
    ld.o1 = XDq + static_cast<size_t>(d.A + d.E) * kRows; ld.o1_gs = xd_gs;
    ld.o2 = e->XO; ld.o2_gs = xo_gs;
    if (e->lstm_mode == 2)
      launch_chain(lstm_mma_kernel, dim3(d.D / kUnitsPerCta, G), kDenseThreads, kMmaSmem, st, ld, pdl);
    else if (e->lstm_bf16)
      launch_chain(dense16_kernel<32, EPI_LSTM, __nv_bfloat16>, dim3(d.D / kUnitsPerCta, G), kDenseThreads,
                   dense_smem<32, __nv_bfloat16>(), st, ld, pdl);
    else
      launch_chain(dense16_kernel<32, EPI_LSTM>, dim3(d.D / kUnitsPerCta, G), kDenseThreads, dense_smem<32>(), st, ld, pdl);
    // [h_d, context] -> frame | stop
    DenseArgs fa = dense(e->WpO, e->bpO, e->XO, xo_gs, d.KO, d.NM + 1, i);
    fa.o1 = e->frameT; fa.o1_gs = static_cast<size_t>(d.NM) * kRows;
    launch_chain(dense16_kernel<8, EPI_FRAME>, dim3((d.NM + 1 + 7) / 8, G), kDenseThreads, dense_smem<8>(), st, fa, pdl);
    // prenet of the new frame -> next frame's XA
    DenseArgs p0 = dense(e->WpP0, e->bpP0, e->frameT, static_cast<size_t>(d.NM) * kRows, d.NM, kP, i);
    p0.o1 = e->x1T; p0.o1_gs = static_cast<size_t>(kP) * kRows; p0.layer = 0;
    launch_chain(dense16_kernel<8, EPI_PRENET>, dim3(kP / 8, G), kDenseThreads, dense_smem<8>(), st, p0, pdl);
    DenseArgs p1 = dense(e->WpP1, e->bpP1, e->x1T, static_cast<size_t>(kP) * kRows, kP, kP, i);
    p1.o1 = XAq; p1.o1_gs = xa_gs; p1.layer = 1;
    launch_chain(dense16_kernel<8, EPI_PRENET>, dim3(kP / 8, G), kDenseThreads, dense_smem<8>(), st, p1, pdl);
  }
  advance_kernel<<<1, 1, 0, st>>>(e->io, frames);
  CK(cudaGetLastError());
}

void decode(wg_taco_engine* e, const float* memory, const int32_t* text_lengths, int B, int S, int max_length,
            int early_stopping, int deterministic, uint64_t seed, float* outputs, float* stops, float* attention,
            int32_t* lengths, int32_t* frames_run, cudaStream_t caller) {
  if (!memory || !text_lengths || !outputs || !stops || !lengths) fail(WG_ERR_INVALID, "wg_taco_decode: NULL buffer");
  if (B < 1 || B > 4096) fail(WG_ERR_INVALID, "wg_taco_decode: B %d outside [1, 4096]", B);
  if (S < 1 || S > 1024) fail(WG_ERR_INVALID, "wg_taco_decode: S %d outside [1, 1024]", S);
  if (max_length < 1) fail(WG_ERR_INVALID, "wg_taco_decode: max_length %d < 1", max_length);
  for (int b = 0; b < B; ++b)
    if (text_lengths[b] < 1 || text_lengths[b] > S)
      fail(WG_ERR_INVALID, "wg_taco_decode: text_lengths[%d] = %d outside [1, %d]", b, text_lengths[b], S);
  wg::DeviceGuard dev_guard(e->device);
  Dims d = e->d;
  d.S = S;
  CK(cudaFuncSetAttribute(energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(energy_smem(d))));
  const int G = (B + kRows - 1) / kRows;
  ensure_state(e, G, S, static_cast<size_t>(B) * S * kAD);
  cudaStream_t st = e->stream;
  CK(cudaEventRecord(e->ev_in, caller));
  CK(cudaStreamWaitEvent(st, e->ev_in, 0));
  // reset the recurrent state (get_initial_state: zeros, :395-417), bookkeeping and per-decode values
  CK(cudaMemsetAsync(e->XA, 0, sizeof(float) * G * 2 * d.KA * kRows, st));
  CK(cudaMemsetAsync(e->XD, 0, sizeof(float) * G * 2 * d.KD * kRows, st));
  CK(cudaMemsetAsync(e->XO, 0, sizeof(float) * G * d.KO * kRows, st));
  CK(cudaMemsetAsync(e->ca, 0, sizeof(float) * G * d.A * kRows, st));
  CK(cudaMemsetAsync(e->cd, 0, sizeof(float) * G * d.D * kRows, st));
  CK(cudaMemsetAsync(e->aw, 0, sizeof(float) * G * kRows * S, st));
  CK(cudaMemsetAsync(e->awc, 0, sizeof(float) * G * kRows * S, st));
  CK(cudaMemsetAsync(e->finished, 0, sizeof(int) * G * kRows, st));
  CK(cudaMemsetAsync(lengths, 0, sizeof(int32_t) * B, st));
  CK(cudaMemcpyAsync(e->text_len, text_lengths, sizeof(int) * B, cudaMemcpyHostToDevice, st));
  TacoIo io{};
  io.memory = memory; io.pm = e->pm; io.text_len = e->text_len; io.outputs = outputs; io.stops = stops;
  io.attn = attention; io.lengths = lengths; io.finished = e->finished; io.seed = seed; io.max_len = max_length;
  io.t_base = 0; io.deterministic = deterministic; io.B = B;
  CK(cudaMemcpyAsync(e->io, &io, sizeof io, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // `io` and text_lengths are host stack/heap: the copies must have left before we return
  processed_memory_kernel<<<B * S, kAD, 0, st>>>(memory, e->Wm, e->pm, d.E);
  CK(cudaGetLastError());

  int chunk = e->graph_chunk;
  int done = 0;
  std::vector<int> fin(B);
  auto all_finished = [&] {
    CK(cudaMemcpyAsync(fin.data(), e->finished, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return std::all_of(fin.begin(), fin.end(), [](int f) { return f != 0; });
  };
  if (chunk <= 0) {
    while (done < max_length) {
      int n = std::min(32, max_length - done);
      if (n > 1 && (n & 1)) --n;   // only the very last batch of frames may be odd (parity, see enqueue_frames)
      enqueue_frames(e, B, S, n, st);
      done += n;
      if (early_stopping && all_finished()) break;
    }
  } else {
    chunk += chunk & 1;   // even, so that chunk-local parity equals frame parity
    auto key = std::make_tuple(B, S, chunk);
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue_frames(e, B, S, chunk, st);
      } catch (...) {
        cudaStreamEndCapture(st, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      CK(cudaStreamEndCapture(st, &graph));
      cudaGraphExec_t exec = nullptr;
      cudaError_t rc = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (rc != cudaSuccess) fail(WG_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(rc));
      if (e->graphs.size() >= TACO_MAX_GRAPHS) {      // evict the least recently used graph
        auto lru = e->graph_used.begin();
        for (auto u = e->graph_used.begin(); u != e->graph_used.end(); ++u)
          if (u->second < lru->second) lru = u;
        auto victim = e->graphs.find(lru->first);
        if (victim != e->graphs.end()) {
          cudaGraphExecDestroy(victim->second);
          e->graphs.erase(victim);
        }
        e->graph_used.erase(lru);
      }
      it = e->graphs.emplace(key, exec).first;
    }
    e->graph_used[key] = ++e->graph_clock;
    while (done < max_length) {
      CK(cudaGraphLaunch(it->second, st));
      done += chunk;
      if (early_stopping && all_finished()) break;
    }
  }
  if (frames_run) *frames_run = std::min(done, max_length);
  CK(cudaEventRecord(e->ev_out, st));
  CK(cudaStreamWaitEvent(caller, e->ev_out, 0));
}

}  // namespace

extern "C" {

int wg_taco_create(const wg_taco_config* cfg, const wg_tensor* tensors, int32_t n_tensors, int32_t device,
                   wg_taco_handle* out) {
  if (out) *out = nullptr;
  if (!cfg || !tensors || !out || n_tensors <= 0) {
    std::lock_guard<std::mutex> g(g_taco_err_mu);
    g_taco_create_err = "wg_taco_create: NULL argument";
    return WG_ERR_INVALID;
  }
  wg_taco_engine* e = new wg_taco_engine();
  int rc = taco_guarded(nullptr, [&] { build_taco(e, cfg, tensors, n_tensors, device); });
  if (rc != WG_OK) {
    destroy_taco(e);
    return rc;
  }
  *out = e;
  return WG_OK;
}

void wg_taco_destroy(wg_taco_handle h) { destroy_taco(h); }

const char* wg_taco_last_error(wg_taco_handle h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> g(g_taco_err_mu);
  static thread_local std::string copy;
  copy = g_taco_create_err;
  return copy.c_str();
}

int wg_taco_decode(wg_taco_handle h, const float* memory, const int32_t* text_lengths, int32_t B, int32_t S,
                   int32_t max_length, int32_t early_stopping, int32_t deterministic, uint64_t seed, float* outputs,
                   float* stop_tokens, float* attention, int32_t* lengths, int32_t* frames_run, void* stream) {
  if (!h) return WG_ERR_INVALID;
  return taco_guarded(h, [&] {
    decode(h, memory, text_lengths, B, S, max_length, early_stopping, deterministic, seed, outputs, stop_tokens,
           attention, lengths, frames_run, static_cast<cudaStream_t>(stream));
  });
}

int wg_taco_set_graph_chunk(wg_taco_handle h, int32_t frames) {
  if (!h || frames < 0 || frames > 1024) return WG_ERR_INVALID;
  h->graph_chunk = frames;
  return WG_OK;
}

}  // extern "C"

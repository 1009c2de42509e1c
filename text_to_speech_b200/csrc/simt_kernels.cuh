// fp32 CUDA-core kernels of the WaveGlow engine (WG_MODE_FP32) and the memory-bound kernels shared
// by both modes (flow boundary: coupling inverse + W^-1 mixing + early-output concat + next start
// conv). Reference arithmetic: architectures/waveglow_arch.py:105-141, :244-306.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace wg {

// ------------------------------------------------------------------------------------------------
// Generic fp32 "row-gather" GEMM:  D[m, :] = sum_s A_s[row(m) + shift_s, 0:K_s] @ W[koff_s : koff_s+K_s, :]
// Rows are (b, l) with l in [0, L); a shifted row outside [0, L) contributes zero -- exactly the
// zero 'same' padding of the dilated conv (waveglow_arch.py:113-118) and the t-j<0 taps of the
// polyphase ConvTranspose (waveglow_arch.py:196-198, :245).
// ------------------------------------------------------------------------------------------------
struct ASeg {
  const float* ptr;  // [B*L, ld]
  int ld;
  int K;             // multiple of 16
  int shift;
};

struct GemmArgs {
  ASeg seg[4];
  int nseg;
  const float* W;     // [sum K_s, N] row-major
  const float* bias;  // [N] (may be null)
  int M, N, L;
  float* out0; int ld0;
  float* out1; int ld1;
  int res_cols;       // EPI_RES_SKIP: columns [0,res_cols) -> out0 += ; rest -> out1
  int skip_init;      // EPI_RES_SKIP: 1 = out1 is written, 0 = accumulated
};

enum { EPI_STORE = 0, EPI_GATE = 1, EPI_RES_SKIP = 2 };

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int EPI>
__global__ void __launch_bounds__(SG_THREADS)
gemm_f32_kernel(const GemmArgs a) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // A loader: 2 float4 per thread: rows ar, ar+64; k-quad akq
  const int ar = tid >> 2, akq = (tid & 3) * 4;
  // B loader: 2 float4 per thread: k rows bk, bk+8; cols bc
  const int bk = tid >> 5, bc = (tid & 31) * 4;

  int koff = 0;
  for (int s = 0; s < a.nseg; ++s) {
    const ASeg sg = a.seg[s];
    // per-thread source rows for this segment (two rows)
    const float* src[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = m0 + ar + 64 * i;
      src[i] = nullptr;
      if (m < a.M) {
        const int b = m / a.L, l = m - b * a.L;
        const int ls = l + sg.shift;
        if (ls >= 0 && ls < a.L) src[i] = sg.ptr + (size_t)(b * (size_t)a.L + ls) * sg.ld;
      }
    }
    for (int k0 = 0; k0 < sg.K; k0 += SG_BK) {
      float4 av[2], bv[2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        av[i] = src[i] ? *reinterpret_cast<const float4*>(src[i] + k0 + akq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n = n0 + bc;
        bv[i] = (n < a.N) ? *reinterpret_cast<const float4*>(a.W + (size_t)(koff + k0 + bk + 8 * i) * a.N + n)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        As[akq + 0][ar + 64 * i] = av[i].x;
        As[akq + 1][ar + 64 * i] = av[i].y;
        As[akq + 2][ar + 64 * i] = av[i].z;
        As[akq + 3][ar + 64 * i] = av[i].w;
        *reinterpret_cast<float4*>(&Bs[bk + 8 * i][bc]) = bv[i];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SG_BK; ++kk) {
        float ra[8], rb[8];
        *reinterpret_cast<float4*>(&ra[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        *reinterpret_cast<float4*>(&ra[4]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        *reinterpret_cast<float4*>(&rb[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
        *reinterpret_cast<float4*>(&rb[4]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
      }
    }
    koff += sg.K;
  }

  const int n = n0 + tx * 8;
  if (n >= a.N) return;
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias[j] = a.bias ? a.bias[n + j] : 0.f;

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= a.M) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bias[j];
    if (EPI == EPI_STORE) {
      float* o = a.out0 + (size_t)m * a.ld0 + n;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (EPI == EPI_GATE) {
      // packed columns come in groups of 8 = [tanh c..c+3 | sigmoid c..c+3]  (waveglow_arch.py:19-24)
      float4 r;
      r.x = tanhf(v[0]) * sigmoid_exact(v[4]);
      r.y = tanhf(v[1]) * sigmoid_exact(v[5]);
      r.z = tanhf(v[2]) * sigmoid_exact(v[6]);
      r.w = tanhf(v[3]) * sigmoid_exact(v[7]);
      *reinterpret_cast<float4*>(a.out0 + (size_t)m * a.ld0 + (n >> 1)) = r;
    } else {  // EPI_RES_SKIP  (waveglow_arch.py:129-139)
      if (n < a.res_cols) {
        float* o = a.out0 + (size_t)m * a.ld0 + n;
        float4 h0 = *reinterpret_cast<float4*>(o), h1 = *reinterpret_cast<float4*>(o + 4);
        *reinterpret_cast<float4*>(o) = make_float4(v[0] + h0.x, v[1] + h0.y, v[2] + h0.z, v[3] + h0.w);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4] + h1.x, v[5] + h1.y, v[6] + h1.z, v[7] + h1.w);
      } else {
        float* o = a.out1 + (size_t)m * a.ld1 + (n - a.res_cols);
        if (a.skip_init) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          float4 s0 = *reinterpret_cast<float4*>(o), s1 = *reinterpret_cast<float4*>(o + 4);
          *reinterpret_cast<float4*>(o) = make_float4(v[0] + s0.x, v[1] + s0.y, v[2] + s0.z, v[3] + s0.w);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4] + s1.x, v[5] + s1.y, v[6] + s1.z, v[7] + s1.w);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// End conv (C -> 2*n_half <= 8), fp32 mode: acc8[m, j] = bias[j] + sum_c skip[m, c] * Wend[c, j]
// (waveglow_arch.py:62-64, :141). One warp per row, HBM-bound (reads skip once).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
end_conv_kernel(const float* __restrict__ skip, const float* __restrict__ Wend /*[C,8] padded*/,
                const float* __restrict__ bend /*[8]*/, float* __restrict__ acc8, int M, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  const float* row = skip + (size_t)warp * C;
  for (int c = lane; c < C; c += 32) {
    const float x = row[c];
    const float4 w0 = *reinterpret_cast<const float4*>(Wend + c * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(Wend + c * 8 + 4);
    s[0] = fmaf(x, w0.x, s[0]); s[1] = fmaf(x, w0.y, s[1]); s[2] = fmaf(x, w0.z, s[2]); s[3] = fmaf(x, w0.w, s[3]);
    s[4] = fmaf(x, w1.x, s[4]); s[5] = fmaf(x, w1.y, s[5]); s[6] = fmaf(x, w1.z, s[6]); s[7] = fmaf(x, w1.w, s[7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  }
  if (lane < 8) {
    float v = s[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) if (lane == j) v = s[j];
    acc8[(size_t)warp * 8 + lane] = v + bend[lane];
  }
}

// ------------------------------------------------------------------------------------------------
// Flow boundary (both modes), HBM-bound, one pass:
//   coupling inverse   a1 = (a1 - b) / exp(s)                 waveglow_arch.py:278-288
//   W^-1 mixing        audio = audio @ Winv                    invertible_conv.py:49-51
//   early re-injection audio = [sigma * z_i, audio]            waveglow_arch.py:292-304
//   next flow's start conv  h = audio_0 @ Wstart + bstart      waveglow_arch.py:58, :108
// `first` replaces the three first steps by audio = sigma * z[:, :n_rem] (waveglow_arch.py:264-275).
// The last flow writes the [M, 8] audio straight into the output waveform ([B, 8 L], :306).
// ------------------------------------------------------------------------------------------------
// Internal row order of the bf16-path buffers. R = 1: position-major (Tp = T = L, row = b*L + l). R = 32: phase-major
// with all utterances of a phase in ONE row sequence, Tp = T + gap rows apart: row = (r*B + b)*Tp + t. The gap rows
// (t >= T) are kept ZERO, so a dilated-conv tap that runs off an utterance reads the reference's zero 'same' padding
// (gap >= max dilation / R frames) and 128-row tiles may span utterances (no ragged last tile per utterance).
struct RowGeom {
  int R, T, Tp, B;
  // Ragged batches (wg_infer_ragged): utterance b has len[b] <= T frames and starts at row off[b] of every phase block;
  // row_b[t] = utterance of phase-block row t, -1 for a gap row; rpp = rows per phase block (sum of len[b] + gap).
  // All three null / rpp == 0: uniform layout (every utterance T frames, Tp rows apart). The caller-side buffers
  // (mel, z, waveform) keep the padded [B, T, ..] shape either way.
  const int* off = nullptr;
  const int* len = nullptr;
  const int* row_b = nullptr;
  int rpp = 0;
  __host__ __device__ int rows_per_phase() const { return row_b ? rpp : B * Tp; }
  __host__ __device__ int rows() const { return R * rows_per_phase(); }
  __device__ int frames(int b) const { return len ? len[b] : T; }
  // phase-block row -> (b, t); returns false for a gap row
  __device__ bool decode_row(int rem, int& b, int& t) const {
    if (row_b) {
      b = row_b[rem];
      if (b < 0) return false;
      t = rem - off[b];
      return true;
    }
    b = rem / Tp;
    t = rem - b * Tp;
    return t < T;
  }
  // internal row -> (b, r, t); returns false for a gap row
  __device__ bool decode(int m, int& b, int& r, int& t) const {
    const int per_r = rows_per_phase();
    r = m / per_r;
    return decode_row(m - r * per_r, b, t);
  }
  __device__ size_t internal(int b, int l) const {   // position l = R*t + r of utterance b
    const int t = l / R, r = l - t * R;
    if (row_b) return static_cast<size_t>(r) * rpp + off[b] + t;
    return (static_cast<size_t>(r) * B + b) * Tp + t;
  }
  __device__ size_t position_major(int b, int r, int t) const { return static_cast<size_t>(b) * R * T + static_cast<size_t>(t) * R + r; }
};

// Ragged geometry tables, one block per utterance: off/len from the kernel parameters (no host buffer has to outlive
// the call, so wg_infer_ragged stays asynchronous and graph-capturable), row_b for the utterance's rows and its gap.
struct GeomChunk {
  int b0, n, gap;
  int off[256];
  int len[256];
};
__global__ void __launch_bounds__(128)
ragged_geom_kernel(const __grid_constant__ GeomChunk c, int* __restrict__ off, int* __restrict__ len, int* __restrict__ row_b) {
  const int j = blockIdx.x;
  if (j >= c.n) return;
  const int o = c.off[j], n = c.len[j];
  if (threadIdx.x == 0) {
    off[c.b0 + j] = o;
    len[c.b0 + j] = n;
  }
  for (int t = threadIdx.x; t < n + c.gap; t += blockDim.x) row_b[o + t] = t < n ? c.b0 + j : -1;
}

struct BoundaryArgs {
  const float* acc8;      // [M,8]: cols [0,n_half) = b, [n_half, 2 n_half) = log s   (null when first)
  const float* audio_in;  // [M,8] (c_in channels used)
  float* audio_out;       // [M,8]
  const float* z;         // [M, n_group] or null (deterministic)
  int n_group;
  int z_off;              // first z channel consumed here
  int n_inject;           // channels of z prepended (first: n_rem; early flows: n_early_size; else 0)
  float sigma;
  int first;
  int c_in;               // channels before injection (2 * n_half)
  float winv[64];         // [c_in, c_in] row-major: out[b] = sum_a in[a] * winv[a*c_in + b]
  // next flow start conv (null Wstart => none)
  const float* Wstart;    // [n_half_next, C]
  const float* bstart;    // [C]
  int n_half_next;
  int C;
  float* h32;             // [M, C]
  __nv_bfloat16* h16;     // [M, C] or null: bf16(h)            (bf16 mode: residual stream = hi + lo)
  __nv_bfloat16* hlo;     // [M, C] or null: bf16(h - bf16(h))
  float* hf_hi;           // [M, C] or null: tf32(h)            (tf32x3 mode: residual stream = hi + lo, both fp32)
  float* hf_lo;           // [M, C] or null: h - tf32(h)
  __nv_bfloat16* hf_b;    // [M, 2C] or null: per 32 channels bf16(hi) x 32 | bf16(lo) x 32, the operands of the cross-term MMAs (tc_tf32_kernels.cuh)
  int acc_parts;          // acc8 is the sum of this many partial accumulators (tf32x3: one per gate chunk), >= 1
  size_t acc_part_stride; // floats between two partials
  int M;
  // bf16 mode: re-arm the folded skip/end accumulator for the next flow (acc8 = bias term) once
  // this flow's value has been consumed
  float* acc8_rearm;      // [M,8] or null
  float acc8_init[8];
  // Row geometry of the internal buffers (RowGeom below): position l = R*t + r of utterance b lives at row
  // (r*B + b)*Tp + t; rows with t >= T are zero gap rows. z and the final waveform are always position-major.
  // M counts INTERNAL rows (geo.rows()).
  RowGeom geo;
  int final_out;          // audio_out is the caller's waveform buffer (position-major)
};

constexpr int FB_ROWS = 64, FB_THREADS = 256;

__global__ void __launch_bounds__(FB_THREADS)
flow_boundary_kernel(const __grid_constant__ BoundaryArgs a) {
  __shared__ float s_a0[FB_ROWS][4];
  __shared__ int s_gap[FB_ROWS];
  const int m0 = blockIdx.x * FB_ROWS;
  const int tid = threadIdx.x;
  const RowGeom& geo = a.geo;
  if (tid < FB_ROWS) {
    const int m = m0 + tid;
    s_gap[tid] = 0;
    int gb, gr, gt;
    if (m < a.M && !geo.decode(m, gb, gr, gt)) {
      s_gap[tid] = 1;       // gap row: stays zero in every buffer (the start conv below writes h = 0, no bias)
      if (!a.final_out) {   // (the caller's waveform buffer has no gap rows)
        *reinterpret_cast<float4*>(a.audio_out + (size_t)m * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(a.audio_out + (size_t)m * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (m < a.M) {
      const size_t prow = geo.position_major(gb, gr, gt);   // position-major row of this internal row
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
      int c = 0;
      if (a.first) {
        c = a.n_inject;
        for (int j = 0; j < c; ++j) x[j] = a.z ? a.sigma * a.z[prow * a.n_group + a.z_off + j] : 0.f;
      } else {
        const int nh = a.c_in >> 1;
        const float4 i0 = *reinterpret_cast<const float4*>(a.audio_in + (size_t)m * 8);
        const float4 i1 = *reinterpret_cast<const float4*>(a.audio_in + (size_t)m * 8 + 4);
        float4 o0 = *reinterpret_cast<const float4*>(a.acc8 + (size_t)m * 8);
        float4 o1 = *reinterpret_cast<const float4*>(a.acc8 + (size_t)m * 8 + 4);
        for (int pp = 1; pp < a.acc_parts; ++pp) {      // partial accumulators, summed in a fixed order
          const float4 q0 = *reinterpret_cast<const float4*>(a.acc8 + pp * a.acc_part_stride + (size_t)m * 8);
          const float4 q1 = *reinterpret_cast<const float4*>(a.acc8 + pp * a.acc_part_stride + (size_t)m * 8 + 4);
          o0.x += q0.x; o0.y += q0.y; o0.z += q0.z; o0.w += q0.w;
          o1.x += q1.x; o1.y += q1.y; o1.z += q1.z; o1.w += q1.w;
        }
        float in[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        const float o[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        for (int j = 0; j < nh; ++j) in[nh + j] = (in[nh + j] - o[j]) / expf(o[nh + j]);
        float y[8];
        for (int bq = 0; bq < a.c_in; ++bq) {
          float acc = 0.f;
          for (int aq = 0; aq < a.c_in; ++aq) acc = fmaf(in[aq], a.winv[aq * a.c_in + bq], acc);
          y[bq] = acc;
        }
        for (int j = 0; j < a.n_inject; ++j)
          x[j] = a.z ? a.sigma * a.z[prow * a.n_group + a.z_off + j] : 0.f;
        for (int j = 0; j < a.c_in; ++j) x[a.n_inject + j] = y[j];
        c = a.c_in + a.n_inject;
      }
      const size_t orow = a.final_out ? prow : static_cast<size_t>(m);
      *reinterpret_cast<float4*>(a.audio_out + orow * 8) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(a.audio_out + orow * 8 + 4) = make_float4(x[4], x[5], x[6], x[7]);
      if (a.acc8_rearm) {
        *reinterpret_cast<float4*>(a.acc8_rearm + (size_t)m * 8) =
            make_float4(a.acc8_init[0], a.acc8_init[1], a.acc8_init[2], a.acc8_init[3]);
        *reinterpret_cast<float4*>(a.acc8_rearm + (size_t)m * 8 + 4) =
            make_float4(a.acc8_init[4], a.acc8_init[5], a.acc8_init[6], a.acc8_init[7]);
        for (int pp = 1; pp < a.acc_parts; ++pp) {
          *reinterpret_cast<float4*>(a.acc8_rearm + pp * a.acc_part_stride + (size_t)m * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(a.acc8_rearm + pp * a.acc_part_stride + (size_t)m * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s_a0[tid][j] = x[j];
      (void)c;
    }
  }
  if (a.Wstart == nullptr) return;
  __syncthreads();
  // start conv of the next flow: coalesced over channels, 8 per thread (16-byte bf16 stores)
  const int C8 = a.C >> 3;
  for (int idx = tid; idx < FB_ROWS * C8; idx += FB_THREADS) {
    const int r = idx / C8, c8 = (idx - r * C8) * 8;
    const int m = m0 + r;
    if (m >= a.M) break;
    float v[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(a.bstart + c8);
      const float4 b1 = *reinterpret_cast<const float4*>(a.bstart + c8 + 4);
      v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    }
    if (s_gap[r]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    for (int j = 0; j < (s_gap[r] ? 0 : a.n_half_next); ++j) {
      const float x = s_a0[r][j];
      const float4 w0 = *reinterpret_cast<const float4*>(a.Wstart + (size_t)j * a.C + c8);
      const float4 w1 = *reinterpret_cast<const float4*>(a.Wstart + (size_t)j * a.C + c8 + 4);
      v[0] = fmaf(x, w0.x, v[0]); v[1] = fmaf(x, w0.y, v[1]); v[2] = fmaf(x, w0.z, v[2]); v[3] = fmaf(x, w0.w, v[3]);
      v[4] = fmaf(x, w1.x, v[4]); v[5] = fmaf(x, w1.y, v[5]); v[6] = fmaf(x, w1.z, v[6]); v[7] = fmaf(x, w1.w, v[7]);
    }
    if (a.h32) {
      *reinterpret_cast<float4*>(a.h32 + (size_t)m * a.C + c8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(a.h32 + (size_t)m * a.C + c8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a.h16) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        const float2 f = __bfloat1622float2(p);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * q] - f.x, v[2 * q + 1] - f.y);
        hi[q] = *reinterpret_cast<const uint32_t*>(&p);
        lo[q] = *reinterpret_cast<const uint32_t*>(&l);
      }
      *reinterpret_cast<uint4*>(a.h16 + (size_t)m * a.C + c8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(a.hlo + (size_t)m * a.C + c8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (a.hf_hi) {
      float hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v[q]));
        hi[q] = __uint_as_float(u);
        lo[q] = v[q] - hi[q];
      }
      *reinterpret_cast<float4*>(a.hf_hi + (size_t)m * a.C + c8) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(a.hf_hi + (size_t)m * a.C + c8 + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<float4*>(a.hf_lo + (size_t)m * a.C + c8) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<float4*>(a.hf_lo + (size_t)m * a.C + c8 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      uint32_t hb[4], lb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 ph = __floats2bfloat162_rn(hi[2 * q], hi[2 * q + 1]);
        const __nv_bfloat162 pl = __floats2bfloat162_rn(lo[2 * q], lo[2 * q + 1]);
        hb[q] = *reinterpret_cast<const uint32_t*>(&ph);
        lb[q] = *reinterpret_cast<const uint32_t*>(&pl);
      }
      // A-operand companion: per K-block of 32 channels [hb x 32 | lb x 32]
      __nv_bfloat16* bp = a.hf_b + (size_t)m * 2 * a.C + 64 * (c8 >> 5) + (c8 & 31);
      *reinterpret_cast<uint4*>(bp) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
      *reinterpret_cast<uint4*>(bp + 32) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Start-conv fold (bf16 phase-major path): the A operand of the FIRST WN layer of a flow.
//
// h0 = start(audio_0) is a rank-n_half (<= 4) function of the audio rows, so layer 0's dilated conv
// (dilation 1) and its residual do not need h0 in HBM at all:
//   conv[l]  = sum_tap h0[l+tap-1] @ Win[tap] = sum_tap [a(l+tap-1), 1] @ [Wstart @ Win[tap]; bstart @ Win[tap]]
//   resid[l] = h0[l]                           =         [a(l), 1]       @ [Wstart; bstart]
// with a zero row (including the ones column) for positions outside the utterance == the reference's
// zero 'same' padding of h0 (waveglow_arch.py:108, :113-118). Row m of `a0` [M, 64] bf16 holds four
// 16-column groups -- taps l-1, l, l+1 and the residual operand l -- each
//   [a_hi(4) | a_lo(4) | a_hi(4) | 1 | 1 | 0 | 0],   a = a_hi + a_lo (bf16 split of the fp32 audio),
// to be multiplied with [G_hi; G_hi; G_lo; g_hi; g_lo] (three-product split, fp32-grade accuracy).
// One thread per (row, group); rows follow the internal order (RowGeom); gap rows are all zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
a0_build_kernel(const float* __restrict__ audio, __nv_bfloat16* __restrict__ a0, const RowGeom geo, int n_half) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= geo.rows() * 4) return;
  const int m = idx >> 2, g = idx & 3;
  int b, r, t;
  const bool valid = geo.decode(m, b, r, t);
  const int l = t * geo.R + r + (g < 3 ? g - 1 : 0);
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = 0u;
  if (valid && l >= 0 && l < geo.R * geo.frames(b)) {
    const size_t src = geo.internal(b, l);
    const float4 v4 = *reinterpret_cast<const float4*>(audio + src * 8);
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j >= n_half) v[j] = 0.f;
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y);
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
    w[0] = *reinterpret_cast<const uint32_t*>(&h01); w[1] = *reinterpret_cast<const uint32_t*>(&h23);
    w[2] = *reinterpret_cast<const uint32_t*>(&l01); w[3] = *reinterpret_cast<const uint32_t*>(&l23);
    w[4] = w[0]; w[5] = w[1];
    w[6] = 0x3F803F80u;   // two bf16 ones
  }
  uint4* dst = reinterpret_cast<uint4*>(a0 + static_cast<size_t>(m) * 64 + g * 16);
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

}  // namespace wg

// libwg_b200.so -- host side of the C ABI declared in include/wg_b200.h.
// Weight repacking (Keras layouts -> kernel layouts), workspace carving, launch sequence of
// WaveGlow.infer (architectures/waveglow_arch.py:244-306) and the C entry points.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wg_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "tc_kernels.cuh"
#include "tc_pair_kernels.cuh"
#include "tc_c512_kernels.cuh"
#include "tc_tf32_kernels.cuh"
#include "tc_tf32_flow_kernel.cuh"

namespace {

using namespace wg;

#ifndef WG_PAIR_DEFAULT
#define WG_PAIR_DEFAULT 1        // the CTA-pair layer kernel is the default where its wave count allows
#endif
constexpr int HOP = 256;         // upsample stride (waveglow_arch.py:197)
constexpr int UPSAMPLE_K = 1024; // upsample kernel size

std::mutex g_err_mu;
std::string g_create_err;

#define CK WG_CK

struct FlowW {
  int n_half = 0, n_rem = 0;
  float* Wstart = nullptr;  // [n_half, C]
  float* bstart = nullptr;  // [C]
  float* Wend8 = nullptr;   // [C, 8] zero padded   (fp32 mode)
  float* bend8 = nullptr;   // [8]
  float bse8[8] = {0};      // bf16 mode: bend + sum_i bskip_i @ Wend (host copy, passed inline)
  float winv[64];
};

struct LayerW {
  // fp32 mode
  float* Wcat = nullptr;  // [3C+S, 2C] gate-interleaved columns
  float* bcat = nullptr;  // [2C]
  float* Wrs = nullptr;   // [C, rs_cols]
  float* brs = nullptr;   // [rs_cols]
  int rs_cols = 0;
  // bf16 mode (fp32 side arrays; the bf16 matrices live in the stacked arrays below)
  float* b1 = nullptr;    // [2C] chunk-packed order
  float* b1_pm = nullptr; // same + the upsample bias pushed through the cond weights (phase-major path)
  float* b2 = nullptr;    // [C]
  std::vector<float> wse_h;  // [C, 8] = Wskip @ Wend (fp32), host copy: passed in the kernel parameter bank
  std::vector<float> wse_p;  // the same in the packed-fold layout [channel pair][column][even, odd] (gate_step2)
  float* wse_d = nullptr;    // tf32x3 mode: device copy of wse_h
};

}  // namespace

struct wg_engine {
  wg_config cfg{};
  int device = 0;
  int C = 0, S = 0, R = 0;  // channels, spect channels (n_mel*n_group), time-groups per frame
  std::vector<FlowW> flows;
  std::vector<LayerW> layers;  // [flow * n_layers + i]
  float* Wup = nullptr;        // fp32: [4*n_mel, R*S]
  float* bup = nullptr;        // [R*S]
  // bf16 mode stacked operand matrices (K-major)
  __nv_bfloat16* Wup16 = nullptr;  // [R*S, Kup_pad]
  __nv_bfloat16* W1 = nullptr;     // [n_flows*n_layers*2C, 3C+S]
  __nv_bfloat16* W2 = nullptr;     // [n_flows*n_layers*C, C]
  __nv_bfloat16* V = nullptr;      // [n_flows*n_layers*R*2C, Kup]: (Wup_r @ Wcond) per layer and upsample phase
  // tf32x3 mode: the same stacked matrices as fp32 (hi, lo) pairs (tc_tf32_kernels.cuh); W1 holds the conv part only
  wg::Tf32Weights t3{};
  // start-conv fold (phase-major bf16 path): layer 0 of a flow consumes the audio rows directly
  __nv_bfloat16* W0 = nullptr;     // [n_flows*2C, 64]: per tap [Wstart@Win_tap hi | same | lo | bstart@Win_tap hi | lo | 0 0], chunk-packed rows
  __nv_bfloat16* H0 = nullptr;     // [n_flows*C, 64]: columns 48..63 = [Wstart hi | same | lo | bstart hi | lo | 0 0]
  bool fold0 = true;               // WG_FOLD0=0 keeps the materialised start conv (A/B and debugging)
  int pm_policy = -1;               // WG_PM: -1 auto (by tile efficiency), 0 never, 1 always
  int Kup = 0;
  std::vector<void*> allocs;
  std::string err;
  int launches = 0;
  int sm_count = 148;
  // wg_infer_host staging
  cudaStream_t stream = nullptr;
  float *pin_mel = nullptr, *pin_z = nullptr, *pin_out = nullptr;
  float *dev_mel = nullptr, *dev_z = nullptr, *dev_out = nullptr;
  void* dev_ws = nullptr;
  size_t cap_mel = 0, cap_z = 0, cap_out = 0, cap_ws = 0;
  // per-kernel profiling (wg_profile_enable / wg_profile_read)
  unsigned long long* timing = nullptr;   // WG_LAYER_TIMING=1: in-kernel cycle counters (debug)
  int dbg_flags = 0;                      // WG_DEBUG_FLAGS, honoured only by a -DWG_PROBES build (WnLayerParams::flags)
  int pair_policy = -1;                   // WG_PAIR: 1 = CTA-pair (cta_group::2) layer kernel, 0 = single-CTA kernel, -1 = default
  int pair_max = 0;                       // CTA pairs that can be resident at once on this device (tc_pair_init)
  int last_flow_kernel = 0;               // the last wg_infer ran its flows on tf32_flow_kernel
  int last_pair = 0;                      // the last wg_infer ran its layers on the CTA-pair kernel
  int pair_epi_warps = 8;                 // WG_PAIR_EPI=16: 16 epilogue warps in the pair kernel
  bool pdl = true;                        // WG_PDL=0: no programmatic dependent launch of the BF16 layer kernels (A/B)
  Tf32FlowState t3_flow;                  // one-launch-per-flow kernel for single-wave tf32x3 calls (tc_tf32_flow_kernel.cuh)
  int t3_flow_policy = -1;                // WG_TF32_FLOW: 0 = per-layer kernels only, 1 / -1 = flow kernel where it fits
  int t3_epi_warps = 0;                   // WG_TF32_EPI=8 / 16: force the epilogue warp count of the tf32x3 kernels (A/B)
  int t3_max_pairs = 0;                   // resident CTA pairs of the tf32x3 kernels (tf32_init)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
  std::vector<int> ev_count;          // per pair: layer launches bracketed by it
  size_t ev_used = 0;
};

namespace {

template <typename T>
T* upload(wg_engine* e, const std::vector<T>& v) {
  T* d = nullptr;
  CK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  e->allocs.push_back(d);
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct TensorView {
  const float* data;
  std::vector<int64_t> shape;
};

const TensorView& need(const std::map<std::string, TensorView>& m, const std::string& name,
                       std::initializer_list<int64_t> shape) {
  auto it = m.find(name);
  if (it == m.end()) fail(WG_ERR_WEIGHTS, "missing weight tensor '%s'", name.c_str());
  std::vector<int64_t> want(shape);
  if (it->second.shape != want) {
    std::string got, exp;
    for (auto d : it->second.shape) got += std::to_string(d) + ",";
    for (auto d : want) exp += std::to_string(d) + ",";
    fail(WG_ERR_WEIGHTS, "weight '%s': shape [%s] != expected [%s]", name.c_str(), got.c_str(), exp.c_str());
  }
  return it->second;
}

// c x c inverse in fp32 (Gauss-Jordan, partial pivoting) -- invertible_conv.py:41-47 uses K.inv on
// the float32 kernel once at weight load.
void invert_f32(const float* W, int c, float* inv) {
  float a[8][16];
  for (int i = 0; i < c; ++i)
    for (int j = 0; j < c; ++j) {
      a[i][j] = W[i * c + j];
      a[i][c + j] = (i == j) ? 1.f : 0.f;
    }
  for (int col = 0; col < c; ++col) {
    int piv = col;
    for (int r = col + 1; r < c; ++r)
      if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
    if (std::fabs(a[piv][col]) < 1e-12f) fail(WG_ERR_WEIGHTS, "invertible 1x1 kernel is singular");
    if (piv != col)
      for (int j = 0; j < 2 * c; ++j) std::swap(a[piv][j], a[col][j]);
    const float d = 1.f / a[col][col];
    for (int j = 0; j < 2 * c; ++j) a[col][j] *= d;
    for (int r = 0; r < c; ++r) {
      if (r == col) continue;
      const float f = a[r][col];
      if (f != 0.f)
        for (int j = 0; j < 2 * c; ++j) a[r][j] -= f * a[col][j];
    }
  }
  for (int i = 0; i < c; ++i)
    for (int j = 0; j < c; ++j) inv[i * c + j] = a[i][c + j];
}

inline __nv_bfloat16 f2bf(float x) { return __float2bfloat16_rn(x); }

struct Ws {  // workspace carving for one (B, T)
  size_t spect = 0, h32 = 0, acts = 0, skip = 0, acc8 = 0, audio0 = 0, audio1 = 0;
  size_t spect16 = 0, h16a = 0, h16b = 0, hlo = 0, aup16 = 0, acts16 = 0, a0 = 0;
  size_t g_off = 0, g_len = 0, g_rowb = 0;   // ragged geometry tables (RowGeom)
  size_t t3_hhi0 = 0, t3_hhi1 = 0, t3_hlo0 = 0, t3_hlo1 = 0, t3_hb0 = 0, t3_hb1 = 0, t3_ahi = 0, t3_ab = 0, t3_chi = 0, t3_cb = 0, t3_sync = 0;
  size_t total = 0;
};

// Ragged batch: per-utterance frame counts (host copy) and the phase-block row offsets derived from them.
struct Ragged {
  std::vector<int> len, off;
  int rpp = 0;   // rows per phase block = sum(len[b] + gap)
};

// rows between two utterances inside a phase block: >= max dilation / R frames, all zero (RowGeom)
int pm_gap(const wg_engine* e) { return ((1 << (e->cfg.n_layers - 1)) + e->R - 1) / e->R; }

// Row layout of the bf16 path, chosen per call by a wave count: phase-major tiles cost 17 + 17 + 6 K-blocks (rank-320
// conditioning, no spect) but come in multiples of R = 32 per 128 frames of the joint sequence; position-major tiles cost
// 22 + 22 + 6 and need the materialised spect. They only differ in speed -- both meet the same parity bar -- and only
// small inputs (a few hundred tiles) ever pick position-major. WG_PM=0/1 forces one.
bool use_pm(const wg_engine* e, int B, int T) {
  if (e->cfg.mode == WG_MODE_TF32X3) return true;    // the tf32x3 kernels only exist for the phase-major layout
  if (e->cfg.mode != WG_MODE_BF16 || !e->V) return false;
  if (e->pm_policy == 0) return false;
  if (e->pm_policy == 1) return true;
  const long sm = e->sm_count > 0 ? e->sm_count : 148;
  const long tiles_pm = (long)e->R * (((long)B * (T + pm_gap(e)) + 127) / 128);
  const long tiles_pos = (long)B * (((long)T * e->R + 127) / 128);
  const long cost_pm = ((tiles_pm + sm - 1) / sm) * 40, cost_pos = ((tiles_pos + sm - 1) / sm) * 50 + 2;
  return cost_pm <= cost_pos;
}

// CTA-pair layer kernels (phase-major; C = 256: tc_pair_kernels.cuh, C = 512: tc512_gate_pair_kernel): same bits, 0-3.4 % /
// 7 % less time per tile (same-box A/Bs, profiles/r02_pair_ab*.jsonl) -- chosen when the wave count (pairs of row tiles on pairs of SMs; an odd tile count per phase
// block leaves a ghost tile) does not eat that gain. WG_PAIR=0/1 forces either kernel.
bool use_pair(const wg_engine* e, const TcPlan& pl) {
  if (e->pair_policy == 0 || e->pair_max < 1) return false;
  if (e->pair_policy == 1) return true;
  if (!WG_PAIR_DEFAULT) return false;
  // a device whose complete TPCs are fewer than sm_count / 2 leaves SMs idle under the pair kernel: the wave count decides
  const long sm = e->sm_count > 0 ? e->sm_count : 148, pairs = e->pair_max;
  const long tiles = (long)pl.tiles_per_row * pl.R, pair_tiles = (long)((pl.tiles_per_row + 1) / 2) * pl.R;
  const double gain = pl.C == 512 ? 0.93 : 0.966;      // measured per-wave time of the pair kernel relative to the single-CTA one
  const double cost_single = (double)((tiles + sm - 1) / sm), cost_pair = gain * (double)((pair_tiles + pairs - 1) / pairs);
  return cost_pair < cost_single;
}

Ws carve(const wg_engine* e, int B, int T, const Ragged* rg = nullptr) {
  Ws w;
  const bool pm = rg ? true : use_pm(e, B, T);
  const size_t rows1 = rg ? (size_t)rg->rpp : (size_t)B * (T + pm_gap(e));           // rows of one phase block
  const size_t M = pm ? rows1 * e->R : (size_t)B * T * e->R;   // internal rows (gap rows included)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  if (e->cfg.mode == WG_MODE_FP32) w.h32 = take(M * e->C * 4);
  const size_t acc_parts = e->cfg.mode == WG_MODE_TF32X3 ? (size_t)(2 * e->C / 256) : 1;   // tf32x3: one partial per gate chunk
  w.acc8 = take(acc_parts * M * 8 * 4);
  w.audio0 = take(M * 8 * 4);
  w.audio1 = take(M * 8 * 4);
  if (e->cfg.mode == WG_MODE_FP32) {
    w.spect = take(M * e->S * 4);
    w.acts = take(M * e->C * 4);
    w.skip = take(M * e->C * 4);
  } else if (e->cfg.mode == WG_MODE_TF32X3) {
    // every MMA operand as tf32(x) in fp32 words + a bf16 companion [.., 2K] (bf16(hi) | bf16(lo)): residual stream x2
    // (ping-pong; + its exact fp32 remainder for the residual add), acts, mel window; one partial fold accumulator per chunk
    w.t3_hhi0 = take(M * e->C * 4); w.t3_hhi1 = take(M * e->C * 4);
    w.t3_hlo0 = take(M * e->C * 4); w.t3_hlo1 = take(M * e->C * 4);
    w.t3_hb0 = take(M * e->C * 4); w.t3_hb1 = take(M * e->C * 4);
    w.t3_ahi = take(M * e->C * 4); w.t3_ab = take(M * e->C * 4);
    w.t3_chi = take(rows1 * e->Kup * 4); w.t3_cb = take(rows1 * e->Kup * 4);
    w.t3_sync = take(T3F_SYNC_WORDS * sizeof(unsigned int));   // barrier words of tf32_flow_kernel: per CALL, a handle may run two at once
    if (rg) {
      w.g_off = take((size_t)B * 4);
      w.g_len = take((size_t)B * 4);
      w.g_rowb = take(rows1 * 4);
    }
  } else {
    if (!pm) w.spect16 = take(M * e->S * 2);
    w.h16a = take(M * e->C * 2);
    w.h16b = take(M * e->C * 2);
    w.hlo = take(M * e->C * 2);
    w.aup16 = take((pm ? rows1 : (size_t)B * T) * e->Kup * 2);
    if (e->C == 512) w.acts16 = take(M * e->C * 2);   // WaveGlow-512: acts travel between the gate and residual kernels
    if (e->W0 && pm) w.a0 = take(M * 64 * 2);   // start fold: A operand of each flow's first layer
    if (rg) {
      w.g_off = take((size_t)B * 4);
      w.g_len = take((size_t)B * 4);
      w.g_rowb = take(rows1 * 4);
    }
  }
  w.total = off;
  return w;
}

void check_shape(const wg_engine* e, int B, int T) {
  if (B <= 0 || T <= 0) fail(WG_ERR_INVALID, "B and T must be positive (got B=%d, T=%d)", B, T);
  const double M = (double)B * (T + (e->cfg.mode != WG_MODE_FP32 ? pm_gap(e) : 0)) * e->R;   // internal rows, gap rows included
  if (M * std::max(e->S, e->C) > 2.0e9) fail(WG_ERR_INVALID, "B*T too large (B=%d, T=%d)", B, T);
}

// Validates per-utterance lengths (1 <= T_b[b] <= T) and lays the utterances out one after the other, gap rows apart.
Ragged make_ragged(const wg_engine* e, int B, int T, const int32_t* T_b) {
  check_shape(e, B, T);
  if (!T_b) fail(WG_ERR_INVALID, "T_b must not be NULL");
  Ragged rg;
  rg.len.assign(T_b, T_b + B);
  rg.off.resize(B);
  const int gap = pm_gap(e);
  long rows = 0;
  for (int b = 0; b < B; ++b) {
    if (T_b[b] <= 0 || T_b[b] > T) fail(WG_ERR_INVALID, "T_b[%d] = %d is outside [1, T=%d]", b, T_b[b], T);
    rg.off[b] = (int)rows;
    rows += T_b[b] + gap;
  }
  rg.rpp = (int)rows;
  return rg;
}

bool is_uniform(const Ragged& rg, int T) {
  for (int l : rg.len)
    if (l != T) return false;
  return true;
}

template <int EPI>
void launch_gemm(wg_engine* e, const GemmArgs& a, cudaStream_t st) {
  dim3 grid((a.N + SG_BN - 1) / SG_BN, (a.M + SG_BM - 1) / SG_BM);
  gemm_f32_kernel<EPI><<<grid, SG_THREADS, 0, st>>>(a);
  CK(cudaGetLastError());
  e->launches++;
}

void launch_boundary(wg_engine* e, const BoundaryArgs& a, cudaStream_t st) {
  flow_boundary_kernel<<<(a.M + FB_ROWS - 1) / FB_ROWS, FB_THREADS, 0, st>>>(a);
  CK(cudaGetLastError());
  e->launches++;
}

// The launch sequence of WaveGlow.infer. stop_flow/stop_layer >= -1 make it a debug prefix run.
void run_infer(wg_engine* e, const float* mel, const float* z, float sigma, int deterministic, int B,
               int T, float* out, void* workspace, size_t ws_bytes, cudaStream_t st, int stop_flow,
               int stop_layer, float* h_out, float* acc_out, const Ragged* rg = nullptr) {
  check_shape(e, B, T);
  if (!mel || (!out && stop_flow < 0)) fail(WG_ERR_INVALID, "mel/out must not be NULL");
  if (!deterministic && !z) fail(WG_ERR_INVALID, "z must be given unless deterministic");
  if (rg && e->cfg.mode == WG_MODE_FP32) fail(WG_ERR_INVALID, "internal: ragged geometry belongs to the phase-major layouts");
  const Ws w = carve(e, B, T, rg);
  if (!workspace || ws_bytes < w.total)
    fail(WG_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total, ws_bytes);
  if ((uintptr_t)workspace % 1024 != 0) fail(WG_ERR_WORKSPACE, "workspace must be 1024-byte aligned");
  DeviceGuard guard(e->device);

  e->launches = 0;
  char* base = static_cast<char*>(workspace);
  const wg_config& c = e->cfg;
  const int C = e->C, S = e->S, R = e->R, L = T * R, M = B * L;
  const bool bf16 = c.mode == WG_MODE_BF16, tf32 = c.mode == WG_MODE_TF32X3, ffma = c.mode == WG_MODE_FP32;
  const bool pm = rg ? true : use_pm(e, B, T);
  // internal row geometry of the bf16 buffers (fp32 mode: position-major)
  RowGeom geo = pm ? RowGeom{R, T, T + pm_gap(e), B} : RowGeom{1, L, L, B};
  if (rg) {
    // per-utterance lengths: tables in the workspace, written by kernels fed from the parameter bank (256 utterances
    // per launch) -- no host buffer outlives this call; the waveform tail beyond 256*T_b[b] is defined as zero
    geo.off = reinterpret_cast<int*>(static_cast<char*>(workspace) + w.g_off);
    geo.len = reinterpret_cast<int*>(static_cast<char*>(workspace) + w.g_len);
    geo.row_b = reinterpret_cast<int*>(static_cast<char*>(workspace) + w.g_rowb);
    geo.rpp = rg->rpp;
    for (int b0 = 0; b0 < B; b0 += 256) {
      GeomChunk gc{};
      gc.b0 = b0; gc.n = std::min(256, B - b0); gc.gap = pm_gap(e);
      std::memcpy(gc.off, rg->off.data() + b0, gc.n * sizeof(int));
      std::memcpy(gc.len, rg->len.data() + b0, gc.n * sizeof(int));
      ragged_geom_kernel<<<gc.n, 128, 0, st>>>(gc, const_cast<int*>(geo.off), const_cast<int*>(geo.len), const_cast<int*>(geo.row_b));
      CK(cudaGetLastError());
      e->launches++;
    }
    if (out) CK(cudaMemsetAsync(out, 0, (size_t)B * L * c.n_group * sizeof(float), st));
  }
  const int Mi = geo.rows();   // internal rows (== M unless phase-major: gap rows)
  float* h32 = ffma ? reinterpret_cast<float*>(base + w.h32) : nullptr;
  float* t3_hhi[2] = {reinterpret_cast<float*>(base + w.t3_hhi0), reinterpret_cast<float*>(base + w.t3_hhi1)};
  float* t3_hlo[2] = {reinterpret_cast<float*>(base + w.t3_hlo0), reinterpret_cast<float*>(base + w.t3_hlo1)};
  __nv_bfloat16* t3_hb[2] = {reinterpret_cast<__nv_bfloat16*>(base + w.t3_hb0), reinterpret_cast<__nv_bfloat16*>(base + w.t3_hb1)};
  const int t3_parts = tf32 ? 2 * C / 256 : 1;                       // partial fold accumulators (one per gate chunk)
  const size_t t3_acc_stride = (size_t)geo.rows() * 8;                // floats between two partials
  float* acc8 = reinterpret_cast<float*>(base + w.acc8);
  float* audio[2] = {reinterpret_cast<float*>(base + w.audio0), reinterpret_cast<float*>(base + w.audio1)};
  float* spect = reinterpret_cast<float*>(base + w.spect);
  float* acts = reinterpret_cast<float*>(base + w.acts);
  float* skip = reinterpret_cast<float*>(base + w.skip);
  __nv_bfloat16* spect16 = reinterpret_cast<__nv_bfloat16*>(base + w.spect16);
  __nv_bfloat16* h16[2] = {reinterpret_cast<__nv_bfloat16*>(base + w.h16a),
                           reinterpret_cast<__nv_bfloat16*>(base + w.h16b)};
  __nv_bfloat16* hlo = reinterpret_cast<__nv_bfloat16*>(base + w.hlo);
  __nv_bfloat16* aup16 = reinterpret_cast<__nv_bfloat16*>(base + w.aup16);
  __nv_bfloat16* acts16 = reinterpret_cast<__nv_bfloat16*>(base + w.acts16);
  __nv_bfloat16* a0 = reinterpret_cast<__nv_bfloat16*>(base + w.a0);
  // start fold: off for a debug prefix that wants h right after the start conv (it would not exist)
  const bool fold0 = bf16 && pm && e->W0 && e->fold0 && !(stop_flow >= 0 && stop_layer == -1);
  CUtensorMap m_acts512;
  const float* zz = deterministic ? nullptr : z;

  TcPlan plan;
  TcPairMaps pmaps;
  Tc512PairMaps pmaps512;
  Tf32Plan plan3;
  Tf32FlowState flow3;
  // ---- (1) upsample + trim + regroup: spect[B*L, S]  (waveglow_arch.py:245-253) ---------------
  if (tf32) {
    RowGeom geo1 = geo;
    geo1.R = 1;
    Tf32Buffers tb{};
    for (int i = 0; i < 2; ++i) { tb.h_hi[i] = t3_hhi[i]; tb.h_lo[i] = t3_hlo[i]; tb.h_b[i] = t3_hb[i]; }
    tb.aup_hi = reinterpret_cast<float*>(base + w.t3_chi); tb.aup_b = reinterpret_cast<__nv_bfloat16*>(base + w.t3_cb);
    tb.acts_hi = reinterpret_cast<float*>(base + w.t3_ahi); tb.acts_b = reinterpret_cast<__nv_bfloat16*>(base + w.t3_ab);
    tb.acc8 = acc8; tb.acc8_stride = t3_acc_stride;
    tf32_prepare(plan3, e->sm_count, C, R, e->Kup, c.n_mel_channels, c.n_flows * c.n_layers, geo.rows_per_phase(), geo1,
                 rg ? 0 : geo.Tp, geo.T, e->t3, tb, e->t3_max_pairs, e->pair_policy, e->t3_epi_warps);
    e->last_pair = plan3.pair ? 1 : 0;
    e->last_flow_kernel = 0;
    flow3 = e->t3_flow;
    flow3.sync = reinterpret_cast<unsigned int*>(base + w.t3_sync);
    if (e->t3_flow_policy != 0 && tf32_flow_fits(plan3, flow3.max_pairs, c.n_layers)) {
      CK(cudaMemsetAsync(flow3.sync, 0, T3F_SYNC_WORDS * sizeof(unsigned int), st));   // arrivals = 0; the barriers re-arm themselves afterwards
      e->launches++;
    }
    e->launches += tf32_upsample(plan3, mel, st);
  } else if (ffma) {
    GemmArgs g{};
    g.nseg = UPSAMPLE_K / HOP;
    for (int j = 0; j < g.nseg; ++j) g.seg[j] = ASeg{mel, c.n_mel_channels, c.n_mel_channels, -j};
    g.W = e->Wup; g.bias = e->bup;
    g.M = B * T; g.N = R * S; g.L = T;
    g.out0 = spect; g.ld0 = R * S;
    launch_gemm<EPI_STORE>(e, g, st);
  } else {
    tc_prepare(plan, e->sm_count, B, T, L, C, S, e->Kup, c.n_mel_channels, c.n_flows * c.n_layers,
               e->Wup16, R * S, e->W1, e->W2, aup16, spect16, h16[0], h16[1], hlo, pm, R, e->V,
               fold0 ? a0 : nullptr, e->W0, e->H0, c.n_flows, pm ? pm_gap(e) : 0, rg ? &geo : nullptr);
    if (const char* to = std::getenv("WG_TILE_ORDER")) plan.tile_order = to[0] != '0';
    plan.pdl = e->pdl;
    e->last_pair = 0;
    if (pm && C == 256 && use_pair(e, plan))
      tc_pair_prepare(pmaps, plan, c.n_flows * c.n_layers, c.n_flows, R, e->W1, e->W2, e->V, e->W0, e->H0, e->pair_max);
    if (pm && C == 512 && use_pair(e, plan)) {
      tc512_pair_prepare(pmaps512, plan, c.n_flows * c.n_layers, c.n_flows, R, e->W1, e->V, e->W0, e->pair_max);
      e->last_pair = 1;
    }
    if (C == 512) make_map_4d(&m_acts512, acts16, 1, pm ? (uint64_t)R : (uint64_t)B, plan.Trows, C, WL_BM);
    e->launches += tc_upsample(plan, mel, e->bup, st);
  }

  // ---- noise -> audio, start conv of the first flow (waveglow_arch.py:264-275, :108) ----------
  const int F = c.n_flows;
  int cur = 0, z_off = 0, hcur = 0;
  {
    BoundaryArgs a{};
    a.geo = geo;
    a.first = 1; a.z = zz; a.n_group = c.n_group; a.z_off = 0; a.n_inject = e->flows[F - 1].n_rem;
    a.sigma = sigma; a.audio_out = audio[cur]; a.M = Mi; a.C = C;
    a.Wstart = fold0 ? nullptr : e->flows[F - 1].Wstart; a.bstart = e->flows[F - 1].bstart;
    a.n_half_next = e->flows[F - 1].n_half; a.h32 = h32; a.h16 = bf16 ? h16[hcur] : nullptr; a.hlo = bf16 ? hlo : nullptr;
    if (tf32) { a.hf_hi = t3_hhi[hcur]; a.hf_lo = t3_hlo[hcur]; a.hf_b = t3_hb[hcur]; }
    a.acc_parts = t3_parts; a.acc_part_stride = t3_acc_stride;
    if (!ffma) { a.acc8_rearm = acc8; std::memcpy(a.acc8_init, e->flows[F - 1].bse8, sizeof a.acc8_init); }
    launch_boundary(e, a, st);
    z_off = a.n_inject;
  }
  auto launch_a0 = [&](const float* audio_rows, int n_half) {
    a0_build_kernel<<<(unsigned)(((size_t)Mi * 4 + 255) / 256), 256, 0, st>>>(audio_rows, a0, geo, n_half);
    CK(cudaGetLastError());
    e->launches++;
  };
  if (fold0) launch_a0(audio[cur], e->flows[F - 1].n_half);

  auto prof_mark = [&](void) {
    if (!e->profiling) return;
    if (e->ev_used == e->ev_pool.size()) {
      cudaEvent_t ev;
      CK(cudaEventCreate(&ev));
      e->ev_pool.push_back(ev);
    }
    CK(cudaEventRecord(e->ev_pool[e->ev_used++], st));
  };

  auto dump = [&](void) {
    if (h_out && ffma) CK(cudaMemcpyAsync(h_out, h32, (size_t)M * C * 4, cudaMemcpyDeviceToDevice, st));
    if (h_out && tf32) {
      const size_t n = (size_t)Mi * C;
      hilo_f32_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t3_hhi[hcur], t3_hlo[hcur], h_out, geo, C);
      CK(cudaGetLastError());
    }
    if (h_out && bf16) {
      const size_t n = (size_t)Mi * C;
      hilo_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h16[hcur], hlo, h_out, geo, C);
      CK(cudaGetLastError());
    }
    if (acc_out) {
      const size_t n = (size_t)Mi * 8;
      unpermute_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc8, acc_out, geo, 8, t3_parts, t3_acc_stride);
      CK(cudaGetLastError());
    }
  };

  bool flow_ok = tf32 && e->t3_flow_policy != 0;   // tf32x3: a flow as ONE persistent launch where the call is one wave
  for (int k = F - 1; k >= 0; --k) {
    const FlowW& fw = e->flows[k];
    if (k == stop_flow && stop_layer == -1) { dump(); return; }
    for (int i = 0; i < c.n_layers; ++i) {
      const LayerW& lw = e->layers[k * c.n_layers + i];
      const int d = 1 << i;
      const bool last = i == c.n_layers - 1;
      if (tf32 && i == 0 && flow_ok && !(k == stop_flow && stop_layer >= 0) &&
          tf32_flow_fits(plan3, flow3.max_pairs, c.n_layers)) {
        // single-wave call: the whole flow as ONE persistent launch (grid barriers instead of kernel boundaries)
        const float *b1s[T3F_MAX_LAYERS], *b2s[T3F_MAX_LAYERS], *wses[T3F_MAX_LAYERS];
        for (int j = 0; j < c.n_layers; ++j) {
          const LayerW& lj = e->layers[k * c.n_layers + j];
          b1s[j] = lj.b1_pm; b2s[j] = lj.b2; wses[j] = lj.wse_d;
        }
        prof_mark();
        const int nl = tf32_wn_flow(plan3, flow3, k * c.n_layers, c.n_layers, hcur, b1s, b2s, wses, st, e->timing);
        if (nl > 0) {
          if (e->profiling) e->ev_count.push_back(c.n_layers);
          e->launches += nl;
          prof_mark();
          if ((c.n_layers - 1) & 1) hcur ^= 1;
          e->last_flow_kernel = 1;
          break;
        }
        // the device would not take the cooperative launch (SMs held by another context, MPS limits ...): this and the
        // remaining flows run on the per-layer kernels -- same bits
        flow_ok = false;
        if (e->profiling) --e->ev_used;
      }
      if (tf32) {
        // one event pair per flow around its back-to-back layer launches (gate + residual kernel per layer)
        if (i == 0) {
          prof_mark();
          if (e->profiling) e->ev_count.push_back(0);
        }
        if (e->profiling) e->ev_count.back() += 1;
        e->launches += tf32_wn_layer(plan3, k * c.n_layers + i, d, last, hcur, lw.b1_pm, lw.b2, lw.wse_d, st, e->timing);
        if (last || (k == stop_flow && i == stop_layer)) prof_mark();
        if (!last) hcur ^= 1;
      } else if (ffma) {
        // in-conv (dilated k=3) + cond 1x1 as one K = 3C+S contraction, gate fused (:113-127, :19-24)
        GemmArgs g{};
        g.nseg = 4;
        g.seg[0] = ASeg{h32, C, C, -d};
        g.seg[1] = ASeg{h32, C, C, 0};
        g.seg[2] = ASeg{h32, C, C, d};
        g.seg[3] = ASeg{spect, S, S, 0};
        g.W = lw.Wcat; g.bias = lw.bcat; g.M = M; g.N = 2 * C; g.L = L;
        g.out0 = acts; g.ld0 = C;
        prof_mark();
        if (e->profiling) e->ev_count.push_back(1);
        launch_gemm<EPI_GATE>(e, g, st);
        prof_mark();
        // res/skip 1x1 + residual add + skip accumulation (:129-139)
        GemmArgs r{};
        r.nseg = 1;
        r.seg[0] = ASeg{acts, C, C, 0};
        r.W = lw.Wrs; r.bias = lw.brs; r.M = M; r.N = lw.rs_cols; r.L = L;
        r.out0 = h32; r.ld0 = C; r.out1 = skip; r.ld1 = C;
        r.res_cols = last ? 0 : C; r.skip_init = (i == 0);
        launch_gemm<EPI_RES_SKIP>(e, r, st);
      } else {
        // bf16: ONE event pair per flow around its n_layers back-to-back layer launches (an event pair per launch
        // costs ~3 % of the step it is meant to measure)
        if (i == 0) {
          prof_mark();
          if (e->profiling) e->ev_count.push_back(0);
        }
        if (e->profiling) e->ev_count.back() += 1;
        if (C == 512)
          e->launches += tc512_wn_layer(plan, m_acts512, k * c.n_layers + i, d, last, hcur, acc8, pm ? lw.b1_pm : lw.b1,
                                        lw.b2, lw.wse_p.data(), st, fold0 && i == 0, &pmaps512);
        else if (pmaps.ready && !(fold0 && i == 0 && e->pair_epi_warps == 8)) {   // FIRST layers: the single-CTA kernel is faster (348 vs 370 us)
          e->last_pair = 1;
          e->launches += tc_wn_layer_pair(plan, pmaps, k * c.n_layers + i, d, last, hcur, acc8, lw.b1_pm, lw.b2,
                                          lw.wse_p.data(), st, fold0 && i == 0, e->timing, e->pair_epi_warps);
        } else
          e->launches += tc_wn_layer(plan, k * c.n_layers + i, d, last, hcur, acc8, pm ? lw.b1_pm : lw.b1, lw.b2,
                                     lw.wse_p.data(), e->timing, e->dbg_flags, st, fold0 && i == 0);
        if (last || (k == stop_flow && i == stop_layer)) prof_mark();
        if (!last) hcur ^= 1;
      }
      if (k == stop_flow && i == stop_layer) {
        if (ffma && acc_out) {
          end_conv_kernel<<<(M * 32 + 255) / 256, 256, 0, st>>>(skip, fw.Wend8, fw.bend8, acc8, M, C);
          CK(cudaGetLastError());
        }
        dump();
        return;
      }
    }
    if (ffma) {
      end_conv_kernel<<<(int)(((size_t)M * 32 + 255) / 256), 256, 0, st>>>(skip, fw.Wend8, fw.bend8, acc8, M, C);
      CK(cudaGetLastError());
      e->launches++;
    }
    // coupling inverse + W^-1 + early re-injection + next start conv (:278-304)
    BoundaryArgs a{};
    a.geo = geo;
    a.first = 0; a.acc8 = acc8; a.audio_in = audio[cur]; a.z = zz; a.n_group = c.n_group;
    a.sigma = sigma; a.c_in = 2 * fw.n_half; a.M = Mi; a.C = C;
    a.acc_parts = t3_parts; a.acc_part_stride = t3_acc_stride;
    std::memcpy(a.winv, fw.winv, sizeof a.winv);
    const bool early = (k % c.n_early_every == 0) && k > 0;
    a.n_inject = early ? c.n_early_size : 0;
    a.z_off = z_off;
    if (early) z_off += c.n_early_size;
    if (k > 0) {
      a.audio_out = audio[cur ^ 1];
      a.Wstart = fold0 ? nullptr : e->flows[k - 1].Wstart; a.bstart = e->flows[k - 1].bstart;
      a.n_half_next = e->flows[k - 1].n_half; a.h32 = h32;
      hcur = 0;
      a.h16 = bf16 ? h16[hcur] : nullptr; a.hlo = bf16 ? hlo : nullptr;
      if (tf32) { a.hf_hi = t3_hhi[hcur]; a.hf_lo = t3_hlo[hcur]; a.hf_b = t3_hb[hcur]; }
      if (!ffma) { a.acc8_rearm = acc8; std::memcpy(a.acc8_init, e->flows[k - 1].bse8, sizeof a.acc8_init); }
    } else {
      a.audio_out = out;  // [B*L, 8] == [B, 8L]  (waveglow_arch.py:306)
      a.final_out = 1;
      a.Wstart = nullptr;
    }
    launch_boundary(e, a, st);
    cur ^= 1;
    if (fold0 && k > 0) launch_a0(audio[cur], e->flows[k - 1].n_half);
  }
}

void build_engine(wg_engine* e, const wg_config* cfg, const wg_tensor* tensors, int n_tensors, int device) {
  const wg_config& c = *cfg;
  e->cfg = c;
  e->device = device;
  if (c.mode != WG_MODE_FP32 && c.mode != WG_MODE_BF16 && c.mode != WG_MODE_TF32X3) fail(WG_ERR_INVALID, "unknown mode %d", c.mode);
  const bool tf32 = c.mode == WG_MODE_TF32X3;
  if (c.kernel_size != 3) fail(WG_ERR_UNSUPPORTED, "kernel_size must be 3 (got %d)", c.kernel_size);
  if (c.n_group < 2 || c.n_group > 8 || c.n_group % 2 || HOP % c.n_group)
    fail(WG_ERR_UNSUPPORTED, "n_group must be an even divisor of 256 that is <= 8 (got %d)", c.n_group);
  if (c.n_mel_channels <= 0 || c.n_mel_channels % 16)
    fail(WG_ERR_UNSUPPORTED, "n_mel_channels must be a positive multiple of 16 (got %d)", c.n_mel_channels);
  if (c.n_channels <= 0 || c.n_channels % 16)
    fail(WG_ERR_UNSUPPORTED, "n_channels must be a positive multiple of 16 (got %d)", c.n_channels);
  if (c.n_flows <= 0 || c.n_layers <= 0 || c.n_layers > 12 || c.n_early_every <= 0 || c.n_early_size < 0 ||
      c.n_early_size % 2)
    fail(WG_ERR_INVALID, "bad flow/layer hparams");
  if (c.mode == WG_MODE_BF16 && c.n_channels != 256 && c.n_channels != 512)
    fail(WG_ERR_UNSUPPORTED, "WG_MODE_BF16 supports n_channels 256 and 512 (got %d); use WG_MODE_FP32",
         c.n_channels);
  if (c.mode == WG_MODE_BF16 && c.n_mel_channels * c.n_group != 640)
    fail(WG_ERR_UNSUPPORTED, "WG_MODE_BF16 requires n_mel_channels * n_group == 640 (got %d)", c.n_mel_channels * c.n_group);
  if (c.mode == WG_MODE_BF16 && (c.n_mel_channels * c.n_group) % 64)
    fail(WG_ERR_UNSUPPORTED, "WG_MODE_BF16 requires n_mel_channels*n_group %% 64 == 0");

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    fail(WG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
         ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
  if (device < 0 || device >= ndev) fail(WG_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
  DeviceGuard guard(device);
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  e->sm_count = prop.multiProcessorCount;
  if (c.mode != WG_MODE_FP32 && prop.major != 10)
    fail(WG_ERR_UNSUPPORTED, "WG_MODE_BF16 / WG_MODE_TF32X3 need an sm_100 GPU (tcgen05/TMEM); device is sm_%d%d", prop.major,
         prop.minor);
  if (tf32 && (c.n_channels % 128 || (UPSAMPLE_K / HOP * c.n_mel_channels) % 32))
    fail(WG_ERR_UNSUPPORTED, "WG_MODE_TF32X3 needs n_channels %% 128 == 0 and 4*n_mel_channels %% 32 == 0 (got %d, %d)",
         c.n_channels, c.n_mel_channels);

  const int C = c.n_channels, NL = c.n_layers, F = c.n_flows;
  e->C = C;
  e->S = c.n_mel_channels * c.n_group;
  e->R = HOP / c.n_group;
  const int S = e->S, R = e->R, NM = c.n_mel_channels, G = c.n_group;

  std::vector<float> wup_f32, ub_vec;       // host copies used while folding the conditioning weights
  std::vector<std::vector<float>> wcond_packed;   // per layer [S, 2C] in chunk-packed column order (bf16 mode)
  std::map<std::string, TensorView> tm;
  for (int i = 0; i < n_tensors; ++i) {
    const wg_tensor& t = tensors[i];
    if (!t.name || !t.data || t.ndim < 1 || t.ndim > 4) fail(WG_ERR_WEIGHTS, "bad tensor entry %d", i);
    TensorView v{t.data, std::vector<int64_t>(t.shape, t.shape + t.ndim)};
    tm[t.name] = v;
  }

  // flow schedule (waveglow_arch.py:202-223)
  e->flows.resize(F);
  {
    int n_half = G / 2, n_rem = G;
    for (int k = 0; k < F; ++k) {
      if (k % c.n_early_every == 0 && k > 0) {
        n_half -= c.n_early_size / 2;
        n_rem -= c.n_early_size;
      }
      if (n_half <= 0) fail(WG_ERR_INVALID, "flow schedule exhausts the channels at flow %d", k);
      e->flows[k].n_half = n_half;
      e->flows[k].n_rem = n_rem;
    }
  }

  // ---- upsample: polyphase packing -----------------------------------------------------------
  // spect[b, R t + r, m*G + g] = bias[m] + sum_{j,i} mel[b, t-j, i] * Wup[256 j + G r + g, m, i]
  {
    const TensorView& uk = need(tm, "upsample/kernel", {UPSAMPLE_K, NM, NM});
    const TensorView& ub = need(tm, "upsample/bias", {NM});
    const int J = UPSAMPLE_K / HOP, N = R * S;
    std::vector<float> bp((size_t)N);
    for (int r = 0; r < R; ++r)
      for (int m = 0; m < NM; ++m)
        for (int g = 0; g < G; ++g) bp[(size_t)r * S + m * G + g] = ub.data[m];
    e->bup = upload(e, bp);
    // fp32 polyphase matrix [4*n_mel, R*S]: the upsample operand of fp32 mode, and the left factor of the
    // folded conditioning weights (Wup_r @ Wcond) of the phase-major bf16 path
    wup_f32.assign((size_t)J * NM * N, 0.f);
    for (int j = 0; j < J; ++j)
      for (int i = 0; i < NM; ++i)
        for (int r = 0; r < R; ++r)
          for (int m = 0; m < NM; ++m)
            for (int g = 0; g < G; ++g)
              wup_f32[((size_t)(j * NM + i)) * N + (size_t)r * S + m * G + g] =
                  uk.data[((size_t)(HOP * j + G * r + g) * NM + m) * NM + i];
    ub_vec.assign(S, 0.f);
    for (int m = 0; m < NM; ++m)
      for (int g = 0; g < G; ++g) ub_vec[m * G + g] = ub.data[m];
    if (c.mode == WG_MODE_FP32) {
      e->Wup = upload(e, wup_f32);
    } else if (tf32) {
      e->Kup = J * NM;      // 320: ten K-blocks of 32 floats
    } else {
      e->Kup = (int)align_up((size_t)J * NM, 64);
      std::vector<__nv_bfloat16> wp((size_t)N * e->Kup, f2bf(0.f));
      for (int j = 0; j < J; ++j)
        for (int i = 0; i < NM; ++i)
          for (int r = 0; r < R; ++r)
            for (int m = 0; m < NM; ++m)
              for (int g = 0; g < G; ++g)
                wp[((size_t)r * S + m * G + g) * e->Kup + (j * NM + i)] =
                    f2bf(uk.data[((size_t)(HOP * j + G * r + g) * NM + m) * NM + i]);
      e->Wup16 = upload(e, wp);
    }
  }

  // ---- flows / layers ------------------------------------------------------------------------
  e->layers.resize((size_t)F * NL);
  const int K1 = 3 * C + S;
  std::vector<__nv_bfloat16> w1all, w2all, w0all, h0all;
  std::vector<float> t3w1h, t3w2h;                    // tf32x3: tf32(w) in fp32 words, stacked over the layers
  std::vector<__nv_bfloat16> t3w1b, t3w2b;            // ... and the bf16 companions [.., 2K] = bf16(hi) | bf16(lo)
  if (tf32) {
    t3w1h.assign((size_t)F * NL * 2 * C * 3 * C, 0.f); t3w1b.assign(2 * t3w1h.size(), f2bf(0.f));
    t3w2h.assign((size_t)F * NL * C * C, 0.f); t3w2b.assign(2 * t3w2h.size(), f2bf(0.f));
  }
  const bool build_fold0 = c.mode == WG_MODE_BF16 && NL > 1;
  if (c.mode == WG_MODE_BF16) {
    w1all.assign((size_t)F * NL * 2 * C * K1, f2bf(0.f));
    w2all.assign((size_t)F * NL * C * C, f2bf(0.f));
    if (build_fold0) {
      w0all.assign((size_t)F * 2 * C * 64, f2bf(0.f));
      h0all.assign((size_t)F * C * 64, f2bf(0.f));
    }
  }
  for (int k = 0; k < F; ++k) {
    FlowW& fw = e->flows[k];
    const std::string p = "block-" + std::to_string(k) + "/";
    const int nh = fw.n_half, nr = fw.n_rem;
    {
      const TensorView& ik = need(tm, "invertible_conv-" + std::to_string(k) + "/conv/kernel", {1, nr, nr});
      // W[o,i] = kernel[0,i,o]; reverse: out[b] = sum_a in[a] * inv(W)[b,a]  (invertible_conv.py:41-51)
      float W[64], Winv[64];
      for (int o = 0; o < nr; ++o)
        for (int i = 0; i < nr; ++i) W[o * nr + i] = ik.data[i * nr + o];
      invert_f32(W, nr, Winv);
      std::memset(fw.winv, 0, sizeof fw.winv);
      for (int a = 0; a < nr; ++a)
        for (int b = 0; b < nr; ++b) fw.winv[a * nr + b] = Winv[b * nr + a];
    }
    const TensorView& sk = need(tm, p + "start_conv/kernel", {1, nh, C});
    const TensorView& sb = need(tm, p + "start_conv/bias", {C});
    fw.Wstart = upload(e, std::vector<float>(sk.data, sk.data + (size_t)nh * C));
    fw.bstart = upload(e, std::vector<float>(sb.data, sb.data + C));
    const TensorView& ek = need(tm, p + "end_conv/kernel", {1, C, 2 * nh});
    const TensorView& eb = need(tm, p + "end_conv/bias", {2 * nh});
    std::vector<float> wend8((size_t)C * 8, 0.f), bend8(8, 0.f);
    for (int cc = 0; cc < C; ++cc)
      for (int j = 0; j < 2 * nh; ++j) wend8[(size_t)cc * 8 + j] = ek.data[(size_t)cc * 2 * nh + j];
    for (int j = 0; j < 2 * nh; ++j) bend8[j] = eb.data[j];
    if (c.mode == WG_MODE_FP32) {
      fw.Wend8 = upload(e, wend8);
      fw.bend8 = upload(e, bend8);
    }
    std::vector<double> bse(8, 0.0);
    for (int j = 0; j < 8; ++j) bse[j] = bend8[j];

    for (int i = 0; i < NL; ++i) {
      LayerW& lw = e->layers[(size_t)k * NL + i];
      const std::string si = std::to_string(i);
      const int rs = (i < NL - 1) ? 2 * C : C;
      lw.rs_cols = rs;
      const TensorView& inw = need(tm, p + "in_conv-" + si + "/kernel", {3, C, 2 * C});
      const TensorView& inb = need(tm, p + "in_conv-" + si + "/bias", {2 * C});
      const TensorView& cw = need(tm, p + "cond_layer-" + si + "/kernel", {1, S, 2 * C});
      const TensorView& cb = need(tm, p + "cond_layer-" + si + "/bias", {2 * C});
      const TensorView& rw = need(tm, p + "res_skip_conv-" + si + "/kernel", {1, C, rs});
      const TensorView& rb = need(tm, p + "res_skip_conv-" + si + "/bias", {rs});
      auto wsrc = [&](int kk, int col) -> float {  // row kk of the [3C+S, 2C] concatenated operand
        return kk < 3 * C ? inw.data[((size_t)kk) * 2 * C + col]           // [3,C,2C] flat = [(j*C+ci), col]
                          : cw.data[((size_t)(kk - 3 * C)) * 2 * C + col];
      };
      if (c.mode == WG_MODE_FP32) {
        // packed column p: group of 8 = [tanh 4c..4c+3 | sigmoid 4c..4c+3]
        std::vector<float> wcat((size_t)K1 * 2 * C), bcat((size_t)2 * C);
        for (int pcol = 0; pcol < 2 * C; ++pcol) {
          const int g4 = pcol >> 3, q = pcol & 7;
          const int col = q < 4 ? 4 * g4 + q : C + 4 * g4 + (q - 4);
          bcat[pcol] = inb.data[col] + cb.data[col];
          for (int kk = 0; kk < K1; ++kk) wcat[(size_t)kk * 2 * C + pcol] = wsrc(kk, col);
        }
        lw.Wcat = upload(e, wcat);
        lw.bcat = upload(e, bcat);
        lw.Wrs = upload(e, std::vector<float>(rw.data, rw.data + (size_t)C * rs));
        lw.brs = upload(e, std::vector<float>(rb.data, rb.data + rs));
      } else {
        // chunk packing: 256-column chunk q = [tanh 128q..128q+127 | sigmoid 128q..128q+127]
        std::vector<float> b1((size_t)2 * C), b2((size_t)C, 0.f);
        lw.wse_h.assign((size_t)C * 8, 0.f);
        std::vector<float>& wse = lw.wse_h;
        __nv_bfloat16* w1 = tf32 ? nullptr : w1all.data() + ((size_t)k * NL + i) * 2 * C * K1;
        for (int pcol = 0; pcol < 2 * C; ++pcol) {
          const int chunk = pcol >> 8, wi = pcol & 255;
          const int col = wi < 128 ? 128 * chunk + wi : C + 128 * chunk + (wi - 128);
          b1[pcol] = inb.data[col] + cb.data[col];
          if (tf32) {
            const size_t row = (((size_t)k * NL + i) * 2 * C + pcol) * 3 * C;
            for (int kk = 0; kk < 3 * C; ++kk) {
              const float v = wsrc(kk, col), hi = tf32_rna_host(v);
              t3w1h[row + kk] = hi;
              const size_t bp = 2 * row + 64 * (size_t)(kk >> 5) + (kk & 31);   // B operand: [lb | hb] per K-block of 32
              t3w1b[bp] = f2bf(v - hi);
              t3w1b[bp + 32] = f2bf(hi);
            }
          } else {
            for (int kk = 0; kk < K1; ++kk) w1[(size_t)pcol * K1 + kk] = f2bf(wsrc(kk, col));
          }
        }
        if (build_fold0 && i == 0) {
          // start fold (a0_build_kernel): 16 K columns per tap = [G hi (4) | G hi (4) | G lo (4) | g hi | g lo | 0 0] with
          // G = Wstart @ Win[tap] ([nh, 2C]) and g = bstart @ Win[tap], folded in double and split into bf16 hi + lo;
          // H0 carries [Wstart; bstart] the same way in columns 48..63 (the residual operand h0 itself).
          auto split = [](double v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
            hi = f2bf((float)v);
            lo = f2bf((float)(v - (double)__bfloat162float(hi)));
          };
          __nv_bfloat16* w0 = w0all.data() + (size_t)k * 2 * C * 64;
          for (int pcol = 0; pcol < 2 * C; ++pcol) {
            const int chunk = pcol >> 8, wi = pcol & 255;
            const int col = wi < 128 ? 128 * chunk + wi : C + 128 * chunk + (wi - 128);
            for (int tap = 0; tap < 3; ++tap) {
              __nv_bfloat16* dst = w0 + (size_t)pcol * 64 + tap * 16;
              for (int j = 0; j <= nh; ++j) {       // j == nh: the bias row
                double acc = 0.0;
                for (int ci = 0; ci < C; ++ci) {
                  const double left = j < nh ? (double)sk.data[(size_t)j * C + ci] : (double)sb.data[ci];
                  acc += left * (double)inw.data[((size_t)tap * C + ci) * 2 * C + col];
                }
                __nv_bfloat16 hi, lo;
                split(acc, hi, lo);
                if (j < nh) { dst[j] = hi; dst[4 + j] = hi; dst[8 + j] = lo; }
                else { dst[12] = hi; dst[13] = lo; }
              }
            }
          }
          __nv_bfloat16* h0 = h0all.data() + (size_t)k * C * 64;
          for (int n = 0; n < C; ++n) {
            __nv_bfloat16* dst = h0 + (size_t)n * 64 + 48;
            for (int j = 0; j <= nh; ++j) {
              const double v = j < nh ? (double)sk.data[(size_t)j * C + n] : (double)sb.data[n];
              __nv_bfloat16 hi, lo;
              split(v, hi, lo);
              if (j < nh) { dst[j] = hi; dst[4 + j] = hi; dst[8 + j] = lo; }
              else { dst[12] = hi; dst[13] = lo; }
            }
          }
        }
        {   // phase-major path: fp32 cond weights in packed column order + bias with the upsample bias folded in
          std::vector<float> wc((size_t)S * 2 * C), b1pm((size_t)2 * C);
          for (int pcol = 0; pcol < 2 * C; ++pcol) {
            const int chunk = pcol >> 8, wi = pcol & 255;
            const int col = wi < 128 ? 128 * chunk + wi : C + 128 * chunk + (wi - 128);
            double acc = b1[pcol];
            for (int sidx = 0; sidx < S; ++sidx) {
              const float wv = cw.data[(size_t)sidx * 2 * C + col];
              wc[(size_t)sidx * 2 * C + pcol] = wv;
              acc += (double)ub_vec[sidx] * (double)wv;
            }
            b1pm[pcol] = (float)acc;
          }
          wcond_packed.push_back(std::move(wc));
          lw.b1_pm = upload(e, b1pm);
        }
        const int skip_off = (i < NL - 1) ? C : 0;
        if (i < NL - 1) {
          __nv_bfloat16* w2 = tf32 ? nullptr : w2all.data() + ((size_t)k * NL + i) * C * C;
          for (int n = 0; n < C; ++n) {
            b2[n] = rb.data[n];
            for (int kk = 0; kk < C; ++kk) {
              const float v = rw.data[(size_t)kk * rs + n];
              if (tf32) {
                const size_t at = (((size_t)k * NL + i) * C + n) * C + kk;
                const size_t rowb = (((size_t)k * NL + i) * C + n) * 2 * C;
                t3w2h[at] = tf32_rna_host(v);
                const size_t bp = rowb + 64 * (size_t)(kk >> 5) + (kk & 31);
                t3w2b[bp] = f2bf(v - t3w2h[at]);
                t3w2b[bp + 32] = f2bf(t3w2h[at]);
              } else {
                w2[(size_t)n * C + kk] = f2bf(v);
              }
            }
          }
        }
        // skip o end fold: acc8 += acts @ (Wskip @ Wend);  bias folded into bse8
        for (int kk = 0; kk < C; ++kk)
          for (int j = 0; j < 2 * nh; ++j) {
            double s = 0.0;
            for (int n = 0; n < C; ++n)
              s += (double)rw.data[(size_t)kk * rs + skip_off + n] * (double)wend8[(size_t)n * 8 + j];
            wse[(size_t)kk * 8 + j] = (float)s;
          }
        for (int j = 0; j < 2 * nh; ++j) {
          double s = 0.0;
          for (int n = 0; n < C; ++n) s += (double)rb.data[skip_off + n] * (double)wend8[(size_t)n * 8 + j];
          bse[j] += s;
        }
        lw.wse_p.assign((size_t)C * 8, 0.f);
        for (int ch = 0; ch < C; ++ch)
          for (int cc = 0; cc < 8; ++cc) lw.wse_p[((size_t)(ch >> 1) * 8 + cc) * 2 + (ch & 1)] = wse[(size_t)ch * 8 + cc];
        lw.b1 = upload(e, b1);
        lw.b2 = upload(e, b2);
        if (tf32) lw.wse_d = upload(e, lw.wse_h);
      }
    }
    if (c.mode != WG_MODE_FP32) {
      for (int j = 0; j < 8; ++j) fw.bse8[j] = (float)bse[j];
    }
  }
  if (tf32) {
    e->t3.W1h = upload(e, t3w1h); e->t3.W1b = upload(e, t3w1b);
    e->t3.W2h = upload(e, t3w2h); e->t3.W2b = upload(e, t3w2b);
    tc_init();
    e->t3_max_pairs = std::min(tf32_init(), e->sm_count / 2);
    if (const char* pr = std::getenv("WG_PAIR")) e->pair_policy = std::atoi(pr);
    e->t3_flow.max_pairs = std::min(tf32_flow_init(), e->sm_count / 2);
    if (const char* fl = std::getenv("WG_TF32_FLOW")) e->t3_flow_policy = std::atoi(fl);
    e->t3_flow.refuse = e->t3_flow_policy == 2;
    // Nsight Compute refuses a cooperative launch of a cluster kernel ("LaunchFailed", which ends the profiled process):
    // under its injection the flow kernel is launched plainly -- ncu serialises kernels, so co-residency holds anyway.
    if (std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || std::getenv("NV_NSIGHT_INJECTION_PORT_BASE")) e->t3_flow.cooperative = false;
    if (const char* co = std::getenv("WG_TF32_COOP")) e->t3_flow.cooperative = co[0] != '0';
    if (const char* ew = std::getenv("WG_TF32_EPI")) e->t3_epi_warps = std::atoi(ew) == 16 ? 16 : (std::atoi(ew) == 8 ? 8 : 0);
    // folded conditioning weights as an fp32 (hi, lo) pair: V[(layer*R + r)*2C + n][k] = sum_s Wup_r[k][s] * Wcond[s][n]
    const int Kw = e->Kup;
    float *d_wup = nullptr, *d_wc = nullptr, *d_tmp = nullptr, *vh = nullptr;
    __nv_bfloat16* vb = nullptr;
    const size_t vcount = (size_t)F * NL * R * 2 * C * Kw;
    std::vector<float> wup_rk((size_t)R * Kw * S);
    for (int r = 0; r < R; ++r)
      for (int kk = 0; kk < Kw; ++kk)
        std::memcpy(&wup_rk[((size_t)r * Kw + kk) * S], &wup_f32[(size_t)kk * R * S + (size_t)r * S], (size_t)S * 4);
    CK(cudaMalloc(&d_wup, wup_rk.size() * 4));
    CK(cudaMalloc(&d_wc, (size_t)S * 2 * C * 4));
    CK(cudaMalloc(&d_tmp, (size_t)R * Kw * 2 * C * 4));
    CK(cudaMalloc(&vh, vcount * 4));
    e->allocs.push_back(vh);
    CK(cudaMalloc(&vb, vcount * 4));
    e->allocs.push_back(vb);
    CK(cudaMemcpy(d_wup, wup_rk.data(), wup_rk.size() * 4, cudaMemcpyHostToDevice));
    for (int li = 0; li < F * NL; ++li) {
      CK(cudaMemcpy(d_wc, wcond_packed[li].data(), (size_t)S * 2 * C * 4, cudaMemcpyHostToDevice));
      GemmArgs g{};
      g.nseg = 1;
      g.seg[0] = ASeg{d_wup, S, S, 0};
      g.W = d_wc; g.bias = nullptr; g.M = R * Kw; g.N = 2 * C; g.L = R * Kw;
      g.out0 = d_tmp; g.ld0 = 2 * C;
      dim3 grid((g.N + SG_BN - 1) / SG_BN, (g.M + SG_BM - 1) / SG_BM);
      gemm_f32_kernel<EPI_STORE><<<grid, SG_THREADS>>>(g);
      const size_t n = (size_t)R * Kw * 2 * C;
      const size_t at = (size_t)li * R * 2 * C * Kw;
      fold_store_tf32_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_tmp, vh + at, vb + 2 * at, R, Kw, 2 * C);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaFree(d_wup); cudaFree(d_wc); cudaFree(d_tmp);
    e->t3.Vh = vh; e->t3.Vb = vb;
  }
  if (c.mode == WG_MODE_BF16) {
    e->W1 = upload(e, w1all);
    e->W2 = upload(e, w2all);
    if (build_fold0) {
      e->W0 = upload(e, w0all);
      e->H0 = upload(e, h0all);
      if (const char* f0 = std::getenv("WG_FOLD0")) e->fold0 = f0[0] != '0';
    }
    tc_init();
    {   // V[(layer*R + r)*2C + n][k] = sum_s Wup_r[k][s] * Wcond_layer[s][n]   (k < 4*n_mel, padded to Kup)
      const int Kw = (UPSAMPLE_K / HOP) * NM;       // 320
      if (Kw % 16 == 0 && Kw == e->Kup) {
        // One fp32 GEMM per layer over ALL phases: rows (r, k) of the phase-stacked upsample matrix against the layer's
        // conditioning weights, [R*Kw, S] @ [S, 2C], then one transposing bf16 store into V (2 launches per layer).
        float* d_wup = nullptr; float* d_wc = nullptr; float* d_tmp = nullptr;
        const size_t vcount = (size_t)F * NL * R * 2 * C * e->Kup;
        std::vector<float> wup_rk((size_t)R * Kw * S);
        for (int r = 0; r < R; ++r)
          for (int k = 0; k < Kw; ++k)
            std::memcpy(&wup_rk[((size_t)r * Kw + k) * S], &wup_f32[(size_t)k * R * S + (size_t)r * S], (size_t)S * 4);
        CK(cudaMalloc(&d_wup, wup_rk.size() * 4));
        CK(cudaMalloc(&d_wc, (size_t)S * 2 * C * 4));
        CK(cudaMalloc(&d_tmp, (size_t)R * Kw * 2 * C * 4));
        CK(cudaMalloc(&e->V, vcount * 2));
        e->allocs.push_back(e->V);
        CK(cudaMemcpy(d_wup, wup_rk.data(), wup_rk.size() * 4, cudaMemcpyHostToDevice));
        for (int li = 0; li < F * NL; ++li) {
          CK(cudaMemcpy(d_wc, wcond_packed[li].data(), (size_t)S * 2 * C * 4, cudaMemcpyHostToDevice));
          GemmArgs g{};
          g.nseg = 1;
          g.seg[0] = ASeg{d_wup, S, S, 0};
          g.W = d_wc; g.bias = nullptr; g.M = R * Kw; g.N = 2 * C; g.L = R * Kw;
          g.out0 = d_tmp; g.ld0 = 2 * C;
          dim3 grid((g.N + SG_BN - 1) / SG_BN, (g.M + SG_BM - 1) / SG_BM);
          gemm_f32_kernel<EPI_STORE><<<grid, SG_THREADS>>>(g);
          const size_t n = (size_t)R * Kw * 2 * C;
          fold_store_bf16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_tmp, e->V + (size_t)li * R * 2 * C * e->Kup, R, Kw, 2 * C);
        }
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        cudaFree(d_wup); cudaFree(d_wc); cudaFree(d_tmp);
      }
      if (const char* pmv = std::getenv("WG_PM")) e->pm_policy = std::atoi(pmv);
    }
    tc512_init();
    e->pair_max = std::min(tc_pair_init(), e->sm_count / 2);
    if (const char* pr = std::getenv("WG_PAIR")) e->pair_policy = std::atoi(pr);
    if (const char* pw = std::getenv("WG_PAIR_EPI")) e->pair_epi_warps = std::atoi(pw) == 16 ? 16 : 8;
    if (const char* pd = std::getenv("WG_PDL")) e->pdl = pd[0] != '0';
#ifdef WG_PROBES
    // A/B probes that deliberately BREAK the result to isolate a cost (profiles/r01_probes.md): compiled only into a
    // -DWG_PROBES build, and even there honoured only when the caller also sets WG_ALLOW_PROBES=1.
    if (const char* f = std::getenv("WG_DEBUG_FLAGS"))
      if (const char* ok = std::getenv("WG_ALLOW_PROBES")) e->dbg_flags = ok[0] == '1' ? std::atoi(f) : 0;
#endif
  }
  if (const char* t = std::getenv("WG_LAYER_TIMING")) {
    if (t[0] == '1' && c.mode != WG_MODE_FP32) {
      CK(cudaMalloc(&e->timing, 128 * sizeof(unsigned long long)));
      e->allocs.push_back(e->timing);
      CK(cudaMemset(e->timing, 0, 128 * sizeof(unsigned long long)));
    }
  }
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
}

void destroy_engine(wg_engine* e) {
  if (!e) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(e->device);
  for (void* p : e->allocs) cudaFree(p);
  if (e->pin_mel) cudaFreeHost(e->pin_mel);
  if (e->pin_z) cudaFreeHost(e->pin_z);
  if (e->pin_out) cudaFreeHost(e->pin_out);
  if (e->dev_mel) cudaFree(e->dev_mel);
  if (e->dev_z) cudaFree(e->dev_z);
  if (e->dev_out) cudaFree(e->dev_out);
  if (e->dev_ws) cudaFree(e->dev_ws);
  if (e->stream) cudaStreamDestroy(e->stream);
  for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
  if (prev >= 0) cudaSetDevice(prev);
  delete e;
}

template <typename F>
int guarded(wg_handle h, F&& f) {
  try {
    f();
    return WG_OK;
  } catch (const Fail& x) {
    if (h) h->err = x.msg;
    else {
      std::lock_guard<std::mutex> g(g_err_mu);
      g_create_err = x.msg;
    }
    return x.code;
  } catch (const std::exception& x) {
    if (h) h->err = x.what();
    return WG_ERR_INVALID;
  }
}

// grows a pinned-host / device buffer pair together (tensorrt_runtime.py:143-177 re-allocates its
// per-shape buffers the same way when the input shapes change)
void ensure_pair(float*& pin, float*& dev, size_t& cap, size_t bytes) {
  if (cap >= bytes && pin && dev) return;
  if (pin) cudaFreeHost(pin);
  if (dev) cudaFree(dev);
  pin = nullptr; dev = nullptr; cap = 0;
  CK(cudaMallocHost(reinterpret_cast<void**>(&pin), bytes));
  CK(cudaMalloc(reinterpret_cast<void**>(&dev), bytes));
  cap = bytes;
}

// WaveGlow.infer on a batch whose utterances have their own frame counts T_b[b] <= T (caller buffers keep the padded
// [B, T, ..] shapes). Every utterance is computed exactly as if it had been passed alone -- no padding frame enters
// any convolution (models/tts/tacotron2.py:183-191 vocodes one trimmed mel at a time).
void infer_ragged(wg_engine* e, const float* mel, const float* z, float sigma, int deterministic, int B, int T,
                  const int32_t* T_b, float* out, void* workspace, size_t ws_bytes, cudaStream_t st) {
  const Ragged rg = make_ragged(e, B, T, T_b);
  if (!mel || !out) fail(WG_ERR_INVALID, "mel/out must not be NULL");
  if (!deterministic && !z) fail(WG_ERR_INVALID, "z must be given unless deterministic");
  if (is_uniform(rg, T)) {
    run_infer(e, mel, z, sigma, deterministic, B, T, out, workspace, ws_bytes, st, -2, -2, nullptr, nullptr);
    return;
  }
  if (e->cfg.mode != WG_MODE_FP32) {
    run_infer(e, mel, z, sigma, deterministic, B, T, out, workspace, ws_bytes, st, -2, -2, nullptr, nullptr, &rg);
    return;
  }
  // FP32 (position-major FFMA) mode: one utterance after the other on the same stream and scratch
  DeviceGuard guard(e->device);
  const size_t Lg = (size_t)T * e->R, G = e->cfg.n_group;
  CK(cudaMemsetAsync(out, 0, (size_t)B * Lg * G * sizeof(float), st));
  int launches = 0;
  for (int b = 0; b < B; ++b) {
    run_infer(e, mel + (size_t)b * T * e->cfg.n_mel_channels, z ? z + (size_t)b * Lg * G : nullptr, sigma, deterministic,
              1, rg.len[b], out + (size_t)b * Lg * G, workspace, ws_bytes, st, -2, -2, nullptr, nullptr);
    launches += e->launches;
  }
  e->launches = launches;
}

void infer_host(wg_engine* h, const float* mel_host, const float* z_host, float sigma, int deterministic, int B, int T,
                const int32_t* T_b, float* out_host) {
  check_shape(h, B, T);
  if (!mel_host || !out_host) fail(WG_ERR_INVALID, "mel_host/out_host must not be NULL");
  if (!deterministic && !z_host) fail(WG_ERR_INVALID, "z_host must be given unless deterministic");
  DeviceGuard guard(h->device);
  size_t ws_b = 0;
  if (T_b) {
    const Ragged rg = make_ragged(h, B, T, T_b);
    int tmax = 0;
    for (int l : rg.len) tmax = std::max(tmax, l);
    ws_b = h->cfg.mode == WG_MODE_FP32 ? carve(h, 1, tmax).total : is_uniform(rg, T) ? carve(h, B, T).total : carve(h, B, T, &rg).total;
  } else {
    ws_b = carve(h, B, T).total;
  }
  const size_t L = (size_t)T * h->R;
  const size_t mel_b = (size_t)B * T * h->cfg.n_mel_channels * 4, z_b = (size_t)B * L * h->cfg.n_group * 4,
               out_b = (size_t)B * L * h->cfg.n_group * 4;
  ensure_pair(h->pin_mel, h->dev_mel, h->cap_mel, mel_b);
  ensure_pair(h->pin_out, h->dev_out, h->cap_out, out_b);
  if (!deterministic) ensure_pair(h->pin_z, h->dev_z, h->cap_z, z_b);
  if (h->cap_ws < ws_b) {
    if (h->dev_ws) cudaFree(h->dev_ws);
    h->dev_ws = nullptr; h->cap_ws = 0;
    CK(cudaMalloc(&h->dev_ws, ws_b));
    h->cap_ws = ws_b;
  }
  std::memcpy(h->pin_mel, mel_host, mel_b);
  CK(cudaMemcpyAsync(h->dev_mel, h->pin_mel, mel_b, cudaMemcpyHostToDevice, h->stream));
  if (!deterministic) {
    std::memcpy(h->pin_z, z_host, z_b);
    CK(cudaMemcpyAsync(h->dev_z, h->pin_z, z_b, cudaMemcpyHostToDevice, h->stream));
  }
  if (T_b)
    infer_ragged(h, h->dev_mel, deterministic ? nullptr : h->dev_z, sigma, deterministic, B, T, T_b, h->dev_out, h->dev_ws,
                 h->cap_ws, h->stream);
  else
    run_infer(h, h->dev_mel, deterministic ? nullptr : h->dev_z, sigma, deterministic, B, T, h->dev_out, h->dev_ws,
              h->cap_ws, h->stream, -2, -2, nullptr, nullptr);
  CK(cudaMemcpyAsync(h->pin_out, h->dev_out, out_b, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::memcpy(out_host, h->pin_out, out_b);
}

}  // namespace

extern "C" {

int wg_abi_version(void) { return WG_ABI_VERSION; }

int wg_create(const wg_config* cfg, const wg_tensor* tensors, int32_t n_tensors, int32_t device, wg_handle* out) {
  if (out) *out = nullptr;
  if (!cfg || !tensors || !out || n_tensors <= 0) {
    std::lock_guard<std::mutex> g(g_err_mu);
    g_create_err = "wg_create: NULL argument";
    return WG_ERR_INVALID;
  }
  wg_engine* e = new wg_engine();
  int rc = guarded(nullptr, [&] { build_engine(e, cfg, tensors, n_tensors, device); });
  if (rc != WG_OK) {
    destroy_engine(e);
    return rc;
  }
  *out = e;
  return WG_OK;
}

void wg_destroy(wg_handle h) { destroy_engine(h); }

const char* wg_last_error(wg_handle h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> g(g_err_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

int wg_workspace_bytes(wg_handle h, int32_t B, int32_t T, size_t* bytes) {
  if (!h || !bytes) return WG_ERR_INVALID;
  return guarded(h, [&] {
    check_shape(h, B, T);
    *bytes = carve(h, B, T).total;
  });
}

int wg_infer(wg_handle h, const float* mel, const float* z, float sigma, int32_t deterministic, int32_t B,
             int32_t T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] {
    run_infer(h, mel, z, sigma, deterministic, B, T, out, workspace, workspace_bytes,
              static_cast<cudaStream_t>(stream), -2, -2, nullptr, nullptr);
  });
}

int wg_workspace_bytes_ragged(wg_handle h, int32_t B, int32_t T, const int32_t* T_b, size_t* bytes) {
  if (!h || !bytes) return WG_ERR_INVALID;
  return guarded(h, [&] {
    const Ragged rg = make_ragged(h, B, T, T_b);
    if (h->cfg.mode == WG_MODE_FP32) {   // FP32 mode runs the utterances one after the other in the same scratch
      int tmax = 0;
      for (int l : rg.len) tmax = std::max(tmax, l);
      *bytes = carve(h, 1, tmax).total;
    } else {
      *bytes = is_uniform(rg, T) ? carve(h, B, T).total : carve(h, B, T, &rg).total;
    }
  });
}

int wg_infer_ragged(wg_handle h, const float* mel, const float* z, float sigma, int32_t deterministic, int32_t B,
                    int32_t T, const int32_t* T_b, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] {
    infer_ragged(h, mel, z, sigma, deterministic, B, T, T_b, out, workspace, workspace_bytes,
                 static_cast<cudaStream_t>(stream));
  });
}

int wg_infer_host(wg_handle h, const float* mel_host, const float* z_host, float sigma, int32_t deterministic,
                  int32_t B, int32_t T, float* out_host) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] { infer_host(h, mel_host, z_host, sigma, deterministic, B, T, nullptr, out_host); });
}

int wg_infer_host_ragged(wg_handle h, const float* mel_host, const float* z_host, float sigma, int32_t deterministic,
                         int32_t B, int32_t T, const int32_t* T_b, float* out_host) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] {
    if (!T_b) fail(WG_ERR_INVALID, "T_b must not be NULL");
    infer_host(h, mel_host, z_host, sigma, deterministic, B, T, T_b, out_host);
  });
}

int wg_last_launch_count(wg_handle h) { return h ? h->launches : 0; }

int wg_profile_enable(wg_handle h, int32_t enable) {
  if (!h) return WG_ERR_INVALID;
  h->profiling = enable != 0;
  h->ev_used = 0;
  h->ev_count.clear();
  return WG_OK;
}

int wg_profile_read(wg_handle h, double* layer_ms_sum, int32_t* layer_launches) {
  if (!h || !layer_ms_sum || !layer_launches) return WG_ERR_INVALID;
  return guarded(h, [&] {
    double sum = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
      CK(cudaEventSynchronize(h->ev_pool[i + 1]));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
      sum += ms;
    }
    *layer_ms_sum = sum;
    int n = 0;
    for (size_t i = 0; i < h->ev_used / 2 && i < h->ev_count.size(); ++i) n += h->ev_count[i];
    *layer_launches = n;
    h->ev_used = 0;
    h->ev_count.clear();
  });
}

int wg_debug_pair_info(wg_handle h, int32_t* max_pairs, int32_t* last_used) {
  if (!h || !max_pairs || !last_used) return WG_ERR_INVALID;
  *max_pairs = h->cfg.mode == WG_MODE_TF32X3 ? h->t3_max_pairs : h->pair_max;
  *last_used = h->last_pair;
  return WG_OK;
}

int wg_debug_read_timing(wg_handle h, uint64_t* out128) {
  if (!h || !out128) return WG_ERR_INVALID;
  return guarded(h, [&] {
    if (!h->timing) fail(WG_ERR_INVALID, "layer timing is off (set WG_LAYER_TIMING=1 before wg_create; BF16 / TF32X3 modes)");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out128, h->timing, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(h->timing, 0, 128 * sizeof(unsigned long long)));
  });
}

int wg_debug_infer_prefix(wg_handle h, const float* mel, const float* z, float sigma, int32_t deterministic,
                          int32_t B, int32_t T, void* workspace, size_t workspace_bytes, void* stream,
                          int32_t stop_flow, int32_t stop_layer, float* h_out, float* acc_out) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] {
    if (stop_flow < 0 || stop_flow >= h->cfg.n_flows || stop_layer < -1 || stop_layer >= h->cfg.n_layers)
      fail(WG_ERR_INVALID, "bad stop point (%d, %d)", stop_flow, stop_layer);
    run_infer(h, mel, z, sigma, deterministic, B, T, nullptr, workspace, workspace_bytes,
              static_cast<cudaStream_t>(stream), stop_flow, stop_layer, h_out, acc_out);
  });
}

int wg_debug_get_spect(wg_handle h, int32_t B, int32_t T, const void* workspace, float* spect_out, void* stream) {
  if (!h) return WG_ERR_INVALID;
  return guarded(h, [&] {
    check_shape(h, B, T);
    if (!workspace || !spect_out) fail(WG_ERR_INVALID, "NULL argument");
    if (use_pm(h, B, T)) fail(WG_ERR_UNSUPPORTED, "the phase-major path does not materialise spect (set WG_PM=0 to inspect it)");
    const Ws w = carve(h, B, T);
    const size_t n = (size_t)B * T * h->R * h->S;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->cfg.mode == WG_MODE_FP32) {
      CK(cudaMemcpyAsync(spect_out, static_cast<const char*>(workspace) + w.spect, n * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      tc_bf16_to_f32(reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(workspace) + w.spect16),
                     spect_out, n, st);
      CK(cudaGetLastError());
    }
  });
}

int wg_debug_gemm_bf16(const void* A, const void* W, const float* bias, float* D, int32_t M, int32_t N, int32_t K,
                       void* stream) {
  std::string msg;
  try {
    tc_init();
    tc_debug_gemm(static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(W), bias, D, M, N, K,
                  static_cast<cudaStream_t>(stream));
    return WG_OK;
  } catch (const Fail& x) {
    std::lock_guard<std::mutex> g(g_err_mu);
    g_create_err = x.msg;
    return x.code;
  }
}

}  // extern "C"

// Log-mel front-end of libwg_b200.so (C ABI: include/wg_mel_b200.h).
//
// Restates TacotronSTFT.mel_spectrogram (reference utils/audio/stft.py:286-319) as ONE kernel:
//   audio --reflect pad, hann window--> 1024-point real FFT per frame (shared memory, fp32)
//         --> magnitude (513 bins) --> sparse mel basis --> log(max(., clip)) --> [B, F, n_mel].
// The reference evaluates the same transform as a 1026-filter strided convolution (stft.py:258-270);
// per frame that is 2.1 MFLOP against 4 KB of input. Here a frame costs ~50 kFLOP, the audio is read
// from HBM once per CTA (a group of 8 frames shares its 2816-sample span through shared memory) and
// only the 80 log-mel values per frame go back: the kernel is sized against the HBM roofline
// (4 B/sample in + 4*n_mel/hop B/sample out), not against the tensor cores.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/wg_mel_b200.h"
#include "common.cuh"

namespace {

using wg::fail;
#define CK WG_CK

constexpr int kNfft = 1024;         // filter_length the kernel implements
constexpr int kM = kNfft / 2;       // complex FFT length (two real samples per complex point)
constexpr int kBins = kM + 1;       // 513
constexpr int kFramesPerCta = 8;    // one warp per frame
constexpr int kThreads = 32 * kFramesPerCta;
constexpr int kMagPitch = 520;      // floats per warp for the magnitude row
constexpr int kZPitch = kM + kM / 8; // float2 per warp for the FFT buffer (see zpad)
constexpr int kTwPasses = 64 + 512; // twiddles of the second and third pass
constexpr int kMaxEll = 8192;       // (widest mel filter) x n_mel entries of the mel basis kept in shared memory

struct MelArgs {
  const float* audio;   // [B, N]
  float* mel;           // [B, F, n_mel]
  long long N;          // samples per row as given
  long long Neff;       // max(N, win_length): the zero-padded length the reference transforms
  int B, F, hop, n_mel, ell_w;
  int groups_per_row;   // ceil(F / kFramesPerCta)
  long long n_groups;
  float clip;
  const float* window;      // [1024]
  const float2* tw512;      // [0,64): exp(-2 pi i r k / 64) at [r*8+k]; [64,576): exp(-2 pi i r k / 512) at 64+[r*64+k]
  const float2* tw1024;     // exp(-2 pi i k / 1024), k <= 256
  const int* mel_cnt;       // [n_mel] non-zeros of every mel channel
  const float2* mel_ell;    // [ell_w][n_mel] {bin (int bits), weight}: entry i of channel m at i*n_mel+m
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// 4-point DFT, natural order in and out
__device__ __forceinline__ void dft4(float2 c0, float2 c1, float2 c2, float2 c3, float2& o0, float2& o1,
                                     float2& o2, float2& o3) {
  float2 s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3), s3 = mul_mi(csub(c1, c3));
  o0 = cadd(s0, s2);
  o1 = cadd(s1, s3);
  o2 = csub(s0, s2);
  o3 = csub(s1, s3);
}

// 8-point DFT, natural order in and out: one radix-2 split into even/odd outputs, then two dft4
__device__ __forceinline__ void dft8(float2* v) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));    // * exp(-i pi/4)
  b2 = mul_mi(b2);                                           // * exp(-i pi/2)
  b3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));   // * exp(-3i pi/4)
  dft4(a0, a1, a2, a3, v[0], v[2], v[4], v[6]);
  dft4(b0, b1, b2, b3, v[1], v[3], v[5], v[7]);
}

// Shared-memory index of complex point i: one float2 of padding after every 8 keeps the stride-8 and
// stride-64 scatter of the Stockham passes spread over all 16 eight-byte banks of a half-warp access.
__device__ __forceinline__ int zpad(int i) { return i + (i >> 3); }

// One Stockham radix-8 pass over the warp's 512 complex points (64 butterflies, two per lane):
// v[r] = in[j + 64 r] * w^(r k), k = j mod Ns, w = exp(-2 pi i / (8 Ns)); out[(j / Ns) 8 Ns + k + r Ns].
// `tw` is laid out [r][k] (k < Ns) so that consecutive lanes read consecutive entries.
// The first pass (Ns = 1) takes its input straight from the staged audio and the window:
// z[n] = (w[2n] x[2n], w[2n+1] x[2n+1]).
template <int Ns>
__device__ __forceinline__ void fft_pass(float2* buf, const float2* tw, int lane, const float* x = nullptr,
                                         const float2* win2 = nullptr) {
  float2 v[2][8];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int j = lane + 32 * jj;
    if (Ns == 1) {
      const bool aligned = (reinterpret_cast<uintptr_t>(x) & 7) == 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = j + 64 * r;
        float2 xv = aligned ? reinterpret_cast<const float2*>(x)[n] : make_float2(x[2 * n], x[2 * n + 1]);
        float2 w = win2[n];
        v[jj][r] = make_float2(xv.x * w.x, xv.y * w.y);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) v[jj][r] = buf[zpad(j + 64 * r)];
    }
    if (Ns > 1) {
      const int k = j & (Ns - 1);
#pragma unroll
      for (int r = 1; r < 8; ++r) v[jj][r] = cmul(v[jj][r], tw[r * Ns + k]);
    }
    dft8(v[jj]);
  }
  __syncwarp();
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int j = lane + 32 * jj;
    const int k = j & (Ns - 1);
    const int base = (j / Ns) * (8 * Ns) + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[zpad(base + r * Ns)] = v[jj][r];
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 2) mel_frames_kernel(MelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_win = reinterpret_cast<float*>(smem_raw);                      // 1024
  float2* s_tw512 = reinterpret_cast<float2*>(s_win + kNfft);             // 576
  float2* s_tw1024 = s_tw512 + kTwPasses;                                 // 264 (257 used)
  float2* s_z = s_tw1024 + 264;                                           // 8 x 576
  float* s_mag = reinterpret_cast<float*>(s_z + kFramesPerCta * kZPitch); // 8 x 520
  int* s_cnt = reinterpret_cast<int*>(s_mag + kFramesPerCta * kMagPitch); // 128 (n_mel used)
  float2* s_ell = reinterpret_cast<float2*>(s_cnt + 128);                 // ell_w x n_mel
  float* s_x = reinterpret_cast<float*>(s_ell + a.ell_w * a.n_mel);       // 7 hop + 1024

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kNfft; i += kThreads) s_win[i] = a.window[i];
  for (int i = tid; i < kTwPasses; i += kThreads) s_tw512[i] = a.tw512[i];
  for (int i = tid; i <= kM / 2; i += kThreads) s_tw1024[i] = a.tw1024[i];
  for (int i = tid; i < a.n_mel; i += kThreads) s_cnt[i] = a.mel_cnt[i];
  for (int i = tid; i < a.ell_w * a.n_mel; i += kThreads) s_ell[i] = a.mel_ell[i];

  float2* zb = s_z + warp * kZPitch;
  float* mg = s_mag + warp * kMagPitch;

  for (long long g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
    const int b = static_cast<int>(g / a.groups_per_row);
    const int f0 = static_cast<int>(g % a.groups_per_row) * kFramesPerCta;
    const int nf = min(kFramesPerCta, a.F - f0);
    const int span = (nf - 1) * a.hop + kNfft;
    const long long j0 = static_cast<long long>(f0) * a.hop - kNfft / 2;
    const float* row = a.audio + static_cast<long long>(b) * a.N;
    __syncthreads();  // previous group's readers of s_x are done (and the tables are in place)
    for (int t = tid; t < span; t += kThreads) {
      long long i = j0 + t;
      if (i < 0) i = -i;                              // reflect, edge sample not repeated
      if (i >= a.Neff) i = 2 * (a.Neff - 1) - i;
      s_x[t] = i < a.N ? __ldg(row + i) : 0.0f;       // [N, Neff) is the reference's zero padding
    }
    __syncthreads();
    if (warp < nf) {
      fft_pass<1>(zb, s_tw512, lane, s_x + warp * a.hop, reinterpret_cast<const float2*>(s_win));
      fft_pass<8>(zb, s_tw512, lane);
      fft_pass<64>(zb, s_tw512 + 64, lane);
      // real-input split: X[k] = E + t O, X[M-k] = conj(E - t O), t = exp(-2 pi i k / 1024)
      for (int k = lane; k <= kM / 2; k += 32) {
        float2 p = zb[zpad(k)], q = zb[zpad((kM - k) & (kM - 1))];
        float2 E = make_float2(0.5f * (p.x + q.x), 0.5f * (p.y - q.y));
        float2 O = make_float2(0.5f * (p.y + q.y), -0.5f * (p.x - q.x));
        float2 tO = cmul(s_tw1024[k], O);
        float2 u = cadd(E, tO), d = csub(E, tO);
        mg[k] = sqrtf(fmaf(u.x, u.x, u.y * u.y));
        mg[kM - k] = sqrtf(fmaf(d.x, d.x, d.y * d.y));
      }
      __syncwarp();
      float* out = a.mel + (static_cast<long long>(b) * a.F + f0 + warp) * a.n_mel;
      for (int m = lane; m < a.n_mel; m += 32) {
        float acc = 0.0f;
        const int cnt = s_cnt[m];
        const float2* e = s_ell + m;
        for (int i = 0; i < cnt; ++i, e += a.n_mel) {
          float2 bw = *e;
          acc = fmaf(mg[__float_as_int(bw.x)], bw.y, acc);
        }
        out[m] = logf(fmaxf(acc, a.clip));
      }
    }
  }
}

size_t mel_smem_bytes(int hop, int ell_entries) {
  size_t floats = kNfft + 2 * kTwPasses + 2 * 264 + 2 * kFramesPerCta * kZPitch + kFramesPerCta * kMagPitch + 128 +
                  2 * static_cast<size_t>(ell_entries) + (kFramesPerCta - 1) * static_cast<size_t>(hop) + kNfft;
  return floats * sizeof(float);
}

}  // namespace

struct wg_mel_engine {
  wg_mel_config cfg{};
  int device = 0;
  int ell_w = 0;
  int n_sm = 148;
  float* d_window = nullptr;
  float2* d_tw512 = nullptr;
  float2* d_tw1024 = nullptr;
  int* d_cnt = nullptr;
  float2* d_ell = nullptr;
  // wg_mel_spectrogram_host staging
  cudaStream_t stream = nullptr;
  float *pin_in = nullptr, *dev_in = nullptr, *pin_out = nullptr, *dev_out = nullptr;
  size_t cap_in = 0, cap_out = 0;
  std::string err;
};

namespace {

std::mutex g_mel_err_mu;
std::string g_mel_create_err;

void destroy_mel(wg_mel_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaFree(e->d_window); cudaFree(e->d_tw512); cudaFree(e->d_tw1024);
  cudaFree(e->d_cnt); cudaFree(e->d_ell);
  if (e->pin_in) cudaFreeHost(e->pin_in);
  if (e->pin_out) cudaFreeHost(e->pin_out);
  cudaFree(e->dev_in); cudaFree(e->dev_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

template <typename F>
int mel_guarded(wg_mel_engine* h, F&& f) {
  try {
    f();
    return WG_OK;
  } catch (const wg::Fail& x) {
    if (h) h->err = x.msg;
    else {
      std::lock_guard<std::mutex> g(g_mel_err_mu);
      g_mel_create_err = x.msg;
    }
    return x.code;
  } catch (const std::exception& x) {
    if (h) h->err = x.what();
    return WG_ERR_INVALID;
  }
}

template <typename T>
T* upload(const std::vector<T>& v) {
  T* d = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&d), std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

void build_mel(wg_mel_engine* e, const wg_mel_config* cfg, const float* window, const float* basis, int device) {
  e->cfg = *cfg;
  e->device = device;
  if (cfg->filter_length != kNfft)
    fail(WG_ERR_UNSUPPORTED, "wg_mel_create: filter_length %d not supported (the kernel implements %d)",
         cfg->filter_length, kNfft);
  if (cfg->hop_length < 1 || cfg->hop_length > kNfft)
    fail(WG_ERR_INVALID, "wg_mel_create: hop_length %d outside [1, %d]", cfg->hop_length, kNfft);
  if (cfg->win_length < 1 || cfg->win_length > kNfft)
    fail(WG_ERR_INVALID, "wg_mel_create: win_length %d outside [1, %d]", cfg->win_length, kNfft);
  if (cfg->n_mel_channels < 1 || cfg->n_mel_channels > 128)
    fail(WG_ERR_INVALID, "wg_mel_create: n_mel_channels %d outside [1, 128]", cfg->n_mel_channels);
  if (!(cfg->clip_val > 0.0f)) fail(WG_ERR_INVALID, "wg_mel_create: clip_val must be > 0");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    fail(WG_ERR_CUDA, "wg_mel_create: no CUDA device (the B200 mel front-end has no CPU fallback)");
  if (device < 0 || device >= n_dev) fail(WG_ERR_INVALID, "wg_mel_create: device %d of %d", device, n_dev);
  wg::DeviceGuard dev_guard(device);
  CK(cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device));

  // mel basis, one column (channel) at a time: the bins with a non-zero weight, stored channel-minor
  // (entry i of channel m at i*n_mel+m) so that the lanes of a warp, one channel each, read consecutively
  const int n_mel = cfg->n_mel_channels;
  std::vector<std::vector<std::pair<int, float>>> cols(n_mel);
  size_t width = 0;
  for (int m = 0; m < n_mel; ++m) {
    for (int k = 0; k < kBins; ++k) {
      float v = basis[static_cast<size_t>(k) * n_mel + m];
      if (!std::isfinite(v)) fail(WG_ERR_INVALID, "wg_mel_create: mel_basis[%d,%d] is not finite", k, m);
      if (v != 0.0f) cols[m].push_back({k, v});
    }
    width = std::max(width, cols[m].size());
  }
  if (width * n_mel > static_cast<size_t>(kMaxEll))
    fail(WG_ERR_UNSUPPORTED, "wg_mel_create: widest mel filter has %zu bins; %zu x %d entries > %d kept in shared memory",
         width, width, n_mel, kMaxEll);
  e->ell_w = static_cast<int>(width);
  std::vector<int> cnt(n_mel);
  std::vector<float2> ell(std::max<size_t>(1, width * n_mel), make_float2(0.0f, 0.0f));
  for (int m = 0; m < n_mel; ++m) {
    cnt[m] = static_cast<int>(cols[m].size());
    for (size_t i = 0; i < cols[m].size(); ++i) {
      int bin = cols[m][i].first;
      float as_float;
      memcpy(&as_float, &bin, sizeof bin);
      ell[i * n_mel + m] = make_float2(as_float, cols[m][i].second);
    }
  }
  std::vector<float> win(window, window + kNfft);
  for (float v : win)
    if (!std::isfinite(v)) fail(WG_ERR_INVALID, "wg_mel_create: window is not finite");
  std::vector<float2> t512(kTwPasses), t1024(kM / 2 + 1);
  const double two_pi = 6.283185307179586476925286766559;
  for (int r = 0; r < 8; ++r) {
    for (int k = 0; k < 8; ++k)
      t512[r * 8 + k] = make_float2(static_cast<float>(std::cos(two_pi * r * k / 64)), static_cast<float>(-std::sin(two_pi * r * k / 64)));
    for (int k = 0; k < 64; ++k)
      t512[64 + r * 64 + k] = make_float2(static_cast<float>(std::cos(two_pi * r * k / kM)), static_cast<float>(-std::sin(two_pi * r * k / kM)));
  }
  for (int k = 0; k <= kM / 2; ++k)
    t1024[k] = make_float2(static_cast<float>(std::cos(two_pi * k / kNfft)), static_cast<float>(-std::sin(two_pi * k / kNfft)));
  e->d_window = upload(win);
  e->d_tw512 = upload(t512);
  e->d_tw1024 = upload(t1024);
  e->d_cnt = upload(cnt);
  e->d_ell = upload(ell);
  size_t smem = mel_smem_bytes(cfg->hop_length, e->ell_w * n_mel);
  if (smem > 227 * 1024) fail(WG_ERR_UNSUPPORTED, "wg_mel_create: %zu bytes of shared memory needed", smem);
  CK(cudaFuncSetAttribute(mel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
}

long long mel_neff(const wg_mel_engine* e, long long n) { return std::max<long long>(n, e->cfg.win_length); }

long long mel_frames(const wg_mel_engine* e, long long n) {
  if (n < 1) fail(WG_ERR_INVALID, "mel: n_samples %lld < 1", n);
  long long neff = mel_neff(e, n);
  if (neff <= kNfft / 2)
    fail(WG_ERR_INVALID, "mel: %lld samples cannot be reflect-padded by %d (win_length %d)", neff, kNfft / 2,
         e->cfg.win_length);
  return neff / e->cfg.hop_length + 1;
}

void launch_mel(wg_mel_engine* e, const float* audio, int B, long long n, float* mel, cudaStream_t st) {
  if (!audio || !mel) fail(WG_ERR_INVALID, "wg_mel_spectrogram: NULL buffer");
  if (B < 1) fail(WG_ERR_INVALID, "wg_mel_spectrogram: B %d < 1", B);
  long long F = mel_frames(e, n);
  if (F > 0x7fffff00LL) fail(WG_ERR_INVALID, "wg_mel_spectrogram: %lld frames per row", F);
  wg::DeviceGuard dev_guard(e->device);
  MelArgs a{};
  a.audio = audio; a.mel = mel; a.N = n; a.Neff = mel_neff(e, n);
  a.B = B; a.F = static_cast<int>(F); a.hop = e->cfg.hop_length; a.n_mel = e->cfg.n_mel_channels; a.ell_w = e->ell_w;
  a.groups_per_row = static_cast<int>((F + kFramesPerCta - 1) / kFramesPerCta);
  a.n_groups = static_cast<long long>(a.groups_per_row) * B;
  a.clip = e->cfg.clip_val;
  a.window = e->d_window; a.tw512 = e->d_tw512; a.tw1024 = e->d_tw1024;
  a.mel_cnt = e->d_cnt; a.mel_ell = e->d_ell;
  size_t smem = mel_smem_bytes(a.hop, a.ell_w * a.n_mel);
  int grid = static_cast<int>(std::min<long long>(a.n_groups, 2LL * e->n_sm));
  mel_frames_kernel<<<grid, kThreads, smem, st>>>(a);
  CK(cudaGetLastError());
}

void ensure_stage(float*& pin, float*& dev, size_t& cap, size_t bytes) {
  if (cap >= bytes && pin && dev) return;
  if (pin) cudaFreeHost(pin);
  if (dev) cudaFree(dev);
  pin = nullptr; dev = nullptr; cap = 0;
  CK(cudaMallocHost(reinterpret_cast<void**>(&pin), bytes));
  CK(cudaMalloc(reinterpret_cast<void**>(&dev), bytes));
  cap = bytes;
}

}  // namespace

extern "C" {

int wg_mel_create(const wg_mel_config* cfg, const float* window, const float* mel_basis, int32_t device,
                  wg_mel_handle* out) {
  if (out) *out = nullptr;
  if (!cfg || !window || !mel_basis || !out) {
    std::lock_guard<std::mutex> g(g_mel_err_mu);
    g_mel_create_err = "wg_mel_create: NULL argument";
    return WG_ERR_INVALID;
  }
  wg_mel_engine* e = new wg_mel_engine();
  int rc = mel_guarded(nullptr, [&] { build_mel(e, cfg, window, mel_basis, device); });
  if (rc != WG_OK) {
    destroy_mel(e);
    return rc;
  }
  *out = e;
  return WG_OK;
}

void wg_mel_destroy(wg_mel_handle h) { destroy_mel(h); }

const char* wg_mel_last_error(wg_mel_handle h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> g(g_mel_err_mu);
  static thread_local std::string copy;
  copy = g_mel_create_err;
  return copy.c_str();
}

int wg_mel_frames(wg_mel_handle h, int64_t n_samples, int64_t* frames) {
  if (!h || !frames) return WG_ERR_INVALID;
  return mel_guarded(h, [&] { *frames = mel_frames(h, n_samples); });
}

int wg_mel_spectrogram(wg_mel_handle h, const float* audio, int32_t B, int64_t n_samples, float* mel, void* stream) {
  if (!h) return WG_ERR_INVALID;
  return mel_guarded(h, [&] { launch_mel(h, audio, B, n_samples, mel, static_cast<cudaStream_t>(stream)); });
}

int wg_mel_spectrogram_host(wg_mel_handle h, const float* audio, int32_t B, int64_t n_samples, float* mel) {
  if (!h) return WG_ERR_INVALID;
  return mel_guarded(h, [&] {
    if (!audio || !mel) fail(WG_ERR_INVALID, "wg_mel_spectrogram_host: NULL buffer");
    if (B < 1) fail(WG_ERR_INVALID, "wg_mel_spectrogram_host: B %d < 1", B);
    long long F = mel_frames(h, n_samples);
    wg::DeviceGuard dev_guard(h->device);
    size_t in_bytes = static_cast<size_t>(B) * n_samples * sizeof(float);
    size_t out_bytes = static_cast<size_t>(B) * F * h->cfg.n_mel_channels * sizeof(float);
    ensure_stage(h->pin_in, h->dev_in, h->cap_in, in_bytes);
    ensure_stage(h->pin_out, h->dev_out, h->cap_out, out_bytes);
    memcpy(h->pin_in, audio, in_bytes);
    CK(cudaMemcpyAsync(h->dev_in, h->pin_in, in_bytes, cudaMemcpyHostToDevice, h->stream));
    launch_mel(h, h->dev_in, B, n_samples, h->dev_out, h->stream);
    CK(cudaMemcpyAsync(h->pin_out, h->dev_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(mel, h->pin_out, out_bytes);
  });
}

}  // extern "C"

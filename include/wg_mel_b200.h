/*
 * wg_mel_b200.h -- C ABI of the B200 log-mel front-end (same shared library, libwg_b200.so).
 *
 * The data format on the INPUT side of the WaveGlow path: the 80-bin log-mel spectrogram the
 * vocoder consumes is produced from audio by the reference's TacotronSTFT. Reference interfaces
 * replaced (paths relative to the reference tree):
 *
 *   utils/audio/stft.py:286-319   TacotronSTFT.__init__/mel_spectrogram/spectral_normalize
 *                                                     -> wg_mel_create / wg_mel_spectrogram
 *   utils/audio/stft.py:189-235   STFT.__init__: windowed Fourier basis (hann, periodic,
 *                                 centre-padded to filter_length)  -> the `window` argument
 *   utils/audio/stft.py:241-280   STFT.transform: reflect pad filter_length/2, strided conv with the
 *                                 basis, magnitude                 -> the arithmetic of the kernel
 *   utils/audio/stft.py:59-68     librosa mel basis [n_bins, n_mel] -> the `mel_basis` argument
 *   utils/audio/stft.py:98-126    MelSTFT.__call__: audio shorter than win_length is zero padded
 *                                 to win_length                     -> done inside wg_mel_spectrogram
 *   utils/audio/stft.py:128-130   get_mel_length                   -> wg_mel_frames (exact count)
 *
 * The kernel computes each frame with a shared-memory radix-8 real FFT in fp32 instead of the
 * reference's 1026-filter convolution (same numbers up to fp32 rounding; tolerance in
 * tests/test_gpu_3_mel.py). Conventions are those of wg_b200.h: status codes (wg_status), no
 * exception crosses the boundary, no CPU fallback.
 */
#ifndef WG_MEL_B200_H_
#define WG_MEL_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "wg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Constructor arguments of TacotronSTFT / MelSTFT (stft.py:28-46, 287-305). */
typedef struct wg_mel_config {
  int32_t sampling_rate;   /* 22050 (informational; the mel basis already encodes it) */
  int32_t n_mel_channels;  /* 80 (1..128) */
  int32_t filter_length;   /* 1024 (the only FFT size the kernel implements) */
  int32_t hop_length;      /* 256 (1..filter_length) */
  int32_t win_length;      /* 1024 (<= filter_length); shorter audio is zero padded to this */
  float clip_val;          /* 1e-5: log(max(mel, clip_val)), stft.py:307-308 */
} wg_mel_config;

typedef struct wg_mel_engine* wg_mel_handle;

/* `window`: host float[filter_length] (already centre-padded); `mel_basis`: host float
 * [filter_length/2+1, n_mel_channels] row-major (= MelSTFT.mel_basis[0], stft.py:68). */
int wg_mel_create(const wg_mel_config* cfg, const float* window, const float* mel_basis, int32_t device,
                  wg_mel_handle* out);
void wg_mel_destroy(wg_mel_handle h);
/* Message of the last failure on this handle (or of the last failed wg_mel_create when h is NULL). */
const char* wg_mel_last_error(wg_mel_handle h);

/* Frames produced for n_samples of audio: max(n_samples, win_length) / hop_length + 1. */
int wg_mel_frames(wg_mel_handle h, int64_t n_samples, int64_t* frames);

/* audio: device float [B, n_samples] (row stride n_samples); mel: device float
 * [B, frames, n_mel_channels]. Asynchronous on `stream` (cudaStream_t), no allocation, no host sync. */
int wg_mel_spectrogram(wg_mel_handle h, const float* audio, int32_t B, int64_t n_samples, float* mel,
                       void* stream);

/* Same with HOST buffers: engine-owned pinned staging, H2D, kernel, D2H, synchronous return. */
int wg_mel_spectrogram_host(wg_mel_handle h, const float* audio, int32_t B, int64_t n_samples, float* mel);

#ifdef __cplusplus
}
#endif
#endif /* WG_MEL_B200_H_ */

/*
 * wg_b200.h -- C ABI of the B200-native WaveGlow inference engine (libwg_b200.so).
 *
 * This is the drop-in boundary for ONE path of yui-mhcp/text_to_speech: WaveGlow vocoder
 * inference. Reference interfaces replaced (paths relative to the reference tree):
 *
 *   utils/keras/runtimes/runtime.py:19-41      Runtime.__init__/load_engine  -> wg_create / wg_destroy
 *   utils/keras/runtimes/runtime.py:34-36      Runtime.__call__              -> wg_infer / wg_infer_host
 *   architectures/waveglow_arch.py:244-306     WaveGlow.infer(inputs, z, sigma, deterministic)
 *                                                                            -> the arithmetic behind wg_infer
 *   architectures/waveglow_arch.py:308-310     set_weights (+ W^-1 rebuild,
 *   architectures/layers/invertible_conv.py:41-47 build_inverse)             -> wg_create (weights in Keras layouts)
 *   utils/keras/runtimes/tensorrt_runtime.py:143-210  per-shape pinned/device buffers, one private
 *                                              stream, synchronous return    -> wg_infer_host
 *
 * Conventions: plain C, no exceptions cross the boundary. Every function returns WG_OK (0) or a
 * negative wg_status; wg_last_error() gives the message (the Python host layer raises RuntimeError
 * with it, as the TensorRT runtime precedent does). A handle is immutable after wg_create, so
 * concurrent wg_infer calls on different streams with different workspaces are safe; wg_infer_host
 * uses engine-owned staging buffers and is NOT re-entrant. Every entry point runs on the engine's
 * device and restores the caller's current CUDA device before it returns. There is no CPU fallback:
 * without a CUDA device wg_create fails with WG_ERR_CUDA.
 */
#ifndef WG_B200_H_
#define WG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WG_ABI_VERSION 2

typedef enum wg_status {
  WG_OK = 0,
  WG_ERR_INVALID = -1,      /* bad argument / shape / hparams */
  WG_ERR_UNSUPPORTED = -2,  /* configuration not supported by the requested mode */
  WG_ERR_CUDA = -3,         /* CUDA runtime/driver failure (message has the CUDA error string) */
  WG_ERR_WORKSPACE = -4,    /* workspace too small or misaligned */
  WG_ERR_WEIGHTS = -5       /* missing tensor / wrong shape in the weight set */
} wg_status;

typedef enum wg_mode {
  WG_MODE_FP32 = 0,  /* fp32 FFMA arithmetic in the reference's op order; <= 1e-4 max-abs vs reference fp32 */
  WG_MODE_BF16 = 1,  /* tcgen05 BF16 operands, fp32 accumulate/residual; <= 2e-2 max-abs, >= 35 dB SNR */
  WG_MODE_TF32X3 = 2 /* fp32-grade on the tensor cores: every fp32 operand split as hi = tf32(x), lo = x - hi, every
                        product issued as a_lo*b_hi + a_hi*b_lo (BF16 MMAs) + a_hi*b_hi (TF32 MMAs), fp32 accumulation in
                        TMEM, fp32 gate; <= 1e-4 max-abs vs reference fp32 (the "fp32/3xTF32 mode"; measured 2e-5) */
} wg_mode;

/* Constructor arguments of architectures.WaveGlow (waveglow_arch.py:164-181) + arithmetic mode. */
typedef struct wg_config {
  int32_t n_mel_channels;  /* 80 */
  int32_t n_flows;         /* 12 */
  int32_t n_group;         /* 8 */
  int32_t n_early_every;   /* 4 */
  int32_t n_early_size;    /* 2 */
  int32_t n_layers;        /* 8 */
  int32_t n_channels;      /* 256 (WaveGlow-256) or 512 (reference default) */
  int32_t kernel_size;     /* 3 */
  int32_t mode;            /* wg_mode */
} wg_config;

/* One named weight tensor: HOST pointer, float32, Keras variable name and Keras layout
 * (Conv1D kernel [k,in,out]; Conv1DTranspose kernel [k,out,in]; see text_to_speech_b200/weights.py). */
typedef struct wg_tensor {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} wg_tensor;

typedef struct wg_engine* wg_handle;

int wg_abi_version(void);

/* Builds an engine on CUDA device `device`: validates hparams/weights, inverts the 1x1 kernels
 * (fp32, once), repacks and uploads the weights. *out is NULL on failure. */
int wg_create(const wg_config* cfg, const wg_tensor* tensors, int32_t n_tensors, int32_t device,
              wg_handle* out);

void wg_destroy(wg_handle h);

/* Message of the last failing call on this handle (h == NULL: last wg_create failure). */
const char* wg_last_error(wg_handle h);

/* Bytes of device scratch wg_infer needs for a [B, T, n_mel] mel. */
int wg_workspace_bytes(wg_handle h, int32_t B, int32_t T, size_t* bytes);

/* WaveGlow.infer. All pointers are DEVICE pointers on the engine's device.
 *   mel  [B, T, n_mel] float32 channels-last          (models/tts/waveglow.py:61-82 input layout)
 *   z    [B, T*256/n_group, n_group] float32, may be NULL iff deterministic != 0
 *        (channels [0,n_rem) seed the first flow, then n_early_size more after each early flow,
 *         waveglow_arch.py:269-270, :298-299)
 *   out  [B, 256*T] float32, caller-owned
 * Asynchronous on `stream` (a cudaStream_t); no allocation, no host synchronisation. */
int wg_infer(wg_handle h, const float* mel, const float* z, float sigma, int32_t deterministic,
             int32_t B, int32_t T, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers: copies through engine-owned pinned staging buffers on the engine's
 * private stream and returns after the waveform is in out_host (tensorrt_runtime.py:193-206). */
int wg_infer_host(wg_handle h, const float* mel_host, const float* z_host, float sigma,
                  int32_t deterministic, int32_t B, int32_t T, float* out_host);

/* Ragged batch: utterance b has T_b[b] frames (1 <= T_b[b] <= T); T_b is a HOST array of B entries that need not
 * outlive the call. The buffers keep the padded shapes of wg_infer (mel [B, T, n_mel], z [B, T*256/n_group, n_group],
 * out [B, 256*T]); only the first T_b[b] frames of an utterance are read and its first 256*T_b[b] samples written --
 * the tail of out is set to zero. No padding frame enters any convolution, so the samples of utterance b equal the
 * ones wg_infer produces for that utterance passed ALONE with T = T_b[b] -- the reference's call pattern, one trimmed
 * mel at a time (models/tts/tacotron2.py:183-191, models/tts/waveglow.py:76-82) -- bit for bit when both calls use the
 * same internal row layout (always, once the WG_PM environment switch pins it; otherwise the layout of a small
 * stand-alone call is picked by a wave count and the two agree to within the mode's tolerance).
 * WG_MODE_BF16 / WG_MODE_TF32X3 pack all utterances into one launch sequence; WG_MODE_FP32 runs them one after the
 * other on `stream`.
 * Asynchronous, no allocation, graph-capturable (the lengths travel in kernel parameters). */
int wg_workspace_bytes_ragged(wg_handle h, int32_t B, int32_t T, const int32_t* T_b, size_t* bytes);
int wg_infer_ragged(wg_handle h, const float* mel, const float* z, float sigma, int32_t deterministic,
                    int32_t B, int32_t T, const int32_t* T_b, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);
int wg_infer_host_ragged(wg_handle h, const float* mel_host, const float* z_host, float sigma,
                         int32_t deterministic, int32_t B, int32_t T, const int32_t* T_b, float* out_host);

/* Number of kernels the last wg_infer on this handle launched (bench.py's gpu_launches). */
int wg_last_launch_count(wg_handle h);

/* ---- test / profiling hooks (used by tests/ and bench.py only) ------------------------------ */

/* Per-kernel device timing of the dominant kernel with CUDA events on the caller's stream: in BF16 mode
 * one event pair per flow around its n_layers back-to-back WN-layer launches (tc_wn_layer_kernel /
 * tc512_*; nothing else is launched inside the bracket), in FP32 mode one pair per in-conv+gate GEMM.
 * wg_profile_read synchronises those events and returns the summed duration (ms) and the number of
 * layer launches it covers since the last read. */
int wg_profile_enable(wg_handle h, int32_t enable);
int wg_profile_read(wg_handle h, double* layer_ms_sum, int32_t* layer_launches);

/* In-kernel cycle counters of the fused WN-layer kernel, summed over CTAs and launches since the last
 * read (only when the engine was created with the environment variable WG_LAYER_TIMING=1, BF16 mode):
 * [0] MMA-warp total, [1] MMA waiting for TMA data, [2] MMA waiting for the epilogue, [3..5] epilogue
 * waiting for chunk a / chunk b / GEMM2 accumulators, [6] gate-epilogue work, [7] residual-epilogue work,
 * [8] TMA producer waiting for free stages. */
int wg_debug_read_timing(wg_handle h, uint64_t* out128);   /* [16..80): MMA wait-for-data cycles by stage position in the tile */

/* CTA-pair (cta_group::2) layer kernel of WG_MODE_BF16, C = 256: *max_pairs = clusters of two CTAs that can be resident at
 * once on this device (74 on a B200 whose 148 SMs form 74 complete TPCs; fewer on parts whose disabled SMs are spread over
 * TPCs -- the engine then keeps the single-CTA kernel), *last_used = 1 if the last wg_infer on this handle ran on it. */
int wg_debug_pair_info(wg_handle h, int32_t* max_pairs, int32_t* last_used);

/* Runs wg_infer but stops after WN layer `stop_layer` of flow `stop_flow` (flows run 11..0) and
 * copies the residual stream h [B*L, C] (float32) and the pre-coupling accumulator [B*L, 8] to
 * the given DEVICE buffers (either may be NULL). stop_layer == -1: stop right after the start
 * conv of that flow. */
int wg_debug_infer_prefix(wg_handle h, const float* mel, const float* z, float sigma,
                          int32_t deterministic, int32_t B, int32_t T, void* workspace,
                          size_t workspace_bytes, void* stream, int32_t stop_flow,
                          int32_t stop_layer, float* h_out, float* acc_out);

/* Copies the upsampled conditioning spect [B*L, n_mel*n_group] of the last wg_infer from the
 * workspace into `spect_out` (DEVICE, float32). */
int wg_debug_get_spect(wg_handle h, int32_t B, int32_t T, const void* workspace, float* spect_out,
                       void* stream);

/* Stand-alone BF16 tcgen05 GEMM self-test: D[M,N] = A[M,K] @ W[N,K]^T + bias, A/W bf16 (K-major),
 * D float32. DEVICE pointers. Used to validate the TMA/UMMA descriptor plumbing in isolation. */
int wg_debug_gemm_bf16(const void* A, const void* W, const float* bias, float* D, int32_t M,
                       int32_t N, int32_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WG_B200_H_ */

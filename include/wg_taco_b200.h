/*
 * wg_taco_b200.h -- C ABI of the B200 Tacotron2 DECODER LOOP (same shared library, libwg_b200.so).
 *
 * The caller side of the WaveGlow path: in the end-to-end tts() workload the autoregressive decoder
 * of the mel producer is the part that dominates (one frame per step, ~35 small library kernels per
 * frame). This entry point replaces that loop -- and only that loop; encoder and postnet stay with
 * the host framework. Reference code restated (paths relative to the reference tree):
 *
 *   architectures/tacotron2_arch.py:609-749   Tacotron2Decoder.infer (zero first frame, while_loop,
 *                                             finished / lengths / stop-token bookkeeping)
 *   architectures/tacotron2_arch.py:422-486   Tacotron2DecoderCell.call (attention LSTM, attention,
 *                                             decoder LSTM, [h, context] output)
 *   architectures/tacotron2_arch.py:188-203   Tacotron2Prenet.call (dense-relu-dropout x2; dropout stays
 *                                             on at inference unless deterministic)
 *   architectures/layers/location_sensitive_attention.py:104-186  LocationSensitiveAttention
 *                                             (concat_mode 2, cumulative, softmax)
 *   keras.layers.LSTMCell                     gate order i, f, c, o; sigmoid recurrent activation
 *
 * Parity status: pinned to the reference's own DECODER source. Tacotron2Prenet / Tacotron2DecoderCell /
 * Tacotron2Decoder.infer and LocationSensitiveAttention are executed unmodified over the Keras shim
 * (oracle/run_reference_taco.py; keras itself is not installable in the authoring environment) and their
 * outputs are committed as fixtures (tests/golden/taco_decoder_*.npz). This decoder reproduces the float64
 * fixture to 2.4e-7 over 24 frames (tests/test_gpu_4_tts.py); the torch restatement used as its day-to-day
 * checker reproduces the same source to 2e-16 (tests/test_oracle_taco.py). The shim's reading of Keras'
 * LSTMCell / Dense / Conv1D is ours, so -- as for WaveGlow -- a real-Keras run stays unpinned. Dropout
 * (deterministic = 0) uses a counter-based hash, not Keras' RNG: same distribution, different stream.
 *
 * Conventions are those of wg_b200.h (wg_status codes, wg_tensor, no exception crosses the boundary,
 * no CPU fallback).
 */
#ifndef WG_TACO_B200_H_
#define WG_TACO_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "wg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* HParamsTacotron2Decoder / HParamsLSA values the kernels are built for (tacotron2_arch.py:108-127,
 * location_sensitive_attention.py:17-24). Sizes marked (fixed) are compile-time constants. */
typedef struct wg_taco_config {
  int32_t n_mel_channels;         /* 80 (<= 128) */
  int32_t prenet_dim;             /* 256: both prenet layers (fixed 256) */
  int32_t embedding_dim;          /* 512: encoder output width (multiple of 4) */
  int32_t attention_rnn_dim;      /* 1024 (multiple of 8) */
  int32_t decoder_rnn_dim;        /* 1024 (multiple of 8) */
  int32_t attention_dim;          /* 128 (fixed) */
  int32_t attention_filters;      /* 32 (fixed) */
  int32_t attention_kernel_size;  /* 31 (odd, <= 63) */
  float prenet_drop_rate;         /* 0.5 */
  int32_t lstm_weight_dtype;      /* how the two LSTM gate GEMMs run:
                                     0: fp32 weights, fp32 FFMA arithmetic (2.4e-7 from the reference-source fixture);
                                     1: weights STORED in bf16 (round to nearest even at create time), fp32 FFMA --
                                        results equal an fp32 model with the rounded weights;
                                     2: split-bf16 tensor-core path (the host layer's default): weights and state are
                                        split into bf16 hi + lo, three mma.sync products per tile, fp32 accumulate --
                                        fp32 accuracy (3.1e-6 from the fixture), 1.7x faster than 0 */
} wg_taco_config;

typedef struct wg_taco_engine* wg_taco_handle;

/* Weights by name, HOST pointers, Keras layouts (Dense [in,out], Conv1D [k,in,out], LSTM kernel [in,4u],
 * recurrent_kernel [u,4u], bias [4u]); names as in text_to_speech_b200/tacotron2.py:
 *   decoder/prenet/layer_{0,1}/kernel, decoder/attention_rnn/{kernel,recurrent_kernel,bias},
 *   decoder/lsa/{query_layer,memory_layer,value_layer,location_conv,location_dense}/kernel,
 *   decoder/decoder_rnn/cell_0/{kernel,recurrent_kernel,bias}, decoder/linear_projection/{kernel,bias},
 *   decoder/gate_output/{kernel,bias}. */
int wg_taco_create(const wg_taco_config* cfg, const wg_tensor* tensors, int32_t n_tensors, int32_t device,
                   wg_taco_handle* out);
void wg_taco_destroy(wg_taco_handle h);
const char* wg_taco_last_error(wg_taco_handle h);

/* Runs the decoder loop for a batch.
 *   memory        device float [B, S, embedding_dim], rows at or beyond text_lengths[b] must be zero
 *   text_lengths  HOST int32 [B]   (the encoder mask: positions >= length get -inf energy)
 *   max_length    frames to allocate / run at most
 *   early_stopping != 0: stop once every utterance has finished (checked every `graph_chunk` frames,
 *                 so frames past that point may be non-zero, unlike the reference's while_loop)
 *   deterministic != 0: no prenet dropout; else masks are a function of (seed, frame, row, unit)
 *   outputs       device float [B, max_length, n_mel]   (decoder_output, before the postnet)
 *   stop_tokens   device float [B, max_length]          (sigmoid probabilities)
 *   attention     device float [B, max_length, S] or NULL
 *   lengths       device int32 [B]                      (frames with finished == false, :672)
 * Frames that were not run stay as the caller initialised them. Runs on the engine's own stream, ordered
 * after the work already queued on `stream` and before work queued on it afterwards; synchronises
 * with the host only when early_stopping is set. *frames_run receives the number of frames executed. */
int wg_taco_decode(wg_taco_handle h, const float* memory, const int32_t* text_lengths, int32_t B, int32_t S,
                   int32_t max_length, int32_t early_stopping, int32_t deterministic, uint64_t seed,
                   float* outputs, float* stop_tokens, float* attention, int32_t* lengths, int32_t* frames_run,
                   void* stream);

/* Frames per CUDA graph replay (default 32; 0 = launch every kernel directly, no graph). */
int wg_taco_set_graph_chunk(wg_taco_handle h, int32_t frames);

#ifdef __cplusplus
}
#endif
#endif /* WG_TACO_B200_H_ */

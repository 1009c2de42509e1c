"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product path (text_to_speech_b200/).

CPU restatement (torch CPU, float32 or float64) of the reference's WaveGlow inference arithmetic,
channels-last, op for op in the reference's order:

  architectures/waveglow_arch.py:19-24     _add_tanh_sigmoid_multiply      -> gate()
  architectures/waveglow_arch.py:105-141   WaveglowBlock.call              -> wn_block()
  architectures/waveglow_arch.py:244-306   WaveGlow.infer                  -> infer()
  architectures/layers/invertible_conv.py:41-51  build_inverse / call(reverse=True) -> w_inverse(), infer()
  models/tts/waveglow.py:61-142, 156-164   wrapper glue / windowing        -> not restated: pinned by the reference's own source
                                           (oracle/run_reference_wrapper.py) and its golden outputs (oracle/gen_golden_wrapper.py)

The arithmetic itself lives in Keras 3 (un-vendored, un-pinned third party: `keras` on a
tensorflow/torch/jax backend; not even listed in the reference's requirements.txt). Keras layer
semantics restated here: Conv1D 'valid' cross-correlation with kernel [k, in, out]
(out[l] = sum_j x[l + j*dilation] @ kernel[j] + bias); Conv1DTranspose 'valid' scatter with kernel
[k, out, in] (out[256 t + kappa, o] += x[t, i] * kernel[kappa, o, i]).

PARITY PIN: the reference holds no golden vector, KAT or test for this path and
keras is not importable in the authoring container. The pin used instead: the reference's OWN
source files (waveglow_arch.py, invertible_conv.py) are executed unmodified over a minimal
Keras-API shim (oracle/keras_shim, torch-backed) by oracle/run_reference.py; this restatement is
asserted equal to that run (tests/test_oracle.py) and the committed fixtures under tests/golden/
are produced by it (oracle/gen_golden.py). The shim is ours, so DESIGN.md states the parity as
"pinned to the reference's source over a Keras shim; real-Keras run unpinned".
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from text_to_speech_b200.weights import WaveGlowHParams, HOP, UPSAMPLE_K


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def conv1x1(x, kernel, bias=None):
    """keras.layers.Conv1D(kernel_size=1): x [B,L,in] @ kernel[0] [in,out] + bias."""
    y = x @ kernel[0]
    return y if bias is None else y + bias


def dilated_conv(x, kernel, bias, dilation):
    """in_layers[i](K.pad(audio, pad)) -- waveglow_arch.py:113-118: zero pad then 'valid' dilated
    cross-correlation. kernel [k, in, out]."""
    k = kernel.shape[0]
    pad_half = (k * dilation - dilation) // 2
    L = x.shape[1]
    xp = F.pad(x, (0, 0, pad_half, pad_half))
    y = None
    for j in range(k):
        term = xp[:, j * dilation: j * dilation + L, :] @ kernel[j]
        y = term if y is None else y + term
    return y + bias


def upsample(mel, kernel, bias):
    """keras.layers.Conv1DTranspose(80, 1024, strides=256) -- waveglow_arch.py:196-198, :245.
    mel [B,T,80]; kernel [1024, out, in] -> [B, 256 T + 768, 80]."""
    B, T, _ = mel.shape
    K, O, _ = kernel.shape
    out = torch.zeros(B, (T - 1) * HOP + K, O, dtype=mel.dtype)
    # y[b,t,kappa,o] = sum_i mel[b,t,i] * kernel[kappa,o,i]
    y = torch.einsum("bti,koi->btko", mel, kernel)
    for j in range(K // HOP):
        # taps kappa = 256 j + rho land on samples 256 (t + j) + rho
        out[:, j * HOP: (j + T) * HOP, :] += y[:, :, j * HOP:(j + 1) * HOP, :].reshape(B, T * HOP, O)
    return out + bias


def gate(in_a, in_b, C):
    """_add_tanh_sigmoid_multiply -- waveglow_arch.py:19-24."""
    in_act = in_a + in_b
    return torch.tanh(in_act[:, :, :C]) * torch.sigmoid(in_act[:, :, C:])


def w_inverse(kernel):
    """Invertible1x1Conv.build_inverse -- invertible_conv.py:41-47. kernel [1, c(in), c(out)].
    Returns M [c, c] such that reverse output = x @ M (M[a, b] = inv(W)[b, a], W[o, i] = kernel[0, i, o])."""
    W = kernel[0].transpose(0, 1)
    return torch.linalg.inv(W).transpose(0, 1).contiguous()


class OracleWaveGlow:
    def __init__(self, hp: WaveGlowHParams, weights, dtype=torch.float32):
        self.hp = hp
        self.dtype = dtype
        self.w = {k: _t(v, dtype) for k, v in weights.items() if not k.startswith("__")}
        # W^-1 is computed once at weight load, in float32 like the reference (K.inv on the fp32
        # kernel), then cast (invertible_conv.py:41-47, waveglow_arch.py:308-310).
        self.w_inv = []
        for k in range(hp.n_flows):
            kern = _t(weights[f"invertible_conv-{k}/conv/kernel"], torch.float32)
            self.w_inv.append(w_inverse(kern).to(dtype))

    # -- waveglow_arch.py:105-141 ---------------------------------------------------------------
    def wn_block(self, k, audio_0, spect, taps=None):
        hp, w = self.hp, self.w
        C, p = hp.n_channels, f"block-{k}/"
        audio = conv1x1(audio_0, w[p + "start_conv/kernel"], w[p + "start_conv/bias"])
        output = None
        for i in range(hp.n_layers):
            dilation = 2 ** i
            in_act = dilated_conv(audio, w[p + f"in_conv-{i}/kernel"], w[p + f"in_conv-{i}/bias"], dilation)
            cond = conv1x1(spect, w[p + f"cond_layer-{i}/kernel"], w[p + f"cond_layer-{i}/bias"])
            acts = gate(in_act, cond, C)
            rs = conv1x1(acts, w[p + f"res_skip_conv-{i}/kernel"], w[p + f"res_skip_conv-{i}/bias"])
            if i < hp.n_layers - 1:
                audio = rs[:, :, :C] + audio
                skip = rs[:, :, C:]
            else:
                skip = rs
            output = skip if i == 0 else skip + output
            if taps is not None:
                taps[f"flow{k}/layer{i}/acts"] = acts
                taps[f"flow{k}/layer{i}/audio"] = audio
                taps[f"flow{k}/layer{i}/skip"] = output
        return conv1x1(output, w[p + "end_conv/kernel"], w[p + "end_conv/bias"])

    def spect(self, mel):
        """waveglow_arch.py:245-253: upsample, trim 768, regroup to [B, L, 640] (channel = mel*8+g)."""
        hp = self.hp
        s = upsample(mel, self.w["upsample/kernel"], self.w["upsample/bias"])
        s = s[:, :-(UPSAMPLE_K - HOP), :]
        B = s.shape[0]
        Lg = s.shape[1] // hp.n_group
        s = s.reshape(B, Lg, hp.n_group, hp.n_mel_channels).permute(0, 1, 3, 2)
        return s.reshape(B, Lg, hp.n_group * hp.n_mel_channels)

    # -- waveglow_arch.py:244-306 ---------------------------------------------------------------
    def infer(self, mel, z=None, sigma=1.0, deterministic=False, taps=None, generator=None):
        hp = self.hp
        mel = _t(mel, self.dtype)
        spect = self.spect(mel)
        B, Lg = spect.shape[0], spect.shape[1]
        n_rem = hp.n_remaining_channels
        if z is not None:
            z = _t(z, self.dtype)
        if deterministic:
            noise = torch.zeros(B, Lg, n_rem, dtype=self.dtype)
        elif z is not None:
            noise = z[:, :, :n_rem]
            z = z[:, :, n_rem:hp.n_group]
        else:
            noise = torch.randn(B, Lg, n_rem, generator=generator).to(self.dtype)
        audio = sigma * noise
        if taps is not None:
            taps["spect"] = spect
        for k in reversed(range(hp.n_flows)):
            n_half = audio.shape[2] // 2
            audio_0, audio_1 = audio[:, :, :n_half], audio[:, :, n_half:]
            output = self.wn_block(k, audio_0, spect, taps)
            s = output[:, :, n_half:]
            b = output[:, :, :n_half]
            audio_1 = (audio_1 - b) / torch.exp(s)
            audio = torch.cat([audio_0, audio_1], dim=2)
            audio = audio @ self.w_inv[k]          # invertible_conv.py:49-51
            if k % hp.n_early_every == 0 and k > 0:
                if deterministic:
                    z_i = torch.zeros(B, Lg, hp.n_early_size, dtype=self.dtype)
                elif z is not None:
                    z_i = z[:, :, :hp.n_early_size]
                    z = z[:, :, hp.n_early_size:hp.n_group]
                else:
                    z_i = torch.randn(B, Lg, hp.n_early_size, generator=generator).to(self.dtype)
                audio = torch.cat([sigma * z_i, audio], dim=2)
            if taps is not None:
                taps[f"flow{k}/audio"] = audio
        return audio.reshape(B, -1)

    __call__ = infer


# -- independent second implementation (NVIDIA channels-first formulation via torch conv ops) ----
def infer_conv_ops(hp: WaveGlowHParams, weights, mel, z, sigma=1.0, dtype=torch.float32):
    """Same function computed with torch's conv1d / conv_transpose1d on channels-first tensors and
    weights transposed [2,1,0] (models/weights_converter.py:252-271). Shares no arithmetic code with
    OracleWaveGlow; tests assert the two agree."""
    w = {k: _t(v, dtype) for k, v in weights.items() if not k.startswith("__")}
    cf = lambda name: w[name].permute(2, 1, 0).contiguous()       # keras [k,in,out] -> torch [out,in,k]
    x = _t(mel, dtype).permute(0, 2, 1)
    up_w = w["upsample/kernel"].permute(2, 1, 0).contiguous()      # keras [k,out,in] -> torch [in,out,k]
    spect = F.conv_transpose1d(x, up_w, w["upsample/bias"], stride=HOP)
    spect = spect[:, :, :-(UPSAMPLE_K - HOP)]
    B = spect.shape[0]
    spect = spect.unfold(2, hp.n_group, hp.n_group).permute(0, 2, 1, 3)
    spect = spect.contiguous().view(B, spect.size(1), -1).permute(0, 2, 1)   # [B, 640, L]
    z = _t(z, dtype).permute(0, 2, 1)
    n_rem = hp.n_remaining_channels
    audio = sigma * z[:, :n_rem]
    z = z[:, n_rem:]
    C = hp.n_channels
    for k in reversed(range(hp.n_flows)):
        n_half = audio.size(1) // 2
        a0, a1 = audio[:, :n_half], audio[:, n_half:]
        p = f"block-{k}/"
        h = F.conv1d(a0, cf(p + "start_conv/kernel"), w[p + "start_conv/bias"])
        out = None
        for i in range(hp.n_layers):
            d = 2 ** i
            in_act = F.conv1d(h, cf(p + f"in_conv-{i}/kernel"), w[p + f"in_conv-{i}/bias"], dilation=d, padding=d)
            in_act = in_act + F.conv1d(spect, cf(p + f"cond_layer-{i}/kernel"), w[p + f"cond_layer-{i}/bias"])
            acts = torch.tanh(in_act[:, :C]) * torch.sigmoid(in_act[:, C:])
            rs = F.conv1d(acts, cf(p + f"res_skip_conv-{i}/kernel"), w[p + f"res_skip_conv-{i}/bias"])
            if i < hp.n_layers - 1:
                h = h + rs[:, :C]
                skip = rs[:, C:]
            else:
                skip = rs
            out = skip if out is None else out + skip
        out = F.conv1d(out, cf(p + "end_conv/kernel"), w[p + "end_conv/bias"])
        s, b = out[:, n_half:], out[:, :n_half]
        a1 = (a1 - b) / torch.exp(s)
        audio = torch.cat([a0, a1], 1)
        Wm = w[f"invertible_conv-{k}/conv/kernel"][0].to(torch.float32).transpose(0, 1)
        Winv = torch.linalg.inv(Wm).to(dtype)
        audio = F.conv1d(audio, Winv[:, :, None])
        if k % hp.n_early_every == 0 and k > 0:
            audio = torch.cat([sigma * z[:, :hp.n_early_size], audio], 1)
            z = z[:, hp.n_early_size:]
    return audio.permute(0, 2, 1).contiguous().view(B, -1)

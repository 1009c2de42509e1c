"""Weight import: NVIDIA-WaveGlow (PyTorch) state_dict -> the Keras-layout weight file the engine loads.

Mirrors what the reference does when it converts the torch.hub checkpoint: Conv1d `[out, in, k]` and
ConvTranspose1d `[in, out, k]` both become Keras `[k, in, out]` / `[k, out, in]` by the `[2, 1, 0]` transpose
(models/weights_converter.py:252-271), weight-norm is removed first (architectures/waveglow_arch.py:327-335
calls `remove_weightnorm`), and the fused `WN.k.cond_layer` of newer checkpoints is split per layer exactly as
`WaveglowBlock(fused=True)` slices it (waveglow_arch.py:119-121). No torch needed: values may be numpy arrays
or anything `np.asarray` accepts (call `.numpy()` / `.cpu()` on tensors first).
"""
from __future__ import annotations

import re

import numpy as np

from .weights import WaveGlowHParams, check_weights


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def _materialise_weight_norm(sd):
    """weight = g * v / ||v|| (norm over all dims but 0), as torch.nn.utils.remove_weight_norm does."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            v = _np(v)
            g = _np(sd[base + ".weight_g"])
            norm = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(axis=1)).reshape((-1,) + (1,) * (v.ndim - 1))
            out[base + ".weight"] = g.reshape(norm.shape) * v / norm
        elif k.endswith(".weight_g"):
            continue
        else:
            out[k] = _np(v)
    return out


def from_nvidia_state_dict(state_dict, *, n_early_every=4, n_early_size=2):
    """Returns (WaveGlowHParams, {keras name: float32 array})."""
    sd = _materialise_weight_norm(state_dict)
    t = lambda w: np.ascontiguousarray(np.transpose(w, (2, 1, 0)))      # torch conv layouts -> keras
    n_flows = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"convinv\.(\d+)\.", k)] if m)
    n_layers = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"WN\.0\.in_layers\.(\d+)\.", k)] if m)
    C = sd["WN.0.start.weight"].shape[0]
    n_group = sd["convinv.0.weight"].shape[0]
    n_mel = sd["upsample.weight"].shape[0]
    ks = sd["WN.0.in_layers.0.weight"].shape[2]
    hp = WaveGlowHParams(n_mel_channels=n_mel, n_flows=n_flows, n_group=n_group, n_early_every=n_early_every,
                         n_early_size=n_early_size, n_layers=n_layers, n_channels=C, kernel_size=ks)
    w = {"upsample/kernel": t(sd["upsample.weight"]), "upsample/bias": sd["upsample.bias"]}
    for k in range(n_flows):
        p, q = f"block-{k}/", f"WN.{k}."
        w[f"invertible_conv-{k}/conv/kernel"] = t(sd[f"convinv.{k}.weight"])
        w[p + "start_conv/kernel"], w[p + "start_conv/bias"] = t(sd[q + "start.weight"]), sd[q + "start.bias"]
        w[p + "end_conv/kernel"], w[p + "end_conv/bias"] = t(sd[q + "end.weight"]), sd[q + "end.bias"]
        fused = q + "cond_layer.weight" in sd
        for i in range(n_layers):
            w[p + f"in_conv-{i}/kernel"] = t(sd[q + f"in_layers.{i}.weight"])
            w[p + f"in_conv-{i}/bias"] = sd[q + f"in_layers.{i}.bias"]
            if fused:
                sl = slice(i * 2 * C, (i + 1) * 2 * C)
                w[p + f"cond_layer-{i}/kernel"] = t(sd[q + "cond_layer.weight"][sl])
                w[p + f"cond_layer-{i}/bias"] = sd[q + "cond_layer.bias"][sl]
            else:
                w[p + f"cond_layer-{i}/kernel"] = t(sd[q + f"cond_layers.{i}.weight"])
                w[p + f"cond_layer-{i}/bias"] = sd[q + f"cond_layers.{i}.bias"]
            w[p + f"res_skip_conv-{i}/kernel"] = t(sd[q + f"res_skip_layers.{i}.weight"])
            w[p + f"res_skip_conv-{i}/bias"] = sd[q + f"res_skip_layers.{i}.bias"]
    w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}
    check_weights(hp, w)
    return hp, w


def to_nvidia_state_dict(hp: WaveGlowHParams, weights, *, fused_cond=False):
    """Inverse mapping (used by the tests and to hand weights to NVIDIA's torch implementation)."""
    t = lambda w: np.ascontiguousarray(np.transpose(np.asarray(w, dtype=np.float32), (2, 1, 0)))
    sd = {"upsample.weight": t(weights["upsample/kernel"]), "upsample.bias": np.asarray(weights["upsample/bias"])}
    for k in range(hp.n_flows):
        p, q = f"block-{k}/", f"WN.{k}."
        sd[f"convinv.{k}.weight"] = t(weights[f"invertible_conv-{k}/conv/kernel"])
        sd[q + "start.weight"], sd[q + "start.bias"] = t(weights[p + "start_conv/kernel"]), weights[p + "start_conv/bias"]
        sd[q + "end.weight"], sd[q + "end.bias"] = t(weights[p + "end_conv/kernel"]), weights[p + "end_conv/bias"]
        cw, cb = [], []
        for i in range(hp.n_layers):
            sd[q + f"in_layers.{i}.weight"] = t(weights[p + f"in_conv-{i}/kernel"])
            sd[q + f"in_layers.{i}.bias"] = weights[p + f"in_conv-{i}/bias"]
            sd[q + f"res_skip_layers.{i}.weight"] = t(weights[p + f"res_skip_conv-{i}/kernel"])
            sd[q + f"res_skip_layers.{i}.bias"] = weights[p + f"res_skip_conv-{i}/bias"]
            if fused_cond:
                cw.append(t(weights[p + f"cond_layer-{i}/kernel"]))
                cb.append(np.asarray(weights[p + f"cond_layer-{i}/bias"]))
            else:
                sd[q + f"cond_layers.{i}.weight"] = t(weights[p + f"cond_layer-{i}/kernel"])
                sd[q + f"cond_layers.{i}.bias"] = weights[p + f"cond_layer-{i}/bias"]
        if fused_cond:
            sd[q + "cond_layer.weight"] = np.concatenate(cw, axis=0)
            sd[q + "cond_layer.bias"] = np.concatenate(cb, axis=0)
    return {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}

"""WaveGlow weight files: hparams, Keras-layout arrays under Keras variable names, seeded generator.

The engine and the oracle share ONE weight file (``.npz``). Arrays are stored exactly as the
reference's Keras layers hold them, so a converted trained checkpoint drops in unchanged:

  upsample/kernel                          [1024, 80(out), 80(in)]   Conv1DTranspose  (waveglow_arch.py:196-198)
  upsample/bias                            [80]
  block-{k}/start_conv/{kernel,bias}       [1, n_half, C], [C]       (waveglow_arch.py:58)
  block-{k}/end_conv/{kernel,bias}         [1, C, 2*n_half], [2*n_half]   (:62-64)
  block-{k}/in_conv-{i}/{kernel,bias}      [3, C, 2C], [2C]          (:73-79)
  block-{k}/cond_layer-{i}/{kernel,bias}   [1, 640, 2C], [2C]        (:81)
  block-{k}/res_skip_conv-{i}/{kernel,bias} [1, C, 2C] (i<n_layers-1) / [1, C, C]   (:83-88)
  invertible_conv-{k}/conv/kernel          [1, c, c]                 (invertible_conv.py:24-32)

Conv1D kernels are ``[k, in, out]``, Conv1DTranspose is ``[k, out, in]`` -- the layouts pinned by
the reference's torch<->keras converter (models/weights_converter.py:252-271).
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import dataclass, asdict

import numpy as np

N_MEL = 80          # base_audio_model.py / TacotronSTFT: 80 mel bins
HOP = 256           # upsample stride (waveglow_arch.py:197)
UPSAMPLE_K = 1024   # upsample kernel size (waveglow_arch.py:197)
SAMPLE_RATE = 22050


@dataclass(frozen=True)
class WaveGlowHParams:
    """Constructor arguments of ``architectures.WaveGlow`` (waveglow_arch.py:164-181)."""
    n_mel_channels: int = N_MEL
    n_flows: int = 12
    n_group: int = 8
    n_early_every: int = 4
    n_early_size: int = 2
    n_layers: int = 8
    n_channels: int = 256
    kernel_size: int = 3

    def flow_channels(self):
        """Per flow k: (n_half, n_remaining) following waveglow_arch.py:202-223."""
        n_half = self.n_group // 2
        n_rem = self.n_group
        out = []
        for k in range(self.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_half -= self.n_early_size // 2
                n_rem -= self.n_early_size
            out.append((n_half, n_rem))
        return out

    @property
    def n_remaining_channels(self):
        return self.flow_channels()[-1][1]

    def to_json(self):
        return json.dumps(asdict(self), sort_keys=True)

    @staticmethod
    def from_json(s):
        return WaveGlowHParams(**json.loads(s))


def weight_names(hp: WaveGlowHParams):
    """Ordered list of (name, shape) for every variable of the model."""
    C, L = hp.n_channels, hp.n_layers
    spect_ch = hp.n_mel_channels * hp.n_group
    names = [("upsample/kernel", (UPSAMPLE_K, hp.n_mel_channels, hp.n_mel_channels)),
             ("upsample/bias", (hp.n_mel_channels,))]
    for k, (n_half, n_rem) in enumerate(hp.flow_channels()):
        names.append((f"invertible_conv-{k}/conv/kernel", (1, n_rem, n_rem)))
        p = f"block-{k}/"
        names.append((p + "start_conv/kernel", (1, n_half, C)))
        names.append((p + "start_conv/bias", (C,)))
        names.append((p + "end_conv/kernel", (1, C, 2 * n_half)))
        names.append((p + "end_conv/bias", (2 * n_half,)))
        for i in range(L):
            rs = 2 * C if i < L - 1 else C
            names.append((p + f"in_conv-{i}/kernel", (hp.kernel_size, C, 2 * C)))
            names.append((p + f"in_conv-{i}/bias", (2 * C,)))
            names.append((p + f"cond_layer-{i}/kernel", (1, spect_ch, 2 * C)))
            names.append((p + f"cond_layer-{i}/bias", (2 * C,)))
            names.append((p + f"res_skip_conv-{i}/kernel", (1, C, rs)))
            names.append((p + f"res_skip_conv-{i}/bias", (rs,)))
    return names


def _glorot_uniform(rng, shape):
    # Keras default kernel initializer. For Conv1D [k, in, out]: fan_in = k*in, fan_out = k*out.
    # For Conv1DTranspose [k, out, in] Keras computes fans on the stored shape the same way.
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def generate_weights(hp: WaveGlowHParams, seed: int = 1234, *, end_std: float = 0.02,
                     bias_std: float = 0.0, convinv: str = "orthogonal"):
    """Seeded random-init WaveGlow (BASELINE.json north_star).

    Kernels: Glorot-uniform (Keras default); the zero-initialised end conv (waveglow_arch.py:62-64)
    is perturbed to N(0, end_std) so the coupling is non-trivial; the invertible 1x1 kernels are
    random orthogonal with det>0 (NVIDIA's own init) unless ``convinv='glorot'``. ``bias_std>0``
    gives every bias a N(0, bias_std) value (Keras default is zeros) so bias handling is exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for name, shape in weight_names(hp):
        if name.endswith("/bias"):
            w[name] = (rng.standard_normal(shape) * bias_std).astype(np.float32)
        elif "end_conv" in name:
            w[name] = (rng.standard_normal(shape) * end_std).astype(np.float32)
        elif "invertible_conv" in name:
            c = shape[-1]
            if convinv == "orthogonal":
                q, _ = np.linalg.qr(rng.standard_normal((c, c)))
                if np.linalg.det(q) < 0:
                    q[:, 0] = -q[:, 0]
                w[name] = q.astype(np.float32)[None]
            else:
                w[name] = _glorot_uniform(rng, shape)
        else:
            w[name] = _glorot_uniform(rng, shape)
    return w


def weights_digest(weights) -> str:
    h = hashlib.sha256()
    for name in sorted(weights):
        if name.startswith("__"):
            continue
        h.update(name.encode())
        h.update(np.ascontiguousarray(weights[name], dtype=np.float32).tobytes())
    return h.hexdigest()


def save_weights(path, hp: WaveGlowHParams, weights):
    arrays = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
    arrays["__hparams__"] = np.frombuffer(hp.to_json().encode(), dtype=np.uint8)
    np.savez(path, **arrays)


def load_weights(path):
    """Returns (hparams, {name: float32 array}). Validates names and shapes against the topology.
    `.npz`: this package's own container; `.h5` / `.weights.h5`: a Keras 3 weights file of the reference's model
    (what `CheckpointManager.load` restores, custom_train_objects/checkpoint_manager.py:169-216)."""
    if str(path).endswith(".h5"):
        from .convert import from_keras_weights_h5
        return from_keras_weights_h5(path)
    with np.load(path) as f:
        if "__hparams__" not in f.files:
            raise ValueError(f"{path}: not a WaveGlow weight file (no __hparams__ entry)")
        hp = WaveGlowHParams.from_json(bytes(f["__hparams__"]).decode())
        w = {k: np.asarray(f[k], dtype=np.float32) for k in f.files if k != "__hparams__"}
    check_weights(hp, w)
    return hp, w


def check_weights(hp, w):
    for name, shape in weight_names(hp):
        if name not in w:
            raise ValueError(f"missing weight {name}")
        if tuple(w[name].shape) != tuple(shape):
            raise ValueError(f"weight {name}: shape {tuple(w[name].shape)} != expected {tuple(shape)}")


def synthetic_mel(rng, B, T, n_mel=N_MEL):
    """Log-mel statistics of the reference's TacotronSTFT golden (tests/__reproduction/
    stft-TacotronSTFT.npy: min -11.513 = log(1e-5), max ~1.16, mean -5.2)."""
    return np.clip(rng.normal(-5.2, 2.0, size=(B, T, n_mel)), -11.513, 1.2).astype(np.float32)


def synthetic_inputs(seed, B, T, hp: WaveGlowHParams = WaveGlowHParams()):
    rng = np.random.Generator(np.random.PCG64(seed))
    mel = synthetic_mel(rng, B, T, hp.n_mel_channels)
    z = rng.standard_normal((B, T * HOP // hp.n_group, hp.n_group)).astype(np.float32)
    return mel, z

"""The plugin boundary: a `Runtime` for the reference's `runtime=` switch.

Mirrors utils/keras/runtimes/runtime.py:19-81 (class `Runtime`: per-path engine cache,
`__call__`, static `load_engine`) and utils/keras/runtimes/__init__.py:23-45 (`build_runtime`,
`_runtimes`). In the reference tree a maintainer registers this class under a key
(`_runtimes['b200'] = B200WaveGlowRuntime`); `BaseModel.compiled_infer` then
returns the Runtime object itself (models/interfaces/base_model.py:366-370) and
`models.tts.WaveGlow.infer` calls it as `self.compiled_infer(mel, **kwargs)[:, :T*256]`
(models/tts/waveglow.py:82).
"""
from __future__ import annotations

import os
from abc import ABCMeta, abstractmethod

import numpy as np

from .engine import WaveGlowEngine
from .weights import HOP


class Runtime(metaclass=ABCMeta):
    """Same contract as the reference's Runtime ABC (runtime.py:19-41)."""
    _engines = {}

    def __init__(self, path, *, engine=None, reload=False, **kwargs):
        if engine is None:
            key = self._engine_key(path, **kwargs)
            if key not in self._engines or reload:
                self._engines[key] = self.load_engine(path, **kwargs)
            engine = self._engines[key]
        self.path = path
        self.engine = engine

    @staticmethod
    def _engine_key(path, **kwargs):
        return path

    def __repr__(self):
        return '<{} path={}>'.format(self.__class__.__name__, self.path)

    @abstractmethod
    def __call__(self, *args, **kwargs):
        """ Performs custom runtime inference """

    @staticmethod
    @abstractmethod
    def load_engine(path, **kwargs):
        """ Loads the custom runtime engine """

    # Optional builders of the reference ABC (runtime.py:44-81). A WaveGlow engine is built from a weight file, not
    # traced from a framework function, so they refuse with the reference's own messages.
    @classmethod
    def build_from(cls, function, path, overwrite=False, **kwargs):
        import os
        if os.path.exists(path) and not overwrite:
            return cls(path, **kwargs)
        if isinstance(function, str):
            if function.endswith('.onnx'):
                return cls.from_onnx(function, path, **kwargs)
            elif function.endswith('.pth'):
                return cls.from_torch(function, path, **kwargs)
            elif os.path.isdir(path):
                return cls.from_tensorflow(function, path, **kwargs)
            raise NotImplementedError('Invalid path : {}'.format(path))
        raise NotImplementedError()

    @classmethod
    def from_tensorflow(cls, function, path, **kwargs):
        raise NotImplementedError('{} cannot be initialized from `tf.function`'.format(cls.__name__))

    @classmethod
    def from_torch(cls, function, path, **kwargs):
        raise NotImplementedError('{} cannot be initialized from `torch.compile`'.format(cls.__name__))

    @classmethod
    def from_onnx(cls, onnx_path, path, **kwargs):
        raise NotImplementedError('{} cannot be initialized from `ONNX`'.format(cls.__name__))


class B200WaveGlowRuntime(Runtime):
    """WaveGlow vocoder runtime on one B200.

    `path` is a WaveGlow weight file (text_to_speech_b200/weights.py format, Keras layouts).
    Call contract = architectures.WaveGlow.infer (waveglow_arch.py:241-244):
        runtime(inputs, z=None, sigma=1.0, deterministic=False, **ignored) -> float32 [B, 256*T]
    `inputs` float32 [B,T,80] (or [T,80]); numpy / anything array-like -> numpy out (host path:
    pinned staging, H2D, kernels, D2H, synchronous return, like tensorrt_runtime.py:193-210);
    a CUDA torch tensor -> CUDA torch tensor out (device path, async on the current stream).
    Unknown keyword arguments are ignored: `graph_compile`'s signature filter (compile.py:68-71) is
    bypassed for non-keras runtimes, so callers' extras (directory=, display=, ...) do arrive here.
    """

    def __init__(self, path, *, engine=None, reload=False, mode="bf16", device=0, seed=None, **kwargs):
        super().__init__(path, engine=engine, reload=reload, mode=mode, device=device)
        self.mode, self.device = mode, device
        self._pinned = {}
        self._gen = None
        self._seed = seed

    @staticmethod
    def _engine_key(path, mode="bf16", device=0, **_):
        return (os.path.abspath(path), mode, int(device))

    @staticmethod
    def load_engine(path, mode="bf16", device=0, **_):
        return WaveGlowEngine.from_file(path, mode=mode, device=device)

    # pinned host staging, re-allocated when shapes grow (tensorrt_runtime.py:143-177)
    def _pin(self, name, shape):
        import torch
        n = int(np.prod(shape))
        buf = self._pinned.get(name)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=torch.float32).pin_memory()
            self._pinned[name] = buf
        return buf[:n].view(*shape)

    def _noise(self, B, Lg, n_group, dev):
        import torch
        if self._gen is None:
            self._gen = torch.Generator(device=dev)
            if self._seed is not None:
                self._gen.manual_seed(int(self._seed))
        return torch.randn(B, Lg, n_group, generator=self._gen, device=dev, dtype=torch.float32)

    def __call__(self, inputs, z=None, sigma=1.0, deterministic=False, **_ignored):
        import torch
        eng = self.engine
        dev = torch.device("cuda", eng.device)
        on_device = isinstance(inputs, torch.Tensor) and inputs.is_cuda
        if on_device:
            mel = inputs.to(torch.float32)
            if mel.dim() == 2:
                mel = mel[None]
        else:
            mel_np = inputs.detach().cpu().numpy() if isinstance(inputs, torch.Tensor) else np.asarray(inputs)
            mel_np = np.asarray(mel_np, dtype=np.float32)
            if mel_np.ndim == 2:
                mel_np = mel_np[None]
            if mel_np.ndim != 3 or mel_np.shape[2] != eng.hp.n_mel_channels:
                raise ValueError(f"inputs must be [B,T,{eng.hp.n_mel_channels}] (channels-last mel), got {mel_np.shape}")
            pm = self._pin("mel", mel_np.shape)
            pm.copy_(torch.from_numpy(np.ascontiguousarray(mel_np)))
            mel = pm.to(dev, non_blocking=True)
        B, T = int(mel.shape[0]), int(mel.shape[1])
        Lg = T * HOP // eng.hp.n_group
        z_dev = None
        if not deterministic:
            if z is None:
                z_dev = self._noise(B, Lg, eng.hp.n_group, dev)   # keras.random.normal stand-in (waveglow_arch.py:272,301)
            elif isinstance(z, torch.Tensor) and z.is_cuda:
                z_dev = z.to(torch.float32)
            else:
                z_np = z.detach().cpu().numpy() if isinstance(z, torch.Tensor) else np.asarray(z)
                z_np = np.ascontiguousarray(z_np, dtype=np.float32)
                if z_np.shape != (B, Lg, eng.hp.n_group):
                    raise ValueError(f"z must be [{B},{Lg},{eng.hp.n_group}], got {z_np.shape}")
                pz = self._pin("z", z_np.shape)
                pz.copy_(torch.from_numpy(z_np))
                z_dev = pz.to(dev, non_blocking=True)
        out = eng.infer_device(mel, z_dev, sigma=float(sigma), deterministic=bool(deterministic))
        if on_device:
            return out
        po = self._pin("out", (B, T * HOP))
        po.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return po.numpy().copy()


# ---- registry, same shape as utils/keras/runtimes/__init__.py:23-45 --------------------------------
_runtimes = {
    'b200': B200WaveGlowRuntime,
}


def build_runtime(runtime, path, *args, **kwargs):
    if runtime not in _runtimes:
        raise ValueError('Unsupported runtime !\n  Accepted : {}\n  Got : {}'.format(
            tuple(_runtimes.keys()), runtime
        ))
    return _runtimes[runtime](path, *args, **kwargs)

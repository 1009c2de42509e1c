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


class MirrorRuntime(metaclass=ABCMeta):
    """Same contract as the reference's Runtime ABC (runtime.py:19-41); the base class used when this package
    runs outside the reference tree (see `Runtime` below)."""
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


def _reference_runtime_abc():
    """Inside the reference tree (`utils.keras.runtimes` importable) the plugin derives from the reference's OWN
    `Runtime` ABC, so `isinstance(rt, Runtime)` checks of the host application hold; elsewhere the mirror is used."""
    try:
        from utils.keras.runtimes.runtime import Runtime as ref            # noqa: PLC0415
        if all(hasattr(ref, a) for a in ("load_engine", "build_from", "_engines")):
            return ref
    except Exception:
        pass
    return None


Runtime = _reference_runtime_abc() or MirrorRuntime


class _Graph:
    """One captured wg_infer launch sequence with its static device buffers (mel, z, out, scratch)."""
    __slots__ = ("graph", "mel", "z", "out", "ws")


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

    Extensions over the reference call (all optional):
      lengths=[T_0, ..]   per-utterance frame counts of a padded batch (wg_infer_ragged): every utterance is
                          computed as if passed alone, which is what the reference's one-sentence-at-a-time loop
                          does (models/tts/tacotron2.py:154-191); the tail of each waveform row is zero.
      copy_outputs        host path: True returns a fresh array per call; False (default) returns a view of a
                          runtime-owned pinned buffer that stays valid for `output_ring - 1` further calls (the
                          TensorRT runtime precedent returns views that the very next call overwrites,
                          tensorrt_runtime.py:208-210).
      out=tensor          device path: write the waveform into this caller-owned [B, 256*T] float32 CUDA tensor.
      graph_max_frames    calls with B*T <= this many frames replay a CUDA graph captured per (B, T, sigma,
                          deterministic, lengths) instead of launching ~120 kernels one by one; 0 disables.
                          `precompile()` captures them ahead of time (the analogue of
                          Tacotron2.precompile_for_stream, models/tts/tacotron2.py:354-356).
    """

    _b200_engines = {}

    def __init__(self, path, *, engine=None, reload=False, mode="bf16", device=0, seed=None, copy_outputs=False,
                 output_ring=2, graph_max_frames=2048, max_graphs=16, **kwargs):
        if engine is None:      # engine cache keyed on (path, mode, device); the ABC's own cache is keyed on the path alone
            key = self._engine_key(path, mode=mode, device=device)
            if key not in B200WaveGlowRuntime._b200_engines or reload:
                B200WaveGlowRuntime._b200_engines[key] = self.load_engine(path, mode=mode, device=device)
            engine = B200WaveGlowRuntime._b200_engines[key]
        super().__init__(path, engine=engine)
        self.mode, self.device = mode, device
        self._pinned = {}
        self._out_ring, self._out_next = [None] * max(1, int(output_ring)), 0
        self.copy_outputs = bool(copy_outputs)
        self.graph_max_frames, self.max_graphs = int(graph_max_frames), int(max_graphs)
        self._graphs = {}          # key -> _Graph, insertion order = LRU order
        self.graph_replays = 0
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

    def _pin_out(self, shape):
        import torch
        n = int(np.prod(shape))
        i = self._out_next
        self._out_next = (i + 1) % len(self._out_ring)
        buf = self._out_ring[i]
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=torch.float32).pin_memory()
            self._out_ring[i] = buf
        return buf[:n].view(*shape)

    def _noise(self, B, Lg, n_group, dev):
        import torch
        if self._gen is None:
            self._gen = torch.Generator(device=dev)
            if self._seed is not None:
                self._gen.manual_seed(int(self._seed))
        return torch.randn(B, Lg, n_group, generator=self._gen, device=dev, dtype=torch.float32)

    # ---- CUDA graphs for small calls (the reference's call pattern is one sentence at a time, B = 1) -------------
    def _graph_for(self, B, T, sigma, deterministic, lengths):
        """Returns the captured graph for this call signature (capturing it on first use), or None."""
        import torch
        if self.graph_max_frames <= 0 or B * T > self.graph_max_frames:
            return None
        key = (B, T, float(sigma), bool(deterministic), None if lengths is None else tuple(int(x) for x in lengths))
        g = self._graphs.pop(key, None)
        if g is None:
            eng = self.engine
            dev = torch.device("cuda", eng.device)
            g = _Graph()
            g.mel = torch.zeros(B, T, eng.hp.n_mel_channels, dtype=torch.float32, device=dev)
            g.z = None if deterministic else torch.zeros(B, T * HOP // eng.hp.n_group, eng.hp.n_group, dtype=torch.float32, device=dev)
            g.out = torch.empty(B, T * HOP, dtype=torch.float32, device=dev)
            need = eng.workspace_bytes(B, T) if lengths is None else eng.workspace_bytes_ragged(B, T, eng._lengths(lengths, B, T))
            g.ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
            saved, eng._ws = eng._ws, g.ws          # the captured launches must keep pointing at THIS scratch
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):        # one eager run first: lazy module loading must not happen under capture
                    eng.infer_device(g.mel, g.z, sigma, deterministic, out=g.out, lengths=lengths)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                g.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g.graph):
                    eng.infer_device(g.mel, g.z, sigma, deterministic, out=g.out, lengths=lengths)
            finally:
                eng._ws = saved
            while len(self._graphs) >= max(1, self.max_graphs):      # LRU: the oldest entry goes
                self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = g
        return g

    def precompile(self, shapes=None, *, multiples=(64, 128), max_frames=None, sigma=1.0, deterministic=False):
        """Pre-sizes staging buffers / scratch and pre-captures the CUDA graphs for the given [(B, T), ..] call shapes,
        so the first real call of a stream pays nothing. Default shapes: B = 1 with T every multiple of `multiples`
        up to `max_frames` (graph_max_frames) -- the padded lengths the reference warms up with
        `padding_multiple` 64 and 128 (models/tts/tacotron2.py:354-356)."""
        import torch
        if shapes is None:
            top = int(max_frames or self.graph_max_frames)
            shapes = sorted({(1, t) for m in multiples for t in range(int(m), top + 1, int(m))})
        eng = self.engine
        done = []
        for B, T in shapes:
            if self._graph_for(int(B), int(T), sigma, deterministic, None) is None:
                eng._workspace(int(B), int(T))
            self._pin("mel", (B, T, eng.hp.n_mel_channels))
            self._pin("z", (B, T * HOP // eng.hp.n_group, eng.hp.n_group))
            for _ in self._out_ring:
                self._pin_out((B, T * HOP))
            done.append((int(B), int(T)))
        torch.cuda.synchronize(torch.device("cuda", eng.device))
        return done

    def __call__(self, inputs, z=None, sigma=1.0, deterministic=False, lengths=None, out=None, **_ignored):
        import torch
        eng = self.engine
        dev = torch.device("cuda", eng.device)
        on_device = isinstance(inputs, torch.Tensor) and inputs.is_cuda
        if on_device:
            mel = inputs.to(torch.float32)
            if mel.dim() == 2:
                mel = mel[None]
            mel_src = mel
        else:
            mel_np = inputs.detach().cpu().numpy() if isinstance(inputs, torch.Tensor) else np.asarray(inputs)
            mel_np = np.asarray(mel_np, dtype=np.float32)
            if mel_np.ndim == 2:
                mel_np = mel_np[None]
            if mel_np.ndim != 3 or mel_np.shape[2] != eng.hp.n_mel_channels:
                raise ValueError(f"inputs must be [B,T,{eng.hp.n_mel_channels}] (channels-last mel), got {mel_np.shape}")
            mel_src = self._pin("mel", mel_np.shape)
            mel_src.copy_(torch.from_numpy(np.ascontiguousarray(mel_np)))
        B, T = int(mel_src.shape[0]), int(mel_src.shape[1])
        if B == 0 or T == 0:
            raise ValueError(f"inputs must hold at least one frame, got shape {tuple(mel_src.shape)}")
        Lg = T * HOP // eng.hp.n_group
        z_src = None
        if not deterministic:
            if z is None:
                z_src = self._noise(B, Lg, eng.hp.n_group, dev)   # keras.random.normal stand-in (waveglow_arch.py:272,301)
            elif isinstance(z, torch.Tensor) and z.is_cuda:
                z_src = z.to(torch.float32)
                if tuple(z_src.shape) != (B, Lg, eng.hp.n_group):
                    raise ValueError(f"z must be [{B},{Lg},{eng.hp.n_group}], got {tuple(z_src.shape)}")
            else:
                z_np = z.detach().cpu().numpy() if isinstance(z, torch.Tensor) else np.asarray(z)
                z_np = np.ascontiguousarray(z_np, dtype=np.float32)
                if z_np.shape != (B, Lg, eng.hp.n_group):
                    raise ValueError(f"z must be [{B},{Lg},{eng.hp.n_group}], got {z_np.shape}")
                z_src = self._pin("z", z_np.shape)
                z_src.copy_(torch.from_numpy(z_np))
        if lengths is not None:
            lengths = [int(x) for x in lengths]
        _out_arg = out          # device path only: caller-owned [B, 256 T] float32 CUDA tensor to write into
        g = self._graph_for(B, T, sigma, deterministic, lengths)
        if g is not None:
            g.mel.copy_(mel_src, non_blocking=True)
            if z_src is not None:
                g.z.copy_(z_src, non_blocking=True)
            g.graph.replay()
            self.graph_replays += 1
            out = g.out
        else:
            mel = mel_src if mel_src.is_cuda else mel_src.to(dev, non_blocking=True)
            z_dev = None if z_src is None else (z_src if z_src.is_cuda else z_src.to(dev, non_blocking=True))
            out = eng.infer_device(mel, z_dev, sigma=float(sigma), deterministic=bool(deterministic), lengths=lengths,
                                   out=out if on_device else None)
        if on_device:
            if g is not None:       # a graph's output buffer is overwritten by its next replay
                return out.clone() if _out_arg is None else _out_arg.copy_(out)
            return out
        po = self._pin_out((B, T * HOP))
        po.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return po.numpy().copy() if self.copy_outputs else po.numpy()


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

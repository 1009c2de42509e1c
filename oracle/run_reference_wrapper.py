"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Executes the reference's OWN vocoder wrapper source, `models/tts/waveglow.py` (WaveGlow.infer: 2-D -> 3-D,
`[:, :256 T]` trim, float / int `win_len`, `hop_len`, `force_pad`, `batch`, `use_slice`, `max_win_len`, overlap
mid-point stitching; `_get_steps`), unmodified, with stand-ins only for what that file imports but the infer path
never computes with: `loggers` (its `timer` decorator -> identity), `utils.pad_to_multiple`, `utils.keras`
(`TensorSpec`; `ops.expand_dims / pad / concatenate / convert_to_numpy` on numpy arrays) and the `BaseAudioModel`
base class. `reference_wrapper_infer(vocoder, mel, runtime=..., **kw)` calls the reference's `infer` on a bare object
carrying the three attributes it reads (`compiled_infer`, `pad_mel_value`, `runtime`).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from .run_reference import REFERENCE_ROOT

_MOD = "_wg_reference_models.tts.waveglow"


def wrapper_reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "tts", "waveglow.py"))


def _np_ops():
    ops = types.SimpleNamespace()
    ops.expand_dims = lambda x, axis: np.expand_dims(np.asarray(x), axis)
    ops.pad = lambda x, pads, constant_values=0: np.pad(np.asarray(x), pads, constant_values=constant_values)
    ops.concatenate = lambda xs, axis=0: np.concatenate([np.asarray(x) for x in xs], axis=axis)
    ops.convert_to_numpy = lambda x: np.asarray(x)
    return ops


def load_reference_wrapper():
    if _MOD in sys.modules:
        return sys.modules[_MOD]
    if not wrapper_reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    saved = {}

    def put(name, mod):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = mod

    identity_timer = lambda fn=None, **kw: fn if callable(fn) else (lambda f: f)     # noqa: E731
    loggers = types.ModuleType("loggers")
    loggers.timer, loggers.Timer = identity_timer, None
    utils = types.ModuleType("utils")
    utils.__path__ = []
    utils.pad_to_multiple = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("XLA-only path"))
    utils_keras = types.ModuleType("utils.keras")
    utils_keras.TensorSpec = lambda *a, **k: None
    utils_keras.ops = _np_ops()
    utils.keras = utils_keras
    for name in ("_wg_reference_models", "_wg_reference_models.tts", "_wg_reference_models.interfaces"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        put(name, pkg)
    base = types.ModuleType("_wg_reference_models.interfaces.base_audio_model")
    base.BaseAudioModel = type("BaseAudioModel", (), {"audio_signature": None})
    put("_wg_reference_models.interfaces.base_audio_model", base)
    put("loggers", loggers), put("utils", utils), put("utils.keras", utils_keras)
    try:
        spec = importlib.util.spec_from_file_location(_MOD, os.path.join(REFERENCE_ROOT, "models", "tts", "waveglow.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[_MOD] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        for name in ("loggers", "utils", "utils.keras"):
            if saved[name] is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = saved[name]


def reference_wrapper_infer(vocoder, mel, runtime="b200", pad_mel_value=-11.0, **kwargs):
    mod = load_reference_wrapper()
    self = types.SimpleNamespace(compiled_infer=vocoder, pad_mel_value=pad_mel_value, runtime=runtime)
    return np.asarray(mod.WaveGlow.infer(self, mel, **kwargs))


def reference_get_steps(length, win_len, hop_len):
    return load_reference_wrapper()._get_steps(length, win_len, hop_len)

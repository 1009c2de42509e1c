"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Executes the reference's OWN WaveGlow source, unmodified, in the authoring container:
``/root/reference/architectures/waveglow_arch.py`` and ``architectures/layers/invertible_conv.py``
are loaded by path under a synthetic package (the real ``architectures/__init__.py`` drags in the
whole monorepo, which needs keras/tensorflow/matplotlib/librosa -- none installable here). If a
real ``keras`` is importable it is used (set KERAS_BACKEND=torch); otherwise the minimal shim in
oracle/keras_shim is put on sys.path. ``which_keras()`` tells which, and every fixture records it.

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or bench.py
calls this module; it is used by oracle/gen_golden.py and by the CPU tests that validate the
restatement (skipped when the reference tree is absent).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("WG_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "keras_shim")
_PKG = "_wg_reference_arch"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "architectures", "waveglow_arch.py"))


def _ensure_keras():
    try:
        import keras  # noqa: F401
    except ModuleNotFoundError:
        sys.path.insert(0, _SHIM_DIR)
        import keras  # noqa: F401
    return sys.modules["keras"]


def which_keras() -> str:
    keras = _ensure_keras()
    return "shim" if getattr(keras, "__version__", "").endswith("shim") else f"keras-{keras.__version__}"


def load_reference_arch():
    """Returns the module object of the reference's waveglow_arch.py (classes WaveGlow, WaveglowBlock)."""
    if _PKG + ".waveglow_arch" in sys.modules:
        return sys.modules[_PKG + ".waveglow_arch"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _ensure_keras()
    arch_dir = os.path.join(REFERENCE_ROOT, "architectures")
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = []           # synthetic package; submodules are registered by hand
    sys.modules[_PKG] = pkg

    def _load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    inv = _load(_PKG + ".layers.invertible_conv", os.path.join(arch_dir, "layers", "invertible_conv.py"))
    layers_pkg = types.ModuleType(_PKG + ".layers")
    layers_pkg.__path__ = []
    layers_pkg.Invertible1x1Conv = inv.Invertible1x1Conv
    sys.modules[_PKG + ".layers"] = layers_pkg
    pkg.layers = layers_pkg
    return _load(_PKG + ".waveglow_arch", os.path.join(arch_dir, "waveglow_arch.py"))


def build_reference_model(hp, weights):
    """Constructs the reference ``WaveGlow`` with ``hp`` and loads ``weights`` (Keras names/layouts)
    through the reference's own ``set_weights`` (which rebuilds every W^-1, waveglow_arch.py:308-310)."""
    arch = load_reference_arch()
    model = arch.WaveGlow(
        n_mel_channels=hp.n_mel_channels, n_flows=hp.n_flows, n_group=hp.n_group,
        n_early_every=hp.n_early_every, n_early_size=hp.n_early_size, n_layers=hp.n_layers,
        n_channels=hp.n_channels, kernel_size=hp.kernel_size, name="WaveGlow")
    if which_keras() == "shim":
        model.set_weights({k: v for k, v in weights.items() if not k.startswith("__")})
    else:   # real keras: positional list in model.weights order, matched by variable path
        ordered = []
        for v in model.weights:
            path = v.path
            key = path.split("/", 1)[1] if path.split("/", 1)[0].lower().startswith("wave") else path
            ordered.append(np.asarray(weights[key]))
        model.set_weights(ordered)
    return model


def reference_infer(hp, weights, mel, z=None, sigma=1.0, deterministic=False, fused=False):
    """Runs the reference's ``WaveGlow.infer`` (fp32) and returns a numpy [B, 256 T] waveform."""
    import torch
    model = build_reference_model(hp, weights)
    with torch.no_grad():
        mel_t = torch.as_tensor(np.asarray(mel), dtype=torch.float32)
        z_t = None if z is None else torch.as_tensor(np.asarray(z), dtype=torch.float32)
        out = model.infer(mel_t, z=z_t, sigma=sigma, deterministic=deterministic)
    return out.detach().cpu().numpy() if hasattr(out, "detach") else np.asarray(out)

"""text_to_speech_b200 -- B200-native WaveGlow vocoder inference behind the `runtime=` plugin
boundary of yui-mhcp/text_to_speech (models/tts/waveglow.py -> architectures/waveglow_arch.py)."""
from .weights import WaveGlowHParams, generate_weights, save_weights, load_weights, synthetic_inputs  # noqa: F401

__all__ = ["WaveGlowHParams", "generate_weights", "save_weights", "load_weights", "synthetic_inputs",
           "WaveGlowEngine", "B200WaveGlowRuntime", "build_runtime", "WaveGlow", "TacotronSTFT", "MelSTFT"]


def __getattr__(name):   # lazy: importing the package must not need torch / the CUDA library
    if name == "WaveGlowEngine":
        from .engine import WaveGlowEngine
        return WaveGlowEngine
    if name in ("B200WaveGlowRuntime", "build_runtime", "Runtime"):
        from . import runtime
        return getattr(runtime, name)
    if name == "WaveGlow":
        from .waveglow import WaveGlow
        return WaveGlow
    if name in ("TacotronSTFT", "MelSTFT"):
        from . import stft
        return getattr(stft, name)
    raise AttributeError(name)

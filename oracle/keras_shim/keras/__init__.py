"""Minimal Keras-3 API shim (torch-backed) -- TEST INFRASTRUCTURE ONLY.

Exists for one purpose: keras is not installable in the authoring container (no network), so the
reference's own WaveGlow source (architectures/waveglow_arch.py, layers/invertible_conv.py) could
not otherwise be executed. This package implements ONLY the symbols those two files touch, with
the semantics documented for Keras 3 (channels-last Conv1D / Conv1DTranspose with 'valid' /
'same' padding, keras.ops elementwise + shape ops). It is never on the product path and is only
ever put on sys.path by oracle/run_reference.py.
"""
from . import ops, layers, saving, random     # noqa: F401
from .layers import Layer, Model              # noqa: F401

__version__ = "3.shim"

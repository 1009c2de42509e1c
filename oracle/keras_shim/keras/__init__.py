"""Minimal Keras-3 API shim (torch-backed) -- TEST INFRASTRUCTURE ONLY.

Exists for one purpose: keras is not installable in the authoring container (no network), so the
reference's own WaveGlow source (architectures/waveglow_arch.py, layers/invertible_conv.py) could
not otherwise be executed; the same goes for the Tacotron2 DECODER (tacotron2_arch.py:143-212, 336-750,
layers/location_sensitive_attention.py, layers/custom_rnn_dropout_cell.py, hparams.py), loaded by
oracle/run_reference_taco.py. This package implements ONLY the symbols those files touch, with
the semantics documented for Keras 3 (channels-last Conv1D / Conv1DTranspose with 'valid' /
'same' padding, keras.ops elementwise + shape ops). It is never on the product path and is only
ever put on sys.path by oracle/run_reference.py.
"""
import contextlib

from . import ops, layers, saving, random, tree, backend     # noqa: F401
from .layers import Layer, Model, Sequential  # noqa: F401

__version__ = "3.shim"


@contextlib.contextmanager
def name_scope(name):
    yield

"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Executes the reference's OWN Tacotron2 DECODER source, unmodified, in the authoring container:

    architectures/tacotron2_arch.py            Tacotron2Prenet, Tacotron2DecoderCell, Tacotron2Decoder(.infer)
    architectures/layers/location_sensitive_attention.py   LocationLayer, LocationSensitiveAttention
    architectures/layers/custom_rnn_dropout_cell.py        CustomRNNDropoutCell
    architectures/hparams.py                   HParams

loaded by path under a synthetic package, over the Keras shim in oracle/keras_shim (or a real keras when
one is importable). What is NOT the reference's: the shim's reading of Keras' Dense / Conv1D / LSTMCell /
StackedRNNCells / keras.ops (third-party, un-pinned), and small stand-ins for modules the decoder never
executes (`utils`, `utils.keras`, `.simple_models`, `.current_blocks._get_var`, the embedding layers used
only by the ENCODER). The encoder and the postnet are built by the reference's generic `simple_cnn` factory
on the functional Keras API and are not covered: this pins the decoder loop -- the part implemented as CUDA
kernels (csrc/taco.cu) -- and nothing else.

The reference calls `Tacotron2Decoder.infer` at batch size 1 (models/tts/tacotron2.py:161-166; its final
`arange(T)[None] <= lengths` only broadcasts for B = 1), so `reference_decode` runs one utterance at a time.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from .run_reference import REFERENCE_ROOT, _ensure_keras, which_keras  # noqa: F401

_PKG = "_wg_reference_taco"

# our weight names (text_to_speech_b200/tacotron2.py) -> variable paths below the reference's Tacotron2Decoder
NAME_MAP = {
    "decoder/prenet/layer_0/kernel": "prenet/layer_0/kernel",
    "decoder/prenet/layer_1/kernel": "prenet/layer_1/kernel",
    "decoder/attention_rnn/kernel": "decoder_cell/attention_rnn/kernel",
    "decoder/attention_rnn/recurrent_kernel": "decoder_cell/attention_rnn/recurrent_kernel",
    "decoder/attention_rnn/bias": "decoder_cell/attention_rnn/bias",
    "decoder/lsa/query_layer/kernel": "decoder_cell/location_sensitive_attention/query_layer/kernel",
    "decoder/lsa/memory_layer/kernel": "decoder_cell/location_sensitive_attention/memory_layer/kernel",
    "decoder/lsa/value_layer/kernel": "decoder_cell/location_sensitive_attention/value_layer/kernel",
    "decoder/lsa/location_conv/kernel": "decoder_cell/location_sensitive_attention/location_layer/location_conv/kernel",
    "decoder/lsa/location_dense/kernel": "decoder_cell/location_sensitive_attention/location_layer/location_dense/kernel",
    "decoder/decoder_rnn/cell_0/kernel": "decoder_cell/decoder_rnn/cell_0/kernel",
    "decoder/decoder_rnn/cell_0/recurrent_kernel": "decoder_cell/decoder_rnn/cell_0/recurrent_kernel",
    "decoder/decoder_rnn/cell_0/bias": "decoder_cell/decoder_rnn/cell_0/bias",
    "decoder/linear_projection/kernel": "linear_projection/kernel",
    "decoder/linear_projection/bias": "linear_projection/bias",
    "decoder/gate_output/kernel": "gate_output/kernel",
    "decoder/gate_output/bias": "gate_output/bias",
}


def taco_reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "architectures", "tacotron2_arch.py"))


def _load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_taco():
    """Returns the module object of the reference's tacotron2_arch.py."""
    if _PKG + ".tacotron2_arch" in sys.modules:
        return sys.modules[_PKG + ".tacotron2_arch"]
    if not taco_reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _ensure_keras()
    arch_dir = os.path.join(REFERENCE_ROOT, "architectures")

    # stand-ins for top-level modules the decoder imports but never executes; removed again after loading
    stubs = {}
    utils = types.ModuleType("utils")
    utils.__path__ = []
    utils.pad_to_multiple = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("stub"))
    file_utils = types.ModuleType("utils.file_utils")
    file_utils.load_json = file_utils.dump_json = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("stub"))
    utils_keras = types.ModuleType("utils.keras")
    utils_keras.TensorSpec = lambda *a, **k: None          # only used as annotations (tacotron2_arch.py:867-874)
    utils_keras.ops = types.SimpleNamespace(is_int=lambda x: isinstance(x, int), is_float=lambda x: isinstance(x, float))
    utils.file_utils, utils.keras = file_utils, utils_keras
    for name, mod in (("utils", utils), ("utils.file_utils", file_utils), ("utils.keras", utils_keras)):
        stubs[name] = sys.modules.get(name)
        sys.modules[name] = mod
    try:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = []
        sys.modules[_PKG] = pkg
        hparams = _load(_PKG + ".hparams", os.path.join(arch_dir, "hparams.py"))
        pkg.hparams = hparams
        layers_pkg = types.ModuleType(_PKG + ".layers")
        layers_pkg.__path__ = []
        sys.modules[_PKG + ".layers"] = layers_pkg
        pkg.layers = layers_pkg
        drop = _load(_PKG + ".layers.custom_rnn_dropout_cell", os.path.join(arch_dir, "layers", "custom_rnn_dropout_cell.py"))
        lsa = _load(_PKG + ".layers.location_sensitive_attention",
                    os.path.join(arch_dir, "layers", "location_sensitive_attention.py"))
        layers_pkg.CustomRNNDropoutCell = drop.CustomRNNDropoutCell
        layers_pkg.HParamsLSA = lsa.HParamsLSA
        layers_pkg.LocationSensitiveAttention = lsa.LocationSensitiveAttention
        # encoder-only layers: never instantiated here
        layers_pkg.CustomEmbedding = layers_pkg.ConcatEmbedding = type("EncoderOnlyLayer", (), {})
        layers_pkg.ConcatMode = types.SimpleNamespace(CONCAT="concat")
        blocks = types.ModuleType(_PKG + ".current_blocks")
        blocks._get_var = lambda v, i, key=None: v[i] if isinstance(v, list) else v     # per-layer value or shared value
        sys.modules[_PKG + ".current_blocks"] = blocks
        simple = types.ModuleType(_PKG + ".simple_models")
        simple.simple_cnn = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("encoder/postnet are not covered"))
        sys.modules[_PKG + ".simple_models"] = simple
        return _load(_PKG + ".tacotron2_arch", os.path.join(arch_dir, "tacotron2_arch.py"))
    finally:
        for name, old in stubs.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def build_reference_decoder(hp, weights):
    """The reference's Tacotron2Decoder with our hyper-parameters and weights, prenet dropout off."""
    import torch
    arch = load_reference_taco()
    dec = arch.Tacotron2Decoder(
        n_mel_channels=hp.n_mel_channels, attention_rnn_dim=hp.attention_rnn_dim, decoder_rnn_dim=hp.decoder_rnn_dim,
        lsa_attention_dim=hp.attention_dim, lsa_attention_filters=hp.attention_filters,
        lsa_attention_kernel_size=hp.attention_kernel_size, prenet_sizes=list(hp.prenet_sizes),
        prenet_drop_rate=hp.prenet_drop_rate, prenet_deterministic=True, name="decoder")
    dec.build([(None, None, hp.n_mel_channels), (None, None, hp.embedding_dim)])
    got = dec.named_variables()
    want = {NAME_MAP[k]: v for k, v in weights.items() if k in NAME_MAP}
    if set(got) != set(want):
        raise KeyError(f"variable mismatch: only in reference {sorted(set(got) - set(want))[:4]}, "
                       f"only in ours {sorted(set(want) - set(got))[:4]}")
    dec.set_weights(want)
    assert dec.prenet.deterministic
    return dec, torch


def reference_decode(hp, weights, memory, max_length, dtype="float32"):
    """memory: numpy [S, E] (one utterance, no padding). Returns dict of numpy arrays: decoder_output [T, n_mel],
    stop_tokens [T], attention_weights [T, S], lengths (int)."""
    import torch
    prev = torch.get_default_dtype()
    torch.set_default_dtype(getattr(torch, dtype))
    try:
        dec, _ = build_reference_decoder(hp, weights)
        mem = torch.as_tensor(np.asarray(memory), dtype=torch.get_default_dtype())[None]
        mask = torch.ones(1, mem.shape[1], dtype=torch.bool)
        with torch.no_grad():
            (outputs, stop_tokens, _), state = dec.infer(encoder_output=mem, encoder_mask=mask, max_length=int(max_length),
                                                        early_stopping=False)
        return {"decoder_output": outputs[0].numpy(), "stop_tokens": stop_tokens[0].numpy(),
                "attention_weights": state.attention_weights[0].numpy(), "lengths": int(state.lengths[0])}
    finally:
        torch.set_default_dtype(prev)

"""CPU: the Tacotron2 DECODER restatement (text_to_speech_b200/tacotron2.py, the checker of csrc/taco.cu) against
fixtures produced by the reference's OWN Tacotron2Decoder.infer source run over the Keras shim
(oracle/run_reference_taco.py, oracle/gen_golden_taco.py), and -- when the reference tree is present -- against a
live run of that source. Encoder and postnet are NOT covered by this pin (functional-Keras `simple_cnn`)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.gen_golden_taco import CASES, decoder_weights, memories
from oracle.run_reference_taco import reference_decode, taco_reference_available
from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights
from text_to_speech_b200.weights import weights_digest


def load_taco_golden(name):
    f = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kw, wseed, gate_bias, mseed, lengths, frames = CASES[name]
    hp = Tacotron2HParams(**kw)
    w = generate_tacotron2_weights(hp, int(f["weight_seed"]))
    w["decoder/gate_output/bias"][:] = float(f["gate_bias"])
    assert weights_digest({k: v for k, v in w.items() if k.startswith("decoder/")}) == bytes(f["weights_sha256"]).decode(), \
        "Tacotron2 weight generator drifted from the fixture"
    mems = memories(hp, int(f["memory_seed"]), [int(s) for s in f["text_lengths"]])
    return hp, w, mems, int(f["frames"]), f


def padded_batch(mems, dtype):
    S = max(len(m) for m in mems)
    memory = torch.zeros(len(mems), S, mems[0].shape[1], dtype=dtype)
    mask = torch.zeros(len(mems), S, dtype=torch.bool)
    for i, m in enumerate(mems):
        memory[i, :len(m)] = torch.as_tensor(m, dtype=dtype)
        mask[i, :len(m)] = True
    return memory, mask


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_reference_fixture(name):
    hp, w, mems, frames, f = load_taco_golden(name)
    assert bytes(f["produced_by"]).decode().startswith("reference Tacotron2Decoder.infer source")
    for dtype, sfx, tol in ((torch.float64, "fp64", 1e-12), (torch.float32, "fp32", 5e-6)):
        model = Tacotron2(hp, w, device="cpu", dtype=dtype)
        # the reference runs one utterance at a time; the restatement runs them as one padded batch: the comparison
        # also pins the padding / attention-mask handling
        memory, mask = padded_batch(mems, dtype)
        out, stops, attn, lengths = model.decode(memory, mask, frames, early_stopping=False, deterministic=True)
        assert lengths.tolist() == [frames] * len(mems)
        for i, m in enumerate(mems):
            assert np.abs(out[i].numpy() - f[f"u{i}_decoder_output_{sfx}"]).max() <= tol
            assert np.abs(stops[i].numpy() - f[f"u{i}_stop_tokens_{sfx}"]).max() <= tol
            assert np.abs(attn[i, :, :len(m)].numpy() - f[f"u{i}_attention_weights_{sfx}"]).max() <= tol
            assert float(attn[i, :, len(m):].abs().max()) == 0.0 if len(m) < attn.shape[2] else True


@pytest.mark.skipif(not taco_reference_available(), reason="reference tree not mounted")
def test_restatement_matches_live_reference_source():
    """Different weights / shapes than the fixtures, straight through the reference's source."""
    hp = Tacotron2HParams(**CASES["taco_decoder_small"][0])
    w = generate_tacotron2_weights(hp, 11)
    w["decoder/gate_output/bias"][:] = -2.0
    dw = {k: v for k, v in w.items() if k.startswith("decoder/")}
    rng = np.random.default_rng(5)
    model = Tacotron2(hp, w, device="cpu", dtype=torch.float64)
    for S, T in ((1, 4), (8, 15), (30, 7)):
        mem = rng.standard_normal((S, hp.embedding_dim)) * 0.7
        ref = reference_decode(hp, dw, mem, T, dtype="float64")
        out, stops, attn, lengths = model.decode(torch.as_tensor(mem)[None], torch.ones(1, S, dtype=torch.bool), T,
                                                 early_stopping=False, deterministic=True)
        assert np.abs(out[0].numpy() - ref["decoder_output"]).max() <= 1e-12
        assert np.abs(stops[0].numpy() - ref["stop_tokens"]).max() <= 1e-12
        assert np.abs(attn[0].numpy() - ref["attention_weights"]).max() <= 1e-12
        assert int(lengths[0]) == ref["lengths"]


@pytest.mark.skipif(not taco_reference_available(), reason="reference tree not mounted")
def test_stop_gate_bookkeeping_matches_reference_source():
    """A gate that fires part-way: `finished`/`lengths` must follow tacotron2_arch.py:671-672 exactly."""
    hp = Tacotron2HParams(**CASES["taco_decoder_small"][0])
    w = generate_tacotron2_weights(hp, 13)
    w["decoder/gate_output/kernel"] *= 40.0          # make the stop probability swing around 0.5
    dw = {k: v for k, v in w.items() if k.startswith("decoder/")}
    mem = np.random.default_rng(2).standard_normal((9, hp.embedding_dim)) * 0.7
    ref = reference_decode(hp, dw, mem, 20, dtype="float64")
    model = Tacotron2(hp, w, device="cpu", dtype=torch.float64)
    out, stops, attn, lengths = model.decode(torch.as_tensor(mem)[None], torch.ones(1, 9, dtype=torch.bool), 20,
                                             early_stopping=False, deterministic=True)
    assert 0 <= ref["lengths"] < 20, "pick weights whose gate fires inside the window"
    assert int(lengths[0]) == ref["lengths"]
    assert np.abs(stops[0].numpy() - ref["stop_tokens"]).max() <= 1e-12

"""CPU: internal consistency of the Tacotron2 mel producer (the CALLER of the path; the decoder is pinned to the
reference source in tests/test_oracle_taco.py, encoder/postnet are not -- see the header of
text_to_speech_b200/tacotron2.py) and of the tts() pipeline's host logic."""
import numpy as np
import pytest
import torch

from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights
from text_to_speech_b200.tts import synthetic_texts

SMALL = Tacotron2HParams(embedding_dim=64, prenet_sizes=(32, 32), attention_rnn_dim=96, decoder_rnn_dim=96,
                         attention_dim=24, attention_filters=8, attention_kernel_size=7, postnet_filters=48)


@pytest.fixture(scope="module")
def small():
    w = generate_tacotron2_weights(SMALL, 3)
    return w, Tacotron2(SMALL, w, device="cpu", dtype=torch.float32), Tacotron2(SMALL, w, device="cpu", dtype=torch.float64)


def test_weight_layouts_follow_keras():
    hp = Tacotron2HParams()
    w = generate_tacotron2_weights(hp, 0)
    assert w["encoder/embeddings"].shape == (148, 512)
    assert w["encoder/conv_0/kernel"].shape == (5, 512, 512)
    assert w["encoder/bi_lstm/forward/kernel"].shape == (512, 1024)
    assert w["decoder/attention_rnn/kernel"].shape == (256 + 512, 4096)
    assert w["decoder/attention_rnn/recurrent_kernel"].shape == (1024, 4096)
    assert w["decoder/decoder_rnn/cell_0/kernel"].shape == (1024 + 512, 4096)
    assert w["decoder/lsa/location_conv/kernel"].shape == (31, 2, 32)
    assert w["decoder/linear_projection/kernel"].shape == (1536, 80)
    assert w["postnet/conv_4/kernel"].shape == (5, 512, 80)
    assert np.all(w["decoder/attention_rnn/bias"][1024:2048] == 1.0)        # unit forget bias
    assert sum(v.size for v in w.values()) == pytest.approx(28.2e6, rel=0.02)  # NVIDIA Tacotron2: 28.2 M parameters


def test_float32_matches_float64_twin(small):
    _, m32, m64 = small
    toks = np.stack(synthetic_texts(2, 1, 12, 12))
    a = m32.infer(toks, max_length=20, early_stopping=False, deterministic=True)
    b = m64.infer(toks, max_length=20, early_stopping=False, deterministic=True)
    assert a.mel.shape == (2, 20, 80) and a.mel.dtype == torch.float32
    assert torch.allclose(a.mel.double(), b.mel, atol=2e-4)
    assert torch.equal(a.lengths, b.lengths)


def test_padding_does_not_change_an_utterance(small):
    _, _, m64 = small
    texts = synthetic_texts(3, 2, 6, 14)
    S = max(len(t) for t in texts)
    batch = np.zeros((3, S), np.int64)
    for j, t in enumerate(texts):
        batch[j, :len(t)] = t
    together = m64.infer(batch, max_length=12, early_stopping=False, deterministic=True)
    for j, t in enumerate(texts):
        alone = m64.infer(t[None], max_length=12, early_stopping=False, deterministic=True)
        assert torch.allclose(alone.mel[0], together.mel[j], atol=1e-9)
        aw = together.attention_weights[j]
        assert torch.allclose(aw.sum(-1), torch.ones_like(aw.sum(-1)), atol=1e-9)
        assert float(aw[:, len(t):].abs().max() if len(t) < S else 0.0) == 0.0      # nothing attends to padding


def _with_gate_bias(w, value):
    w = dict(w)
    w["decoder/gate_output/kernel"] = np.zeros_like(w["decoder/gate_output/kernel"])
    w["decoder/gate_output/bias"] = np.full(1, value, np.float32)
    return w


def test_stop_and_length_logic(small):
    w, _, _ = small
    toks = np.stack(synthetic_texts(2, 4, 9, 9))
    never = Tacotron2(SMALL, _with_gate_bias(w, -20.0), device="cpu").infer(toks, max_length=15, deterministic=True)
    assert never.lengths.tolist() == [15, 15] and never.mel.shape[1] == 15          # runs to max_length
    at_once = Tacotron2(SMALL, _with_gate_bias(w, 20.0), device="cpu").infer(toks, max_length=15, deterministic=True)
    assert at_once.lengths.tolist() == [0, 0]                                        # finished on the first frame
    assert float(at_once.decoder_output[:, 1:].abs().max()) == 0.0                   # loop stopped: rest stays zero
    ratio = Tacotron2(SMALL, _with_gate_bias(w, -20.0), device="cpu").infer(toks, max_length=2.0, deterministic=True)
    assert ratio.mel.shape[1] == 18                                                  # float = frames per token


def test_prenet_dropout_is_on_unless_deterministic(small):
    _, m32, _ = small
    toks = np.stack(synthetic_texts(1, 5, 8, 8))
    torch.manual_seed(0)
    a = m32.infer(toks, max_length=6, early_stopping=False)
    b = m32.infer(toks, max_length=6, early_stopping=False)
    c = m32.infer(toks, max_length=6, early_stopping=False, deterministic=True)
    d = m32.infer(toks, max_length=6, early_stopping=False, deterministic=True)
    assert not torch.equal(a.mel, b.mel) and torch.equal(c.mel, d.mel)


def test_max_length_is_required(small):
    _, m32, _ = small
    with pytest.raises(ValueError):
        m32.infer(np.ones((1, 4), np.int64))


def test_taco_header_binding_and_library_agree(lib_built):
    import os
    import re
    from conftest import ROOT
    from text_to_speech_b200 import _lib
    src = open(os.path.join(ROOT, "include", "wg_taco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(wg_taco_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(_lib.TACO_EXPORTS)
    for name in declared:
        assert hasattr(lib_built, name)


def test_b200_decoder_has_no_cpu_fallback(small):
    _, m32, _ = small
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m32.infer(np.ones((1, 4), np.int64), max_length=3, decoder="b200")

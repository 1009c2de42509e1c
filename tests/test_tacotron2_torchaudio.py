"""CPU: the Tacotron2 encoder / postnet restatement (text_to_speech_b200/tacotron2.py) against an INDEPENDENT third-party
implementation of the same architecture: torchaudio.models.tacotron2 (`_Encoder`, `_Postnet`), itself a port of NVIDIA's
Tacotron2 -- the model the reference says it copies (architectures/tacotron2_arch.py:214-333 builds the same
embedding -> 3 x (conv k=5, batch-norm, relu) -> BiLSTM and 5 x (conv k=5, batch-norm, tanh) stacks through `simple_cnn`).

The reference's own encoder / postnet source cannot be executed here (functional Keras API, automatic mask propagation),
so this is what pins the two restatements: same weights (Keras layouts converted to torch's), inference mode, sequences
WITHOUT padding (torchaudio does not zero padded positions between the convolutions as the reference's MaskedConv1D
does; the reference's one-sentence-at-a-time call has no padding either)."""
import numpy as np
import pytest
import torch

from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights

ta = pytest.importorskip("torchaudio.models.tacotron2")


def _load_conv_bn(seq, w, conv, bn):
    t = lambda k: torch.from_numpy(np.asarray(w[k], dtype=np.float32))
    with torch.no_grad():
        seq[0].weight.copy_(t(conv + "/kernel").permute(2, 1, 0))          # Keras [k, in, out] -> torch [out, in, k]
        seq[0].bias.copy_(t(conv + "/bias"))
        seq[1].weight.copy_(t(bn + "/gamma"))
        seq[1].bias.copy_(t(bn + "/beta"))
        seq[1].running_mean.copy_(t(bn + "/moving_mean"))
        seq[1].running_var.copy_(t(bn + "/moving_variance"))


@pytest.fixture(scope="module")
def model():
    hp = Tacotron2HParams()
    w = generate_tacotron2_weights(hp, 11)
    for i in range(hp.encoder_n_conv):                     # non-trivial conv biases (the generator leaves them at zero)
        w[f"encoder/conv_{i}/bias"] = (0.05 * np.random.default_rng(i).standard_normal(hp.embedding_dim)).astype(np.float32)
    return hp, w, Tacotron2(hp, w, device="cpu")


def test_encoder_matches_torchaudio(model):
    hp, w, ours = model
    enc = ta._Encoder(hp.embedding_dim, hp.encoder_n_conv, hp.encoder_kernel_size).eval()
    for i, seq in enumerate(enc.convolutions):
        assert seq[1].eps == hp.bn_epsilon
        _load_conv_bn(seq, w, f"encoder/conv_{i}", f"encoder/bn_{i}")
    t = lambda k: torch.from_numpy(np.asarray(w[k], dtype=np.float32))
    with torch.no_grad():
        for d, sfx in (("forward", ""), ("backward", "_reverse")):
            # Keras LSTM: kernel [in, 4u], recurrent [u, 4u], ONE bias, gates i, f, c, o == torch's i, f, g, o
            getattr(enc.lstm, "weight_ih_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/kernel").T)
            getattr(enc.lstm, "weight_hh_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/recurrent_kernel").T)
            getattr(enc.lstm, "bias_ih_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/bias"))
            getattr(enc.lstm, "bias_hh_l0" + sfx).zero_()
    rng = np.random.default_rng(0)
    for B, S in ((1, 37), (3, 20), (2, 1)):
        tokens = torch.from_numpy(rng.integers(1, hp.vocab_size, size=(B, S)))
        with torch.no_grad():
            want = enc(t("encoder/embeddings")[tokens].transpose(1, 2), torch.full((B,), S))
            got, mask = ours.encode(tokens)
        assert got.shape == want.shape == (B, S, hp.embedding_dim) and bool(mask.all())
        err = float((got - want).abs().max())
        assert err <= 2e-5, (B, S, err)


def test_postnet_matches_torchaudio(model):
    hp, w, ours = model
    post = ta._Postnet(hp.n_mel_channels, hp.postnet_filters, hp.postnet_kernel_size, hp.postnet_n_conv).eval()
    for i, seq in enumerate(post.convolutions):
        _load_conv_bn(seq, w, f"postnet/conv_{i}", f"postnet/bn_{i}")
    g = torch.Generator().manual_seed(3)
    for B, T in ((1, 50), (2, 9)):
        x = torch.randn(B, T, hp.n_mel_channels, generator=g)
        with torch.no_grad():
            want = post(x.transpose(1, 2)).transpose(1, 2)
            got = ours.postnet(x, torch.ones(B, T, dtype=torch.bool))
        err = float((got - want).abs().max())
        assert got.shape == want.shape and err <= 2e-5, (B, T, err)


def test_padding_is_masked_the_reference_way(model):
    """Where the two differ by design: a padded batch. Ours zeroes padded positions between the convolutions
    (MaskedConv1D), so the VALID part of a padded utterance equals the same utterance run alone."""
    hp, w, ours = model
    rng = np.random.default_rng(1)
    tokens = torch.from_numpy(rng.integers(1, hp.vocab_size, size=(2, 30)))
    tokens[1, 17:] = hp.pad_token
    with torch.no_grad():
        both, mask = ours.encode(tokens)
        alone, _ = ours.encode(tokens[1:2, :17])
    assert int(mask[1].sum()) == 17 and not both[1, 17:].any()
    assert float((both[1, :17] - alone[0]).abs().max()) <= 2e-6

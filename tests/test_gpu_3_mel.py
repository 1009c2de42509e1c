"""GPU: the log-mel front-end kernel (wg_mel_spectrogram, through the C ABI) against the oracle and
against the reference's own fixture.

Tolerances (log-mel units, natural log):
  * vs the reference fixture: 2e-3, the reference's own bound for it (tests/test_utils_audio.py:109-111);
  * vs the float64 oracle: 5e-4 max-abs. The kernel is an fp32 FFT, the reference an fp32 1024-term
    convolution; the float32 oracle itself sits 1.3e-4 from the float64 one on the fixture.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import mel_oracle

pytestmark = pytest.mark.gpu

TOL64 = 5e-4


@pytest.fixture(scope="module")
def stft(lib_built):
    from text_to_speech_b200.stft import TacotronSTFT
    return TacotronSTFT()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "mel_tacotron_stft.npz"))


def _speechlike(rng, B, N):
    t = np.arange(N) / 22050.0
    x = np.zeros((B, N))
    for b in range(B):
        for _ in range(6):
            f, a, ph = rng.uniform(80, 7000), rng.uniform(0.02, 0.3), rng.uniform(0, 6.28)
            x[b] += a * np.sin(2 * np.pi * f * t + ph) * (0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(1, 6) * t))
        x[b] += 0.01 * rng.standard_normal(N)
    return np.clip(x, -1, 1).astype(np.float32)


def test_reference_fixture(stft, golden):
    mel = stft(golden["audio_22050"])
    assert mel.shape == (1, 350, 80) and mel.dtype == np.float32
    err_ref = np.abs(mel[0] - golden["mel_reference"]).max()
    err_64 = np.abs(mel - mel_oracle.tacotron_mel(golden["audio_22050"])).max()
    print(f"\nfixture: vs reference {err_ref:.2e} (bound 2e-3), vs float64 oracle {err_64:.2e}")
    assert err_ref <= float(golden["reference_max_err"])
    assert err_64 <= TOL64


@pytest.mark.parametrize("B,N", [(1, 1024), (1, 1025), (3, 5000), (2, 22050), (5, 8191), (1, 256 * 40), (4, 256 * 17 + 255)])
def test_against_oracle_shapes(stft, B, N):
    x = _speechlike(np.random.default_rng(B * 100003 + N), B, N)
    want = mel_oracle.tacotron_mel(x)
    got = stft.mel_spectrogram(x)
    assert got.shape == want.shape == (B, N // 256 + 1, 80)
    assert stft.n_frames(N) == want.shape[1]
    assert np.abs(got - want).max() <= TOL64


def test_device_tensor_path_matches_host_path(stft):
    x = _speechlike(np.random.default_rng(9), 3, 30000)
    host = stft.mel_spectrogram(x)
    dev = stft.mel_spectrogram(torch.from_numpy(x).cuda())
    assert dev.is_cuda and dev.shape == host.shape
    assert np.array_equal(dev.cpu().numpy(), host)
    assert isinstance(stft.mel_spectrogram(torch.from_numpy(x)), torch.Tensor)


def test_short_audio_zero_padded_like_reference(stft):
    x = _speechlike(np.random.default_rng(3), 2, 700)
    got = stft.mel_spectrogram(x)
    want = mel_oracle.tacotron_mel(x)
    assert got.shape == want.shape == (2, 5, 80)
    assert np.abs(got - want).max() <= TOL64
    assert np.array_equal(got, stft.mel_spectrogram(np.pad(x, [(0, 0), (0, 324)])))


def test_silence_hits_the_clip_exactly(stft):
    mel = stft.mel_spectrogram(np.zeros((2, 4096), np.float32))
    assert np.array_equal(mel, np.full_like(mel, np.log(np.float32(1e-5))))


def test_one_dimensional_call_and_batch_independence(stft):
    x = _speechlike(np.random.default_rng(4), 4, 12345)
    all_ = stft(x)
    for b in range(4):
        assert np.array_equal(stft(x[b])[0], all_[b])


def test_linearity_of_the_magnitude(stft):
    # exp(log-mel) is homogeneous of degree 1 in the audio wherever the clip is inactive
    x = _speechlike(np.random.default_rng(6), 1, 20000)
    a, b = stft(x), stft(0.25 * x)
    live = (a > -9.0) & (b > -9.0)
    assert live.mean() > 0.5
    assert np.abs((a - b)[live] - np.log(4.0)).max() <= 2e-4


def test_long_batch_checksum_against_oracle(stft):
    # 16 x 10 s (the K2 utterance shape seen from the audio side)
    x = _speechlike(np.random.default_rng(8), 16, 860 * 256)
    got = stft.mel_spectrogram(torch.from_numpy(x).cuda()).cpu().numpy()
    want = mel_oracle.tacotron_mel(x)
    assert got.shape == (16, 861, 80)
    assert np.abs(got - want).max() <= TOL64
    assert abs(got.astype(np.float64).sum() - want.sum()) <= 1e-6 * np.abs(want).sum()


def test_other_hop_and_mel_sizes(lib_built):
    from text_to_speech_b200.stft import TacotronSTFT
    x = _speechlike(np.random.default_rng(11), 2, 9000)
    for kw in (dict(hop_length=200, win_length=800, n_mel_channels=64, mel_fmin=50.0, mel_fmax=7600.0),
               dict(hop_length=512, n_mel_channels=128, mel_fmax=11025.0),
               dict(hop_length=255, n_mel_channels=40)):
        t = TacotronSTFT(**kw)
        want = mel_oracle.tacotron_mel(x, n_mel_channels=t.n_mel_channels, hop_length=t.hop_length,
                                       win_length=t.win_length, mel_fmin=t.mel_fmin, mel_fmax=t.mel_fmax)
        got = t.mel_spectrogram(x)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= TOL64, kw


def test_errors_are_loud(stft):
    with pytest.raises(ValueError):
        stft.mel_spectrogram(np.zeros((2, 3, 4), np.float32))
    from text_to_speech_b200.stft import TacotronSTFT
    with pytest.raises(RuntimeError, match="filter_length"):
        TacotronSTFT(filter_length=400, win_length=400, hop_length=160)
    with pytest.raises(NotImplementedError):
        TacotronSTFT(window="hamming")

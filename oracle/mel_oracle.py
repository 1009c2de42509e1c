"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product path (text_to_speech_b200/).

CPU restatement (numpy, float32 or float64) of the reference's Tacotron log-mel front-end, in the
reference's own formulation (a strided convolution with a windowed Fourier basis, NOT an FFT):

  utils/audio/stft.py:189-235   STFT.__init__  (fft(eye) basis, real rows then imaginary rows, hann
                                window from scipy.signal.get_window(fftbins=periodic), centre padded)
                                                                      -> forward_basis()
  utils/audio/stft.py:241-280   STFT.transform (reflect pad, conv1d stride hop, sqrt(re^2+im^2))
                                                                      -> stft_magnitude()
  utils/audio/stft.py:59-68     librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)   -> mel_filter_bank()
  utils/audio/stft.py:98-126    MelSTFT.__call__ (zero pad short audio to win_length) -> tacotron_mel()
  utils/audio/stft.py:307-314   log(max(mag @ mel_basis, 1e-5))       -> tacotron_mel()

Third-party pieces restated from their published definitions (absent here, un-pinned in the
reference's requirements.txt): `librosa.filters.mel` defaults (Slaney scale: linear below 1 kHz at
200/3 Hz per mel, log above with step ln(6.4)/27; triangles normalised by 2/(f[m+2]-f[m])) and
scipy's periodic hann window 0.5 - 0.5 cos(2 pi n / N).

PARITY PIN: this one IS pinned by the reference's own fixture pair -- input
tests/__reproduction/audio_resample.npy (= load_audio(audio_test.wav, 22050), test_utils_audio.py:62-64),
output tests/__reproduction/stft-TacotronSTFT.npy, asserted by the reference at max_err 2e-3
(tests/test_utils_audio.py:109-111). Both are committed as tests/golden/mel_tacotron_stft.npz by
oracle/gen_golden_mel.py; tests/test_oracle_mel.py holds this restatement to the same 2e-3
(observed 6.7e-4 max, 9e-5 median; float32 and float64 agree with each other to 1e-5, so the residual
is the fixture's own provenance, not this code).
"""
from __future__ import annotations

import numpy as np


def hann(win_length, filter_length, periodic=True, dtype=np.float64):
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (win_length if periodic else win_length - 1))
    pad = filter_length - win_length
    return np.pad(w, (pad // 2, pad - pad // 2)).astype(dtype)


def forward_basis(filter_length=1024, win_length=1024, periodic=True, dtype=np.float64):
    """[2*cutoff, filter_length]: rows 0..cutoff-1 real part, rows cutoff.. imaginary part (stft.py:206-228)."""
    cutoff = filter_length // 2 + 1
    four = np.fft.fft(np.eye(filter_length))
    basis = np.vstack([np.real(four[:cutoff]), np.imag(four[:cutoff])])
    # the reference casts the basis to float32 BEFORE windowing it (stft.py:214, 226)
    basis = basis.astype(np.float32).astype(np.float64) * hann(win_length, filter_length, periodic)
    return basis.astype(dtype)


def stft_magnitude(audio, filter_length=1024, hop_length=256, win_length=1024, periodic=True, dtype=np.float64):
    """audio [B, N] -> magnitude [B, F, cutoff] (stft.py:241-280)."""
    audio = np.asarray(audio, dtype=dtype)
    x = np.pad(audio, [(0, 0), (filter_length // 2, filter_length // 2)], mode="reflect")
    F = (x.shape[1] - filter_length) // hop_length + 1
    s0, s1 = x.strides
    frames = np.lib.stride_tricks.as_strided(x, (x.shape[0], F, filter_length), (s0, hop_length * s1, s1))
    basis = forward_basis(filter_length, win_length, periodic, dtype)
    out = np.einsum("bfk,ck->bfc", frames, basis, optimize=True).astype(dtype)
    cutoff = filter_length // 2 + 1
    re, im = out[..., :cutoff], out[..., cutoff:]
    return np.sqrt(re * re + im * im)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + 27.0 * np.log(np.maximum(f, 1e-10) / 1000.0) / np.log(6.4), 3.0 * f / 200.0)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) * (m - 15.0) / 27.0), 200.0 * m / 3.0)


def mel_filter_bank(sr=22050, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0):
    """[n_mels, n_fft/2+1] float32, librosa.filters.mel defaults (htk=False, norm='slaney')."""
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fb = np.zeros((n_mels, len(freqs)))
    for m in range(n_mels):
        lo, mid, hi = edges[m], edges[m + 1], edges[m + 2]
        up = (freqs - lo) / (mid - lo)
        down = (hi - freqs) / (hi - mid)
        fb[m] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


def tacotron_mel(audio, sampling_rate=22050, n_mel_channels=80, filter_length=1024, hop_length=256, win_length=1024,
                 mel_fmin=0.0, mel_fmax=8000.0, clip_val=1e-5, periodic=True, dtype=np.float64):
    """audio [N] or [B, N] -> log-mel [B, F, n_mel] (MelSTFT.__call__ + TacotronSTFT.mel_spectrogram)."""
    audio = np.asarray(audio, dtype=dtype)
    if audio.ndim == 1:
        audio = audio[None]
    if audio.shape[1] < win_length:
        audio = np.pad(audio, [(0, 0), (0, win_length - audio.shape[1])])
    mag = stft_magnitude(audio, filter_length, hop_length, win_length, periodic, dtype)
    basis = mel_filter_bank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax).T.astype(dtype)
    return np.log(np.maximum(mag @ basis, dtype(clip_val)))

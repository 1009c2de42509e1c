"""CPU: the log-mel oracle against the reference's OWN fixture pair (tests/__reproduction, committed as
tests/golden/mel_tacotron_stft.npz), the host-side parameters of the product against the oracle's, and
the mel half of the C ABI (symbols, argument checks, no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from oracle import mel_oracle
from text_to_speech_b200 import _lib
from text_to_speech_b200 import stft as host


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "mel_tacotron_stft.npz"))


def test_oracle_matches_reference_fixture_within_reference_tolerance(golden):
    # the reference asserts this pair at max_err 2e-3 (tests/test_utils_audio.py:109-111)
    tol = float(golden["reference_max_err"])
    assert tol == 2e-3
    for dtype in (np.float32, np.float64):
        mel = mel_oracle.tacotron_mel(golden["audio_22050"], dtype=dtype)
        assert mel.shape == (1, 350, 80) and mel.dtype == dtype
        err = np.abs(mel[0] - golden["mel_reference"]).max()
        assert err <= tol, f"{dtype.__name__}: {err}"
        assert err <= 1e-3          # observed 6.7e-4: keep the margin visible


def test_oracle_float32_and_float64_agree(golden):
    m32 = mel_oracle.tacotron_mel(golden["audio_22050"], dtype=np.float32)
    m64 = mel_oracle.tacotron_mel(golden["audio_22050"], dtype=np.float64)
    assert np.abs(m32 - m64).max() <= 5e-4


def test_convolution_form_equals_fft_form():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3000)) * 0.3
    mag = mel_oracle.stft_magnitude(x, dtype=np.float64)
    xp = np.pad(x, [(0, 0), (512, 512)], mode="reflect")
    win = mel_oracle.hann(1024, 1024)
    for f in (0, 5, mag.shape[1] - 1):
        ref = np.abs(np.fft.rfft(xp[:, 256 * f:256 * f + 1024] * win, axis=1))
        # the reference rounds its Fourier basis to float32 (stft.py:214): 1e-7 relative on a norm-32 frame
        assert np.abs(mag[:, f] - ref).max() <= 1e-5


def test_filter_bank_properties():
    fb = mel_oracle.mel_filter_bank()
    assert fb.shape == (80, 513) and fb.dtype == np.float32 and (fb >= 0).all()
    nz = [np.nonzero(r)[0] for r in fb]
    assert all(len(i) and (np.diff(i) == 1).all() for i in nz)            # one contiguous triangle per channel
    assert all(nz[m][0] <= nz[m + 1][0] for m in range(79))
    assert fb[:, 373:].max() == 0.0                                       # nothing above fmax = 8 kHz (bin 371.5)
    # Slaney normalisation: every triangle has unit area in Hz
    assert np.allclose(fb.sum(axis=1) * (22050 / 1024), 1.0, atol=0.08)


def test_short_audio_is_zero_padded_to_win_length():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(700).astype(np.float32) * 0.1
    a = mel_oracle.tacotron_mel(x)
    b = mel_oracle.tacotron_mel(np.pad(x, (0, 324)))
    assert a.shape == (1, 5, 80) and np.array_equal(a, b)


def test_host_parameters_match_oracle():
    assert np.array_equal(host.slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0), mel_oracle.mel_filter_bank())
    assert np.array_equal(host.slaney_mel_basis(16000, 1024, 64, 50.0, 7600.0),
                          mel_oracle.mel_filter_bank(16000, 1024, 64, 50.0, 7600.0))
    for wl in (1024, 800):
        assert np.abs(host.hann_window(wl, 1024) - mel_oracle.hann(wl, 1024)).max() < 1e-7


def _header_functions():
    src = open(os.path.join(ROOT, "include", "wg_mel_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_mel_[a-z0-9_]+)\s*\(", src)))


def test_mel_header_binding_and_library_agree(lib_built):
    assert _header_functions() == sorted(_lib.MEL_EXPORTS)
    for name in _header_functions():
        assert hasattr(lib_built, name), f"{name} not exported"


def test_mel_null_arguments_are_rejected(lib_built):
    h = ctypes.c_void_p()
    assert lib_built.wg_mel_create(None, None, None, 0, ctypes.byref(h)) == -1
    assert b"NULL" in lib_built.wg_mel_last_error(None)
    n = ctypes.c_int64()
    assert lib_built.wg_mel_frames(None, 1000, ctypes.byref(n)) == -1
    assert lib_built.wg_mel_spectrogram(None, None, 1, 1000, None, None) == -1


def test_unsupported_fft_size_is_an_error_not_a_fallback(lib_built):
    cfg = _lib.WgMelConfig(16000, 80, 400, 160, 400, 1e-5)
    win = np.ones(400, np.float32)
    basis = np.ones((201, 80), np.float32)
    f32p = ctypes.POINTER(ctypes.c_float)
    h = ctypes.c_void_p()
    rc = lib_built.wg_mel_create(ctypes.byref(cfg), win.ctypes.data_as(f32p), basis.ctypes.data_as(f32p), 0,
                                 ctypes.byref(h))
    assert rc == -2 and b"filter_length 400 not supported" in lib_built.wg_mel_last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib_built):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        host.TacotronSTFT()


def test_product_does_not_import_the_oracle():
    src = open(os.path.join(ROOT, "text_to_speech_b200", "stft.py")).read()
    assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M)

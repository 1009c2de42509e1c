"""Host-side mirror of the reference's log-mel front-end (`utils/audio/stft.py`): the data format on
the input side of the WaveGlow path.

`TacotronSTFT` keeps the reference's constructor arguments, attributes and call contract
(stft.py:28-46, 98-126, 286-319) and routes the arithmetic to `wg_mel_spectrogram` in libwg_b200.so
(`include/wg_mel_b200.h`). Like the reference, the window and the mel filter bank are parameters
built once on the host in numpy (there: `scipy.signal.get_window` + `librosa.util.pad_center`,
stft.py:220-223, and `librosa.filters.mel`, stft.py:61-68; neither library is a dependency here, so
the two published formulas are written out below). There is no CPU fallback for the transform
itself: without the CUDA library / a GPU the constructor raises.
"""
from __future__ import annotations

import ctypes
import json
import math

import numpy as np

from . import _lib

__all__ = ["TacotronSTFT", "MelSTFT", "hann_window", "slaney_mel_basis"]


def hann_window(win_length: int, filter_length: int, periodic: bool = True) -> np.ndarray:
    """`get_window('hann', win_length, fftbins=periodic)` zero-padded symmetrically to `filter_length`
    (stft.py:220-223)."""
    n = np.arange(win_length, dtype=np.float64)
    denom = win_length if periodic else max(win_length - 1, 1)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / denom)
    lpad = (filter_length - win_length) // 2
    out = np.zeros(filter_length, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out.astype(np.float32)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """The filter bank `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` returns with its defaults
    (Slaney mel scale, area-normalised triangles), shape [n_mels, n_fft/2+1], float32."""
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class MelSTFT:
    """Attribute/constructor mirror of the reference's base class (stft.py:27-68, 83-96, 128-133,
    147-170). Only the Tacotron flavour has a kernel; `create` refuses the others by name."""

    def __init__(self, sampling_rate, n_mel_channels=80, *, win_length=1024, hop_length=256, filter_length=1024,
                 mel_fmin=0.0, mel_fmax=8000.0, normalize_mode=None, pre_emph=0.0, **kwargs):
        if normalize_mode is not None:
            raise NotImplementedError("normalize_mode is not used on the WaveGlow path (reference default None)")
        if pre_emph:
            raise NotImplementedError("pre_emph is not used on the WaveGlow path (reference default 0.)")
        self.n_mel_channels = int(n_mel_channels)
        self.sampling_rate = int(sampling_rate)
        # fractions of a second are accepted like the reference does (stft.py:51-55)
        self.win_length = int(win_length if win_length > 1.0 else win_length * sampling_rate)
        self.hop_length = int(hop_length if hop_length > 1.0 else hop_length * sampling_rate)
        self.filter_length = int(filter_length if filter_length > 1.0 else filter_length * sampling_rate)
        self.mel_fmin, self.mel_fmax = mel_fmin, mel_fmax
        self.pre_emph, self.normalize_mode = pre_emph, normalize_mode
        basis = slaney_mel_basis(self.sampling_rate, self.filter_length, self.n_mel_channels, mel_fmin, mel_fmax)
        self.mel_basis = np.ascontiguousarray(basis.T[None])          # [1, n_bins, n_mel] (stft.py:68)

    @property
    def rate(self):
        return self.sampling_rate

    def get_mel_length(self, audio_length):
        return int(math.ceil(max(self.filter_length, audio_length) / self.hop_length))

    def get_audio_length(self, mel_length):
        return mel_length * self.hop_length

    def get_config(self):
        return {
            "class_name": self.__class__.__name__, "n_mel_channels": self.n_mel_channels,
            "sampling_rate": self.sampling_rate, "win_length": self.win_length, "hop_length": self.hop_length,
            "filter_length": self.filter_length, "mel_fmin": self.mel_fmin, "mel_fmax": self.mel_fmax,
            "pre_emph": self.pre_emph, "normalize_mode": self.normalize_mode,
        }

    def save(self, filename):
        if not filename.endswith(".json"):
            filename += ".json"
        with open(filename, "w") as f:
            json.dump(self.get_config(), f, indent=4)
        return filename

    save_to_file = save

    @classmethod
    def load_from_file(cls, filename):
        with open(filename) as f:
            return MelSTFT.create(**json.load(f))

    @staticmethod
    def create(class_name, *args, **kwargs):
        if class_name == "TacotronSTFT":
            return TacotronSTFT(*args, **kwargs)
        raise ValueError(f"Unknown Mel STFT class !\n  Accepted : ('TacotronSTFT',)\n  Got : {class_name}")

    def __eq__(self, other):
        return isinstance(other, MelSTFT) and self.get_config() == other.get_config()


class TacotronSTFT(MelSTFT):
    """`TacotronSTFT(sampling_rate=22050, n_mel_channels=80, window='hann', periodic=True, ...)`
    (stft.py:286-319). `__call__(audio)` takes `[length]` or `[B, length]` float audio in [-1, 1]
    (numpy, torch CPU or torch CUDA) and returns `[B, frames, n_mel]` log-mel in the same kind of
    container, exactly as `MelSTFT.__call__` does (stft.py:98-126)."""

    def __init__(self, sampling_rate=22050, n_mel_channels=80, *, window="hann", periodic=True, device=0,
                 clip_val=1e-5, **kwargs):
        super().__init__(sampling_rate=sampling_rate, n_mel_channels=n_mel_channels, **kwargs)
        if window != "hann":
            raise NotImplementedError(f"window {window!r}: only 'hann' is implemented")
        self.window, self.periodic, self.device, self.clip_val = window, bool(periodic), int(device), float(clip_val)
        self._window = hann_window(self.win_length, self.filter_length, self.periodic)
        self._lib = _lib.load_library()
        cfg = _lib.WgMelConfig(self.sampling_rate, self.n_mel_channels, self.filter_length, self.hop_length,
                               self.win_length, self.clip_val)
        f32p = ctypes.POINTER(ctypes.c_float)
        basis = np.ascontiguousarray(self.mel_basis[0], dtype=np.float32)
        h = ctypes.c_void_p()
        rc = self._lib.wg_mel_create(ctypes.byref(cfg), self._window.ctypes.data_as(f32p),
                                     basis.ctypes.data_as(f32p), self.device, ctypes.byref(h))
        if rc != _lib.WG_OK:
            raise RuntimeError(f"wg_mel_create failed ({rc}): {self._lib.wg_mel_last_error(None).decode()}")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.wg_mel_destroy(h)

    def _check(self, rc, what):
        if rc != _lib.WG_OK:
            raise RuntimeError(f"{what} failed ({rc}): {self._lib.wg_mel_last_error(self._h).decode()}")

    def n_frames(self, n_samples: int) -> int:
        """Exact frame count of the transform (the reference's `get_mel_length` is an estimate)."""
        out = ctypes.c_int64()
        self._check(self._lib.wg_mel_frames(self._h, int(n_samples), ctypes.byref(out)), "wg_mel_frames")
        return out.value

    def mel_spectrogram(self, audio):
        """`[B, samples]` -> `[B, frames, n_mel]` (stft.py:310-314)."""
        try:
            import torch
        except ImportError:  # pragma: no cover
            torch = None
        if torch is not None and isinstance(audio, torch.Tensor):
            if audio.dim() != 2:
                raise ValueError(f"audio must be [B, samples], got {tuple(audio.shape)}")
            if audio.is_cuda:
                if audio.device.index != self.device:
                    raise ValueError(f"audio is on {audio.device}, the transform on cuda:{self.device}")
                x = audio.detach().to(torch.float32).contiguous()
                B, N = x.shape
                out = torch.empty((B, self.n_frames(N), self.n_mel_channels), dtype=torch.float32, device=x.device)
                stream = torch.cuda.current_stream(x.device).cuda_stream
                self._check(self._lib.wg_mel_spectrogram(self._h, x.data_ptr(), B, N, out.data_ptr(), stream),
                            "wg_mel_spectrogram")
                return out
            return torch.from_numpy(self.mel_spectrogram(audio.detach().numpy()))
        x = np.ascontiguousarray(audio, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError(f"audio must be [B, samples], got {x.shape}")
        B, N = x.shape
        out = np.empty((B, self.n_frames(N), self.n_mel_channels), dtype=np.float32)
        f32p = ctypes.POINTER(ctypes.c_float)
        self._check(self._lib.wg_mel_spectrogram_host(self._h, x.ctypes.data_as(f32p), B, N, out.ctypes.data_as(f32p)),
                    "wg_mel_spectrogram_host")
        return out

    def __call__(self, audio, **kwargs):
        if len(audio.shape) == 1:
            audio = audio[None]
        return self.mel_spectrogram(audio)

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"filter_length": self.filter_length, "hop_length": self.hop_length, "win_length": self.win_length,
                    "window": self.window, "periodic": self.periodic})
        return cfg

"""Drop-in for the reference's model wrapper `models.tts.WaveGlow` on the infer path
(models/tts/waveglow.py:61-144): same inputs ([T,80] / [B,T,80] / '.npy' path), same windowing
options (win_len / hop_len / force_pad / batch / use_slice / max_win_len), same stitching at the
overlap midpoints, same `[:, :T*256]` trim. The vocoder itself is a `Runtime` (runtime.py)."""
from __future__ import annotations

import math

import numpy as np

from .runtime import build_runtime

PAD_MEL_VALUE = -11.0   # models/tts/waveglow.py:29


def _get_steps(length, win_len, hop_len):
    """models/tts/waveglow.py:156-164."""
    num_steps = int(math.ceil((length - win_len) / hop_len)) + 1
    if num_steps == 1:
        return [0]
    max_step = length - win_len
    actual_step_size = max_step / (num_steps - 1)
    return np.round(np.arange(num_steps) * actual_step_size).astype(np.int32)


def _to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


class WaveGlow:
    """`WaveGlow(path=..., runtime='b200', mode='bf16')(mel, sigma=0.6, z=z)` -> waveform [B, 256*T]."""

    def __init__(self, *, path, runtime="b200", pad_mel_value=PAD_MEL_VALUE, **runtime_kwargs):
        if runtime == "keras":
            raise ValueError("this package only provides non-keras runtimes (runtime='b200')")
        self.runtime = runtime
        self.pad_mel_value = pad_mel_value
        self.model = build_runtime(runtime, path, **runtime_kwargs)

    @property
    def compiled_infer(self):
        # base_model.py:366-370: for runtime != 'keras' the Runtime object itself is the callable
        return self.model

    def infer(self, mel, *, win_len=None, hop_len=-64, force_pad=None, batch=False, use_slice=False,
              max_win_len=None, **kwargs):
        if isinstance(mel, str):
            mel = np.load(mel)
        if len(mel.shape) == 2:
            mel = mel[None]
        seq_len = mel.shape[1]
        audio_len = seq_len * 256
        if win_len is None:
            return self.compiled_infer(mel, **kwargs)[:, :audio_len]

        if isinstance(win_len, float):
            if not use_slice:
                win_len = int(math.ceil(seq_len / win_len) * win_len)
            else:
                win_len = max(1, seq_len // win_len) * int(win_len)
        if max_win_len is not None:
            win_len = min(max_win_len, win_len)
        kwargs['padding_multiple'] = win_len

        if seq_len <= win_len:
            if force_pad is None:
                force_pad = self.runtime == 'keras'
            if not force_pad:
                return self.compiled_infer(mel)     # reference drops kwargs here (waveglow.py:96)
            win_len = max(win_len, seq_len)
            mel_np = _to_numpy(mel)
            padded = np.pad(mel_np, [(0, 0), (0, win_len - seq_len), (0, 0)], constant_values=self.pad_mel_value)
            return self.compiled_infer(padded, **kwargs)[:, :audio_len]
        elif mel.shape[0] > 1:
            return self.compiled_infer(mel, **kwargs)

        if isinstance(hop_len, float):
            hop_len = int(win_len * hop_len)
        if hop_len < 0:
            hop_len = win_len + hop_len
        starts = _get_steps(seq_len, win_len, hop_len)
        parts = [mel[:, start: start + win_len] for start in starts]
        overlaps = ((starts[:-1] + win_len) - starts[1:]) * 256
        if batch:
            stacked = np.concatenate([_to_numpy(p) for p in parts], axis=0)
            audio_parts = _to_numpy(self.compiled_infer(stacked, **kwargs))
        else:
            audio_parts = [_to_numpy(self.compiled_infer(p, **kwargs)[0]) for p in parts]
        audio = []
        for i, part in enumerate(audio_parts):
            start = 0 if i == 0 else overlaps[i - 1] // 2
            end = None if i == len(audio_parts) - 1 else -overlaps[i] // 2
            audio.append(part[start:end])
        return np.concatenate(audio, axis=-1)

    __call__ = infer

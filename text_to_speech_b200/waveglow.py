"""Drop-in for the reference's model wrapper `models.tts.WaveGlow` on the infer path
(models/tts/waveglow.py:61-144): same inputs ([T,80] / [B,T,80] / '.npy' path), same windowing
options (win_len / hop_len / force_pad / batch / use_slice / max_win_len), same stitching at the
overlap midpoints, same `[:, :T*256]` trim. The vocoder itself is a `Runtime` (runtime.py)."""
from __future__ import annotations

import math

import numpy as np

from .runtime import build_runtime

PAD_MEL_VALUE = -11.0   # models/tts/waveglow.py:29


def window_starts(n_frames, window, hop):
    """First frame of every window: evenly spread so that the last one ends exactly at `n_frames`
    (behaviour of `_get_steps`, models/tts/waveglow.py:156-164)."""
    count = int(math.ceil((n_frames - window) / hop)) + 1
    if count == 1:
        return [0]
    stride = (n_frames - window) / (count - 1)
    return np.round(np.arange(count) * stride).astype(np.int32)


_get_steps = window_starts      # the reference's name for it


def _to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _window_length(n_frames, win_len, use_slice, max_win_len):
    """A float `win_len` means "a multiple of this many frames": rounded up to cover the mel, or (use_slice) down to
    whole slices; `max_win_len` caps either (waveglow.py:84-90)."""
    if isinstance(win_len, float):
        if use_slice:
            win_len = max(1, n_frames // win_len) * int(win_len)
        else:
            win_len = int(math.ceil(n_frames / win_len) * win_len)
    return win_len if max_win_len is None else min(max_win_len, win_len)


def _stitch(pieces, starts, window, hop_samples=256):
    """Joins per-window waveforms at the midpoints of their overlaps (waveglow.py:136-142)."""
    overlap = ((starts[:-1] + window) - starts[1:]) * hop_samples
    last = len(pieces) - 1
    kept = []
    for i, piece in enumerate(pieces):
        head = overlap[i - 1] // 2 if i else 0
        tail = -overlap[i] // 2 if i < last else None
        kept.append(piece[head:tail])
    return np.concatenate(kept, axis=-1)


class WaveGlow:
    """`WaveGlow(path=..., runtime='b200', mode='bf16')(mel, sigma=0.6, z=z)` -> waveform [B, 256*T].

    Behaviour of `models.tts.WaveGlow.infer` (waveglow.py:61-144), case by case:
      no `win_len`                 one call on the whole mel, trimmed to 256*T samples
      mel not longer than a window one call; WITHOUT the caller's kwargs and untrimmed unless `force_pad` (the reference
                                   pads only for the keras runtime, :95-96), else padded with `pad_mel_value` and trimmed
      several utterances           one direct call, untrimmed (:108-112)
      one long utterance           windows of `win_len` frames every `hop_len` (negative = overlap, float = fraction),
                                   vocoded one by one or as one batch, stitched at the overlap midpoints
    """

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
        vocode = self.compiled_infer
        mel = np.load(mel) if isinstance(mel, str) else mel
        mel = mel[None] if len(mel.shape) == 2 else mel
        n_frames = mel.shape[1]
        n_samples = 256 * n_frames
        if win_len is None:
            return vocode(mel, **kwargs)[:, :n_samples]

        window = _window_length(n_frames, win_len, use_slice, max_win_len)
        kwargs["padding_multiple"] = window
        if n_frames <= window:
            pad = (self.runtime == "keras") if force_pad is None else force_pad
            if not pad:
                return vocode(mel)
            extra = max(window, n_frames) - n_frames
            padded = np.pad(_to_numpy(mel), [(0, 0), (0, extra), (0, 0)], constant_values=self.pad_mel_value)
            return vocode(padded, **kwargs)[:, :n_samples]
        if mel.shape[0] > 1:
            return vocode(mel, **kwargs)

        hop = int(window * hop_len) if isinstance(hop_len, float) else hop_len
        hop = window + hop if hop < 0 else hop
        starts = window_starts(n_frames, window, hop)
        windows = [mel[:, s: s + window] for s in starts]
        if batch:
            pieces = _to_numpy(vocode(np.concatenate([_to_numpy(w) for w in windows], axis=0), **kwargs))
        else:
            pieces = [_to_numpy(vocode(w, **kwargs)[0]) for w in windows]
        return _stitch(pieces, starts, window)

    __call__ = infer

"""Output side of the path: the waveform writer the reference's `AudioSaver` calls
(`utils/callbacks/file_saver.py:118-125` -> `utils/audio/audio_io.py:346-369` `write_audio` ->
`utils/audio/audio_processing.py:51-62` `normalize_audio`). Host-only byte work, restated in numpy.

`write_audio(filename, audio, rate, normalize=True, factor=32767)`: remove the mean, scale the peak to
`factor`, cast to int16 (truncation, like `ndarray.astype`) and write a PCM `.wav` with
`scipy.io.wavfile.write`, which is what the reference's `write_wav` does (:366-369). Other containers
(`.mp3` ... through pydub/ffmpeg, :371-400) are not reproduced: neither tool exists here.
"""
from __future__ import annotations

import numpy as np

__all__ = ["normalize_audio", "write_audio"]


def normalize_audio(audio, max_val=32767, dtype=np.int16):
    """audio_processing.py:51-62: int16 by default, float32 in [-1, 1] when `max_val <= 1.`."""
    if max_val <= 1.0:
        dtype = np.float32
    audio = np.asarray(audio)
    audio = audio - np.mean(audio)
    peak = np.max(np.abs(audio))
    if peak <= 1e-9:
        return audio.astype(dtype)
    return (audio * (max_val / peak)).astype(dtype)


def write_audio(filename, audio, rate, normalize=True, factor=32767):
    """audio_io.py:346-364 for the `.wav` extension. Returns `filename`."""
    ext = filename.split(".")[-1]
    if ext != "wav":
        raise ValueError("Unsupported file extension !\n  Accepted : ('wav',)\n  Got : {}".format(filename))
    audio = np.asarray(audio.detach().cpu().numpy() if hasattr(audio, "detach") else audio)
    if normalize and len(audio) > 0:
        audio = normalize_audio(audio, max_val=factor)
    from scipy.io.wavfile import write
    write(filename, rate, audio)
    return filename

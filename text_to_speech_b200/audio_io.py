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

__all__ = ["normalize_audio", "write_audio", "AudioSaver", "JSONSaver"]


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


class AudioSaver:
    """`utils/callbacks/file_saver.py:118-125` (+ `FileSaver.apply` :100-109): writes `output['audio']` to
    `<directory>/audios/audio-<n>.wav` with `write_audio(filename, audio, rate=output['rate'])` and records the path in
    `infos['audio']`. The reference's default container is `.mp3` (pydub / ffmpeg, absent here): `.wav` is the one written."""

    def __init__(self, directory, key="audio", file_format="audio-{}.wav", subdir="audios"):
        import os
        self.key, self.directory = key, os.path.join(directory, subdir) if subdir else directory
        self.file_format = file_format
        os.makedirs(self.directory, exist_ok=True)
        self._index = 0

    def apply(self, infos, output, **_):
        import os
        if isinstance(output.get(self.key), str):
            infos.setdefault(self.key, output[self.key])
            return
        if infos.get(self.key) is None:
            infos[self.key] = os.path.join(self.directory, self.file_format.format(self._index))
            self._index += 1
        write_audio(infos[self.key], output[self.key], rate=output["rate"])

    def join(self):
        pass


class JSONSaver:
    """The `map.json` of a prediction directory (`example_outputs/en/map.json`): {key text: {'text', 'cleaned'?,
    'splitted'?, 'rate', 'time', 'audio': path}} -- everything of an entry except the arrays ('mel', 'attention', raw
    'audio'; `models/tts/tacotron2.py:226-230`). Rewritten after every entry so a stream can be followed on disk."""

    def __init__(self, directory, filename="map.json"):
        import json
        import os
        os.makedirs(directory, exist_ok=True)
        self.path = os.path.join(directory, filename)
        self.data = {}
        if os.path.exists(self.path):
            with open(self.path, encoding="utf-8") as f:
                self.data = json.load(f)

    def apply(self, infos, output=None, **_):
        import json
        entry = {k: v for k, v in infos.items() if not isinstance(v, np.ndarray) and k not in ("mel", "attention")}
        self.data[infos["text"]] = entry
        with open(self.path, "w", encoding="utf-8") as f:
            json.dump(self.data, f, indent=4, default=float)

    def join(self):
        pass

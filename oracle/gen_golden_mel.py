"""ORACLE -- TEST INFRASTRUCTURE ONLY. Copies the reference's own fixture pair for the Tacotron
log-mel transform into tests/golden/mel_tacotron_stft.npz (run in the authoring container; the
reference tree does not exist on the GPU box).

    input : /root/reference/tests/__reproduction/audio_resample.npy  (test_utils_audio.py:62-64)
    output: /root/reference/tests/__reproduction/stft-TacotronSTFT.npy (test_utils_audio.py:109-111, max_err 2e-3)

    python -m oracle.gen_golden_mel
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPRO = "/root/reference/tests/__reproduction"


def main():
    audio = np.load(os.path.join(REPRO, "audio_resample.npy"))
    mel = np.load(os.path.join(REPRO, "stft-TacotronSTFT.npy"))
    assert audio.dtype == np.float32 and audio.ndim == 1 and mel.shape == (len(audio) // 256 + 1, 80)
    out = os.path.join(ROOT, "tests", "golden", "mel_tacotron_stft.npz")
    np.savez_compressed(out, audio_22050=audio, mel_reference=mel, reference_max_err=np.float64(2e-3),
                        source=np.frombuffer(b"reference tests/__reproduction (audio_resample.npy, stft-TacotronSTFT.npy)",
                                             dtype=np.uint8))
    print(out, os.path.getsize(out), "bytes; audio", audio.shape, "mel", mel.shape)


if __name__ == "__main__":
    main()

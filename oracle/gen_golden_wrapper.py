"""TEST INFRASTRUCTURE: golden outputs of the reference's OWN wrapper source (models/tts/waveglow.py `WaveGlow.infer`,
executed unmodified through oracle/run_reference_wrapper.py) over a deterministic stand-in vocoder, so that the product
wrapper is pinned to the reference's outputs also where /root/reference does not exist (the GPU box).

    python -m oracle.gen_golden_wrapper        # writes tests/golden/wrapper_cases.npz

Each case: mel = default_rng(seed).normal(size=(B, T, 80)) float32, kwargs as a python literal, and either shape / dtype /
sha256 / every 997th sample of the waveform the reference returns or the name of the exception it raises (its float-slice quirk with use_slice=True is part of the
contract)."""
import hashlib
import os

import numpy as np

from .run_reference_wrapper import reference_get_steps, reference_wrapper_infer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KWARGS = [dict(), dict(win_len=128), dict(win_len=64, hop_len=-16), dict(win_len=128, batch=True), dict(win_len=0.5),
          dict(win_len=3.0, use_slice=True), dict(win_len=512), dict(win_len=512, force_pad=True),
          dict(win_len=100, hop_len=0.5, max_win_len=80), dict(win_len=96, hop_len=-32, batch=True)]
SHAPES = [(1, 37), (1, 200), (1, 333), (3, 300)]
STEPS = [(37, 16, 12), (200, 64, 48), (200, 128, 64), (333, 128, 64), (1000, 256, 192), (300, 300, 100)]


class FakeVocoder:
    """sample s of frame t = mel[t, 0] * 1000 + position in window (same as tests/test_host_logic.py::_FakeRuntime)."""
    def __call__(self, mel, **kw):
        mel = np.asarray(mel)
        B, T, _ = mel.shape
        return np.repeat(mel[:, :, 0], 256, axis=1) * 1000.0 + np.arange(T * 256)[None] * 1e-3


def main():
    out = {}
    n = 0
    for (B, T) in SHAPES:
        for kw in KWARGS:
            if B > 1 and kw.get("win_len") != 128:
                continue
            seed = 1000 * B + T
            mel = np.random.default_rng(seed).normal(size=(B, T, 80)).astype(np.float32)
            out[f"case{n}_meta"] = np.frombuffer(repr((seed, B, T, kw)).encode(), dtype=np.uint8)
            try:
                wave = np.ascontiguousarray(reference_wrapper_infer(FakeVocoder(), mel, **kw))
                # the waveform itself is megabytes: keep its dtype / shape, a sha256 of its bytes (the comparison is bit-exact)
                # and every 997th sample for a readable diff
                out[f"case{n}_shape"] = np.asarray(wave.shape, dtype=np.int64)
                out[f"case{n}_dtype"] = np.frombuffer(str(wave.dtype).encode(), dtype=np.uint8)
                out[f"case{n}_sha256"] = np.frombuffer(hashlib.sha256(wave.tobytes()).hexdigest().encode(), dtype=np.uint8)
                out[f"case{n}_probe"] = wave.reshape(-1)[::997].copy()
            except Exception as e:
                out[f"case{n}_error"] = np.frombuffer(type(e).__name__.encode(), dtype=np.uint8)
            n += 1
    for j, a in enumerate(STEPS):
        out[f"steps{j}_args"] = np.asarray(a, dtype=np.int64)
        out[f"steps{j}"] = np.asarray(reference_get_steps(*a), dtype=np.int64)
    out["n_cases"], out["n_steps"] = np.asarray(n), np.asarray(len(STEPS))
    out["produced_by"] = np.frombuffer(b"reference models/tts/waveglow.py executed by oracle/run_reference_wrapper.py", dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wrapper_cases.npz"), **out)
    print(f"{n} wrapper cases, {len(STEPS)} step cases")


if __name__ == "__main__":
    main()

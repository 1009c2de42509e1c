import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Returns (hparams, weights, fixture dict). Weights are regenerated from the recorded seed and
    checked against the sha256 the fixture pins."""
    from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, weights_digest
    f = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    hp = WaveGlowHParams.from_json(bytes(f["hparams"]).decode())
    kw = dict(eval(bytes(f["weight_kwargs"]).decode()))
    w = generate_weights(hp, int(f["weight_seed"]), **kw)
    assert weights_digest(w) == bytes(f["weights_sha256"]).decode(), "weight generator drifted from the fixture"
    return hp, w, f


GOLDEN_CASES = ["tiny_c16", "nvidia_c32", "wg256_t24", "wg256_bias_t33", "wg256_k1", "wg512_t16"]


@pytest.fixture(scope="session")
def lib_built():
    from text_to_speech_b200 import _lib
    _lib.build_library()
    return _lib.load_library()


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(x, dtype=np.float64) - ref
    return 10.0 * np.log10((ref ** 2).mean() / max((err ** 2).mean(), 1e-300))

"""ORACLE -- TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz in the authoring container.

Each fixture = inputs (mel, z, sigma), the sha256 of the generated weight set, and the waveform
produced by the REFERENCE'S OWN SOURCE (oracle/run_reference.py; real keras if importable, else
the Keras shim -- recorded in the fixture) plus the float64 restatement as arbiter.

    python -m oracle.gen_golden            # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from text_to_speech_b200.weights import (WaveGlowHParams, generate_weights, synthetic_inputs,  # noqa: E402
                                         weights_digest)
from oracle.run_reference import reference_infer, which_keras                                # noqa: E402
from oracle.waveglow_oracle import OracleWaveGlow                                            # noqa: E402

# name -> (hparams, weight seed, weight kwargs, input seed, B, T, sigma, store_weights)
CASES = {
    # tiny topology (weights are regenerated from the seed; the fixture pins their sha256)
    "tiny_c16": (WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=3, n_channels=16),
                 7, dict(bias_std=0.05), 3, 2, 6, 0.6, False),
    # NVIDIA hparams, 12 flows / 8 layers, narrow channels: every flow-schedule branch, small file
    "nvidia_c32": (WaveGlowHParams(n_channels=32), 11, dict(bias_std=0.05), 4, 2, 40, 1.0, False),
    # WaveGlow-256, short utterance, batch 2 with ragged-free shapes
    "wg256_t24": (WaveGlowHParams(), 1234, dict(), 5, 1, 24, 0.6, False),
    "wg256_bias_t33": (WaveGlowHParams(), 4321, dict(bias_std=0.05), 6, 2, 33, 0.6, False),
    # BASELINE.json configs[0]: WaveGlow-256, one 200-frame mel, batch 1, fixed z, sigma 0.6
    "wg256_k1": (WaveGlowHParams(), 1234, dict(), 2024, 1, 200, 0.6, False),
    # WaveGlow-512 (reference default width)
    "wg512_t16": (WaveGlowHParams(n_channels=512), 99, dict(), 8, 1, 16, 0.6, False),
}


def main(out_dir=os.path.join(ROOT, "tests", "golden")):
    os.makedirs(out_dir, exist_ok=True)
    for name, (hp, wseed, wkw, iseed, B, T, sigma, store_w) in CASES.items():
        w = generate_weights(hp, wseed, **wkw)
        mel, z = synthetic_inputs(iseed, B, T, hp)
        ref = reference_infer(hp, w, mel, z, sigma=sigma)
        o64 = OracleWaveGlow(hp, w, torch.float64)(mel, z, sigma).numpy()
        det = reference_infer(hp, w, mel, None, sigma=sigma, deterministic=True)
        arrays = dict(
            hparams=np.frombuffer(hp.to_json().encode(), dtype=np.uint8),
            weight_seed=np.int64(wseed), weight_kwargs=np.frombuffer(repr(sorted(wkw.items())).encode(), dtype=np.uint8),
            weights_sha256=np.frombuffer(weights_digest(w).encode(), dtype=np.uint8),
            input_seed=np.int64(iseed), mel=mel, z=z, sigma=np.float64(sigma),
            wave_reference_fp32=ref.astype(np.float32), wave_oracle_fp64=o64.astype(np.float64),
            wave_reference_deterministic=det.astype(np.float32),
            produced_by=np.frombuffer(("reference source over " + which_keras()).encode(), dtype=np.uint8),
        )
        if store_w:
            arrays.update({"w:" + k: v for k, v in w.items()})
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(f"{name}: B={B} T={T} C={hp.n_channels} |wave|max={np.abs(ref).max():.3f} "
              f"ref-vs-fp64 {np.abs(ref - o64).max():.2e} ({which_keras()})")


if __name__ == "__main__":
    main()

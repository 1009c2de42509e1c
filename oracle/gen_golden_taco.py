"""ORACLE -- TEST INFRASTRUCTURE ONLY. Generates tests/golden/taco_decoder_*.npz in the authoring container.

Each fixture = the encoder memories of a few utterances (seeded), the sha256 of the seeded decoder weights, and
what the REFERENCE'S OWN Tacotron2Decoder.infer source (oracle/run_reference_taco.py; batch size 1 per utterance,
as the reference calls it) produces for them: decoder_output, stop_tokens, attention_weights, lengths -- in
float32 and in float64 (arbiter).

    python -m oracle.gen_golden_taco
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.run_reference import which_keras                                                   # noqa: E402
from oracle.run_reference_taco import reference_decode                                         # noqa: E402
from text_to_speech_b200.tacotron2 import Tacotron2HParams, generate_tacotron2_weights         # noqa: E402
from text_to_speech_b200.weights import weights_digest                                         # noqa: E402

SMALL = dict(embedding_dim=64, prenet_sizes=(32, 32), attention_rnn_dim=96, decoder_rnn_dim=96, attention_dim=24,
             attention_filters=8, attention_kernel_size=7, postnet_filters=48)
# name -> (hparams kwargs, weight seed, stop-gate bias, memory seed, text lengths, frames)
CASES = {
    "taco_decoder_small": (SMALL, 3, -3.0, 0, (11, 5, 17), 12),
    # NVIDIA dimensions (the ones csrc/taco.cu is built for): 1024-d LSTMs, 128-d attention, 31-tap location conv
    "taco_decoder_nvidia": ({}, 77, -10.0, 1, (23, 9, 40), 24),
}


def decoder_weights(hp, seed, gate_bias):
    w = generate_tacotron2_weights(hp, seed)
    w["decoder/gate_output/bias"][:] = gate_bias
    return {k: v for k, v in w.items() if k.startswith("decoder/")}


def memories(hp, seed, lengths):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal((s, hp.embedding_dim)) * 0.5).astype(np.float32) for s in lengths]


def main(out_dir=os.path.join(ROOT, "tests", "golden")):
    for name, (kw, wseed, gate_bias, mseed, lengths, frames) in CASES.items():
        hp = Tacotron2HParams(**kw)
        w = decoder_weights(hp, wseed, gate_bias)
        arrays = dict(hparams=np.frombuffer(repr(sorted(kw.items())).encode(), dtype=np.uint8), weight_seed=np.int64(wseed),
                      gate_bias=np.float64(gate_bias), weights_sha256=np.frombuffer(weights_digest(w).encode(), dtype=np.uint8),
                      memory_seed=np.int64(mseed), text_lengths=np.asarray(lengths, np.int64), frames=np.int64(frames),
                      produced_by=np.frombuffer(("reference Tacotron2Decoder.infer source over " + which_keras()).encode(),
                                                dtype=np.uint8))
        for i, mem in enumerate(memories(hp, mseed, lengths)):
            r32 = reference_decode(hp, w, mem, frames, dtype="float32")
            r64 = reference_decode(hp, w, mem, frames, dtype="float64")
            assert r32["lengths"] == r64["lengths"] == frames
            for key in ("decoder_output", "stop_tokens", "attention_weights"):
                arrays[f"u{i}_{key}_fp32"] = r32[key].astype(np.float32)
                arrays[f"u{i}_{key}_fp64"] = r64[key].astype(np.float64)
            print(f"{name} u{i}: S={len(mem)} frames={frames} |out|max={np.abs(r64['decoder_output']).max():.3f} "
                  f"fp32-fp64 {np.abs(r32['decoder_output'] - r64['decoder_output']).max():.2e}")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, os.path.getsize(os.path.join(out_dir, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()

"""Soak: repeat the hot paths many times and require bit-identical results every time (races in the mbarrier /
cp.async protocols would show up as rare mismatches). python tools/soak.py [n_wg] [n_taco]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine
from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights
from text_to_speech_b200.tts import synthetic_texts
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs

n_wg = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_taco = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hp = WaveGlowHParams()
eng = WaveGlowEngine(hp, generate_weights(hp, 1234), mode="bf16", device=0)
bad = 0
for B, T in ((16, 860), (3, 333), (1, 37)):
    mel, z = synthetic_inputs(1, B, T, hp)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    ref = eng.infer_device(mel_d, z_d, 0.6).clone()
    n = n_wg if B == 16 else 3 * n_wg
    for i in range(n):
        out = eng.infer_device(mel_d, z_d, 0.6)
        if not torch.equal(out, ref):
            bad += 1
            print(f"WaveGlow {B}x{T}: run {i} differs, max {float((out - ref).abs().max()):.3e}")
    torch.cuda.synchronize()
    print(f"WaveGlow {B}x{T}: {n} runs, finite {bool(torch.isfinite(ref).all())}")
thp = Tacotron2HParams()
tw = generate_tacotron2_weights(thp, 77)
tw["decoder/gate_output/bias"][:] = -10.0
m = Tacotron2(thp, tw, device="cuda")
toks = np.stack(synthetic_texts(16, 99, 86, 86))
ref = m.infer(toks, max_length=200, early_stopping=False, deterministic=True, decoder="b200").decoder_output.clone()
for i in range(n_taco):
    out = m.infer(toks, max_length=200, early_stopping=False, deterministic=True, decoder="b200").decoder_output
    if not torch.equal(out, ref):
        bad += 1
        print(f"decoder: run {i} differs, max {float((out - ref).abs().max()):.3e}")
print(f"decoder: {n_taco} runs of 200 frames")
print("SOAK", "FAILED" if bad else "OK", bad)
sys.exit(1 if bad else 0)

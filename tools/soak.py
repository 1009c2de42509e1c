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
weights = generate_weights(hp, 1234)
bad = 0
# (engine mode, WG_PAIR, shapes): the BF16 layer kernel as the engine picks it (CTA pairs for the big batch), the single-CTA
# kernel forced, a ragged batch on either, and the tf32x3 kernels: one persistent launch per flow with grid barriers (1 x 200,
# the ragged 5 x 90), the per-layer CTA-pair kernels (2 x 150, 4 x 860) and the single-CTA ones
for mode, pair, cases in (("bf16", "-1", ((16, 860, None), (3, 333, None), (1, 37, None), (12, 300, "ragged"))),
                          ("bf16", "0", ((16, 860, None), (12, 300, "ragged"))),
                          ("tf32x3", "-1", ((1, 200, None), (5, 90, "ragged"), (2, 150, None), (4, 860, None))),
                          ("tf32x3", "0", ((1, 200, None), (2, 150, None)))):
    os.environ["WG_PAIR"] = pair
    eng = WaveGlowEngine(hp, weights, mode=mode, device=0)
    for B, T, ragged in cases:
        mel, z = synthetic_inputs(1, B, T, hp)
        lengths = [max(1, T - 23 * b) for b in range(B)] if ragged else None
        mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
        ref = eng.infer_device(mel_d, z_d, 0.6, lengths=lengths).clone()
        n = (n_wg if B * T > 5000 else 3 * n_wg) // (1 if mode == "bf16" or B * T <= 500 else 6)
        for i in range(n):
            out = eng.infer_device(mel_d, z_d, 0.6, lengths=lengths)
            if not torch.equal(out, ref):
                bad += 1
                print(f"WaveGlow {mode} {B}x{T}: run {i} differs, max {float((out - ref).abs().max()):.3e}")
        torch.cuda.synchronize()
        print(f"WaveGlow {mode} WG_PAIR={pair} {B}x{T}{' ragged' if ragged else ''}: {n} runs, pair kernel {eng.pair_info()[1]}, "
              f"launches per infer {eng.last_launch_count}, finite {bool(torch.isfinite(ref).all())}", flush=True)
    eng.close()
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

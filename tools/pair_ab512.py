"""Same-box A/B of the CTA-pair WaveGlow-512 gate kernel (tc512_gate_pair_kernel) against the single-CTA one: bitwise
comparison on small shapes (ghost tile, ragged), then K3 (32 x 860) timing, alternating runs."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
hp = WaveGlowHParams(n_channels=512)
w = generate_weights(hp, 1234)
os.environ["WG_PM"] = "1"
os.environ["WG_PAIR"] = "0"
single = WaveGlowEngine(hp, w, mode="bf16")
os.environ["WG_PAIR"] = "1"
pair = WaveGlowEngine(hp, w, mode="bf16")
del w


def run(eng, mel, z, lengths=None):
    out = eng.infer_device(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda(), 0.6, lengths=lengths)
    torch.cuda.synchronize()
    return out.cpu().numpy()


ok = True
for (B, T, lengths) in [(1, 12, None), (2, 150, None), (3, 300, None), (4, 97, [97, 5, 33, 64])]:
    mel, z = synthetic_inputs(B * 1000 + T, B, T, hp)
    a, b = run(single, mel, z, lengths), run(pair, mel, z, lengths)
    same = bool(np.array_equal(a, b))
    ok &= same
    print(json.dumps({"B": B, "T": T, "ragged": lengths is not None, "bitwise_equal": same, "max_abs_diff": float(np.abs(a - b).max()),
                      "finite": bool(np.isfinite(b).all()), "pair_used": pair.pair_info()[1]}), flush=True)
g = torch.Generator(device="cuda").manual_seed(11)
md = torch.clamp(torch.randn(32, 860, 80, generator=g, device="cuda") * 2.0 - 5.2, -11.513, 1.2)
zd = torch.randn(32, 860 * 32, 8, generator=g, device="cuda")
out = torch.empty(32, 860 * 256, device="cuda")
res = {"single": [], "pair": []}
for name, eng in (("single", single), ("pair", pair)) * 2:
    for _ in range(2):
        eng.infer_device(md, zd, 0.6, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.infer_device(md, zd, 0.6, out=out)
    e1.record()
    torch.cuda.synchronize()
    res[name].append(e0.elapsed_time(e1) / reps)
    if name == "single":
        ref = out.clone()
    else:
        ok &= bool(torch.equal(out, ref))
print(json.dumps({"k3_ms_single": res["single"], "k3_ms_pair": res["pair"], "all_bitwise_equal": ok,
                  "resident_cta_pairs": pair.pair_info()[0]}), flush=True)

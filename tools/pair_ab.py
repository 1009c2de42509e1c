"""Same-box A/B of the CTA-pair (cta_group::2) WN-layer kernel against the single-CTA kernel: bitwise comparison of the
waveforms on a few shapes (incl. an odd tile count per phase block = a ghost tile, and a ragged batch), then K2 timing,
alternating runs.   python tools/pair_ab.py [reps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
hp = WaveGlowHParams()
w = generate_weights(hp, 1234)
os.environ["WG_PM"] = "1"
os.environ["WG_PAIR"] = "0"
single = WaveGlowEngine(hp, w, mode="bf16")
os.environ["WG_PAIR"] = "1"
pair = WaveGlowEngine(hp, w, mode="bf16")
os.environ["WG_PAIR_EPI"] = "16"
wide = WaveGlowEngine(hp, w, mode="bf16")      # CTA pairs with 16 epilogue warps (a -DWG_PROBES build; else the same as `pair`)


def run(eng, mel, z, lengths=None):
    out = eng.infer_device(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda(), 0.6, lengths=lengths)
    torch.cuda.synchronize()
    return out.cpu().numpy()


ok = True
for (B, T, lengths) in [(1, 12, None), (2, 150, None), (3, 300, None), (1, 200, None), (4, 97, [97, 5, 33, 64]), (16, 860, None)]:
    mel, z = synthetic_inputs(B * 1000 + T, B, T, hp)
    a, b, c = run(single, mel, z, lengths), run(pair, mel, z, lengths), run(wide, mel, z, lengths)
    same = bool(np.array_equal(a, b)) and bool(np.array_equal(a, c))
    ok &= same
    print(json.dumps({"B": B, "T": T, "ragged": lengths is not None, "bitwise_equal": same,
                      "max_abs_diff": float(max(np.abs(a - b).max(), np.abs(a - c).max())), "finite": bool(np.isfinite(c).all()),
                      "launches": pair.last_launch_count}), flush=True)
mel, z = synthetic_inputs(2024, 16, 860, hp)
md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
out = torch.empty(16, 860 * 256, device="cuda")
res = {"single": [], "pair": [], "pair16": []}
for name, eng in (("single", single), ("pair", pair), ("pair16", wide)) * 2:
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
print(json.dumps({"k2_ms_single": res["single"], "k2_ms_pair": res["pair"], "k2_ms_pair_16_epilogue_warps": res["pair16"], "all_bitwise_equal": ok,
                  "resident_cta_pairs": pair.pair_info()[0], "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}), flush=True)

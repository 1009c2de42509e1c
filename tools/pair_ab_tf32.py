"""Same-box A/B of the tf32x3 CTA-pair kernels (cta_group::2 kind::tf32) against the single-CTA kernels: bitwise comparison
of the waveforms on a few shapes (incl. an odd tile count per phase block = a ghost tile, and a ragged batch), then timing
of K1 (1 x 200 frames), 1 x 860 and 8 x 860, alternating runs.   python tools/pair_ab_tf32.py [reps]"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
hp = WaveGlowHParams()
w = generate_weights(hp, 1234)
engines = {}
os.environ["WG_TF32_FLOW"] = "0"       # the per-layer kernels; `default` below may run a flow as one persistent launch
for pr in ("0", "1"):
    os.environ["WG_PAIR"] = pr
    engines["pair" if pr == "1" else "single"] = WaveGlowEngine(hp, w, mode="tf32x3")
os.environ["WG_PAIR"] = "1"
for ew in ("8", "16"):
    os.environ["WG_TF32_EPI"] = ew                 # default: by shape (16 epilogue warps when every CTA runs one item)
    engines["pair_ew" + ew] = WaveGlowEngine(hp, w, mode="tf32x3")
del os.environ["WG_PAIR"], os.environ["WG_TF32_EPI"], os.environ["WG_TF32_FLOW"]
engines["default"] = WaveGlowEngine(hp, w, mode="tf32x3")


def run(eng, mel, z, lengths=None):
    out = eng.infer_device(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda(), 0.6, lengths=lengths)
    torch.cuda.synchronize()
    return out.cpu().numpy()


ok = True
for (B, T, lengths) in [(1, 12, None), (2, 150, None), (3, 300, None), (1, 200, None), (4, 97, [97, 5, 33, 64]), (8, 860, None)]:
    mel, z = synthetic_inputs(B * 1000 + T, B, T, hp)
    outs = {k: run(e, mel, z, lengths) for k, e in engines.items()}
    a = outs["single"]
    same = all(bool(np.array_equal(a, o)) for o in outs.values())
    ok &= same
    print(json.dumps({"B": B, "T": T, "ragged": lengths is not None, "bitwise_equal": same,
                      "sha256_default": hashlib.sha256(outs["default"].tobytes()).hexdigest()[:16],
                      "max_abs_diff": float(max(np.abs(a - o).max() for o in outs.values())),
                      "finite": all(bool(np.isfinite(o).all()) for o in outs.values()),
                      "default_used_pairs": engines["default"].pair_info()[1]}), flush=True)
for (B, T) in [(1, 200), (1, 860), (8, 860)]:
    mel, z = synthetic_inputs(2024, B, T, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    out = torch.empty(B, T * 256, device="cuda")
    res = {k: [] for k in engines}
    for name, eng in list(engines.items()) * 2:
        for _ in range(3):
            eng.infer_device(md, zd, 0.6, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.infer_device(md, zd, 0.6, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(round(e0.elapsed_time(e1) / reps, 3))
    print(json.dumps({"B": B, "T": T, "ms": res, "default_used_pairs": engines["default"].pair_info()[1]}), flush=True)
print(json.dumps({"all_bitwise_equal": ok, "resident_cta_pairs": engines["pair"].pair_info()[0],
                  "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}), flush=True)

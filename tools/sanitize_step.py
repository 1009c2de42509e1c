"""Small infers through every kernel path -- bf16 single-CTA / CTA-pair (forced) / ragged, tf32x3 uniform + ragged,
WaveGlow-512 ragged -- meant for `compute-sanitizer --tool memcheck`. That tool is closed on this pool; out-of-bounds WRITES
are covered by the canary test (tests/test_gpu_5_ragged.py::test_no_write_outside_the_workspace_or_the_output) instead."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

os.environ["WG_PM"] = "1"
for C, mode, pair in ((256, "bf16", "0"), (256, "bf16", "1"), (256, "tf32x3", "0"), (512, "bf16", "0")):
    os.environ["WG_PAIR"] = pair
    hp = WaveGlowHParams(n_channels=C, n_flows=4 if C == 512 else 12)
    w = generate_weights(hp, 7)
    eng = WaveGlowEngine(hp, w, mode=mode)
    mel, z = synthetic_inputs(3, 3, 40, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    a = eng.infer_device(md, zd, 0.6)
    b = eng.infer_device(md, zd, 0.6, lengths=[40, 7, 33])
    torch.cuda.synchronize()
    print(C, mode, "pair" if pair == "1" else "single", float(a.abs().max()), float(b.abs().max()), eng.last_launch_count, flush=True)
    assert bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all())
    eng.close()
print("ok")

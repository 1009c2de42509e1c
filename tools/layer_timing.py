"""In-kernel cycle counters of the WN-layer kernels (WG_LAYER_TIMING=1) on K2, single-CTA vs CTA-pair kernel.
Slots: [0] MMA warp total, [1] MMA waiting for TMA data, [2] MMA waiting for the epilogue, [3..5] epilogue waiting for
chunk a / chunk b / GEMM2, [6] gate-epilogue work, [7] residual-epilogue work, [8] TMA producer waiting for a free stage."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

hp = WaveGlowHParams()
w = generate_weights(hp, 1234)
mel, z = synthetic_inputs(2024, 16, 860, hp)
md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
os.environ["WG_LAYER_TIMING"] = "1"
for pair in ("0", "1"):
    os.environ["WG_PAIR"] = pair
    eng = WaveGlowEngine(hp, w, mode="bf16")
    for _ in range(2):
        eng.infer_device(md, zd, 0.6)
    torch.cuda.synchronize()
    eng.read_layer_timing()
    eng.infer_device(md, zd, 0.6)
    torch.cuda.synchronize()
    t = eng.read_layer_timing()
    tot = max(t[0], 1)
    n_mma_warps = 148 if pair == "0" else 74
    print(json.dumps({"pair": pair == "1", "mma_total_cycles_per_cta": t[0] / n_mma_warps / 96,
                      "mma_wait_tma_pct": 100 * t[1] / tot, "mma_wait_epilogue_pct": 100 * t[2] / tot,
                      "epi_wait_chunk_a_pct": 100 * t[3] / tot * (n_mma_warps / n_mma_warps), "epi_wait_chunk_b_pct": 100 * t[4] / tot,
                      "epi_wait_gemm2_pct": 100 * t[5] / tot, "epi_gate_work_pct": 100 * t[6] / tot,
                      "epi_resid_work_pct": 100 * t[7] / tot, "producer_wait_free_stage_pct": 100 * t[8] / tot}), flush=True)
    eng.close()

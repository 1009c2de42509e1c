"""In-kernel cycle counters of the tf32x3 kernels (WG_LAYER_TIMING=1), single-CTA vs CTA-pair, at K1 (1 x 200) and 8 x 860.
Slots (Tf32Params::timing): gate 16.., residual 32..: +0 MMA warp total, +1 waiting for TMA data, +2 waiting for the epilogue,
+3 epilogue waiting for an accumulator, +4 epilogue work, +5 kernel entry -> MMA loop, +6 producer waiting for a free
stage, +7 MMA-issuing CTAs (summed over the 96 / 84 launches)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402

hp = WaveGlowHParams()
w = generate_weights(hp, 1234)
os.environ["WG_LAYER_TIMING"] = "1"
os.environ["WG_TF32_FLOW"] = "0"
for (B, T) in [(1, 200), (8, 860)]:
    mel, z = synthetic_inputs(2024, B, T, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    for pair in ("0", "1"):
        os.environ["WG_PAIR"] = pair
        eng = WaveGlowEngine(hp, w, mode="tf32x3")
        for _ in range(2):
            eng.infer_device(md, zd, 0.6)
        torch.cuda.synchronize()
        eng.read_layer_timing()
        eng.infer_device(md, zd, 0.6)
        torch.cuda.synchronize()
        t = eng.read_layer_timing()
        for name, o in (("gate", 16), ("residual", 32)):
            ctas = max(t[o + 7], 1)     # MMA-issuing CTAs x launches
            print(json.dumps({"B": B, "T": T, "pair": pair == "1", "kernel": name, "mma_ctas_x_launches": t[o + 7],
                              "cycles_per_cta": {"mma_total": t[o] / ctas, "mma_wait_tma": t[o + 1] / ctas,
                                                 "mma_wait_epilogue": t[o + 2] / ctas, "entry_to_mma_loop": t[o + 5] / ctas,
                                                 "epi_wait_acc": t[o + 3] / ctas, "epi_work": t[o + 4] / ctas,
                                                 "producer_wait_free_stage": t[o + 6] / ctas}}), flush=True)
        eng.close()

# the one-launch-per-flow kernel (single-wave calls): slots 48.. summed over CTAs x 12 flows
os.environ["WG_TF32_FLOW"] = "1"
os.environ.pop("WG_PAIR", None)
mel, z = synthetic_inputs(2024, 1, 200, hp)
md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
eng = WaveGlowEngine(hp, w, mode="tf32x3")
for _ in range(2):
    eng.infer_device(md, zd, 0.6)
torch.cuda.synchronize()
eng.read_layer_timing()
eng.infer_device(md, zd, 0.6)
torch.cuda.synchronize()
t = eng.read_layer_timing()
n = max(t[58], 1)          # CTAs x flows
L = 8                      # layers per flow
print(json.dumps({"kernel": "tf32_flow_kernel", "ctas_x_flows": t[58], "cycles_per_cta_per_layer": {
    "kernel_total": t[56] / n / L, "producer_wait_h_barrier": t[48] / n / L, "producer_wait_acts_barrier": t[57] / n / L,
    "mma_wait_gate_operands": t[49] / (n / 2) / L, "mma_wait_residual_operands": t[50] / (n / 2) / L,
    "epi_wait_gate_acc": t[51] / n / L, "epi_gate_work": t[52] / n / L, "epi_acts_store_and_arrive": t[53] / n / L,
    "epi_arrive_to_residual_acc": t[54] / n / L, "epi_residual_work_store_arrive": t[55] / n / L},
    "note": "sums over CTAs (MMA slots: leader CTAs) x 12 flows, divided by CTAs x 8 layers"}), flush=True)

"""One warm-up + one infer (default: K2, BF16): the command ncu wraps for the per-kernel captures
(optional args: B T C mode, e.g. `1 200 256 tf32x3` = K1 on the tf32x3 kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 860)
C = int(sys.argv[3]) if len(sys.argv) > 3 else 256          # python tools/profile_step.py 32 860 512 = K3
hp = WaveGlowHParams(n_channels=C); w = generate_weights(hp, 1234)
mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
eng = WaveGlowEngine(hp, w, mode=mode)
mel, z = synthetic_inputs(1, B, T, hp)
md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
for _ in range(2):
    out = eng.infer_device(md, zd, 0.6)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))

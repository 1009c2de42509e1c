"""Where the end-to-end K2 step spends its host-visible time: staging copies, H2D, infer, D2H (each with a sync, so the parts
do not overlap as they do in the real call)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from text_to_speech_b200.weights import WaveGlowHParams, synthetic_inputs
hp = WaveGlowHParams()
mel, z = synthetic_inputs(1, 16, 860, hp)
pm = torch.empty(mel.size, dtype=torch.float32).pin_memory().view(mel.shape)
pz = torch.empty(z.size, dtype=torch.float32).pin_memory().view(z.shape)
po = torch.empty(16 * 860 * 256, dtype=torch.float32).pin_memory()
dm, dz, do = torch.empty_like(pm, device="cuda"), torch.empty_like(pz, device="cuda"), torch.empty(16 * 860 * 256, device="cuda")
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e3, 3)
res = {"threads": torch.get_num_threads()}
res["stage_mel_ms"] = t(lambda: pm.copy_(torch.from_numpy(mel)))
res["stage_z_ms"] = t(lambda: pz.copy_(torch.from_numpy(z)))
res["stage_z_numpy_ms"] = t(lambda: np.copyto(pz.numpy(), z))
res["h2d_mel_z_ms"] = t(lambda: (dm.copy_(pm, non_blocking=True), dz.copy_(pz, non_blocking=True)))
res["d2h_out_ms"] = t(lambda: po.copy_(do, non_blocking=True))
for nt in (1, 4, 16):
    torch.set_num_threads(nt)
    res[f"stage_z_ms_threads{nt}"] = t(lambda: pz.copy_(torch.from_numpy(z)))
print(json.dumps(res))

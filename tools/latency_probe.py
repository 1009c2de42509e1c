"""Single-utterance latency of the vocoder (the reference's call pattern, B = 1; models/tts/tacotron2.py:154-191):
device-resident launch sequence, eager vs CUDA-graph replay, and end to end through the plugin call with host buffers.
    python tools/latency_probe.py [--modes bf16,tf32x3,fp32] [--frames 200,860]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.engine import WaveGlowEngine  # noqa: E402
from text_to_speech_b200.runtime import B200WaveGlowRuntime  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs, save_weights  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="bf16,tf32x3,fp32")
    ap.add_argument("--frames", default="200,860")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    path = f"/tmp/wg_latency_{os.getpid()}.npz"
    save_weights(path, hp, w)
    for mode in args.modes.split(","):
        eng = WaveGlowEngine(hp, w, mode=mode, device=0)
        eager = B200WaveGlowRuntime(path, engine=eng, mode=mode, graph_max_frames=0)
        graph = B200WaveGlowRuntime(path, engine=eng, mode=mode, graph_max_frames=4096)
        for T in [int(x) for x in args.frames.split(",")]:
            mel, z = synthetic_inputs(7, 1, T, hp)
            md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
            out = torch.empty(1, T * 256, device="cuda")
            for _ in range(3):
                eng.infer_device(md, zd, 0.6, out=out)
                eager(mel, z=z, sigma=0.6)
                graph(mel, z=z, sigma=0.6)
            dev_ms = timed(lambda: eng.infer_device(md, zd, 0.6, out=out), args.reps)
            g = graph._graph_for(1, T, 0.6, False, None)
            graph_ms = timed(lambda: g.graph.replay(), args.reps)
            res = {}
            for name, rt in (("eager", eager), ("graph", graph)):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    rt(mel, z=z, sigma=0.6)
                res[name] = (time.perf_counter() - t0) / args.reps * 1e3
            a, b = eager(mel, z=z, sigma=0.6).copy(), graph(mel, z=z, sigma=0.6).copy()
            print(json.dumps({"mode": mode, "B": 1, "T": T, "audio_s": T * 256 / 22050, "launches": eng.last_launch_count,
                              "device_eager_ms": dev_ms, "device_graph_ms": graph_ms, "e2e_host_eager_ms": res["eager"],
                              "e2e_host_graph_ms": res["graph"], "graph_equals_eager": bool(np.array_equal(a, b))}), flush=True)
        eng.close()
    os.remove(path)


if __name__ == "__main__":
    main()

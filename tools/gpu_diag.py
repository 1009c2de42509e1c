"""Staged GPU diagnostics (development aid, run under gpurun): each stage prints error statistics
against the oracle so one remote call localises a bug. Usage: python tools/gpu_diag.py <stage>...
stages: fp32 | gemm | bf16 | time"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.waveglow_oracle import OracleWaveGlow                      # noqa: E402
from text_to_speech_b200.engine import WaveGlowEngine, debug_gemm_bf16  # noqa: E402
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs  # noqa: E402


def stats(name, got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    snr = 10 * np.log10((ref ** 2).mean() / max((err ** 2).mean(), 1e-300))
    print(f"  {name:34s} max|ref| {np.abs(ref).max():9.4f}  max err {err.max():.3e}  rms err {np.sqrt((err**2).mean()):.3e}  "
          f"SNR {snr:6.1f} dB  nan {int(np.isnan(got).sum())}", flush=True)
    return err.max()


def stage_fp32():
    for C, B, T in ((64, 2, 5), (256, 1, 24)):
        hp = WaveGlowHParams(n_channels=C)
        w = generate_weights(hp, 1234, bias_std=0.05)
        mel, z = synthetic_inputs(5, B, T, hp)
        taps = {}
        ref = OracleWaveGlow(hp, w).infer(mel, z, 0.6, taps=taps).numpy()
        eng = WaveGlowEngine(hp, w, mode="fp32")
        md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
        out = eng.infer_device(md, zd, sigma=0.6)
        torch.cuda.synchronize()
        print(f"fp32 C={C} B={B} T={T} launches={eng.last_launch_count}")
        stats("spect", eng.debug_spect(B, T).cpu().numpy(), taps["spect"].reshape(-1, 640).numpy())
        for (k, i) in [(11, -1), (11, 0), (11, 7), (10, 7), (0, 7)]:
            h, acc = eng.debug_prefix(md, zd, 0.6, k, i)
            torch.cuda.synchronize()
            if i >= 0:
                stats(f"h flow{k} layer{i}", h.cpu().numpy(), taps[f"flow{k}/layer{i}/audio"].reshape(-1, C).numpy())
        stats("waveform", out.cpu().numpy(), ref)


def stage_gemm():
    for (M, N, K) in ((128, 256, 64), (128, 256, 128), (256, 512, 320), (1000, 1024, 1408)):
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
        W = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
        bias = torch.randn(N, generator=g, device="cuda")
        D = debug_gemm_bf16(A, W, bias)
        torch.cuda.synchronize()
        ref = A.float() @ W.float().t() + bias
        print(f"gemm {M}x{N}x{K}")
        stats("D", D.cpu().numpy(), ref.cpu().numpy())
        if M == 128 and K == 64:
            e = (D - ref).abs().cpu().numpy()
            print("   err by row block of 8:", np.round(e.reshape(16, 8, N).max(axis=(1, 2)), 3))
            print("   err by col block of 32:", np.round(e.reshape(M, 8, 32).max(axis=(0, 2)), 3))


def stage_bf16():
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234, bias_std=0.05)
    B, T = 2, 7
    mel, z = synthetic_inputs(5, B, T, hp)
    taps = {}
    o = OracleWaveGlow(hp, w)
    ref = o.infer(mel, z, 0.6, taps=taps).numpy()
    eng = WaveGlowEngine(hp, w, mode="bf16")
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    print("bf16: upsample + start conv   WG_PM =", os.environ.get("WG_PM"))
    h, acc = eng.debug_prefix(md, zd, 0.6, 11, -1)
    torch.cuda.synchronize()
    if os.environ.get("WG_PM") != "1":
        stats("spect", eng.debug_spect(B, T).cpu().numpy(), taps["spect"].reshape(-1, 640).numpy())
    for (k, i) in [(11, 0), (11, 1), (11, 6), (11, 7), (10, 0), (10, 7), (0, 7)]:
        h, acc = eng.debug_prefix(md, zd, 0.6, k, i)
        torch.cuda.synchronize()
        nh = hp.flow_channels()[k][0]
        ref_acc = (o.w[f"block-{k}/end_conv/bias"] + taps[f"flow{k}/layer{i}/skip"] @ o.w[f"block-{k}/end_conv/kernel"][0])
        stats(f"h flow{k} layer{i}", h.cpu().numpy(), taps[f"flow{k}/layer{i}/audio"].reshape(-1, 256).numpy())
        stats(f"acc8 flow{k} layer{i}", acc.cpu().numpy()[:, :2 * nh], ref_acc.reshape(-1, 2 * nh).numpy())
    out = eng.infer_device(md, zd, sigma=0.6)
    torch.cuda.synchronize()
    stats("waveform", out.cpu().numpy(), ref)
    print("launches", eng.last_launch_count)


def stage_accuracy():
    """bf16 engine vs the CPU oracle on a larger sample (slow: CPU oracle)."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    B, T = 4, 860
    mel, z = synthetic_inputs(2024, B, T, hp)
    t = time.time()
    ref = OracleWaveGlow(hp, w)(mel, z, 0.6).numpy()
    print(f"oracle {B}x{T}: {time.time()-t:.1f}s on {torch.get_num_threads()} threads")
    for mode in ("bf16", "fp32"):
        eng = WaveGlowEngine(hp, w, mode=mode)
        out = eng.infer_device(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda(), sigma=0.6)
        torch.cuda.synchronize()
        stats(f"waveform {mode} {B}x{T}", out.cpu().numpy(), ref)
        e = np.abs(out.cpu().numpy() - ref)
        print("   |err| quantiles 99.9/99.99/99.999%:", np.quantile(e, [0.999, 0.9999, 0.99999]))
        eng.close()


def stage_c512():
    hp = WaveGlowHParams(n_channels=512)
    w = generate_weights(hp, 99, bias_std=0.05)
    B, T = 2, 9
    mel, z = synthetic_inputs(5, B, T, hp)
    taps = {}
    o = OracleWaveGlow(hp, w)
    ref = o.infer(mel, z, 0.6, taps=taps).numpy()
    t = time.time()
    eng = WaveGlowEngine(hp, w, mode="bf16")
    print(f"engine create {time.time()-t:.1f}s  WG_PM={os.environ.get('WG_PM')}")
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    for (k, i) in [(11, 0), (11, 1), (11, 7), (0, 7)]:
        h, acc = eng.debug_prefix(md, zd, 0.6, k, i)
        torch.cuda.synchronize()
        nh = hp.flow_channels()[k][0]
        ref_acc = (o.w[f"block-{k}/end_conv/bias"] + taps[f"flow{k}/layer{i}/skip"] @ o.w[f"block-{k}/end_conv/kernel"][0])
        stats(f"h flow{k} layer{i}", h.cpu().numpy(), taps[f"flow{k}/layer{i}/audio"].reshape(-1, 512).numpy())
        if i == 7:
            stats(f"acc8 flow{k} layer{i}", acc.cpu().numpy()[:, :2 * nh], ref_acc.reshape(-1, 2 * nh).numpy())
    out = eng.infer_device(md, zd, sigma=0.6)
    torch.cuda.synchronize()
    stats("waveform", out.cpu().numpy(), ref)


def stage_latency():
    """Single-utterance latency (B=1), eager launches vs a CUDA graph of the same call."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    eng = WaveGlowEngine(hp, w, mode="bf16")
    for T in (100, 300, 860):
        mel, z = synthetic_inputs(1, 1, T, hp)
        md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
        out = torch.empty(1, T * 256, device="cuda")
        for _ in range(3):
            eng.infer_device(md, zd, 0.6, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.infer_device(md, zd, 0.6, out=out)
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng.infer_device(md, zd, 0.6, out=out)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / 10
        print(f"latency B=1 T={T} ({T*256/22050:.1f} s of audio): eager {eager:.2f} ms, CUDA graph {graph:.2f} ms "
              f"-> {T*256/22050/(graph*1e-3):.0f} xRT")


def stage_time():
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    bt = tuple(int(x) for x in os.environ.get("WG_BT", "16,860").split(","))
    for mode, (B, T) in (("bf16", bt), ("fp32", (1, 200))):
        eng = WaveGlowEngine(hp, w, mode=mode)
        mel, z = synthetic_inputs(1, B, T, hp)
        md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
        out = torch.empty(B, T * 256, device="cuda")
        for _ in range(2):
            eng.infer_device(md, zd, 0.6, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            eng.infer_device(md, zd, 0.6, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flop = 20308852.0 * B * T * 256
        print(f"time {mode} B={B} T={T}: {ms:.2f} ms/infer, {B*T*256/ms*1e3/1e6:.2f} Msamples/s, {flop/ms/1e9:.1f} TFLOP/s algorithmic")
        if mode == "bf16" and os.environ.get("WG_LAYER_TIMING") == "1":
            t = eng.read_layer_timing()      # summed over CTAs and all launches of the timed + warm-up infers
            names = ["mma total", "mma wait TMA", "mma wait epilogue", "epi wait chunk a", "epi wait chunk b",
                     "epi wait gemm2", "epi gate work", "epi residual work", "tma wait free stage"]
            tot = max(t[0], 1)
            for nm, v in zip(names, t):
                print(f"    {nm:22s} {v/1e6:12.1f} Mcycles  {100.0*v/tot:6.1f}% of mma total")
            print(f"    MMA thread: issuing MMAs {100.0*t[80]/tot:5.1f}%  commit instructions {100.0*t[81]/tot:5.1f}%  syncwarp {100.0*t[82]/tot:5.1f}% (GEMM1 stages only)")
            hist = t[16:16 + 52]
            if sum(hist):
                print("    MMA wait-for-data by stage position in the tile (% of all such waiting):")
                tot_h = sum(hist)
                print("     ", " ".join(f"{100.0 * v / tot_h:4.1f}" for v in hist))
            if t[10]:
                print(f"    commit issue -> producer awake: {t[9]/t[10]:8.0f} cycles avg over {t[10]} stages")
            if t[12]:
                print(f"    TMA issue -> MMA thread sees data: {t[11]/t[12]:8.0f} cycles avg (all), "
                      f"{t[13]/max(t[14],1):8.0f} cycles avg over the {t[14]} stages the MMA waited for")
        eng.close()


if __name__ == "__main__":
    for st in sys.argv[1:]:
        t = time.time()
        print(f"===== stage {st} =====", flush=True)
        globals()["stage_" + st]()
        print(f"===== stage {st} done in {time.time()-t:.1f}s =====", flush=True)

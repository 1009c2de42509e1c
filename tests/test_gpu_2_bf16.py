"""GPU parity, WG_MODE_BF16 (tcgen05/TMEM/TMA path). Tolerances from BASELINE.json north_star:
max-abs <= 2e-2 and SNR >= 35 dB against the reference's fp32 waveform."""
import numpy as np
import pytest
import torch

from conftest import load_golden, snr_db
from oracle.waveglow_oracle import OracleWaveGlow
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs

pytestmark = pytest.mark.gpu
TOL_BF16_ABS, TOL_BF16_SNR = 2e-2, 35.0


def _engine(hp, w, mode="bf16"):
    from text_to_speech_b200.engine import WaveGlowEngine
    return WaveGlowEngine(hp, w, mode=mode, device=0)


def _run(eng, mel, z, sigma, deterministic=False):
    mel_d = torch.from_numpy(np.ascontiguousarray(mel)).cuda()
    z_d = None if z is None else torch.from_numpy(np.ascontiguousarray(z)).cuda()
    out = eng.infer_device(mel_d, z_d, sigma=sigma, deterministic=deterministic)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 320), (300, 512, 256), (1000, 1024, 1408)])
def test_tcgen05_gemm_against_torch_fp32(lib_built, M, N, K):
    from text_to_speech_b200.engine import debug_gemm_bf16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    D = debug_gemm_bf16(A, W, bias)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias          # plain PyTorch fp32 reference of the same op
    err = (D - ref).abs().max().item()
    assert err <= 2e-3 * (K ** 0.5), f"GEMM {M}x{N}x{K}: max err {err}"


@pytest.mark.parametrize("case", ["wg256_t24", "wg256_bias_t33", "wg256_k1"])
def test_bf16_matches_golden(lib_built, case):
    hp, w, f = load_golden(case)
    eng = _engine(hp, w)
    out = _run(eng, f["mel"], f["z"], float(f["sigma"]))
    ref = f["wave_reference_fp32"]
    err, snr = np.abs(out - ref).max(), snr_db(ref, out)
    print(f"{case}: bf16 max-abs {err:.3e}, SNR {snr:.1f} dB")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    eng.close()


def test_bf16_intermediates_against_oracle_taps(lib_built, monkeypatch):
    monkeypatch.setenv("WG_PM", "0")     # spect is only materialised by the position-major path
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234, bias_std=0.05)
    mel, z = synthetic_inputs(5, 2, 7, hp)        # L = 224: a full tile + a ragged tile per utterance
    taps = {}
    o = OracleWaveGlow(hp, w)
    o.infer(mel, z, 0.6, taps=taps)
    eng = _engine(hp, w)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    _run(eng, mel, z, 0.6)
    spect = eng.debug_spect(2, 7).cpu().numpy()
    ref_s = taps["spect"].reshape(spect.shape).numpy()
    assert np.abs(spect - ref_s).max() <= 2e-2 * max(1.0, np.abs(ref_s).max())
    for (k, i) in [(11, 0), (11, 1), (11, 6), (11, 7), (3, 7)]:
        h, acc = eng.debug_prefix(mel_d, z_d, 0.6, k, i)
        torch.cuda.synchronize()
        ref_h = taps[f"flow{k}/layer{i}/audio"].reshape(-1, hp.n_channels).numpy()
        nh = hp.flow_channels()[k][0]
        ref_acc = o.w[f"block-{k}/end_conv/bias"] + taps[f"flow{k}/layer{i}/skip"] @ o.w[f"block-{k}/end_conv/kernel"][0]
        ref_acc = ref_acc.reshape(-1, 2 * nh).numpy()
        eh = np.abs(h.cpu().numpy() - ref_h).max()
        ea = np.abs(acc.cpu().numpy()[:, :2 * nh] - ref_acc).max()
        print(f"flow {k} layer {i}: h err {eh:.3e} (|h| {np.abs(ref_h).max():.2f}), acc err {ea:.3e}")
        assert eh <= 5e-2 * max(1.0, np.abs(ref_h).max()), (k, i)
        # the folded skip/end accumulator carries the bias terms of ALL layers from the start, so it is
        # comparable with end(skip) only once the last layer has been added
        assert i != hp.n_layers - 1 or ea <= 3e-2, (k, i)


@pytest.mark.parametrize("B,T", [(1, 1), (1, 4), (2, 5), (3, 13), (1, 37)])
def test_bf16_ragged_shapes_against_oracle(lib_built, B, T):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(200 + B * 10 + T, B, T, hp)
    ref = OracleWaveGlow(hp, w)(mel, z, 0.6).numpy()
    out = _run(_engine(hp, w), mel, z, 0.6)
    assert np.abs(out - ref).max() <= TOL_BF16_ABS and snr_db(ref, out) >= TOL_BF16_SNR


@pytest.mark.parametrize("pm", ["0", "1"])
@pytest.mark.parametrize("B,T", [(2, 5), (3, 33), (1, 130), (2, 200)])
def test_bf16_both_row_layouts_against_oracle(lib_built, monkeypatch, pm, B, T):
    """WG_PM=0: position-major rows + materialised spect; WG_PM=1: phase-major rows, conditioning folded to
    its rank-320 form (Wup_r @ Wcond). Both must meet the BF16 bar on ragged and multi-tile shapes."""
    monkeypatch.setenv("WG_PM", pm)
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234, bias_std=0.05)
    mel, z = synthetic_inputs(300 + B * 10 + T, B, T, hp)
    ref = OracleWaveGlow(hp, w)(mel, z, 0.6).numpy()
    eng = _engine(hp, w)
    out = _run(eng, mel, z, 0.6)
    err, snr = np.abs(out - ref).max(), snr_db(ref, out)
    print(f"pm={pm} B={B} T={T}: max-abs {err:.3e} SNR {snr:.1f} dB")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    taps = {}
    OracleWaveGlow(hp, w).infer(mel, z, 0.6, taps=taps)
    h, acc = eng.debug_prefix(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda(), 0.6, 11, 7)
    torch.cuda.synchronize()
    ref_h = taps["flow11/layer7/audio"].reshape(-1, hp.n_channels).numpy()
    assert np.abs(h.cpu().numpy() - ref_h).max() <= 5e-2 * max(1.0, np.abs(ref_h).max())
    eng.close()


@pytest.mark.parametrize("C,B,T", [(256, 2, 5), (256, 1, 130), (256, 3, 200), (512, 2, 5), (512, 1, 140)])
def test_bf16_start_fold_against_oracle_and_unfolded_path(lib_built, monkeypatch, C, B, T):
    """Phase-major path with the start conv folded into each flow's first WN layer (WG_FOLD0=1, default): h0 is never
    materialised; layer 0 reads the audio rows through a0_build_kernel. Must meet the BF16 bar, agree with the
    unfolded path (WG_FOLD0=0) far inside it, and give the right residual stream right after layer 0 (edges included:
    the tap rows l-1 / l+1 outside the utterance are the reference's zero padding of h0, bias and all)."""
    monkeypatch.setenv("WG_PM", "1")
    hp = WaveGlowHParams(n_channels=C)
    w = generate_weights(hp, 1234, bias_std=0.05)
    mel, z = synthetic_inputs(900 + B * 10 + T, B, T, hp)
    taps = {}
    ref = OracleWaveGlow(hp, w).infer(mel, z, 0.6, taps=taps).numpy()
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    outs, h0s = {}, {}
    for fold in ("1", "0"):
        monkeypatch.setenv("WG_FOLD0", fold)
        eng = _engine(hp, w)
        outs[fold] = _run(eng, mel, z, 0.6)
        launches = eng.last_launch_count
        for (k, i) in [(11, 0), (5, 0), (0, 0), (0, 1)]:
            h, _ = eng.debug_prefix(mel_d, z_d, 0.6, k, i)
            torch.cuda.synchronize()
            h0s[fold, k, i] = h.cpu().numpy()
        eng.close()
        err, snr = np.abs(outs[fold] - ref).max(), snr_db(ref, outs[fold])
        print(f"WG_FOLD0={fold} C={C} B={B} T={T}: max-abs {err:.3e} SNR {snr:.1f} dB, {launches} launches")
        assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    assert np.abs(outs["1"] - outs["0"]).max() <= TOL_BF16_ABS
    for (k, i) in [(11, 0), (5, 0), (0, 0), (0, 1)]:
        ref_h = taps[f"flow{k}/layer{i}/audio"].reshape(-1, hp.n_channels).numpy()
        scale = max(1.0, np.abs(ref_h).max())
        e1, e0 = np.abs(h0s["1", k, i] - ref_h).max(), np.abs(h0s["0", k, i] - ref_h).max()
        print(f"flow {k} layer {i}: h err folded {e1:.3e} / unfolded {e0:.3e} (|h| {scale:.2f})")
        assert e1 <= 5e-2 * scale and e0 <= 5e-2 * scale
        # first and last position of every utterance: the zero-padded taps
        L = T * 32
        edge = np.concatenate([np.arange(B) * L, np.arange(B) * L + L - 1])
        assert np.abs(h0s["1", k, i][edge] - ref_h[edge]).max() <= 5e-2 * scale


@pytest.mark.parametrize("pm", ["0", "1"])
def test_bf16_waveglow512_matches_golden(lib_built, monkeypatch, pm):
    """WaveGlow-512 (reference default width, BASELINE.json configs[2]) on the two-kernel tcgen05 layer."""
    monkeypatch.setenv("WG_PM", pm)
    hp, w, f = load_golden("wg512_t16")
    eng = _engine(hp, w)
    out = _run(eng, f["mel"], f["z"], float(f["sigma"]))
    ref = f["wave_reference_fp32"]
    err, snr = np.abs(out - ref).max(), snr_db(ref, out)
    print(f"wg512_t16 pm={pm}: bf16 max-abs {err:.3e}, SNR {snr:.1f} dB")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    eng.close()


def test_bf16_waveglow512_multi_tile_against_oracle(lib_built, monkeypatch):
    monkeypatch.setenv("WG_PM", "1")   # the automatic row layout depends on (B, T); bit-identity holds within one layout
    hp = WaveGlowHParams(n_channels=512)
    w = generate_weights(hp, 99, bias_std=0.05)
    mel, z = synthetic_inputs(77, 2, 150, hp)          # phase-major: 3 tiles of 128 rows per phase over 2 x (150 + 4) rows
    ref = OracleWaveGlow(hp, w)(mel, z, 0.6).numpy()
    eng = _engine(hp, w)
    out = _run(eng, mel, z, 0.6)
    err, snr = np.abs(out - ref).max(), snr_db(ref, out)
    print(f"wg512 2x150: bf16 max-abs {err:.3e}, SNR {snr:.1f} dB")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    assert np.array_equal(_run(eng, mel, z, 0.6), out)                       # reproducible
    assert np.array_equal(_run(eng, mel[1:2], z[1:2], 0.6)[0], out[1])       # utterances independent
    eng.close()


def test_bf16_agrees_with_fp32_engine_and_is_reproducible(lib_built):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(9, 3, 21, hp)
    e16, e32 = _engine(hp, w, "bf16"), _engine(hp, w, "fp32")
    a, b = _run(e16, mel, z, 0.6), _run(e32, mel, z, 0.6)
    assert np.abs(a - b).max() <= TOL_BF16_ABS and snr_db(b, a) >= TOL_BF16_SNR
    assert np.array_equal(_run(e16, mel, z, 0.6), a)                 # bit-reproducible run to run
    for bi in range(3):                                              # utterances never interact
        assert np.array_equal(_run(e16, mel[bi:bi + 1], z[bi:bi + 1], 0.6)[0], a[bi])


def test_bf16_full_size_properties(lib_built):
    """BASELINE.json configs[1] (WaveGlow-256, 16 x 860 frames): too big for the CPU oracle in a test,
    so check size-independent properties: utterance b of the batch == the same utterance run alone
    (bit-identical), finite output, and a CPU-oracle spot check on one short utterance embedded in it."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(2024, 16, 860, hp)
    eng = _engine(hp, w)
    full = _run(eng, mel, z, 0.6)
    assert full.shape == (16, 860 * 256) and np.isfinite(full).all()
    for b in (0, 7, 15):
        assert np.array_equal(_run(eng, mel[b:b + 1], z[b:b + 1], 0.6)[0], full[b])
    # the first 40 frames of an utterance only see frames < 40 + receptive field; compare the first
    # 8 frames (2048 samples) with the oracle run on a 120-frame prefix
    ref = OracleWaveGlow(hp, w)(mel[3:4, :120], z[3:4, :120 * 32], 0.6).numpy()
    assert np.abs(full[3, :2048] - ref[0, :2048]).max() <= TOL_BF16_ABS


def test_bf16_full_size_parity_against_fp32_engine(lib_built):
    """The headline configuration (WaveGlow-256, 16 x 860 frames) end to end: the BF16 tcgen05 path against the
    fp32 engine (itself pinned to the reference fp32 waveform at <= 1e-4 by test_gpu_1_fp32.py) on all
    3.5 M samples. Bars from BASELINE.json: max-abs <= 2e-2, SNR >= 35 dB."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(2024, 16, 860, hp)
    a = _run(_engine(hp, w, "bf16"), mel, z, 0.6)
    b = _run(_engine(hp, w, "fp32"), mel, z, 0.6)
    err, snr = np.abs(a - b).max(), snr_db(b, a)
    print(f"K2 full size: bf16 vs fp32 engine max-abs {err:.3e}, SNR {snr:.1f} dB, |wave|max {np.abs(b).max():.2f}")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR


def test_waveglow512_full_size_parity_against_fp32_engine(lib_built):
    """BASELINE.json configs[2] at its full size (WaveGlow-512, 32 x 860 frames, 7.0 M samples): the two-kernel tcgen05
    layer against the fp32 engine (pinned to the reference fp32 waveform at <= 1e-4 on the wg512_t16 fixture) on ALL
    samples, a CPU-oracle check of an utterance prefix, run-to-run reproducibility and batch independence."""
    hp = WaveGlowHParams(n_channels=512)
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(2026, 32, 860, hp)
    e16 = _engine(hp, w, "bf16")
    a = _run(e16, mel, z, 0.6)
    assert a.shape == (32, 860 * 256) and np.isfinite(a).all()
    assert np.array_equal(_run(e16, mel, z, 0.6), a)
    assert np.array_equal(_run(e16, mel[17:18], z[17:18], 0.6)[0], a[17])
    e16.close()
    e32 = _engine(hp, w, "fp32")
    b = np.concatenate([_run(e32, mel[i:i + 8], z[i:i + 8], 0.6) for i in range(0, 32, 8)])   # 8 at a time: fp32 scratch is 9 KB/row
    e32.close()
    err, snr = np.abs(a - b).max(), snr_db(b, a)
    print(f"K3 full size: bf16 vs fp32 engine max-abs {err:.3e}, SNR {snr:.1f} dB, |wave|max {np.abs(b).max():.2f}")
    assert err <= TOL_BF16_ABS and snr >= TOL_BF16_SNR
    # the first 8 frames only see frames < 8 + receptive field (12 flows x 255 positions = 96 frames): oracle on a 120-frame prefix
    ref = OracleWaveGlow(hp, w)(mel[5:6, :120], z[5:6, :120 * 32], 0.6).numpy()
    assert np.abs(a[5, :2048] - ref[0, :2048]).max() <= TOL_BF16_ABS
    assert np.abs(b[5, :2048] - ref[0, :2048]).max() <= 1e-4


def test_runtime_plugin_end_to_end(lib_built, tmp_path):
    """The call a user of the reference makes: WaveGlow(runtime='b200', path=...)(mel, sigma=..., z=...)
    with host numpy buffers, plus the extra kwargs the reference's callers pass along."""
    from text_to_speech_b200.weights import save_weights
    from text_to_speech_b200.waveglow import WaveGlow
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    path = str(tmp_path / "wg256.npz")
    save_weights(path, hp, w)
    mel, z = synthetic_inputs(11, 1, 30, hp)
    voc = WaveGlow(path=path, runtime="b200", mode="bf16")
    out = voc(mel[0], sigma=0.6, z=z, directory="ignored", display=False)
    ref = OracleWaveGlow(hp, w)(mel, z, 0.6).numpy()
    assert out.shape == (1, 30 * 256) and np.abs(out - ref).max() <= TOL_BF16_ABS
    # sampling path (z omitted): finite, right shape, different draws differ
    s1, s2 = voc(mel, sigma=0.6), voc(mel, sigma=0.6)
    assert s1.shape == (1, 30 * 256) and np.isfinite(s1).all() and not np.array_equal(s1, s2)
    # device-resident call returns a CUDA tensor
    d = voc.model(torch.from_numpy(mel).cuda(), z=torch.from_numpy(z).cuda(), sigma=0.6)
    assert d.is_cuda and np.abs(d.cpu().numpy() - ref).max() <= TOL_BF16_ABS
    # windowed inference: the SAME wrapper (pinned to the reference's wrapper source by tests/test_host_logic.py) over the
    # B200 fp32 runtime and over a CPU-oracle vocoder -- windows, per-window calls and stitching included
    fp = WaveGlow(path=path, runtime="b200", mode="fp32")
    long_mel, _ = synthetic_inputs(12, 1, 100, hp)
    got = fp(long_mel, win_len=64, hop_len=-16, deterministic=True)
    cpu = WaveGlow.__new__(WaveGlow)
    cpu.runtime, cpu.pad_mel_value = "b200", -11.0
    cpu.model = lambda m, **kw: OracleWaveGlow(hp, w)(np.asarray(m), None, kw.get("sigma", 1.0), deterministic=True).numpy()
    want = cpu(long_mel, win_len=64, hop_len=-16, deterministic=True)
    # (the reference stitches the per-window `[0]` rows: a windowed single utterance comes back 1-D, waveglow.py:130-142)
    assert got.shape == want.shape == (100 * 256,) and np.abs(got - want).max() <= 1e-4


def test_wg_infer_is_cuda_graph_capturable(lib_built):
    """wg_infer does no allocation and no host synchronisation, so a caller can capture it into a CUDA graph
    (the launch-bound single-utterance case) and replay it with new data in the same buffers."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    eng = _engine(hp, w)
    mel, z = synthetic_inputs(41, 1, 150, hp)
    mel2, z2 = synthetic_inputs(42, 1, 150, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    out = torch.empty(1, 150 * 256, device="cuda")
    eager = eng.infer_device(md, zd, 0.6).clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.infer_device(md, zd, 0.6, out=out)            # warm-up on the side stream (workspace allocation)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.infer_device(md, zd, 0.6, out=out)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    md.copy_(torch.from_numpy(mel2)); zd.copy_(torch.from_numpy(z2))
    g.replay()
    torch.cuda.synchronize()
    ref = OracleWaveGlow(hp, w)(mel2, z2, 0.6).numpy()
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL_BF16_ABS


def test_maximum_sizes(lib_built):
    """Edge of the supported shape range: a 64 x 860-frame batch (4 x K2, 1.76 M rows, ~7 GB of workspace) must equal
    the same utterances run in K2-sized batches bit for bit (32-bit row/element indexing holds), a long single
    utterance (1 x 6000 frames = 70 s) must equal itself run inside a batch, and a shape whose element count
    would overflow the kernels' 32-bit indexing is refused before anything is launched."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    eng = _engine(hp, w)
    mel, z = synthetic_inputs(77, 64, 860, hp)
    big = _run(eng, mel, z, 0.6)
    assert big.shape == (64, 860 * 256) and np.isfinite(big).all()
    for b0 in (0, 48):
        assert np.array_equal(_run(eng, mel[b0:b0 + 16], z[b0:b0 + 16], 0.6), big[b0:b0 + 16])
    del big
    mel, z = synthetic_inputs(78, 2, 6000, hp)
    pair = _run(eng, mel, z, 0.6)
    assert np.array_equal(_run(eng, mel[1:2], z[1:2], 0.6)[0], pair[1])
    with pytest.raises(RuntimeError, match="too large"):
        eng.workspace_bytes(200, 1000)          # 6.4 M rows x 640 conditioning channels > 2^31 elements
    with pytest.raises(RuntimeError, match="positive"):
        eng.workspace_bytes(0, 10)


def test_concurrent_wg_infer_on_two_streams(lib_built):
    """include/wg_b200.h: a handle is immutable after wg_create, so concurrent wg_infer calls on different streams with
    different workspaces are safe. Two host threads drive the same handle through the raw C ABI; each result must be
    bit-identical to the serial run."""
    import threading
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    eng = _engine(hp, w)
    lib, h = eng._lib, eng._h
    jobs = []
    for seed, (B, T) in enumerate([(3, 120), (2, 333)]):
        mel, z = synthetic_inputs(500 + seed, B, T, hp)
        serial = _run(eng, mel, z, 0.6)
        ws = torch.empty(eng.workspace_bytes(B, T) + 1024, dtype=torch.uint8, device="cuda")
        ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        jobs.append(dict(B=B, T=T, serial=serial, mel=torch.from_numpy(mel).cuda(), z=torch.from_numpy(z).cuda(),
                         out=torch.zeros(B, T * 256, device="cuda"), ws=ws, ws_ptr=ws_ptr,
                         ws_bytes=eng.workspace_bytes(B, T), stream=torch.cuda.Stream(), rc=[]))
    torch.cuda.synchronize()

    def worker(j):
        for _ in range(5):
            j["rc"].append(lib.wg_infer(h, j["mel"].data_ptr(), j["z"].data_ptr(), 0.6, 0, j["B"], j["T"],
                                        j["out"].data_ptr(), j["ws_ptr"], j["ws_bytes"], j["stream"].cuda_stream))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    for j in jobs:
        assert j["rc"] == [0] * 5, lib.wg_last_error(h)
        assert np.array_equal(j["out"].cpu().numpy(), j["serial"])


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_profile_hooks_count_every_layer_launch(lib_built, mode):
    """wg_profile_enable / wg_profile_read (bench.py's roofline timing): one infer = n_flows * n_layers layer launches,
    a positive device time, and the counters re-arm after a read."""
    hp = WaveGlowHParams()
    eng = _engine(hp, generate_weights(hp, 1234), mode=mode)
    mel, z = synthetic_inputs(3, 2, 9, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    eng.infer_device(md, zd, 0.6)
    eng.profile_enable(True)
    for _ in range(2):
        eng.infer_device(md, zd, 0.6)
    ms, n = eng.profile_read()
    assert n == 2 * hp.n_flows * hp.n_layers and ms > 0.0
    eng.infer_device(md, zd, 0.6)
    ms1, n1 = eng.profile_read()
    assert n1 == hp.n_flows * hp.n_layers and 0.0 < ms1 < ms
    eng.profile_enable(False)
    eng.infer_device(md, zd, 0.6)
    assert eng.profile_read()[1] == 0
    eng.close()

"""GPU parity, WG_MODE_TF32X3: the fp32-grade mode on the tensor cores (tcgen05 kind::tf32, three products per
multiply on fp32 (hi, lo) operand pairs). Bar from BASELINE.json north_star: "an fp32/3xTF32 mode within 1e-4 max-abs
of the reference's fp32 waveform"."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.waveglow_oracle import OracleWaveGlow
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _engine(hp, w, mode="tf32x3"):
    from text_to_speech_b200.engine import WaveGlowEngine
    return WaveGlowEngine(hp, w, mode=mode, device=0)


def _run(eng, mel, z, sigma, deterministic=False, lengths=None):
    mel_d = torch.from_numpy(np.ascontiguousarray(mel)).cuda()
    z_d = None if z is None else torch.from_numpy(np.ascontiguousarray(z)).cuda()
    out = eng.infer_device(mel_d, z_d, sigma=sigma, deterministic=deterministic, lengths=lengths)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("case", ["wg256_t24", "wg256_bias_t33", "wg256_k1", "wg512_t16"])
def test_tf32x3_matches_golden(lib_built, case):
    """The reference-source fixtures (oracle/gen_golden.py), incl. BASELINE.json configs[0] (wg256_k1: 1 x 200 frames)
    and the 512-channel default width."""
    hp, w, f = load_golden(case)
    eng = _engine(hp, w)
    sigma = float(f["sigma"])
    out = _run(eng, f["mel"], f["z"], sigma)
    assert out.shape == f["wave_reference_fp32"].shape
    err = np.abs(out - f["wave_reference_fp32"]).max()
    err64 = np.abs(out - f["wave_oracle_fp64"]).max()
    print(f"{case}: tf32x3 err vs reference fp32 {err:.2e}, vs fp64 {err64:.2e}, launches {eng.last_launch_count}")
    assert err <= TOL and err64 <= TOL
    det = _run(eng, f["mel"], None, sigma, deterministic=True)
    assert np.abs(det - f["wave_reference_deterministic"]).max() <= TOL
    eng.close()


def test_tf32x3_intermediates_against_oracle_taps(lib_built):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 32, bias_std=0.05)
    mel, z = synthetic_inputs(5, 2, 7, hp)          # 2 x (7 + 4) phase-block rows: one partially filled tile per phase
    taps = {}
    OracleWaveGlow(hp, w).infer(mel, z, 0.6, taps=taps)
    eng = _engine(hp, w)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    for (k, i) in [(11, -1), (11, 0), (11, 7), (7, 3), (0, 6)]:
        h, acc = eng.debug_prefix(mel_d, z_d, 0.6, k, i)
        torch.cuda.synchronize()
        if i >= 0 and i < 7:
            ref_h = taps[f"flow{k}/layer{i}/audio"].reshape(-1, hp.n_channels).numpy()
            err = np.abs(h.cpu().numpy() - ref_h).max()
            print(f"flow {k} layer {i}: residual stream max-abs err {err:.2e} (|h|max {np.abs(ref_h).max():.2f})")
            assert err <= 2e-5 * max(1.0, np.abs(ref_h).max()), (k, i)
    eng.close()


@pytest.mark.parametrize("B,T", [(1, 1), (2, 5), (3, 37), (2, 150), (4, 860)])      # 4 x 860: many waves of CTA pairs, 10 s utterances
def test_tf32x3_shapes_against_fp32_engine_and_oracle(lib_built, B, T):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 31, bias_std=0.05)
    mel, z = synthetic_inputs(100 + B * 10 + T, B, T, hp)
    e3 = _engine(hp, w)
    out = _run(e3, mel, z, 0.8)
    ffma = _run(_engine(hp, w, "fp32"), mel, z, 0.8)
    err = np.abs(out - ffma).max()
    print(f"{B} x {T}: tf32x3 vs the FFMA fp32 engine {err:.2e}")
    assert err <= TOL
    if B * T <= 120:
        assert np.abs(out - OracleWaveGlow(hp, w)(mel, z, 0.8).numpy()).max() <= TOL
    assert np.array_equal(_run(e3, mel, z, 0.8), out)                                  # reproducible
    assert np.array_equal(_run(e3, mel[B - 1:B], z[B - 1:B], 0.8)[0], out[B - 1])      # utterances never interact
    e3.close()


def test_tf32x3_ragged_batch(lib_built):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    lengths = [40, 9, 23]
    mel, z = synthetic_inputs(3, 3, 40, hp)
    eng = _engine(hp, w)
    out = _run(eng, mel, z, 0.6, lengths=lengths)
    oracle = OracleWaveGlow(hp, w)
    for b, n in enumerate(lengths):
        alone = _run(eng, mel[b:b + 1, :n], z[b:b + 1, :n * 32], 0.6)[0]
        assert np.array_equal(out[b, :n * 256], alone) and not out[b, n * 256:].any()
        ref = oracle(mel[b:b + 1, :n], z[b:b + 1, :n * 32], 0.6).numpy()[0]
        assert np.abs(out[b, :n * 256] - ref).max() <= TOL
    eng.close()


def test_tf32x3_refuses_what_it_cannot_run(lib_built):
    from text_to_speech_b200.engine import WaveGlowEngine, WaveGlowError
    hp = WaveGlowHParams(n_channels=64)
    with pytest.raises(WaveGlowError, match="TF32X3"):
        WaveGlowEngine(hp, generate_weights(hp, 1), mode="tf32x3")


@pytest.mark.parametrize("B,T,lengths", [(1, 12, None), (3, 300, None), (1, 200, None), (4, 97, [97, 5, 33, 64])])
def test_tf32x3_kernel_variants_give_the_same_bits(lib_built, monkeypatch, B, T, lengths):
    """The CTA-pair kernels (cta_group::2 kind::tf32, an odd tile count per phase block = a ghost tile included) against
    the single-CTA kernels, 8 against 16 epilogue warps, and the one-launch-per-flow kernel (grid barriers) against the
    per-layer kernels: the engine picks among them by shape, so a ragged batch is bit-identical to its stand-alone
    utterances only if all of them produce the same bits."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(B * 1000 + T, B, T, hp)
    outs = {}
    monkeypatch.setenv("WG_TF32_FLOW", "0")          # the per-layer kernels
    for pair in ("0", "1"):
        for ew in ("8", "16"):
            monkeypatch.setenv("WG_PAIR", pair)
            monkeypatch.setenv("WG_TF32_EPI", ew)
            eng = _engine(hp, w)
            outs[(pair, ew)] = _run(eng, mel, z, 0.6, lengths=lengths)
            assert eng.pair_info()[1] == (pair == "1")
            eng.close()
    monkeypatch.delenv("WG_PAIR")
    monkeypatch.delenv("WG_TF32_EPI")
    monkeypatch.delenv("WG_TF32_FLOW")               # default: a flow as ONE persistent launch where the call is one wave
    eng = _engine(hp, w)
    ref = _run(eng, mel, z, 0.6, lengths=lengths)
    # one wave (<= 2 row tiles per phase: 64 CTA pairs): 12 flow launches + 12 boundaries + geometry / im2col instead of
    # 180 layer launches; 3 x 300 frames is 8 tiles per phase and stays on the per-layer kernels
    rows = sum(n + 4 for n in (lengths or [T] * B))          # phase-block rows incl. the 4 gap rows per utterance
    pair_items = -(-(-(-rows // 128)) // 2) * 32 * 2         # tile pairs per phase x 32 phases x 2 chunks
    one_wave = pair_items <= eng.pair_info()[0]              # every item has its own resident CTA pair (74 on a full B200)
    assert (eng.last_launch_count < 40) == one_wave, (eng.last_launch_count, pair_items, eng.pair_info())
    again = _run(eng, mel, z, 0.6, lengths=lengths)  # the grid barrier re-arms itself
    assert np.array_equal(ref, again)
    eng.close()
    monkeypatch.setenv("WG_TF32_FLOW", "2")          # the device refuses the cooperative launch -> per-layer kernels, same bits
    eng = _engine(hp, w)
    assert np.array_equal(ref, _run(eng, mel, z, 0.6, lengths=lengths)) and eng.last_launch_count > 150
    eng.close()
    assert np.isfinite(ref).all()
    for k, o in outs.items():
        assert np.array_equal(ref, o), k


def test_tf32x3_flow_kernel_under_a_cuda_graph_and_on_two_streams(lib_built):
    """tf32_flow_kernel is a cooperative launch: it must survive CUDA-graph capture / replay (the runtime's call path for
    short utterances) and two engines running on two streams at once (co-residency is the launch's, not the caller's,
    problem)."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    mel, z = synthetic_inputs(9, 1, 64, hp)
    md, zd = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    e1, e2 = _engine(hp, w), _engine(hp, w)
    ref = _run(e1, mel, z, 0.6)
    if e1.pair_info()[0] >= 64:                        # 1 x 64 frames = one tile per phase = 64 pair items: the flow kernel
        assert e1.last_launch_count < 40
    out = torch.empty(1, 64 * 256, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        e1.infer_device(md, zd, 0.6, out=out)          # warm-up on a side stream (workspace allocation)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        e1.infer_device(md, zd, 0.6, out=out)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    o1, o2 = torch.empty_like(out), torch.empty_like(out)
    for _ in range(5):
        with torch.cuda.stream(s1):
            e1.infer_device(md, zd, 0.6, out=o1)
        with torch.cuda.stream(s2):
            e2.infer_device(md, zd, 0.6, out=o2)
    torch.cuda.synchronize()
    assert np.array_equal(o1.cpu().numpy(), ref) and np.array_equal(o2.cpu().numpy(), ref)
    e1.close()
    e2.close()


def test_tf32x3_one_handle_on_two_streams(lib_built):
    """include/wg_b200.h: a handle is immutable after wg_create -- concurrent wg_infer calls with different workspaces
    are safe. For the flow kernel that means its grid-barrier words live in the CALLER's workspace, not in the handle:
    two host threads drive one handle (one call per flow kernel, one on the per-layer kernels), every result
    bit-identical to the serial run."""
    import threading
    hp = WaveGlowHParams()
    eng = _engine(hp, generate_weights(hp, 1234))
    lib, h = eng._lib, eng._h
    jobs = []
    for seed, (B, T) in enumerate([(1, 200), (1, 120), (2, 333)]):
        mel, z = synthetic_inputs(700 + seed, B, T, hp)
        serial = _run(eng, mel, z, 0.6)
        nbytes = eng.workspace_bytes(B, T)
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device="cuda")
        jobs.append(dict(B=B, T=T, serial=serial, mel=torch.from_numpy(mel).cuda(), z=torch.from_numpy(z).cuda(),
                         out=torch.zeros(B, T * 256, device="cuda"), ws=ws, ws_ptr=(ws.data_ptr() + 1023) // 1024 * 1024,
                         ws_bytes=nbytes, stream=torch.cuda.Stream(), rc=[]))
    torch.cuda.synchronize()

    def worker(j):
        for _ in range(8):
            j["rc"].append(lib.wg_infer(h, j["mel"].data_ptr(), j["z"].data_ptr(), 0.6, 0, j["B"], j["T"],
                                        j["out"].data_ptr(), j["ws_ptr"], j["ws_bytes"], j["stream"].cuda_stream))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    for j in jobs:
        assert j["rc"] == [0] * 8, lib.wg_last_error(h)
        assert np.array_equal(j["out"].cpu().numpy(), j["serial"])
    eng.close()

"""GPU parity, WG_MODE_FP32: the CUDA path (through the C ABI) against the oracle and the golden
fixtures. Tolerance from BASELINE.json north_star: max-abs <= 1e-4 vs the reference's fp32 waveform."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle.waveglow_oracle import OracleWaveGlow
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-4


def _engine(hp, w, mode="fp32"):
    from text_to_speech_b200.engine import WaveGlowEngine
    return WaveGlowEngine(hp, w, mode=mode, device=0)


def _run(eng, mel, z, sigma, deterministic=False):
    mel_d = torch.from_numpy(np.ascontiguousarray(mel)).cuda()
    z_d = None if z is None else torch.from_numpy(np.ascontiguousarray(z)).cuda()
    out = eng.infer_device(mel_d, z_d, sigma=sigma, deterministic=deterministic)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fp32_matches_golden(lib_built, case):
    hp, w, f = load_golden(case)
    eng = _engine(hp, w)
    sigma = float(f["sigma"])
    out = _run(eng, f["mel"], f["z"], sigma)
    assert out.shape == f["wave_reference_fp32"].shape
    err = np.abs(out - f["wave_reference_fp32"]).max()
    err64 = np.abs(out - f["wave_oracle_fp64"]).max()
    print(f"{case}: err vs reference fp32 {err:.2e}, vs fp64 {err64:.2e}")
    assert err <= TOL_FP32 and err64 <= TOL_FP32
    det = _run(eng, f["mel"], None, sigma, deterministic=True)
    assert np.abs(det - f["wave_reference_deterministic"]).max() <= TOL_FP32
    eng.close()


@pytest.mark.parametrize("B,T", [(1, 1), (1, 3), (2, 5), (3, 7), (1, 37), (5, 2)])
def test_fp32_ragged_shapes_against_oracle(lib_built, B, T):
    # L = 32 T is never a multiple of the 128-row tiles here; T=1 is the smallest legal input
    hp = WaveGlowHParams(n_channels=64)
    w = generate_weights(hp, 31, bias_std=0.05)
    mel, z = synthetic_inputs(100 + B * 10 + T, B, T, hp)
    ref = OracleWaveGlow(hp, w)(mel, z, 0.8).numpy()
    out = _run(_engine(hp, w), mel, z, 0.8)
    assert np.abs(out - ref).max() <= TOL_FP32


def test_fp32_intermediates_against_oracle_taps(lib_built):
    hp = WaveGlowHParams(n_channels=64)
    w = generate_weights(hp, 32, bias_std=0.05)
    mel, z = synthetic_inputs(5, 2, 6, hp)
    taps = {}
    OracleWaveGlow(hp, w).infer(mel, z, 0.6, taps=taps)
    eng = _engine(hp, w)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    _run(eng, mel, z, 0.6)
    spect = eng.debug_spect(2, 6).cpu().numpy()
    assert np.abs(spect - taps["spect"].reshape(spect.shape).numpy()).max() <= 1e-5
    for (k, i) in [(11, 0), (11, 7), (7, 3), (0, 7)]:
        h, acc = eng.debug_prefix(mel_d, z_d, 0.6, k, i)
        torch.cuda.synchronize()
        ref_h = taps[f"flow{k}/layer{i}/audio"].reshape(-1, hp.n_channels).numpy()
        assert np.abs(h.cpu().numpy() - ref_h).max() <= 1e-4, (k, i)


def test_fp32_batch_independence_and_sigma(lib_built):
    hp = WaveGlowHParams(n_channels=64)
    w = generate_weights(hp, 33)
    mel, z = synthetic_inputs(6, 3, 9, hp)
    eng = _engine(hp, w)
    full = _run(eng, mel, z, 1.0)
    for b in range(3):
        single = _run(eng, mel[b:b + 1], z[b:b + 1], 1.0)
        assert np.array_equal(single[0], full[b])           # utterances never interact: bit-identical
    assert np.array_equal(_run(eng, mel, z, 1.0), full)     # idempotent / deterministic
    assert not np.allclose(_run(eng, mel, z, 0.5), full)
    # deterministic=True == explicit zero noise
    assert np.array_equal(_run(eng, mel, None, 0.7, deterministic=True), _run(eng, mel, np.zeros_like(z), 0.7))


def test_host_entry_point_and_errors(lib_built):
    from text_to_speech_b200.engine import WaveGlowError
    hp = WaveGlowHParams(n_channels=32)
    w = generate_weights(hp, 34)
    mel, z = synthetic_inputs(7, 2, 4, hp)
    eng = _engine(hp, w)
    a = eng.infer_host(mel, z, 0.6)
    b = _run(eng, mel, z, 0.6)
    assert np.array_equal(a, b)
    assert eng.last_launch_count > 0
    with pytest.raises(ValueError):
        eng.infer_device(torch.zeros(2, 4, 79, device="cuda"), torch.from_numpy(z).cuda())
    with pytest.raises(ValueError):
        eng.infer_device(torch.from_numpy(mel).cuda(), torch.from_numpy(z[:, :-1]).cuda())
    with pytest.raises(WaveGlowError, match="workspace"):
        import ctypes
        rc = eng._lib.wg_infer(eng._h, 1, 1, 1.0, 0, 1, 1, 1, 0, 0, 0)
        eng._check(rc, "wg_infer")


def test_every_argument_error_is_reported_before_any_launch(lib_built):
    """The wg_status / message of each refusal (wg_b200.h), through the raw C ABI; none of them may launch a kernel."""
    from text_to_speech_b200.engine import WaveGlowEngine, WaveGlowError
    hp = WaveGlowHParams(n_channels=32)
    w = generate_weights(hp, 35)
    eng = _engine(hp, w)
    lib, h = eng._lib, eng._h
    mel, z = synthetic_inputs(8, 1, 4, hp)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    out = torch.zeros(1, 4 * 256, device="cuda")
    need = eng.workspace_bytes(1, 4)
    ws = torch.empty(need + 2048, dtype=torch.uint8, device="cuda")
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    _run(eng, mel, z, 0.6)
    launches = eng.last_launch_count

    def call(mel_p=mel_d.data_ptr(), z_p=z_d.data_ptr(), det=0, B=1, T=4, out_p=out.data_ptr(), ws_p=base, ws_n=need):
        rc = lib.wg_infer(h, mel_p, z_p, 0.6, det, B, T, out_p, ws_p, ws_n, 0)
        return rc, lib.wg_last_error(h).decode()

    assert call()[0] == 0
    for kwargs, code, text in [(dict(mel_p=0), -1, "NULL"), (dict(out_p=0), -1, "NULL"), (dict(z_p=0), -1, "z must be given"),
                               (dict(B=0), -1, "positive"), (dict(T=-3), -1, "positive"),
                               (dict(ws_n=need - 1), -4, "too small"), (dict(ws_p=base + 8), -4, "aligned"),
                               (dict(B=4096, T=4096), -1, "too large")]:
        rc, msg = call(**kwargs)
        assert rc == code and text in msg, (kwargs, rc, msg)
    assert call(z_p=0, det=1)[0] == 0                        # z may be NULL when deterministic
    torch.cuda.synchronize()
    assert eng.last_launch_count == launches                 # the failed calls launched nothing; the good ones the same count
    # create-time refusals
    bad = dict(w)
    bad["invertible_conv-0/conv/kernel"] = np.zeros_like(w["invertible_conv-0/conv/kernel"])
    with pytest.raises(WaveGlowError, match="singular"):
        WaveGlowEngine(hp, bad, mode="fp32", device=0)
    with pytest.raises(WaveGlowError, match="n_channels 256 and 512"):
        WaveGlowEngine(hp, w, mode="bf16", device=0)
    with pytest.raises(WaveGlowError, match="out of range"):
        WaveGlowEngine(hp, w, mode="fp32", device=99)

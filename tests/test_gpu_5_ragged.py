"""GPU parity of wg_infer_ragged (per-utterance lengths) and of the small-call CUDA-graph path.

The reference vocodes ONE trimmed mel at a time (models/tts/tacotron2.py:183-191 -> models/tts/waveglow.py:76-82), so
the parity statement for a batch is: utterance b of the batch == the reference's stand-alone call on that utterance.
WaveGlow is not causal (receptive field ~ 1 s), so a padded batch cannot give that for the tail of a short utterance;
the ragged entry point does -- bit for bit against the engine's own stand-alone run (same row layout), and within the
mode's tolerance against the CPU oracle run on each utterance alone."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import snr_db
from oracle.waveglow_oracle import OracleWaveGlow
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs, save_weights

pytestmark = pytest.mark.gpu
TOL_BF16_ABS, TOL_BF16_SNR, TOL_FP32 = 2e-2, 35.0, 1e-4


def _engine(hp, w, mode="bf16"):
    from text_to_speech_b200.engine import WaveGlowEngine
    return WaveGlowEngine(hp, w, mode=mode, device=0)


def _run(eng, mel, z, sigma, lengths=None, deterministic=False):
    mel_d = torch.from_numpy(np.ascontiguousarray(mel)).cuda()
    z_d = None if z is None else torch.from_numpy(np.ascontiguousarray(z)).cuda()
    out = eng.infer_device(mel_d, z_d, sigma=sigma, deterministic=deterministic, lengths=lengths)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _alone(eng, mel, z, b, T_b, sigma=0.6):
    return _run(eng, mel[b:b + 1, :T_b], z[b:b + 1, :T_b * 32], sigma)[0]


@pytest.mark.parametrize("C,lengths", [(256, [150, 33, 97, 150, 5, 128, 1]), (256, [40, 300]), (512, [70, 150, 9])])
def test_ragged_batch_equals_every_utterance_alone(lib_built, monkeypatch, C, lengths):
    monkeypatch.setenv("WG_PM", "1")     # one row layout for every call: bit-identity is defined within a layout
    hp = WaveGlowHParams(n_channels=C)
    w = generate_weights(hp, 1234, bias_std=0.05)
    B, T = len(lengths), max(lengths)
    mel, z = synthetic_inputs(31, B, T, hp)
    for b, n in enumerate(lengths):      # poison the padding: nothing beyond an utterance's own frames may be read
        mel[b, n:] = 1e4
        z[b, n * 32:] = -1e4
    eng = _engine(hp, w)
    out = _run(eng, mel, z, 0.6, lengths=lengths)
    assert out.shape == (B, T * 256) and np.isfinite(out).all()
    oracle = OracleWaveGlow(hp, w)
    for b, n in enumerate(lengths):
        assert np.array_equal(out[b, :n * 256], _alone(eng, mel, z, b, n)), f"utterance {b} (T={n}) differs from its stand-alone run"
        assert not out[b, n * 256:].any(), f"tail of utterance {b} is not zero"
        if C == 256 or n <= 70:          # the reference's own call on this utterance
            ref = oracle(mel[b:b + 1, :n], z[b:b + 1, :n * 32], 0.6).numpy()[0]
            err = np.abs(out[b, :n * 256] - ref).max()
            assert err <= TOL_BF16_ABS and (n < 4 or snr_db(ref, out[b, :n * 256]) >= TOL_BF16_SNR), (b, n, err)
    assert np.array_equal(_run(eng, mel, z, 0.6, lengths=lengths), out)      # reproducible
    eng.close()


def test_padded_batch_differs_from_the_reference_call_where_ragged_does_not(lib_built):
    """Why the entry point exists: the same short utterance inside a PADDED batch (pad value -11, as
    models/tts/waveglow.py:52-58 pads) leaves the tolerance near its end; the ragged call does not."""
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    n, T = 40, 120
    mel, z = synthetic_inputs(5, 2, T, hp)
    mel[1, n:] = -11.0
    ref = OracleWaveGlow(hp, w)(mel[1:2, :n], z[1:2, :n * 32], 0.6).numpy()[0]
    eng = _engine(hp, w)
    padded = _run(eng, mel, z, 0.6)[1, :n * 256]
    ragged = _run(eng, mel, z, 0.6, lengths=[T, n])[1, :n * 256]
    e_pad, e_rag = np.abs(padded - ref).max(), np.abs(ragged - ref).max()
    print(f"short utterance in a padded batch: max-abs {e_pad:.3e} vs the stand-alone reference; ragged: {e_rag:.3e}")
    assert e_rag <= TOL_BF16_ABS
    assert e_pad > e_rag
    eng.close()


def test_uniform_lengths_are_the_plain_call(lib_built):
    hp = WaveGlowHParams()
    eng = _engine(hp, generate_weights(hp, 1234))
    mel, z = synthetic_inputs(8, 3, 50, hp)
    assert np.array_equal(_run(eng, mel, z, 0.6, lengths=[50, 50, 50]), _run(eng, mel, z, 0.6))
    eng.close()


def test_ragged_fp32_mode_runs_utterance_by_utterance(lib_built):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234, bias_std=0.05)
    lengths = [24, 7, 16]
    mel, z = synthetic_inputs(12, 3, 24, hp)
    eng = _engine(hp, w, "fp32")
    out = _run(eng, mel, z, 0.6, lengths=lengths)
    oracle = OracleWaveGlow(hp, w)
    for b, n in enumerate(lengths):
        assert np.array_equal(out[b, :n * 256], _alone(eng, mel, z, b, n))
        assert not out[b, n * 256:].any()
        ref = oracle(mel[b:b + 1, :n], z[b:b + 1, :n * 32], 0.6).numpy()[0]
        assert np.abs(out[b, :n * 256] - ref).max() <= TOL_FP32
    eng.close()


def test_ragged_deterministic_and_host_entry_point(lib_built):
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    lengths = [30, 11]
    mel, z = synthetic_inputs(4, 2, 30, hp)
    eng = _engine(hp, w)
    dev = _run(eng, mel, z, 0.6, lengths=lengths)
    assert np.array_equal(eng.infer_host(mel, z, 0.6, lengths=lengths), dev)          # wg_infer_host_ragged
    det = _run(eng, mel, None, 0.6, lengths=lengths, deterministic=True)
    ref = OracleWaveGlow(hp, w)(mel[1:2, :11], None, 0.6, deterministic=True).numpy()[0]
    assert np.abs(det[1, :11 * 256] - ref).max() <= TOL_BF16_ABS
    eng.close()


def test_ragged_refusals_through_the_raw_c_abi(lib_built):
    hp = WaveGlowHParams()
    eng = _engine(hp, generate_weights(hp, 1234))
    lib, h = eng._lib, eng._h
    B, T = 2, 20
    mel, z = synthetic_inputs(4, B, T, hp)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    out = torch.zeros(B, T * 256, device="cuda")
    ws = torch.empty(eng.workspace_bytes(B, T) * 2 + 2048, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
    before = lib.wg_last_launch_count(h)

    def call(lens, ws_bytes=ws.numel() - 1024):
        arr = None if lens is None else (ctypes.c_int32 * B)(*lens)
        return lib.wg_infer_ragged(h, mel_d.data_ptr(), z_d.data_ptr(), 0.6, 0, B, T, arr, out.data_ptr(), ws_ptr, ws_bytes, 0)

    assert call(None) == -1 and b"T_b" in lib.wg_last_error(h)
    assert call([20, 0]) == -1 and b"outside" in lib.wg_last_error(h)
    assert call([21, 5]) == -1
    assert call([20, 5], ws_bytes=1024) == -4
    n = ctypes.c_size_t()
    assert lib.wg_workspace_bytes_ragged(h, B, T, (ctypes.c_int32 * B)(20, 5), ctypes.byref(n)) == 0 and n.value > 0
    assert lib.wg_workspace_bytes_ragged(h, B, T, (ctypes.c_int32 * B)(20, 99), ctypes.byref(n)) == -1
    assert lib.wg_last_launch_count(h) == before          # no refusal launched anything
    assert call([20, 5]) == 0
    torch.cuda.synchronize()
    assert np.isfinite(out.cpu().numpy()).all()
    eng.close()


def test_sharded_ragged_sweep_through_the_real_runtime(lib_built, tmp_path):
    """K5 in miniature through the product path: sharding.plan_batches(ragged) -> run_rank -> B200WaveGlowRuntime ->
    wg_infer_ragged, both ranks of a 2-rank plan run here one after the other; every utterance must equal the
    reference's stand-alone call (CPU oracle, deterministic noise) on it."""
    from text_to_speech_b200 import sharding
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    path = str(tmp_path / "wg.npz")
    save_weights(path, hp, w)
    rng = np.random.default_rng(3)
    lengths = [int(x) for x in rng.integers(6, 90, size=14)]
    mels = [synthetic_inputs(100 + i, 1, n, hp)[0][0] for i, n in enumerate(lengths)]
    plan = sharding.plan_batches(lengths, 2, max_frames=200, max_batch=4, ragged=True)
    rt = B200WaveGlowRuntime(path, mode="bf16", device=0)
    got = {}
    for rank in range(2):
        got.update(sharding.run_rank(rt, mels, plan[rank], ragged=True, sigma=0.6, deterministic=True))
    assert sorted(got) == list(range(len(lengths)))
    oracle = OracleWaveGlow(hp, w)
    for i, n in enumerate(lengths):
        ref = oracle(mels[i][None], None, 0.6, deterministic=True).numpy()[0]
        assert got[i].shape == (n * 256,)
        assert np.abs(got[i] - ref).max() <= TOL_BF16_ABS, i


def test_runtime_graph_replay_equals_eager(lib_built, tmp_path):
    """Small calls replay a CUDA graph captured per call signature: same bits as the eager launch sequence, host and
    device inputs, ragged lengths, LRU eviction, precompile()."""
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    path = str(tmp_path / "wg.npz")
    save_weights(path, hp, w)
    eager = B200WaveGlowRuntime(path, mode="bf16", device=0, graph_max_frames=0)
    graph = B200WaveGlowRuntime(path, engine=eager.engine, mode="bf16", device=0, graph_max_frames=512, max_graphs=2)
    mel, z = synthetic_inputs(21, 1, 64, hp)
    a = eager(mel, z=z, sigma=0.6).copy()
    for _ in range(3):                                   # capture, then replays
        assert np.array_equal(graph(mel, z=z, sigma=0.6), a)
    assert graph.graph_replays == 3 and eager.graph_replays == 0 and len(graph._graphs) == 1
    mel2, z2 = synthetic_inputs(22, 1, 64, hp)           # same signature, new data: the replay reads the new inputs
    assert np.array_equal(graph(mel2, z=z2, sigma=0.6), eager(mel2, z=z2, sigma=0.6))
    d = graph(torch.from_numpy(mel).cuda(), z=torch.from_numpy(z).cuda(), sigma=0.6)
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), a)
    mel3, z3 = synthetic_inputs(23, 3, 40, hp)
    lens = [40, 13, 25]
    assert np.array_equal(graph(mel3, z=z3, sigma=0.6, lengths=lens), eager(mel3, z=z3, sigma=0.6, lengths=lens))
    assert np.array_equal(graph(mel3, z=z3, sigma=0.6, lengths=lens), eager(mel3, z=z3, sigma=0.6, lengths=lens))
    graph(mel3, z=z3, sigma=1.0)                         # third signature: the oldest graph is evicted
    assert len(graph._graphs) == 2
    big_mel, big_z = synthetic_inputs(24, 2, 300, hp)    # above graph_max_frames: eager launches
    n = graph.graph_replays
    assert np.array_equal(graph(big_mel, z=big_z, sigma=0.6), eager(big_mel, z=big_z, sigma=0.6))
    assert graph.graph_replays == n
    pre = B200WaveGlowRuntime(path, engine=eager.engine, mode="bf16", device=0, graph_max_frames=256)
    assert pre.precompile() == [(1, 64), (1, 128), (1, 192), (1, 256)]
    assert len(pre._graphs) == 4
    assert np.array_equal(pre(mel, z=z, sigma=1.0), eager(mel, z=z, sigma=1.0)) and pre.graph_replays == 1


def test_output_ring_views_and_copies(lib_built, tmp_path):
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    hp = WaveGlowHParams()
    path = str(tmp_path / "wg.npz")
    save_weights(path, hp, generate_weights(hp, 1234))
    rt = B200WaveGlowRuntime(path, mode="bf16", device=0, output_ring=2)
    mel, z = synthetic_inputs(1, 1, 16, hp)
    a = rt(mel, z=z, sigma=0.6)
    keep = a.copy()
    b = rt(mel, z=-z, sigma=0.6)
    assert np.array_equal(a, keep) and not np.array_equal(a, b)       # one further call leaves the previous view intact
    rc = B200WaveGlowRuntime(path, engine=rt.engine, mode="bf16", device=0, copy_outputs=True)
    c = rc(mel, z=z, sigma=0.6)
    for _ in range(3):
        rc(mel, z=-z, sigma=0.6)
    assert np.array_equal(c, keep)


def test_pipelined_sweep_equals_the_plain_one(lib_built, tmp_path):
    """sharding.run_rank_pipelined (pinned rings, copy stream, preallocated device buffers, runtime(out=...)) must return
    exactly what the plain batch-by-batch loop returns, sweep after sweep."""
    from text_to_speech_b200 import sharding
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    hp = WaveGlowHParams()
    path = str(tmp_path / "wg.npz")
    save_weights(path, hp, generate_weights(hp, 1234))
    rng = np.random.default_rng(9)
    lengths = [int(x) for x in rng.integers(4, 60, size=11)]
    mels = [synthetic_inputs(300 + i, 1, n, hp)[0][0] for i, n in enumerate(lengths)]
    plan = sharding.plan_batches(lengths, 1, max_frames=120, max_batch=4, ragged=True)[0]
    rt = B200WaveGlowRuntime(path, mode="bf16", device=0)
    want = sharding.run_rank(rt, mels, plan, ragged=True, sigma=0.6, deterministic=True)
    for _ in range(2):
        got, stats = sharding.run_rank_pipelined(rt, mels, plan, sigma=0.6, deterministic=True)
        assert sorted(got) == sorted(want) and all(np.array_equal(got[i], want[i]) for i in want)
        assert stats["h2d_bytes"] > 0 and stats["d2h_bytes"] == sum(len(b.indices) * b.T * 256 * 4 for b in plan)
    d = torch.empty(2, 8 * 256, device="cuda")
    mel, z = synthetic_inputs(1, 2, 8, hp)
    r = rt(torch.from_numpy(mel).cuda(), z=torch.from_numpy(z).cuda(), sigma=0.6, out=d)      # graph path + out=
    assert r.data_ptr() == d.data_ptr() and np.array_equal(d.cpu().numpy(), rt(mel, z=z, sigma=0.6))


@pytest.mark.parametrize("C,mode,pair,flow", [(256, "bf16", "0", "1"), (256, "bf16", "1", "1"), (256, "tf32x3", "0", "1"),
                                              (256, "tf32x3", "1", "0"), (256, "tf32x3", "1", "1"), (512, "bf16", "0", "1"),
                                              (256, "fp32", "0", "1")])
def test_no_write_outside_the_workspace_or_the_output(lib_built, monkeypatch, C, mode, pair, flow):
    """Canary check through the raw C ABI (compute-sanitizer is not available on this pool): the scratch the engine asks
    for and the caller's waveform buffer sit between guard bands; after uniform and ragged infers on every kernel path
    (single-CTA, CTA pair incl. a ghost tile, tf32x3 per-layer kernels and the one-launch-per-flow kernel, WaveGlow-512,
    FFMA) the bands must be untouched."""
    monkeypatch.setenv("WG_PM", "1")
    monkeypatch.setenv("WG_PAIR", pair)
    monkeypatch.setenv("WG_TF32_FLOW", flow)
    hp = WaveGlowHParams(n_channels=C, n_flows=4 if C == 512 else 12)
    eng = _engine(hp, generate_weights(hp, 7), mode)
    lib, h = eng._lib, eng._h
    B, T, lens = 3, 40, [40, 7, 33]
    mel, z = synthetic_inputs(3, B, T, hp)
    mel_d, z_d = torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda()
    arr = (ctypes.c_int32 * B)(*lens)
    n1, n2 = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.wg_workspace_bytes(h, B, T, ctypes.byref(n1)) == 0
    assert lib.wg_workspace_bytes_ragged(h, B, T, arr, ctypes.byref(n2)) == 0
    G = 1 << 20
    for ragged, need in ((False, n1.value), (True, n2.value)):
        need_al = (need + 1023) // 1024 * 1024
        ws = torch.full((G + need_al + G + 1024,), 0xA5, dtype=torch.uint8, device="cuda")
        off = (-ws.data_ptr()) % 1024 + G
        out = torch.full((G // 4 + B * T * 256 + G // 4,), 7.5, dtype=torch.float32, device="cuda")
        out_ptr = out.data_ptr() + G
        if ragged:
            rc = lib.wg_infer_ragged(h, mel_d.data_ptr(), z_d.data_ptr(), 0.6, 0, B, T, arr, out_ptr, ws.data_ptr() + off, need, 0)
        else:
            rc = lib.wg_infer(h, mel_d.data_ptr(), z_d.data_ptr(), 0.6, 0, B, T, out_ptr, ws.data_ptr() + off, need, 0)
        assert rc == 0, lib.wg_last_error(h)
        torch.cuda.synchronize()
        assert bool((ws[:off] == 0xA5).all()) and bool((ws[off + need_al:] == 0xA5).all()), "write outside the workspace"
        assert bool((out[:G // 4] == 7.5).all()) and bool((out[G // 4 + B * T * 256:] == 7.5).all()), "write outside the waveform"
        body = out[G // 4:G // 4 + B * T * 256]
        assert bool(torch.isfinite(body).all()) and float(body.abs().max()) > 0
    eng.close()


def test_ragged_batch_of_hundreds_of_short_utterances(lib_built):
    """More than 256 utterances (the geometry tables travel in kernel parameters, 256 per launch) of 1-6 frames each:
    spot-checked against stand-alone runs, every tail zero."""
    hp = WaveGlowHParams()
    eng = _engine(hp, generate_weights(hp, 1234))
    rng = np.random.default_rng(4)
    B, T = 300, 6
    lengths = [int(x) for x in rng.integers(1, T + 1, size=B)]
    mel, z = synthetic_inputs(77, B, T, hp)
    out = _run(eng, mel, z, 0.6, lengths=lengths)
    assert np.isfinite(out).all() and eng.last_launch_count == 122 + 2
    for b in (0, 1, 255, 256, 257, 299):
        n = lengths[b]
        assert np.array_equal(out[b, :n * 256], _alone(eng, mel, z, b, n)) and not out[b, n * 256:].any()
    eng.close()


def test_engine_from_keras_weights_h5_equals_engine_from_npz(lib_built, tmp_path):
    """`WaveGlow(path='<name>.weights.h5')`: the reference's checkpoint format (checkpoint_manager.py:169-216) through
    h5lite / convert.from_keras_weights_h5 gives the same engine as the .npz of the same weights, bit for bit."""
    from oracle.h5_writer import write_h5, keras3_waveglow_layout
    from text_to_speech_b200.waveglow import WaveGlow
    hp = WaveGlowHParams()
    w = generate_weights(hp, 1234)
    npz, h5 = str(tmp_path / "wg.npz"), str(tmp_path / "waveglow.weights.h5")
    save_weights(npz, hp, w)
    write_h5(h5, keras3_waveglow_layout(hp, w))
    mel, z = synthetic_inputs(6, 2, 20, hp)
    a = np.array(WaveGlow(path=npz, runtime="b200", mode="bf16")(mel, z=z, sigma=0.6))
    b = np.array(WaveGlow(path=h5, runtime="b200", mode="bf16")(mel, z=z, sigma=0.6))
    assert a.shape == (2, 20 * 256) and np.array_equal(a, b)

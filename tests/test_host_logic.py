"""CPU: host-side mirror of the reference interface -- Runtime registry, wrapper windowing/stitching
(models/tts/waveglow.py:61-164), weight files, utterance sharding."""
import math

import os

import numpy as np
import pytest

from text_to_speech_b200 import sharding
from text_to_speech_b200.runtime import Runtime, build_runtime, _runtimes
from text_to_speech_b200.waveglow import WaveGlow, _get_steps
from text_to_speech_b200.weights import (WaveGlowHParams, generate_weights, load_weights, save_weights,
                                         weight_names, weights_digest)


def test_unknown_runtime_raises_like_the_reference():
    # utils/keras/runtimes/__init__.py:24-27
    with pytest.raises(ValueError, match="Unsupported runtime"):
        build_runtime("nope", "x.npz")
    assert "b200" in _runtimes and issubclass(_runtimes["b200"], Runtime)


def test_runtime_engine_cache_contract():
    # runtime.py:20-29: engines cached per path, `reload` forces a new load, `engine=` bypasses both
    loads = []

    class Dummy(Runtime):
        _engines = {}

        def __call__(self, x):
            return self.engine

        @staticmethod
        def load_engine(path, **kw):
            loads.append(path)
            return object()

    a, b = Dummy("p"), Dummy("p")
    assert a.engine is b.engine and loads == ["p"]
    c = Dummy("p", reload=True)
    assert c.engine is not a.engine and loads == ["p", "p"]
    d = Dummy("q", engine="E")
    assert d.engine == "E" and loads == ["p", "p"]
    assert repr(a) == "<Dummy path=p>"


@pytest.mark.parametrize("length,win,hop", [(100, 50, 40), (1000, 256, 192), (257, 256, 192), (300, 300, 100)])
def test_window_starts_cover_the_mel(length, win, hop):
    a = _get_steps(length, win, hop)
    assert a[0] == 0 and (len(a) == 1 or a[-1] == length - win)
    assert all(0 < a[i + 1] - a[i] <= hop for i in range(len(a) - 1))          # evenly spread, never further apart than `hop`


class _FakeRuntime:
    """Deterministic stand-in vocoder: sample s of frame t = mel[t, 0] * 1000 + position in window."""
    def __call__(self, mel, **kw):
        mel = np.asarray(mel)
        B, T, _ = mel.shape
        base = np.repeat(mel[:, :, 0], 256, axis=1) * 1000.0
        return base + np.arange(T * 256)[None] * 1e-3


def _wrapper_with_fake():
    w = WaveGlow.__new__(WaveGlow)
    w.runtime, w.pad_mel_value, w.model = "b200", -11.0, _FakeRuntime()
    return w


def _wrapper_golden():
    import hashlib
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "wrapper_cases.npz"))
    cases = []
    for n in range(int(f["n_cases"])):
        seed, B, T, kw = eval(bytes(f[f"case{n}_meta"]).decode())
        if f"case{n}_error" in f.files:
            cases.append((seed, B, T, kw, None, bytes(f[f"case{n}_error"]).decode()))
        else:
            cases.append((seed, B, T, kw, dict(shape=tuple(f[f"case{n}_shape"]), dtype=bytes(f[f"case{n}_dtype"]).decode(),
                                               sha256=bytes(f[f"case{n}_sha256"]).decode(), probe=f[f"case{n}_probe"]), None))
    steps = [(tuple(int(x) for x in f[f"steps{j}_args"]), [int(x) for x in f[f"steps{j}"]]) for j in range(int(f["n_steps"]))]
    return cases, steps, hashlib


def test_wrapper_matches_the_golden_cases_of_the_reference_source():
    """tests/golden/wrapper_cases.npz = outputs of the reference's OWN models/tts/waveglow.py (oracle/gen_golden_wrapper.py)
    over the stand-in vocoder: every windowing / padding / stitching / batch branch, bit for bit (sha256 of the bytes), and
    the exceptions the reference raises (its float-slice quirk with use_slice=True) raised here too."""
    cases, steps, hashlib = _wrapper_golden()
    assert len(cases) >= 30
    for seed, B, T, kw, want, err in cases:
        mel = np.random.default_rng(seed).normal(size=(B, T, 80)).astype(np.float32)
        if err is not None:
            with pytest.raises(Exception) as ei:
                _wrapper_with_fake()(mel, **kw)
            assert type(ei.value).__name__ == err, (T, kw)
            continue
        got = np.ascontiguousarray(_wrapper_with_fake()(mel, **kw))
        assert got.shape == want["shape"] and str(got.dtype) == want["dtype"], (T, kw)
        assert np.array_equal(got.reshape(-1)[::997], want["probe"]), (T, kw)
        assert hashlib.sha256(got.tobytes()).hexdigest() == want["sha256"], (T, kw)
    for args, want in steps:
        assert [int(x) for x in _get_steps(*args)] == want, args


def test_wrapper_accepts_2d_and_batches():
    rng = np.random.default_rng(1)
    w = _wrapper_with_fake()
    mel2 = rng.normal(size=(50, 80)).astype(np.float32)
    assert w(mel2).shape == (1, 50 * 256)
    melb = rng.normal(size=(3, 300, 80)).astype(np.float32)
    assert w(melb, win_len=128).shape == (3, 300 * 256)      # batch > 1: direct inference (waveglow.py:108-112)


def test_weight_file_roundtrip(tmp_path):
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=2, n_channels=16)
    w = generate_weights(hp, 3, bias_std=0.1)
    p = tmp_path / "w.npz"
    save_weights(p, hp, w)
    hp2, w2 = load_weights(p)
    assert hp2 == hp and weights_digest(w2) == weights_digest(w)
    assert [n for n, _ in weight_names(hp)] == sorted(w, key=[n for n, _ in weight_names(hp)].index)


def test_flow_schedule_matches_reference_topology():
    # waveglow_arch.py:202-223 with NVIDIA hparams: flows 0-3 (4/8), 4-7 (3/6), 8-11 (2/4)
    fc = WaveGlowHParams().flow_channels()
    assert fc[:4] == [(4, 8)] * 4 and fc[4:8] == [(3, 6)] * 4 and fc[8:] == [(2, 4)] * 4
    assert WaveGlowHParams().n_remaining_channels == 4


def test_lpt_sharding_covers_everything_once_and_balances():
    rng = np.random.default_rng(5)
    lengths = rng.integers(172, 1724, size=1024)
    for ws in (1, 2, 4, 8):
        shards = sharding.assign_utterances(lengths, ws)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(1024))
        loads = [int(lengths[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= lengths.max()
    batches = sharding.make_batches(shards[0], lengths, max_frames=16 * 860)
    assert sorted(i for b in batches for i in b.indices) == sorted(shards[0])
    assert all(len(b.indices) * b.T <= 16 * 860 or len(b.indices) == 1 for b in batches)
    assert all(b.T == max(lengths[i] for i in b.indices) for b in batches)
    assert sharding.padding_waste(batches, lengths) < 0.12


def test_global_batch_plan_is_balanced_and_tight():
    rng = np.random.default_rng(5)
    lengths = rng.integers(172, 1724, size=1024)
    for ws in (1, 2, 4, 8):
        plan = sharding.plan_batches(lengths, ws, max_frames=16 * 860)
        flat = sorted(i for r in plan for b in r for i in b.indices)
        assert flat == list(range(1024))
        loads = [sum(b.T * len(b.indices) for b in r) for r in plan]
        assert (max(loads) - min(loads)) / max(loads) < 0.03
        assert sharding.padding_waste([b for r in plan for b in r], lengths) < 0.02
        assert plan == sharding.plan_batches(lengths, ws, max_frames=16 * 860)   # deterministic


def test_write_audio_matches_reference_normalisation(tmp_path):
    """audio_processing.py:51-62 / audio_io.py:346-369: mean removed, peak scaled to 32767, int16 truncation, PCM wav."""
    from scipy.io import wavfile
    from text_to_speech_b200.audio_io import normalize_audio, write_audio
    rng = np.random.default_rng(0)
    wave = (rng.standard_normal(5000) * 0.3 + 0.05).astype(np.float32)
    path = write_audio(str(tmp_path / "a.wav"), wave, 22050)
    rate, pcm = wavfile.read(path)
    assert rate == 22050 and pcm.dtype == np.int16 and pcm.shape == wave.shape
    centred = wave - np.mean(wave)
    want = (centred * (32767 / np.max(np.abs(centred)))).astype(np.int16)
    assert np.array_equal(pcm, want) and np.abs(pcm).max() == 32767
    f = normalize_audio(wave, max_val=1.0)
    assert f.dtype == np.float32 and abs(np.abs(f).max() - 1.0) < 1e-6
    assert np.array_equal(normalize_audio(np.zeros(10, np.float32)), np.zeros(10, np.int16))    # silent input: no division
    raw = write_audio(str(tmp_path / "b.wav"), wave, 22050, normalize=False)
    assert wavfile.read(raw)[1].dtype == np.float32
    with pytest.raises(ValueError, match="Unsupported file extension"):
        write_audio(str(tmp_path / "c.mp3"), wave, 22050)


REF_RUNTIME = "/root/reference/utils/keras/runtimes/runtime.py"


@pytest.mark.skipif(not __import__("os").path.isfile(REF_RUNTIME), reason="reference tree not mounted")
def test_runtime_mirror_is_interchangeable_with_the_reference_abc():
    """Loads the reference's REAL Runtime ABC (a stand-alone file) and checks (a) the mirror exposes the same public
    surface with the same signatures and (b) B200WaveGlowRuntime's methods satisfy the real ABC, i.e. the class
    INTEGRATION.md registers under `_runtimes['b200']` can inherit from the reference's base unchanged."""
    import importlib.util
    import inspect
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    spec = importlib.util.spec_from_file_location("_ref_runtime", REF_RUNTIME)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Ref = mod.Runtime
    public = [n for n in vars(Ref) if not n.startswith("_") or n in ("__init__", "__call__", "__repr__")]
    for name in public:
        assert hasattr(Runtime, name), f"mirror lacks {name}"
        if callable(getattr(Ref, name)):
            assert str(inspect.signature(getattr(Ref, name))) == str(inspect.signature(getattr(Runtime, name))), name
    assert Ref.__abstractmethods__ == Runtime.__abstractmethods__ == frozenset({"__call__", "load_engine"})
    for name in ("build_from", "from_tensorflow", "from_torch", "from_onnx"):
        assert isinstance(inspect.getattr_static(Ref, name), classmethod) and isinstance(inspect.getattr_static(Runtime, name), classmethod)

    class OnRealBase(Ref):                       # what a maintainer would write inside the reference tree
        __call__ = B200WaveGlowRuntime.__call__
        load_engine = staticmethod(B200WaveGlowRuntime.load_engine)

    fake_engine = object()
    rt = OnRealBase("weights.npz", engine=fake_engine)
    assert rt.engine is fake_engine and rt.path == "weights.npz" and repr(rt) == "<OnRealBase path=weights.npz>"
    with pytest.raises(NotImplementedError, match="cannot be initialized from `ONNX`"):
        OnRealBase.from_onnx("m.onnx", "weights.npz")
    with pytest.raises(NotImplementedError, match="cannot be initialized from `ONNX`"):
        B200WaveGlowRuntime.from_onnx("m.onnx", "weights.npz")


def _ref_wrapper_available():
    from oracle.run_reference_wrapper import wrapper_reference_available
    return wrapper_reference_available()


@pytest.mark.skipif(not _ref_wrapper_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("T", [37, 200, 333])
@pytest.mark.parametrize("kw", [dict(), dict(win_len=128), dict(win_len=64, hop_len=-16), dict(win_len=128, batch=True),
                                dict(win_len=0.5), dict(win_len=3.0, use_slice=True), dict(win_len=512),
                                dict(win_len=512, force_pad=True), dict(win_len=100, hop_len=0.5, max_win_len=80),
                                dict(win_len=96, hop_len=-32, batch=True)])
def test_wrapper_matches_the_reference_wrapper_source(T, kw):
    """models/tts/waveglow.py executed unmodified (oracle/run_reference_wrapper.py) against our wrapper, with the same
    stand-in vocoder behind both: every windowing / padding / stitching branch must give the same samples."""
    from oracle.run_reference_wrapper import reference_get_steps, reference_wrapper_infer
    rng = np.random.default_rng(T)
    mel = rng.normal(size=(1, T, 80)).astype(np.float32)
    try:
        want = reference_wrapper_infer(_FakeRuntime(), mel, **kw)
    except Exception as e:      # the reference's own quirks (float slice bounds with use_slice=True) are part of the contract
        with pytest.raises(type(e)):
            _wrapper_with_fake()(mel, **kw)
        return
    got = np.asarray(_wrapper_with_fake()(mel, **kw))
    assert got.shape == want.shape and np.array_equal(got, want)
    for args in ((T, 64, 48), (T, 128, 64), (1000, 256, 192)):
        if args[0] > args[1]:
            assert list(reference_get_steps(*args)) == list(_get_steps(*args))


@pytest.mark.skipif(not _ref_wrapper_available(), reason="reference tree not mounted")
def test_wrapper_2d_input_and_batch_branch_match_the_reference_source():
    from oracle.run_reference_wrapper import reference_wrapper_infer
    rng = np.random.default_rng(2)
    mel2 = rng.normal(size=(50, 80)).astype(np.float32)
    assert np.array_equal(np.asarray(_wrapper_with_fake()(mel2)), reference_wrapper_infer(_FakeRuntime(), mel2))
    melb = rng.normal(size=(3, 300, 80)).astype(np.float32)
    assert np.array_equal(np.asarray(_wrapper_with_fake()(melb, win_len=128)),
                          reference_wrapper_infer(_FakeRuntime(), melb, win_len=128))


REF_AUDIO_PROC = "/root/reference/utils/audio/audio_processing.py"


@pytest.mark.skipif(not __import__("os").path.isfile(REF_AUDIO_PROC), reason="reference tree not mounted")
def test_normalize_audio_equals_the_reference_function():
    """utils/audio/audio_processing.py::normalize_audio executed from the reference file (its unrelated imports --
    librosa.util, loggers.timer, the dispatch wrapper -- replaced by inert stand-ins) against audio_io.normalize_audio."""
    import importlib.util
    import sys
    import types
    from text_to_speech_b200.audio_io import normalize_audio
    saved = {k: sys.modules.get(k) for k in ("librosa", "librosa.util", "loggers", "_ref_audio", "_ref_audio.wrappers")}
    try:
        librosa = types.ModuleType("librosa")
        librosa.util = types.ModuleType("librosa.util")
        loggers = types.ModuleType("loggers")
        loggers.timer = lambda fn=None, **kw: fn if callable(fn) else (lambda f: f)
        pkg = types.ModuleType("_ref_audio")
        pkg.__path__ = []
        wrappers = types.ModuleType("_ref_audio.wrappers")

        def dispatch_wrapper(*a, **k):
            def deco(fn):
                fn.dispatch = lambda *aa, **kk: (aa[0] if aa and callable(aa[0]) else (lambda f: f))
                return fn
            return deco
        wrappers.dispatch_wrapper = dispatch_wrapper
        sub = types.ModuleType("_ref_audio.audio")
        sub.__path__ = []
        sys.modules.update({"librosa": librosa, "librosa.util": librosa.util, "loggers": loggers, "_ref_audio": pkg,
                            "_ref_audio.wrappers": wrappers, "_ref_audio.audio": sub})
        spec = importlib.util.spec_from_file_location("_ref_audio.audio.audio_processing", REF_AUDIO_PROC)
        ref = importlib.util.module_from_spec(spec)
        sys.modules["_ref_audio.audio.audio_processing"] = ref
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(4)
    for wave in (rng.standard_normal(4000).astype(np.float32) * 0.2 + 0.03, np.zeros(100, np.float32),
                 (rng.standard_normal(777) * 3).astype(np.float32), rng.standard_normal(50)):
        for kw in (dict(), dict(max_val=1.0), dict(max_val=20000)):
            a, b = normalize_audio(wave, **kw), ref.normalize_audio(wave, **kw)
            assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("ragged", [False, True])
def test_plan_batches_honours_its_hard_limits(ragged):
    """many short utterances used to collapse into one giant batch once max_batch forced early cuts."""
    from text_to_speech_b200 import sharding
    cases = [([10] * 10000, 1, 20000, 64), ([100] * 1000, 8, 4000, 16), ([5, 900, 7, 33, 860, 12] * 40, 4, 16 * 860, 64),
             ([3000, 10, 10], 2, 1000, 8)]
    for lengths, ws, max_frames, max_batch in cases:
        plan = sharding.plan_batches(lengths, ws, max_frames, max_batch, ragged=ragged)
        assert len(plan) == ws
        batches = [b for r in plan for b in r]
        assert sorted(i for b in batches for i in b.indices) == list(range(len(lengths)))
        for b in batches:
            assert len(b.indices) <= max_batch
            assert b.T == max(lengths[i] for i in b.indices)
            frames = sum(lengths[i] for i in b.indices) if ragged else b.T * len(b.indices)
            assert frames <= max_frames or len(b.indices) == 1, (len(b.indices), b.T, frames)
        loads = [sum((sum(lengths[i] for i in b.indices) if ragged else b.T * len(b.indices)) for b in r) for r in plan]
        if len(batches) >= 4 * ws:
            assert max(loads) <= 1.15 * (sum(loads) / ws), loads


def test_run_rank_passes_true_lengths_when_ragged():
    from text_to_speech_b200 import sharding
    seen = []

    def vocoder(x, lengths=None, **kw):
        seen.append((x.shape, lengths, kw))
        return np.repeat(x[:, :, 0], 256, axis=1)

    mels = [np.full((n, 80), float(i), np.float32) for i, n in enumerate([5, 9, 2])]
    plan = sharding.plan_batches([5, 9, 2], 1, max_frames=100, ragged=True)[0]
    out = sharding.run_rank(vocoder, mels, plan, ragged=True, sigma=0.6)
    assert all(l is not None and kw == {"sigma": 0.6} for _, l, kw in seen)
    assert sorted(x for _, l, _ in seen for x in l) == [2, 5, 9]
    for i, n in enumerate([5, 9, 2]):
        assert out[i].shape == (n * 256,) and (out[i] == float(i)).all()


def test_audio_and_json_savers_write_the_reference_layout(tmp_path):
    """AudioSaver / JSONSaver (utils/callbacks/file_saver.py:100-125, example_outputs/en/map.json): audios/audio-<n>.wav
    and a map.json whose entries hold everything but the arrays, with the audio replaced by its file path."""
    import json
    from scipy.io import wavfile
    from text_to_speech_b200.audio_io import AudioSaver, JSONSaver, normalize_audio
    d = str(tmp_path / "out")
    savers = [AudioSaver(d), JSONSaver(d)]
    rng = np.random.default_rng(0)
    for n, text in enumerate(["hello world", "second sentence"]):
        audio = rng.standard_normal(2205).astype(np.float32) * 0.1
        output = {"text": text, "mel": np.zeros((3, 80), np.float32), "audio": audio, "rate": 22050, "time": 0.1}
        infos = {k: v for k, v in output.items() if k not in ("mel", "audio")}
        for s in savers:
            s.apply(infos, output)
        rate, data = wavfile.read(infos["audio"])
        assert rate == 22050 and np.array_equal(data, normalize_audio(audio))
        assert infos["audio"].endswith(f"audios/audio-{n}.wav")
    m = json.load(open(tmp_path / "out" / "map.json"))
    assert list(m) == ["hello world", "second sentence"]
    assert set(m["hello world"]) == {"text", "rate", "time", "audio"} and m["second sentence"]["audio"].endswith("audio-1.wav")
    again = JSONSaver(d)                      # an existing map is extended, not overwritten
    again.apply({"text": "third", "rate": 22050, "time": 0.2, "audio": "x.wav"})
    assert list(json.load(open(tmp_path / "out" / "map.json"))) == ["hello world", "second sentence", "third"]

"""GPU: the tts() pipeline (Tacotron2 producer -> B200 WaveGlow runtime) and the CUDA-graph decode loop.
Parity of the decoder kernels: against the reference-source fixture (last test) and the torch restatement; the
encoder/postnet are torch library code with unpinned parity (see text_to_speech_b200/tacotron2.py)."""
import os

import numpy as np
import pytest
import torch

from text_to_speech_b200.weights import HOP, WaveGlowHParams, generate_weights, save_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(lib_built, tmp_path_factory):
    from text_to_speech_b200.runtime import B200WaveGlowRuntime
    from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights
    hp = WaveGlowHParams()
    path = str(tmp_path_factory.mktemp("w") / "wg256.npz")
    save_weights(path, hp, generate_weights(hp, 1234))
    thp = Tacotron2HParams()
    tw = generate_tacotron2_weights(thp, 5)
    tw["decoder/gate_output/bias"][:] = -10.0
    return Tacotron2(thp, tw, device="cuda"), B200WaveGlowRuntime(path, mode="bf16", device=0, seed=3), tw, thp


def test_graph_decode_equals_eager(models):
    from text_to_speech_b200.tts import synthetic_texts
    taco = models[0]
    toks = np.stack(synthetic_texts(3, 1, 20, 20))
    a = taco.infer(toks, max_length=70, early_stopping=False, deterministic=True)
    b = taco.infer(toks, max_length=70, early_stopping=False, deterministic=True, use_graph=True, graph_chunk=32)
    assert torch.equal(a.lengths, b.lengths) and a.lengths.tolist() == [70, 70, 70]
    assert torch.allclose(a.mel, b.mel, atol=1e-5) and torch.allclose(a.attention_weights, b.attention_weights, atol=1e-6)
    c = taco.infer(toks, max_length=70, early_stopping=False, deterministic=True, use_graph=True, graph_chunk=32)
    assert torch.equal(b.mel, c.mel)                      # replaying the cached graph is deterministic


def test_gpu_producer_matches_cpu_float64(models):
    from text_to_speech_b200.tacotron2 import Tacotron2
    from text_to_speech_b200.tts import synthetic_texts
    taco, _, tw, thp = models
    toks = np.stack(synthetic_texts(2, 2, 15, 15))
    ref = Tacotron2(thp, tw, device="cpu", dtype=torch.float64).infer(toks, max_length=25, early_stopping=False, deterministic=True)
    got = taco.infer(toks, max_length=25, early_stopping=False, deterministic=True)
    assert torch.allclose(got.mel.cpu().double(), ref.mel, atol=1e-4)     # true fp32 (cuDNN TF32 off) vs float64, 25 frames


def test_pipeline_returns_the_vocoder_output_for_the_producer_mels(models):
    from text_to_speech_b200.tts import synthetic_texts, tts
    taco, voc, _, _ = models
    texts = synthetic_texts(5, 7, 8, 16)
    tm = {}
    out = tts(texts, taco, voc, max_length=4.0, batch_size=4, early_stopping=False, deterministic=True,
              use_graph=False, timings=tm, vocoder_max_frames=150)
    assert sorted(out) == list(range(5)) and tm["utterances"] == 5
    for i, o in out.items():
        T = o["mel"].shape[0]
        assert T > 0 and o["audio"].shape == (T * HOP,) and o["rate"] == 22050 and np.isfinite(o["audio"]).all()
        assert abs(o["time"] - T * HOP / 22050) < 1e-9
    # deterministic vocoder on the same mel must give the same waveform as the pipeline's noise-free run
    i = 2
    mel = torch.from_numpy(out[i]["mel"]).cuda()[None]
    again = tts(texts, taco, voc, max_length=4.0, batch_size=4, early_stopping=False, deterministic=True,
                use_graph=False, vocoder_max_frames=150)
    assert np.array_equal(again[i]["mel"], out[i]["mel"])
    z = torch.zeros(1, mel.shape[1] * 32, 8, device="cuda")
    w0 = voc(mel, z=z, sigma=0.6).cpu().numpy()[0]
    assert w0.shape == out[i]["audio"].shape


def test_sharding_covers_every_text_once(models):
    from text_to_speech_b200.tts import synthetic_texts, tts
    taco, voc, _, _ = models
    texts = synthetic_texts(6, 9, 6, 12)
    parts = [tts(texts, taco, voc, max_length=3.0, batch_size=4, early_stopping=False, deterministic=True,
                 use_graph=False, rank=r, world_size=2) for r in range(2)]
    assert sorted(list(parts[0]) + list(parts[1])) == list(range(6)) and not set(parts[0]) & set(parts[1])


def test_finished_at_first_frame_gives_silence(models):
    from text_to_speech_b200.tacotron2 import Tacotron2
    from text_to_speech_b200.tts import synthetic_texts, tts
    _, voc, tw, thp = models
    tw = dict(tw)
    tw["decoder/gate_output/kernel"] = np.zeros_like(tw["decoder/gate_output/kernel"])
    tw["decoder/gate_output/bias"] = np.full(1, 10.0, np.float32)
    out = tts(synthetic_texts(2, 3, 6, 6), Tacotron2(thp, tw, device="cuda"), voc, max_length=3.0, use_graph=False)
    assert all(o["mel"].shape[0] == 0 and len(o["audio"]) == int(0.15 * 22050) and not o["audio"].any() for o in out.values())


# ---- the B200 decoder loop (csrc/taco.cu through wg_taco_decode) against the torch restatement --------------------

def _ragged_tokens(n, seed, lo, hi):
    from text_to_speech_b200.tts import synthetic_texts
    texts = synthetic_texts(n, seed, lo, hi)
    S = max(len(t) for t in texts)
    toks = np.zeros((n, S), np.int64)
    for j, t in enumerate(texts):
        toks[j, :len(t)] = t
    return toks


@pytest.mark.parametrize("B,lo,hi,T", [(3, 12, 20, 70), (17, 5, 9, 23), (1, 1, 1, 9), (16, 40, 40, 64), (33, 150, 300, 10)])
def test_b200_decoder_matches_torch_restatement(models, B, lo, hi, T):
    taco = models[0]
    toks = _ragged_tokens(B, 100 + B, lo, hi)
    ref = taco.infer(toks, max_length=T, early_stopping=False, deterministic=True)
    got = taco.infer(toks, max_length=T, early_stopping=False, deterministic=True, decoder="b200")
    assert got.lengths.tolist() == ref.lengths.tolist() == [T] * B
    # fp32 recurrences with different summation orders drift apart slowly: tolerance grows with the frame count
    tol = 2e-5 * T + 1e-4
    assert float((got.decoder_output - ref.decoder_output).abs().max()) <= tol
    assert float((got.attention_weights - ref.attention_weights).abs().max()) <= tol
    assert float((got.stop_tokens - ref.stop_tokens).abs().max()) <= tol
    assert float((got.mel - ref.mel).abs().max()) <= 4 * tol
    first = float((got.decoder_output[:, :2] - ref.decoder_output[:, :2]).abs().max())
    assert first <= 2e-5, first                                  # the first frames agree to fp32 rounding


def test_b200_decoder_against_float64_twin(models):
    from text_to_speech_b200.tacotron2 import Tacotron2
    taco, _, tw, thp = models
    toks = _ragged_tokens(4, 31, 10, 18)
    ref = Tacotron2(thp, tw, device="cpu", dtype=torch.float64).infer(toks, max_length=30, early_stopping=False, deterministic=True)
    got = taco.infer(toks, max_length=30, early_stopping=False, deterministic=True, decoder="b200")
    assert float((got.decoder_output.cpu().double() - ref.decoder_output).abs().max()) <= 1e-3
    assert float((got.attention_weights.cpu().double() - ref.attention_weights).abs().max()) <= 1e-3


def test_b200_decoder_graph_and_direct_launch_are_bitwise_equal(models):
    taco = models[0]
    toks = _ragged_tokens(5, 8, 9, 14)
    outs = [taco.infer(toks, max_length=45, early_stopping=False, deterministic=True, decoder="b200", graph_chunk=c)
            for c in (0, 32, 6, 32)]
    for o in outs[1:]:
        assert torch.equal(o.decoder_output, outs[0].decoder_output) and torch.equal(o.attention_weights, outs[0].attention_weights)
        assert torch.equal(o.lengths, outs[0].lengths)


def test_b200_decoder_dropout_is_seeded(models):
    taco = models[0]
    toks = _ragged_tokens(2, 9, 10, 10)
    kw = dict(max_length=12, early_stopping=False, decoder="b200")
    a, b, c = taco.infer(toks, seed=1, **kw), taco.infer(toks, seed=1, **kw), taco.infer(toks, seed=2, **kw)
    d = taco.infer(toks, deterministic=True, **kw)
    assert torch.equal(a.decoder_output, b.decoder_output)
    assert not torch.equal(a.decoder_output, c.decoder_output) and not torch.equal(a.decoder_output, d.decoder_output)
    assert torch.equal(a.decoder_output[:, 0], d.decoder_output[:, 0])      # frame 0 sees prenet(0) = 0 either way


def test_b200_decoder_stop_logic(models):
    from text_to_speech_b200.tacotron2 import Tacotron2
    _, _, tw, thp = models
    tw = dict(tw)
    tw["decoder/gate_output/kernel"] = np.zeros_like(tw["decoder/gate_output/kernel"])
    tw["decoder/gate_output/bias"] = np.full(1, 10.0, np.float32)
    m = Tacotron2(thp, tw, device="cuda")
    toks = _ragged_tokens(3, 12, 6, 8)
    out = m.infer(toks, max_length=200, early_stopping=True, deterministic=True, decoder="b200")
    assert out.lengths.tolist() == [0, 0, 0]
    assert float(out.decoder_output[:, 32:].abs().max()) == 0.0               # stopped after the first 32-frame chunk
    assert float(out.stop_tokens[:, 0].min()) > 0.99


def test_b200_decoder_rejects_bad_arguments(models):
    taco = models[0]
    toks = _ragged_tokens(2, 3, 5, 5)
    toks[0, 2] = 0                                                             # a hole in the mask
    with pytest.raises(ValueError, match="prefix mask"):
        taco.infer(toks, max_length=4, decoder="b200")
    with pytest.raises(ValueError, match="decoder must be"):
        taco.infer(_ragged_tokens(1, 1, 4, 4), max_length=4, decoder="tpu")


def test_pipeline_with_the_b200_decoder(models):
    from text_to_speech_b200.tts import synthetic_texts, tts
    taco, voc, _, _ = models
    texts = synthetic_texts(5, 7, 8, 16)
    a = tts(texts, taco, voc, max_length=4.0, batch_size=4, early_stopping=False, deterministic=True, use_graph=False)
    b = tts(texts, taco, voc, max_length=4.0, batch_size=4, early_stopping=False, deterministic=True, decoder="b200")
    for i in range(5):
        assert a[i]["mel"].shape == b[i]["mel"].shape and np.abs(a[i]["mel"] - b[i]["mel"]).max() < 2e-3
        assert b[i]["audio"].shape == (b[i]["mel"].shape[0] * HOP,)


def test_b200_decoder_bf16_lstm_weights_equal_fp32_math_on_rounded_weights(models):
    """lstm_weight_dtype = 1 stores the two LSTM matrices in bf16 and keeps fp32 arithmetic: the result must be that of
    the fp32 restatement run on the SAME rounded weights (tight), and stay close to the unrounded model (loose)."""
    from text_to_speech_b200.tacotron2 import Tacotron2
    taco, _, tw, thp = models
    rounded = dict(tw)
    for k in ("decoder/attention_rnn/kernel", "decoder/attention_rnn/recurrent_kernel",
              "decoder/decoder_rnn/cell_0/kernel", "decoder/decoder_rnn/cell_0/recurrent_kernel"):
        rounded[k] = torch.from_numpy(tw[k]).to(torch.bfloat16).to(torch.float32).numpy()
    toks = _ragged_tokens(6, 77, 10, 20)
    T = 40
    ref = Tacotron2(thp, rounded, device="cuda").infer(toks, max_length=T, early_stopping=False, deterministic=True)
    got = Tacotron2(thp, tw, device="cuda", b200_lstm_weights="bf16").infer(toks, max_length=T, early_stopping=False,
                                                                             deterministic=True, decoder="b200")
    tol = 2e-5 * T + 1e-4
    assert float((got.decoder_output - ref.decoder_output).abs().max()) <= tol
    assert float((got.attention_weights - ref.attention_weights).abs().max()) <= tol
    full = taco.infer(toks, max_length=T, early_stopping=False, deterministic=True)
    dev = float((got.mel - full.mel).abs().max())
    print(f"\nbf16-stored LSTM weights vs fp32 weights: max |mel diff| over {T} frames = {dev:.3e}")
    assert dev <= 5e-2


@pytest.mark.parametrize("lstm", ["fp32", "split_bf16"])
def test_b200_decoder_against_reference_source_fixture(lib_built, lstm):
    """csrc/taco.cu against what the reference's OWN Tacotron2Decoder.infer source produced (fixture generated over
    the Keras shim, oracle/gen_golden_taco.py): three utterances the reference decoded one by one, run here as one
    padded batch."""
    from test_oracle_taco import load_taco_golden, padded_batch
    from text_to_speech_b200.tacotron2 import Tacotron2
    hp, w, mems, frames, f = load_taco_golden("taco_decoder_nvidia")
    model = Tacotron2(hp, w, device="cuda", b200_lstm_weights=lstm)
    memory, mask = padded_batch(mems, torch.float32)
    out, stops, attn, lengths = model.decode_b200(memory.cuda(), mask.cuda(), frames, early_stopping=False, deterministic=True)
    assert lengths.tolist() == [frames] * len(mems)
    worst = 0.0
    for i, m in enumerate(mems):
        e_out = np.abs(out[i].cpu().numpy() - f[f"u{i}_decoder_output_fp64"]).max()
        e_stop = np.abs(stops[i].cpu().numpy() - f[f"u{i}_stop_tokens_fp64"]).max()
        e_att = np.abs(attn[i, :, :len(m)].cpu().numpy() - f[f"u{i}_attention_weights_fp64"]).max()
        worst = max(worst, e_out, e_stop, e_att)
        assert float(attn[i, :, len(m):].abs().max()) == 0.0 if len(m) < attn.shape[2] else True
    print(f"\nCUDA decoder [{lstm} LSTM path] vs reference-source fixture (float64), {frames} frames: max abs error {worst:.2e}")
    assert worst <= (1e-5 if lstm == "fp32" else 1e-4)


def test_stream_one_sentence_at_a_time_with_savers(models, tmp_path):
    """tts.stream (Tacotron2.stream, models/tts/tacotron2.py:354-367): pre-warm, then one sentence per call from a
    queue; every audio equals the vocoder's stand-alone call on the produced mel; audios/*.wav + map.json appear."""
    import json
    import queue
    from text_to_speech_b200.tts import stream, synthetic_texts
    taco, voc = models[0], models[1]
    texts = synthetic_texts(3, 5, 6, 20)
    q = queue.Queue()
    for i, t in enumerate(texts):
        q.put((f"sentence {i}", t))
    q.put(None)
    res = list(stream(q, taco, voc, directory=str(tmp_path), max_length=24, early_stopping=False, deterministic=True,
                      multiples=(64,), max_tokens=64, max_frames=64))
    assert [r["text"] for r in res] == ["sentence 0", "sentence 1", "sentence 2"]
    for r in res:
        assert r["mel"].shape == (24, 80) and r["audio"].shape == (24 * 256,) and np.isfinite(r["audio"]).all()
        alone = voc(torch.from_numpy(r["mel"][None]).cuda(), sigma=0.6, deterministic=True)[0].cpu().numpy()
        assert np.array_equal(r["audio"], alone)
        assert r["infos"]["audio"].endswith(".wav")
    m = json.load(open(tmp_path / "map.json"))
    assert list(m) == ["sentence 0", "sentence 1", "sentence 2"] and abs(m["sentence 1"]["time"] - 24 * 256 / 22050) < 1e-9

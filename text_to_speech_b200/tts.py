"""End-to-end `tts()` pipeline around the WaveGlow path (BASELINE.json configs[3]):
token sequences -> Tacotron2 mel producer -> B200 WaveGlow runtime -> waveforms, utterance-sharded.

Mirrors the data flow of the reference's `Tacotron2.infer` wrapper (`models/tts/tacotron2.py:104-241`):
each text is synthesised to a mel trimmed to its predicted length (`outputs.mel[0, :lengths[0]]`,
:183), vocoded (`vocoder(mels[-1], **kwargs)`, :187), and returned as `{'mel', 'audio', 'rate', 'time'}`
(:196-203; an utterance with no frame gives `silence_time` seconds of zeros, :205-210). Where the
reference walks the sentences one by one at batch 1, this pipeline batches them: texts are sharded over
ranks by length (no collective, sharding.py), synthesised in batches, and the mels -- still on the
device -- are re-bucketed by frame count for the vocoder.
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence

import numpy as np

from .sharding import assign_utterances, make_batches
from .weights import HOP, SAMPLE_RATE

PAD_MEL_VALUE = -11.0     # models/tts/waveglow.py:52-58 pads mels with this value


def synthetic_texts(n: int, seed: int, min_len: int = 40, max_len: int = 120, vocab_size: int = 148) -> List[np.ndarray]:
    """Seeded random token sequences (ids 1..vocab-1; 0 is the pad token)."""
    rng = np.random.default_rng(seed)
    return [rng.integers(1, vocab_size, size=int(rng.integers(min_len, max_len + 1))).astype(np.int64) for _ in range(n)]


def tts(token_seqs: Sequence[np.ndarray], tacotron, vocoder, *, max_length=10.0, batch_size: int = 16,
        vocoder_max_frames: int = 16 * 860, sigma: float = 0.6, silence_time: float = 0.15, rank: int = 0,
        world_size: int = 1, early_stopping: bool = True, deterministic: bool = False, use_graph: bool = True,
        pad_token: int = 0, timings: dict | None = None, decoder: str = "torch", seed: int = 0) -> Dict[int, dict]:
    """Synthesises this rank's share of `token_seqs`; returns {utterance index: {'mel' [T,80] numpy,
    'audio' [256 T] numpy float32, 'rate', 'time'}}. `vocoder` is a B200WaveGlowRuntime (called with
    device tensors); `tacotron` a text_to_speech_b200.tacotron2.Tacotron2."""
    import torch
    lengths = [len(s) for s in token_seqs]
    mine = assign_utterances(lengths, world_size)[rank]
    t0 = time.perf_counter()
    mels = {}
    for batch in make_batches(mine, lengths, max_frames=batch_size * max(lengths), max_batch=batch_size):
        S = batch.T
        toks = np.full((len(batch.indices), S), pad_token, dtype=np.int64)
        for j, i in enumerate(batch.indices):
            toks[j, :lengths[i]] = token_seqs[i]
        out = tacotron.infer(toks, max_length=max_length, early_stopping=early_stopping, deterministic=deterministic,
                             use_graph=use_graph, decoder=decoder, seed=seed, return_attention=False)
        n_frames = out.lengths.cpu().tolist()
        for j, i in enumerate(batch.indices):
            mels[i] = out.mel[j, :n_frames[j]]                     # tacotron2.py:183
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    results = {}
    frames = {i: int(m.shape[0]) for i, m in mels.items()}
    voiced = [i for i in mine if frames[i] > 0]
    for i in mine:
        if frames[i] == 0:                                         # tacotron2.py:205-210
            results[i] = {"mel": np.zeros((0, 80), np.float32), "audio": np.zeros(int(silence_time * SAMPLE_RATE), np.float32),
                          "rate": SAMPLE_RATE, "time": silence_time}
    flens = [frames.get(i, 0) for i in range(len(token_seqs))]
    for batch in make_batches(voiced, flens, max_frames=vocoder_max_frames):
        x = torch.full((len(batch.indices), batch.T, 80), PAD_MEL_VALUE, dtype=torch.float32, device=mels[batch.indices[0]].device)
        for j, i in enumerate(batch.indices):
            x[j, :frames[i]] = mels[i]
        # true frame counts: each utterance is vocoded as if alone (tacotron2.py:183-187), no padding frame is computed
        wave_d = vocoder(x, sigma=sigma, lengths=[frames[i] for i in batch.indices])
        # waveforms and mels leave the device through pinned staging (one asynchronous copy each, one sync)
        wave_h = torch.empty(wave_d.shape, dtype=torch.float32, pin_memory=True)
        mel_h = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
        wave_h.copy_(wave_d, non_blocking=True)
        mel_h.copy_(x, non_blocking=True)
        torch.cuda.current_stream(x.device).synchronize()
        wave, mel_np = wave_h.numpy(), mel_h.numpy()
        for j, i in enumerate(batch.indices):
            audio = wave[j, :frames[i] * HOP].copy()               # models/tts/waveglow.py:82
            results[i] = {"mel": mel_np[j, :frames[i]].copy(), "audio": audio, "rate": SAMPLE_RATE,
                          "time": len(audio) / SAMPLE_RATE}
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if timings is not None:
        timings.update(synthesizer_s=t1 - t0, vocoder_s=t2 - t1, utterances=len(mine),
                       samples=int(sum(len(r["audio"]) for r in results.values())))
    return results


def precompile_for_stream(tacotron, vocoder, *, multiples=(64, 128), max_frames=None, max_tokens=128, max_length=10.0,
                          decoder="b200", graph_chunk=32, **_):
    """`Tacotron2.precompile_for_stream` (models/tts/tacotron2.py:354-356) warms the reference's compiled graphs with
    one dummy sentence per padding multiple (64, 128). The B200 analogue pre-pays everything a first call would:
    the vocoder's scratch, pinned staging and CUDA graphs for B = 1 at every multiple of `multiples` up to
    `max_frames` (B200WaveGlowRuntime.precompile), and the synthesizer's decoder graphs for token counts at those
    multiples up to `max_tokens`."""
    shapes = vocoder.precompile(multiples=multiples, max_frames=max_frames)
    tok_lens = sorted({m for mult in multiples for m in range(int(mult), int(max_tokens) + 1, int(mult))})
    for S in tok_lens:
        toks = np.ones((1, S), dtype=np.int64)
        tacotron.infer(toks, max_length=2 * graph_chunk, early_stopping=False, deterministic=True, use_graph=True,
                       decoder=decoder, graph_chunk=graph_chunk, return_attention=False)
    return {"vocoder_shapes": shapes, "token_lengths": tok_lens}


def stream(items, tacotron, vocoder, *, callbacks=None, directory=None, precompile=True, max_length=10.0, sigma=0.6,
           silence_time=0.15, decoder="b200", padding_multiple=64, pad_token=0, early_stopping=True,
           deterministic=False, seed=0, **kwargs):
    """`Tacotron2.stream` (models/tts/tacotron2.py:364-367 -> BaseModel.predict, base_model.py:676-713): pre-warm, then
    take sentences one at a time from `items` -- an iterable (list, generator, `queue.Queue` drained until `None`) of
    token arrays or `(text, tokens)` pairs -- synthesise, vocode and hand every result to the callbacks
    (`AudioSaver` + `JSONSaver` when `directory` is given: audios/audio-<n>.wav and map.json). One sentence per call,
    batch 1, like the reference's loop (:154-191). Tokens are padded to `padding_multiple` (the synthesizer masks the
    padding), so the pre-captured decoder graphs are hit; the vocoder gets the true frame count, so the audio is the
    stand-alone result. Yields the result dicts ({'text', 'mel', 'audio', 'rate', 'time', 'infos'})."""
    import queue as _queue
    from .audio_io import AudioSaver, JSONSaver
    callbacks = list(callbacks or [])
    if directory is not None:
        callbacks += [AudioSaver(directory), JSONSaver(directory)]
    if precompile:
        precompile_for_stream(tacotron, vocoder, max_length=max_length, decoder=decoder, **kwargs)

    def source():
        if isinstance(items, _queue.Queue):
            while True:
                it = items.get()
                if it is None:
                    return
                yield it
        else:
            yield from items

    for n, item in enumerate(source()):
        text, tokens = item if isinstance(item, tuple) else (str(n), item)
        tokens = np.asarray(tokens, dtype=np.int64)
        S = len(tokens)
        Sp = -(-S // padding_multiple) * padding_multiple if padding_multiple else S
        toks = np.full((1, Sp), pad_token, dtype=np.int64)
        toks[0, :S] = tokens
        t0 = time.perf_counter()
        ml = int(S * max_length) if isinstance(max_length, float) else int(max_length)
        out = tacotron.infer(toks, max_length=ml, early_stopping=early_stopping, deterministic=deterministic,
                             use_graph=True, decoder=decoder, seed=seed + n, return_attention=False)
        frames = int(out.lengths[0])
        if frames == 0:                                            # tacotron2.py:205-210
            audio = np.zeros(int(silence_time * SAMPLE_RATE), np.float32)
            mel = np.zeros((0, 80), np.float32)
        else:
            mel_d = out.mel[:1, :frames]                           # tacotron2.py:183
            audio = np.array(vocoder(mel_d, sigma=sigma, deterministic=deterministic)[0, :frames * HOP].cpu().numpy())
            mel = mel_d[0].cpu().numpy()
        result = {"text": text, "mel": mel, "audio": audio, "rate": SAMPLE_RATE, "time": len(audio) / SAMPLE_RATE,
                  "generation_time": time.perf_counter() - t0}
        infos = {k: v for k, v in result.items() if k not in ("mel", "audio")}
        for cb in callbacks:
            cb.apply(infos, result)
        result["infos"] = infos                                    # the map.json entry ('audio' = file path once saved)
        yield result
    for cb in callbacks:
        cb.join()

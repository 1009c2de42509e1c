"""Utterance sharding across the GPUs of one box.

The path has no cross-utterance dependency (each output sample depends only on its own mel, z and
the weights), so whole utterances are partitioned: longest-processing-time-first over ranks, then
length-bucketed batches inside a rank. No collective is involved; results stay with their rank
(pinned host buffers) and are concatenated by utterance index by whoever needs them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


def assign_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy LPT: sort by length (desc, index as tie-break), give each utterance to the least
    loaded rank. Deterministic, so every rank computes the same partition without talking."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    return shards


@dataclass
class Batch:
    indices: List[int]   # utterance ids (global)
    T: int               # padded frame count of the batch (max length in it)


def make_batches(indices: Sequence[int], lengths: Sequence[int], max_frames: int, max_batch: int = 64) -> List[Batch]:
    """Length-bucketed batches: utterances sorted by length, cut so that B * T_max <= max_frames.
    Padding frames cost FLOPs but are trimmed like models/tts/waveglow.py:82 trims to T*256."""
    idx = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    batches, cur = [], []
    for i in idx:
        tmax = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and ((len(cur) + 1) * tmax > max_frames or len(cur) >= max_batch):
            batches.append(Batch(cur, int(lengths[cur[0]])))
            cur = []
        cur.append(i)
    if cur:
        batches.append(Batch(cur, int(lengths[cur[0]])))
    return batches


def plan_batches(lengths: Sequence[int], world_size: int, max_frames: int, max_batch: int = 64,
                 ragged: bool = False) -> List[List[Batch]]:
    """Global plan used by bench.py / the multi-GPU driver. ALL utterances are bucketed by length first
    (so the spread inside a batch, hence the padding waste, does not grow with the number of ranks);
    the target number of batches is a multiple of world_size and the cuts are placed at equal shares of
    the total frame count, so every rank gets the same number of near-equal batches; batches then go to
    ranks longest-processing-time-first. Hard limits always cut: a batch never holds more than
    `max_batch` utterances nor (unless a single utterance is longer) more than `max_frames` frames --
    B * T_max padded frames, or the sum of the true lengths when `ragged` (wg_infer_ragged computes no
    padding frames). Deterministic: every rank computes the same plan without communicating."""
    if world_size <= 0 or max_frames <= 0 or max_batch <= 0:
        raise ValueError("world_size, max_frames and max_batch must be positive")
    n = len(lengths)
    idx = sorted(range(n), key=lambda i: (-int(lengths[i]), i))
    total = sum(int(l) for l in lengths)
    nb = max(-(-total // max_frames), -(-n // max_batch), 1)
    nb = -(-nb // world_size) * world_size
    if n >= world_size:
        nb = min(nb, n)

    def cost(ids):
        return sum(int(lengths[i]) for i in ids) if ragged else int(lengths[ids[0]]) * len(ids)

    batches, cur, acc, k = [], [], 0, 1
    for i in idx:
        if cur and (len(cur) >= max_batch or cost(cur + [i]) > max_frames):     # hard limits
            batches.append(cur)
            cur = []
        cur.append(i)
        acc += int(lengths[i])
        if acc * nb >= total * k and len(batches) < nb - 1:                       # equal-share cut
            batches.append(cur)
            cur = []
        while acc * nb >= total * k:
            k += 1
    if cur:
        batches.append(cur)
    out = [Batch(b, int(lengths[b[0]])) for b in batches]
    order = sorted(range(len(out)), key=lambda j: (-cost(out[j].indices), j))
    load = [0] * world_size
    per_rank: List[List[Batch]] = [[] for _ in range(world_size)]
    for j in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        per_rank[r].append(out[j])
        load[r] += cost(out[j].indices)
    return per_rank


def padding_waste(batches: Sequence[Batch], lengths: Sequence[int]) -> float:
    real = sum(int(lengths[i]) for b in batches for i in b.indices)
    padded = sum(b.T * len(b.indices) for b in batches)
    return 0.0 if padded == 0 else 1.0 - real / padded


def run_rank(vocoder, mels, plan_for_rank, pad_mel_value=-11.0, hop=256, ragged=False, **vocoder_kwargs):
    """Runs one rank's batches through `vocoder(mel[B,T,80]) -> [B, hop*T]` and returns {utterance id:
    waveform trimmed to its own length} (models/tts/waveglow.py:82 trims the same way). `mels` is the
    global list of [T_i, 80] arrays; only this rank's utterances are touched.
    ragged=True passes the true frame counts (`lengths=`, B200WaveGlowRuntime -> wg_infer_ragged): every
    utterance then equals the reference's one-at-a-time call on it (models/tts/tacotron2.py:183-191) instead
    of a call on the padded mel, whose last ~second differs (WaveGlow is not causal)."""
    import numpy as np
    out = {}
    for batch in plan_for_rank:
        n_mel = mels[batch.indices[0]].shape[1]
        x = np.full((len(batch.indices), batch.T, n_mel), pad_mel_value, dtype=np.float32)
        for j, i in enumerate(batch.indices):
            x[j, :mels[i].shape[0]] = mels[i]
        kw = dict(vocoder_kwargs)
        if ragged:
            kw["lengths"] = [int(mels[i].shape[0]) for i in batch.indices]
        y = np.asarray(vocoder(x, **kw))
        for j, i in enumerate(batch.indices):
            out[i] = y[j, :mels[i].shape[0] * hop].copy()
    return out


def bind_rank_to_cpus(local_rank: int, local_world: int):
    """Gives each rank of a box its own slice of the CPUs this process may run on (its staging memcpys and the
    launch thread then never share a core with another rank's). Returns the CPU list, or None if unsupported."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    if local_world <= 1 or len(cpus) < 2 * local_world:
        return cpus
    per = len(cpus) // local_world
    mine = cpus[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cpus
    return mine


def run_rank_pipelined(runtime, mels, plan_for_rank, pad_mel_value=-11.0, hop=256, ragged=True, sigma=1.0, z_seed=None,
                       deterministic=False):
    """One rank's share of a sharded sweep with the transfers off the critical path (inputs scattered /
    waveforms gathered through pinned host buffers): while the kernels of batch i run, the mels of batch i+1 are
    staged and copied host->device and the waveforms of batch i-1 travel device->host on a copy stream into a
    two-deep ring of pinned buffers; the host only blocks on the copy it is about to consume. `runtime` is a
    B200WaveGlowRuntime (device tensors in, device tensors out); noise is drawn on the device (z=None, the
    reference's default call) unless `deterministic`. Returns ({utterance id: waveform [hop * T_i] numpy},
    stats: h2d / d2h bytes and where the host spent its time)."""
    import time
    import numpy as np
    import torch
    dev = torch.device("cuda", runtime.engine.device)
    main = torch.cuda.current_stream(dev)
    n_mel = runtime.engine.hp.n_mel_channels
    cap_in = max((len(b.indices) * b.T * n_mel for b in plan_for_rank), default=0)
    cap_out = max((len(b.indices) * b.T * hop for b in plan_for_rank), default=0)
    st = getattr(runtime, "_sweep_state", None)          # pinned rings and the copy stream live with the runtime
    if st is None or st["cap_in"] < cap_in or st["cap_out"] < cap_out:
        st = {"cap_in": cap_in, "cap_out": cap_out, "copy_stream": torch.cuda.Stream(device=dev),
              "pin_in": [torch.empty(cap_in, dtype=torch.float32, pin_memory=True) for _ in range(2)],
              "pin_out": [torch.empty(cap_out, dtype=torch.float32, pin_memory=True) for _ in range(2)],
              # device-side rings too: no allocator traffic (and none of its implicit synchronisations) inside the loop
              "dev_in": [torch.empty(cap_in, dtype=torch.float32, device=dev) for _ in range(2)],
              "dev_out": [torch.empty(cap_out, dtype=torch.float32, device=dev) for _ in range(2)]}
        runtime._sweep_state = st
    copy_stream, pin_in, pin_out = st["copy_stream"], st["pin_in"], st["pin_out"]
    dev_in, dev_out, out_free = st["dev_in"], st["dev_out"], [None, None]
    in_free = [None, None]        # event: the H2D copy that read pin_in[i] has finished
    if z_seed is not None:
        runtime._seed, runtime._gen = int(z_seed), None
    out, pending = {}, []         # pending: (batch, pinned view, event) whose D2H is in flight
    stats = {"h2d_bytes": 0, "d2h_bytes": 0, "host_stage_s": 0.0, "host_wait_d2h_s": 0.0, "host_copy_out_s": 0.0,
             "host_launch_s": 0.0}

    def drain(upto):
        while len(pending) > upto:
            batch, view, ev = pending.pop(0)
            t0 = time.perf_counter()
            ev.synchronize()
            t1 = time.perf_counter()
            vn = view.numpy()
            for j, i in enumerate(batch.indices):
                out[i] = vn[j, :mels[i].shape[0] * hop].copy()
            stats["host_wait_d2h_s"] += t1 - t0
            stats["host_copy_out_s"] += time.perf_counter() - t1

    for k, batch in enumerate(plan_for_rank):
        B, T = len(batch.indices), batch.T
        slot = k % 2
        t0 = time.perf_counter()
        if in_free[slot] is not None:
            in_free[slot].synchronize()
        x = pin_in[slot][:B * T * n_mel].view(B, T, n_mel)
        xn = x.numpy()
        for j, i in enumerate(batch.indices):
            n = mels[i].shape[0]
            xn[j, :n] = mels[i]
            xn[j, n:] = pad_mel_value
        x_d = dev_in[slot][:B * T * n_mel].view(B, T, n_mel)
        x_d.copy_(x, non_blocking=True)
        in_free[slot] = torch.cuda.Event()
        in_free[slot].record(main)
        stats["h2d_bytes"] += x.numel() * 4
        t1 = time.perf_counter()
        lengths = [int(mels[i].shape[0]) for i in batch.indices] if ragged else None
        if out_free[slot] is not None:
            main.wait_event(out_free[slot])           # the D2H copy that read dev_out[slot] two batches ago
        y_d = runtime(x_d, sigma=sigma, deterministic=deterministic, lengths=lengths,
                      out=dev_out[slot][:B * T * hop].view(B, T * hop))
        done = torch.cuda.Event()
        done.record(main)
        t2 = time.perf_counter()
        stats["host_stage_s"] += t1 - t0
        stats["host_launch_s"] += t2 - t1
        drain(1)                   # at most one older D2H in flight: its pinned slot is the one reused next
        view = pin_out[slot][:B * T * hop].view(B, T * hop)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            view.copy_(y_d, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        out_free[slot] = ev
        pending.append((batch, view, ev))
        stats["d2h_bytes"] += B * T * hop * 4
    drain(0)
    return out, stats

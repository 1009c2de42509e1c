"""CPU, world_size 2 over gloo: the N>1 path -- every rank derives the same utterance plan without
communicating, runs only its own batches, the union of the ranks' results equals the single-process
result, and the timing reduction bench.py uses (barrier + MAX over ranks) works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from text_to_speech_b200 import sharding


def _fake_vocoder(x):
    # deterministic per-utterance stand-in (no cross-utterance or padding dependence)
    return np.repeat(x[:, :, 0] * 3.0 + x[:, :, 1], 256, axis=1)


def _make_mels(n=37, seed=3):
    rng = np.random.default_rng(seed)
    return [rng.normal(size=(int(t), 80)).astype(np.float32) for t in rng.integers(5, 60, size=n)]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mels = _make_mels()
    lengths = [m.shape[0] for m in mels]
    plan = sharding.plan_batches(lengths, world, max_frames=200)
    mine = sharding.run_rank(_fake_vocoder, mels, plan[rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.tobytes() for k, v in mine.items()})
    t = torch.tensor([0.5 + rank], dtype=torch.float64)          # bench.py: MAX over ranks of the device time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret["owners"] = [sorted(g.keys()) for g in gathered]
        ret["union"] = {k: v for g in gathered for k, v in g.items()}
        ret["tmax"] = float(t.item())
    dist.destroy_process_group()


def test_two_ranks_cover_the_workload_exactly_once():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    mels = _make_mels()
    single = sharding.run_rank(_fake_vocoder, mels, sharding.plan_batches([m.shape[0] for m in mels], 1, 200)[0])
    owners = ret["owners"]
    assert sorted(owners[0] + owners[1]) == list(range(len(mels)))          # every utterance exactly once
    assert not set(owners[0]) & set(owners[1])
    for i, m in enumerate(mels):
        got = np.frombuffer(ret["union"][i], dtype=np.float32)
        assert got.shape == (m.shape[0] * 256,) and np.array_equal(got, single[i])
    assert ret["tmax"] == 1.5

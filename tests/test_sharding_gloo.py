"""world_size-2 gloo test (CPU) of the N>1 host logic: frame sharding + gather of the padded keypoint records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperpose_b200 import capi, sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_results(rank, B, cap):
    rng = np.random.default_rng(100 + rank)
    counts = rng.integers(0, cap + 1, size=B).astype(np.int32)
    rec = np.zeros((B, cap), capi.HUMAN_DT)
    for i in range(B):
        rec["score"][i, :counts[i]] = rng.random(counts[i]) + rank
        rec["parts"]["x"][i, :counts[i]] = rng.random((counts[i], 18))
    return rec, counts


def _worker(rank, world, port, B, cap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec, counts = _fake_results(rank, B, cap)
    h = torch.from_numpy(np.frombuffer(rec.tobytes(), np.uint8).copy())
    c = torch.from_numpy(counts)
    gh, gc = sharding.gather_records(h, c, world)
    if rank == 0:
        q.put((gh.numpy().copy(), gc.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2():
    world, B, cap = 2, 4, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, cap, q)) for r in range(world)]
    for p in procs: p.start()
    gh, gc = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    frames = sharding.unpack_records(gh, gc, cap)
    assert len(frames) == world * B
    for r in range(world):
        rec, counts = _fake_results(r, B, cap)
        for i in range(B):
            assert frames[r * B + i].tobytes() == rec[i, :counts[i]].tobytes()


@pytest.mark.parametrize("n,world", [(128, 8), (17, 4), (3, 8), (16, 1)])
def test_shard_range_partitions_frames(n, world):
    seen = []
    for r in range(world):
        lo, hi = sharding.shard_range(n, world, r)
        seen += list(range(lo, hi))
        assert hi - lo in (n // world, n // world + 1)
    assert seen == list(range(n))

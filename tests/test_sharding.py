"""CPU tests of the N>1 host logic: frame sharding + end-of-clip gather over gloo, world_size 2."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depth_pro import video


def test_shard_frames_partition():
    for n in (0, 1, 7, 240):
        for world in (1, 2, 4, 8):
            parts = [video.shard_frames(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        video.shard_frames(10, 2, 2)


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = video.shard_frames(n_frames, rank, world)
    # stand-in for the per-frame work: a deterministic function of the frame index only
    recs = [{"index": i, "checksum": float(torch.arange(i + 3, dtype=torch.float64).sum()), "rank": rank} for i in mine]
    out = video.gather_records(recs, rank, world)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r["index"] for r in out] == list(range(n))
    assert all(r["rank"] == r["index"] % 2 for r in out)
    # sharding must not change any result: identical to a single-rank run
    single = [float(torch.arange(i + 3, dtype=torch.float64).sum()) for i in range(n)]
    assert [r["checksum"] for r in out] == single


def test_colormap_lut_shape():
    lut = video.colormap_lut("turbo")
    assert lut.shape == (256, 3) and lut.dtype.name == "uint8"
    with pytest.raises(ValueError):
        video.colormap_lut("nope")

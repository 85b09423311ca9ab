"""world_size-2 gloo test of the frame-sharding host logic (no GPU, no collective on the data path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcdet_b200 import sharding


def test_partition_covers_every_frame_once():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            seen = sorted(f for r in range(world) for f in sharding.frames_of_rank(n, r, world))
            assert seen == list(range(n))
            sizes = [len(sharding.frames_of_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.batches_of_rank(32, 1, 8, 4) == [[1, 9, 17, 25]]
    assert sharding.batches_of_rank(10, 0, 2, 2) == [[0, 2], [4, 6], [8]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        own = sharding.frames_of_rank(n_frames, rank, world)
        # every rank "processes" its frames independently: the stand-in result depends on the frame only
        local = [np.arange(f % 5 + 1) + 100 * f for f in own]
        allres = sharding.gather_frame_results(local, n_frames)
        ok = all(np.array_equal(allres[f], np.arange(f % 5 + 1) + 100 * f) for f in range(n_frames))
        t = sharding.max_over_ranks(1.0 + rank)
        dist.barrier()
        q.put((rank, ok, t, len(own)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_and_max_timing():
    world, n_frames = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]          # max over ranks on both
    assert sum(r[3] for r in res) == n_frames

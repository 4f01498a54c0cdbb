"""N>1 path on CPU: world_size-2 gloo processes exercise the batch sharding,
ragged token gathering and max-over-ranks timing reduction used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import frx


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        images = torch.arange(n_items, dtype=torch.float32).view(n_items, 1)
        mine = frx.sharding.shard(images)
        # stand-in for the decode: token row i = [i, 2i, 3i] -- checks ORDER after the gather
        local = torch.stack([mine[:, 0].long() * k for k in (1, 2, 3)], 1)
        full = frx.sharding.gather_tokens(local, n_items)
        slowest = frx.sharding.max_over_ranks(10.0 + rank)
        ret[rank] = (mine[:, 0].tolist(), full.tolist(), slowest)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch_exactly():
    for n in (0, 1, 5, 7, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [frx.sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding_and_gather():
    n_items, world = 7, 2   # ragged: 4 + 3
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, ret), nprocs=world, join=True)
    assert ret[0][0] == [0.0, 1.0, 2.0, 3.0] and ret[1][0] == [4.0, 5.0, 6.0]
    expect = [[i, 2 * i, 3 * i] for i in range(n_items)]
    assert ret[0][1] == expect and ret[1][1] == expect
    assert ret[0][2] == 11.0 and ret[1][2] == 11.0     # max over ranks on every rank


def _grad_worker(rank, world, port, ret):
    """Training data parallelism: the bucketed, callback-driven all-reduce of the flat gradient buffer (what
    EfficientSATRN.train_step does over NCCL) equals one all-reduce of the whole buffer, whatever the buckets cover."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        g = torch.Generator().manual_seed(rank)
        flat = torch.randn(n, generator=g)
        mine = flat.clone()
        red = frx.sharding.GradBucketReducer(flat)
        red.begin()
        # the library reports buckets in backward completion order (decoder first, stem last); leave two gaps
        for off, cnt in ((700, 300), (400, 250), (100, 300)):
            red.on_bucket(off, cnt)
        gaps = red.uncovered()
        n_works = red.finish()
        want = mine.clone()
        dist.all_reduce(want)
        ret[rank] = (gaps, n_works, bool(torch.equal(flat, want)), flat.sum().item())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucketed_gradient_allreduce():
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_grad_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        gaps, n_works, same, _ = ret[r]
        assert gaps == [(0, 100), (650, 700)]
        assert n_works == 5 and same
    assert ret[0][3] == ret[1][3]      # both ranks hold the same summed gradient

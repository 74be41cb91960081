"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, result gather, max-over-ranks timing.
The per-pair compute is replaced by a deterministic stand-in (the kernels themselves need a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_is_balanced_and_covers():
    from rnc.dist import shard_range
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]      # BASELINE configs[3]: 64 pairs -> 8 per GPU
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_model(im1, im2, iters=1):
    lo = (im1[:, :2, ::8, ::8] - im2[:, :2, ::8, ::8]) * iters
    return lo, im1[:, :2] + im2[:, :2]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rnc.dist import infer_sharded, max_over_ranks
        g = torch.Generator().manual_seed(0)
        im1, im2 = torch.rand(n, 3, 16, 24, generator=g), torch.rand(n, 3, 16, 24, generator=g)
        lo, up = infer_sharded(_fake_model, im1, im2, gather=True, iters=3)
        rlo, rup = _fake_model(im1, im2, iters=3)
        ok = torch.equal(lo, rlo) and torch.equal(up, rup)
        local = infer_sharded(_fake_model, im1, im2, gather=False, iters=3)
        n_local = 0 if local is None else local[0].shape[0]
        t = max_over_ranks(10.0 + rank)
        q.put((rank, ok, n_local, t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8, 1])
def test_infer_sharded_world2_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)                                   # gathered result == single-process result
    assert sum(nl for _, _, nl, _ in res) == n                               # shards partition the batch
    assert all(t == 11.0 for _, _, _, t in res)                              # max over ranks

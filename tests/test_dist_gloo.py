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


# ----------------------------------------------------------------------------- training step under DDP (config #5 host logic)


class _TinyFlowNet(torch.nn.Module):
    """CPU stand-in with RAFT's training call signature: model(image1, image2, iters) -> list of predictions."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.conv = torch.nn.Conv2d(6, 2, 3, padding=1)

    def forward(self, image1, image2, iters=12):
        f = self.conv(torch.cat([image1, image2], 1) / 255.0)
        return [f * (i + 1) / iters for i in range(iters)]


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from rnc.dist import shard_range
        from rnc.train import fetch_optimizer, train_step
        g = torch.Generator().manual_seed(1)
        im1, im2 = torch.rand(4, 3, 16, 24, generator=g) * 255, torch.rand(4, 3, 16, 24, generator=g) * 255
        gt, valid = torch.randn(4, 2, 16, 24, generator=g), torch.ones(4, 16, 24)
        lo, hi = shard_range(4, world, rank)
        net = DDP(_TinyFlowNet())
        opt, sched = fetch_optimizer(net, lr=1e-2, num_steps=10)
        loss, _ = train_step(net, opt, sched, im1[lo:hi], im2[lo:hi], gt[lo:hi], valid[lo:hi], iters=3, clip=1.0)
        # plain lists: a tensor in the queue is shared through a file descriptor that dies with this process
        q.put((rank, [p.detach().tolist() for p in net.module.parameters()], float(loss)))
    finally:
        dist.destroy_process_group()


def test_train_step_under_ddp_equals_the_full_batch_step():
    """Two ranks with half the batch each, gradients averaged by DDP's all-reduce == one process on the whole batch (the loss is
    a mean over equal shards): the replicas stay identical and match the single-process update (train.py:203-227 semantics)."""
    from rnc.train import fetch_optimizer, train_step
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    g = torch.Generator().manual_seed(1)
    im1, im2 = torch.rand(4, 3, 16, 24, generator=g) * 255, torch.rand(4, 3, 16, 24, generator=g) * 255
    gt, valid = torch.randn(4, 2, 16, 24, generator=g), torch.ones(4, 16, 24)
    net = _TinyFlowNet()
    opt, sched = fetch_optimizer(net, lr=1e-2, num_steps=10)
    train_step(net, opt, sched, im1, im2, gt, valid, iters=3, clip=1.0)
    for a, b, c in zip(res[0][1], res[1][1], net.parameters()):
        a, b = torch.tensor(a), torch.tensor(b)
        assert torch.equal(a, b)                                  # replicas in sync
        assert torch.allclose(a, c.detach(), atol=1e-6)           # == full-batch step

"""world_size-2 gloo test (CPU) of the bucketed all-reduce host logic: averaged gradients equal the
single-process gradient of the concatenated batch; buckets are launched in backward order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_vit_b200.ddp import BucketedAllReduce, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                               torch.nn.Linear(64, 2))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _net()
        if rank == 1:                      # perturb: the constructor must broadcast rank 0's weights
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        ddp = BucketedAllReduce(net, bucket_mb=0.004)   # ~4 KB buckets -> several buckets
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 16, generator=g), torch.randint(0, 2, (8,), generator=g)
        xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
        for _ in range(2):                 # two steps: bucket state must reset
            net.zero_grad(set_to_none=False)
            for p in net.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            loss = torch.nn.functional.cross_entropy(net(xs), ys)
            loss.backward()
            ddp.finish()
        grads = [p.grad.clone() for p in net.parameters()]
        launched = ddp.launched
        # gradient accumulation: two half-batches, the first under no_sync(), must give the same averaged gradient
        net.zero_grad(set_to_none=True)
        h = xs.shape[0] // 2
        with ddp.no_sync():
            (0.5 * torch.nn.functional.cross_entropy(net(xs[:h]), ys[:h])).backward()
            ddp.finish()                   # no-op inside no_sync
        (0.5 * torch.nn.functional.cross_entropy(net(xs[h:]), ys[h:])).backward()
        ddp.finish()
        acc = [p.grad.clone().numpy() for p in net.parameters()]
        # a second backward without no_sync() while buckets are in flight must raise, not corrupt the buckets
        net.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(xs), ys).backward()
        raised = False
        try:
            torch.nn.functional.cross_entropy(net(xs), ys).backward()
        except RuntimeError as e:
            raised = "no_sync" in str(e)
        ddp.finish()
        q.put((rank, len(ddp.buckets), launched, [g.numpy() for g in grads], acc, raised))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = _net()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 16, generator=g), torch.randint(0, 2, (8,), generator=g)
    # mean over ranks of per-rank mean losses == mean over the full batch (equal shard sizes)
    torch.nn.functional.cross_entropy(net(x), y).backward()
    ref = [p.grad for p in net.parameters()]
    for rank, nb, launched, grads, acc, raised in results:
        assert nb >= 3 and launched == 2 * nb
        assert raised, "double backward without no_sync() must raise"
        for a, b, c in zip(grads, ref, acc):
            assert torch.allclose(torch.from_numpy(a), b, atol=1e-6), rank
            assert torch.allclose(torch.from_numpy(c), b, atol=1e-6), rank


def test_requires_initialised_process_group():
    with pytest.raises(RuntimeError, match="not initialised"):
        BucketedAllReduce(_net())

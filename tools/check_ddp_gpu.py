"""Numerical check of the sink-mode data-parallel path on real GPUs (run under torchrun on >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_gpu.py

Every rank takes its shard of one batch, runs forward + CE + backward with `BucketedAllReduce` (gradients written
straight into the NCCL buckets, all_reduce(AVG) overlapped with backward) and `FusedAdamW`; rank 0 also computes the
full-batch gradients with a second, non-distributed model.  Checks: (1) the averaged gradients equal the full-batch
gradients, (2) after two optimizer steps (the second with gradient accumulation under no_sync) all ranks hold
bit-identical parameters and they match a single-process run of the same schedule.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_vit_b200 as nv  # noqa: E402
from neural_vit_b200.ddp import shard_batch  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    kw = dict(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=3, dropout=0.0,
              attention_dropout=0.0, drop_path=0.0)
    cfg = nv.Temporal3DViTConfig(**kw)
    for precision, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        torch.manual_seed(100 + rank)                      # different init per rank: the constructor must broadcast
        model = nv.Temporal3DViT(cfg, precision=precision).to(dev).train()
        ddp = nv.BucketedAllReduce(model, bucket_mb=0.25)
        opt = nv.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.01, model=model)
        assert len(ddp.buckets) >= 3 and ddp.sink is not None
        g = torch.Generator().manual_seed(7)
        B = 8 * world
        x = torch.randn(B, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(dev)
        y = torch.randint(0, 2, (B,), generator=g).to(dev)
        xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
        # reference: same (rank-0) weights, full batch, no DDP, stock optimizer
        ref = nv.Temporal3DViT(cfg, precision=precision).to(dev).train()
        ref.load_state_dict(model.state_dict())
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.01)

        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(xs), ys).backward()
        ddp.finish()
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        worst = max(rel(p.grad, q.grad) for p, q in zip(model.parameters(), ref.parameters()))
        assert worst < tol, (precision, "averaged gradient vs full batch", worst)
        opt.step()
        ropt.step()
        # step 2: gradient accumulation over two half-shards, the first under no_sync()
        h = xs.shape[0] // 2
        opt.zero_grad()
        with ddp.no_sync():
            (0.5 * torch.nn.functional.cross_entropy(model(xs[:h]), ys[:h])).backward()
        (0.5 * torch.nn.functional.cross_entropy(model(xs[h:]), ys[h:])).backward()
        ddp.finish()
        opt.step()
        ropt.zero_grad()
        idx = torch.cat([torch.arange(r, B, world)[:h] for r in range(world)] + [torch.arange(r, B, world)[h:] for r in range(world)])
        torch.nn.functional.cross_entropy(ref(x[idx]), y[idx]).backward()
        ropt.step()
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), (precision, "ranks diverged")
        pw = max(rel(p, q) for p, q in zip(model.parameters(), ref.parameters()))
        assert pw < (1e-4 if precision == "fp32" else 3e-2), (precision, "parameters after 2 steps", pw)
        if rank == 0:
            print(f"[{precision}] world={world} buckets={len(ddp.buckets)}: averaged-gradient error {worst:.2e}, "
                  f"ranks bit-identical after 2 steps, parameters vs single-process run {pw:.2e}", flush=True)
        ddp.remove()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

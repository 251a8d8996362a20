"""Batch-sharded data parallelism for the drop-in model: bucketed gradient all-reduce over NCCL
(NVLink 5 / NVSwitch), overlapped with backward.

The reference has no distributed code; samples are independent, so the only exchange
step of the path is the gradient all-reduce.  One process per GPU (torchrun); every rank holds the
full fp32 parameters.  Gradients are produced block by block (one autograd Function per encoder
block), each parameter's post-accumulate hook copies its gradient into a flat bucket view, and as
soon as a bucket is complete its all-reduce is launched asynchronously on NCCL's stream while the
remaining backward kernels keep running.  ``finish()`` waits for the outstanding buckets; after it
``p.grad`` of every parameter is a view into the averaged flat bucket (no copy back).

The class is backend-agnostic (it only uses torch.distributed), so the host logic is unit-tested on
CPU with gloo at world_size 2.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        self.views = []
        off = 0
        for p in params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.pending = len(params)
        self.handle = None


class BucketedAllReduce:
    """Gradient averaging across the default (or given) process group with compute/communication overlap.

    Usage (mirrors the four hot-loop lines of the reference, train.py:223-227):
        ddp = BucketedAllReduce(model)            # once
        optimizer.zero_grad(); loss = crit(model(x), y); loss.backward(); ddp.finish(); optimizer.step()
    """

    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None,
                 broadcast_parameters: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        params = [p for p in module.parameters() if p.requires_grad]
        if broadcast_parameters:
            for p in module.parameters():
                dist.broadcast(p.data, src=0, group=process_group)
        # buckets in reverse registration order == the order backward produces gradients
        limit = int(bucket_mb * 1024 * 1024)
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > limit or cur[0].dtype != p.dtype):
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(_Bucket(cur))
        self._where = {}
        self._hooks = []
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._where[p] = (bi, pi)
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched = 0

    # called by autograd right after p.grad has been accumulated
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi, pi = self._where[p]
        b = self.buckets[bi]
        view = b.views[pi]
        if p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
            p.grad = view
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket) -> None:
        if self.world > 1:
            b.handle = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched += 1

    def finish(self) -> None:
        """Wait for every outstanding bucket and turn sums into means.  Call after backward()."""
        for b in self.buckets:
            if b.pending != 0:
                # parameters that received no gradient this step (unused) contribute zeros
                for p, v in zip(b.params, b.views):
                    if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                        if p.grad is None:
                            v.zero_()
                        else:
                            v.copy_(p.grad)
                        p.grad = v
                self._launch(b)
            if b.handle is not None:
                b.handle.wait()
                b.handle = None
            if self.world > 1:
                b.flat.div_(self.world)
            b.pending = len(b.params)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r takes samples r, r+world, ... (SURVEY.md section 8e partitioning)."""
    return x[rank::world]

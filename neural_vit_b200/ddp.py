"""Batch-sharded data parallelism for the drop-in model: bucketed gradient all-reduce over NCCL
(NVLink 5 / NVSwitch), overlapped with backward.

The reference has no distributed code; samples are independent, so the only exchange
step of the path is the gradient all-reduce.  One process per GPU (torchrun); every rank holds the full fp32
parameters.

Two modes, chosen by what the wrapped module offers:

* **sink mode** (``Temporal3DViT``): the model's backward kernels accumulate every parameter gradient *directly* into
  the flat fp32 buckets of a ``GradSink`` (gradsink.py) -- no AccumulateGrad pass, no copy into a communication
  buffer -- and tell the sink when a block's gradient kernels have been enqueued; as soon as a bucket is complete its
  ``all_reduce(AVG)`` is launched asynchronously on NCCL's stream while the remaining backward kernels keep running.
  There is no separate averaging pass (NCCL ``AVG``; gloo, used by the CPU tests, falls back to ``SUM`` + one scale).
* **hook mode** (any other ``nn.Module``, e.g. the CPU tests): post-accumulate-grad hooks copy each gradient into its
  bucket view.

``finish()`` waits for the outstanding buckets; after it ``p.grad`` of every parameter is a view into the averaged
flat bucket.  ``no_sync()`` suppresses communication for gradient accumulation; calling ``backward`` twice without it
raises instead of silently reducing a bucket that is already in flight.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional

import torch
import torch.distributed as dist

from .gradsink import Bucket, GradSink


def _bump(p: torch.Tensor) -> None:
    torch.autograd.graph.increment_version(p)


class BucketedAllReduce:
    """Gradient averaging across the default (or given) process group with compute/communication overlap.

    Usage (mirrors the four hot-loop lines of the reference, train.py:223-227):
        ddp = BucketedAllReduce(model)            # once
        optimizer.zero_grad(); loss = crit(model(x), y); loss.backward(); ddp.finish(); optimizer.step()
    """

    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None,
                 broadcast_parameters: bool = True, measure_exposed: bool = False):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.module = module
        backend = dist.get_backend(process_group)
        self._avg = backend == "nccl"                    # in-collective averaging: no division pass afterwards
        self._sync = True
        self.launched = 0
        self.measure_exposed = measure_exposed
        self.exposed_events: List[tuple] = []            # (start, end) CUDA events around the waits in finish()
        params = [p for p in module.parameters() if p.requires_grad]
        if broadcast_parameters:
            with torch.no_grad():
                for p in module.parameters():
                    dist.broadcast(p, src=0, group=process_group)
                    _bump(p)                              # the bf16 operand shadows are keyed on _version
            if hasattr(module, "invalidate_shadows"):
                module.invalidate_shadows()
        self.sink: Optional[GradSink] = None
        self._hooks = []
        if hasattr(module, "attach_grad_sink") and all(p.is_cuda for p in params):
            self.sink = module.attach_grad_sink(bucket_mb=bucket_mb)
            self.sink.on_bucket_ready = self._on_bucket_ready
            self.buckets: List[Bucket] = self.sink.buckets
        else:
            self.buckets = self._make_buckets(params, bucket_mb)
            self._where = {}
            for bi, b in enumerate(self.buckets):
                for pi, p in enumerate(b.params):
                    self._where[p] = (bi, pi)
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- hook mode --------------------------------------------------------------------------------------------
    @staticmethod
    def _make_buckets(params, bucket_mb: float) -> List[Bucket]:
        limit = int(bucket_mb * 1024 * 1024)
        buckets, cur, cur_bytes = [], [], 0
        for p in reversed(params):      # reverse registration order == the order backward produces gradients
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > limit or cur[0].dtype != p.dtype):
                buckets.append(_HookBucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            buckets.append(_HookBucket(cur))
        return buckets

    def _on_grad(self, p: torch.nn.Parameter) -> None:     # autograd: right after p.grad has been accumulated
        bi, pi = self._where[p]
        b = self.buckets[bi]
        if b.launched:
            raise RuntimeError("backward() ran again while this bucket's all-reduce was in flight; wrap the extra "
                               "backward passes of gradient accumulation in `with ddp.no_sync():`")
        view = b.views[pi]
        if p.grad.data_ptr() != view.data_ptr():
            if b.accumulated:
                view.add_(p.grad)       # a no_sync() pass left earlier gradients in the bucket
            else:
                view.copy_(p.grad)
            p.grad = view
        b.pending -= 1
        if b.pending == 0:
            if self._sync:
                self._launch(b)
            else:
                b.accumulated = True
                b.pending = len(b.params)

    # ---- sink mode --------------------------------------------------------------------------------------------
    def _on_bucket_ready(self, b: Bucket) -> None:
        if b.launched:
            raise RuntimeError("backward() ran again while this bucket's all-reduce was in flight; wrap the extra "
                               "backward passes of gradient accumulation in `with ddp.no_sync():`")
        if self._sync:
            self._launch(b)

    # ---- common -----------------------------------------------------------------------------------------------
    def _launch(self, b) -> None:
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            b.handle = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        b.launched = True
        self.launched += 1

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside accumulate locally; the next pass outside reduces the accumulated total."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def finish(self) -> None:
        """Wait for every outstanding bucket; afterwards p.grad holds the cross-rank mean.  Call after backward()."""
        if not self._sync:
            return
        ev = None
        if self.measure_exposed and torch.cuda.is_available():
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        for b in self.buckets:
            if not b.launched:
                if self.sink is None:
                    # parameters that received no gradient this step (unused) contribute zeros
                    for p, v in zip(b.params, b.views):
                        if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                            if p.grad is None:
                                if not b.accumulated:
                                    v.zero_()
                            elif b.accumulated:
                                v.add_(p.grad)
                            else:
                                v.copy_(p.grad)
                            p.grad = v
                self._launch(b)
            if b.handle is not None:
                b.handle.wait()
                b.handle = None
            if self.world > 1 and not self._avg:
                b.flat.mul_(1.0 / self.world)
            b.pending = len(b.params)
            b.launched = False
            if self.sink is None:
                b.accumulated = False
        if ev is not None:
            ev[1].record()
            self.exposed_events.append(ev)
        if self.sink is not None:
            self.sink.attach_grads()
            self.sink.consumed = True

    def exposed_ms(self) -> float:
        """Mean time the compute stream spent waiting for the collectives in finish() (not overlapped)."""
        if not self.exposed_events:
            return 0.0
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in self.exposed_events) / len(self.exposed_events)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self.sink is not None:
            self.sink.on_bucket_ready = None


class _HookBucket:
    """Bucket of the hook mode (any device / dtype; no alignment padding)."""

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
        self.launched = False
        self.accumulated = False


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r takes samples r, r+world, ... (SURVEY.md section 8e partitioning)."""
    return x[rank::world]

"""Flat gradient buffers that the backward kernels accumulate into directly.

SURVEY.md section 8(e): "wgrad kernels can write straight into the bucket (grad-as-bucket-view) to avoid a copy ...
wgrad epilogue -> bucket buffer -> NCCL with no intermediate pass".  A ``GradSink`` owns one flat fp32 buffer per
bucket (buckets in reverse registration order = the order backward produces gradients); every parameter's ``.grad``
is a view into its bucket.  The model's autograd Functions hand those views to the weight-gradient GEMMs (split-K
``red`` accumulation), the bias column sums, the LayerNorm / LayerScale / positional reductions -- all of which
*add* to their destination -- and return ``None`` to autograd for those inputs, so there is no AccumulateGrad pass,
no per-step ``zeros`` allocation and no copy into a communication buffer.  ``FusedAdamW`` reads the same buffers
(one launch per bucket) and ``BucketedAllReduce`` all-reduces them in place.

Zero-filling is lazy: whoever consumes the gradients (``FusedAdamW.step`` / ``BucketedAllReduce.finish``) marks the
sink consumed and the next training forward clears the buffers with one memset per bucket; several backward passes
before that accumulate, which is what gradient accumulation / ``no_sync`` needs.  ``zero_grad(set_to_none=True)`` of
a stock optimizer is honoured too (a dropped ``.grad`` counts as consumed and is re-attached after backward).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

ALIGN = 64  # elements: every parameter starts on a 256-byte boundary of its bucket (vector / TMA alignment)


class Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.offsets: List[int] = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=params[0].device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, params)]
        self.pending = len(params)
        self.handle = None          # in-flight all-reduce (BucketedAllReduce)
        self.launched = False


class GradSink:
    def __init__(self, params: List[torch.nn.Parameter], bucket_mb: float = 32.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("GradSink needs at least one trainable parameter")
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("GradSink holds fp32 CUDA parameters only")
        limit = int(bucket_mb * 1024 * 1024)
        self.buckets: List[Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(params):                      # backward order
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > limit:
                self.buckets.append(Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(Bucket(cur))
        self._where: Dict[int, tuple] = {}
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._where[id(p)] = (bi, pi)
                p.grad = b.views[pi]
        self.consumed = False
        self.on_bucket_ready: Optional[Callable[[Bucket], None]] = None

    # ------------------------------------------------------------------------------------------
    def view(self, p: torch.nn.Parameter) -> Optional[torch.Tensor]:
        w = self._where.get(id(p))
        return None if w is None else self.buckets[w[0]].views[w[1]]

    def zero(self) -> None:
        for b in self.buckets:
            b.flat.zero_()
        self.consumed = False

    def begin_pass(self) -> None:
        """Called by the model at the start of every training forward."""
        if self.consumed or self.buckets[0].params[0].grad is None:
            self.zero()
        for b in self.buckets:
            b.pending = len(b.params)

    def mark_ready(self, params) -> None:
        """The gradient kernels of ``params`` have been enqueued on the current stream."""
        for p in params:
            w = self._where.get(id(p))
            if w is None:
                continue
            b = self.buckets[w[0]]
            if p.grad is None:
                p.grad = b.views[w[1]]
            b.pending -= 1
            if b.pending == 0 and self.on_bucket_ready is not None:
                self.on_bucket_ready(b)

    def attach_grads(self) -> None:
        for b in self.buckets:
            for p, v in zip(b.params, b.views):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v

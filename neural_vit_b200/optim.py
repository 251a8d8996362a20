"""Fused AdamW for the drop-in model.

Replaces ``torch.optim.AdamW(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay)`` of the reference
(train.py:154-156, step at :227; train_hptune.py:321-325) with the same update rule, hyper-parameter names and
``param_groups`` interface, but

* parameters, gradients and both moments live in flat fp32 buffers, bucket by bucket -- the *same* buckets the
  backward kernels accumulate into (gradsink.py) and ``BucketedAllReduce`` all-reduces -- so one ``tvit_adamw`` launch
  per bucket updates everything (a C2 model is 3 launches instead of ~110 tensors x several foreach kernels);
* the same kernel writes the bf16 operand shadow of every parameter it updates (the ``W`` [out, in] copies the forward
  tensor-core GEMMs read), and one ``tvit_shadow_t_multi`` launch rebuilds all transposed, LayerScale-prescaled
  ``(gamma (.) W)^T`` copies the input-gradient GEMMs read; the model's shadow cache adopts them, so no cast kernels
  run in the next forward;
* ``zero_grad`` is one memset per bucket.

Differences from torch.optim.AdamW, all irrelevant for this model: a single parameter group; ``amsgrad``,
``maximize`` and ``capturable`` are not offered; a parameter that received no gradient in a step still sees weight
decay (its slot in the flat gradient buffer is simply zero).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .gradsink import GradSink


def _bump_versions(tensors) -> None:
    try:
        torch.autograd.graph.increment_version(tensors)
    except TypeError:                                    # older signature: one tensor at a time
        for t in tensors:
            torch.autograd.graph.increment_version(t)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, *,
                 model: Optional[nn.Module] = None, bucket_mb: float = 32.0):
        if isinstance(params, nn.Module):
            model, params = params, list(params.parameters())
        params = list(params)
        if params and isinstance(params[0], dict):
            if len(params) != 1:
                raise ValueError("FusedAdamW supports a single parameter group")
            group0 = dict(params[0])
            params = list(group0.pop("params"))
            lr, betas = group0.get("lr", lr), group0.get("betas", betas)
            eps, weight_decay = group0.get("eps", eps), group0.get("weight_decay", weight_decay)
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        trainable = [p for p in params if p.requires_grad]
        if not trainable or not all(p.is_cuda and p.dtype == torch.float32 for p in trainable):
            raise RuntimeError("FusedAdamW needs fp32 CUDA parameters (move the model to the GPU first; "
                               "there is no CPU fallback)")
        self.model = model if (model is not None and hasattr(model, "attach_grad_sink")) else None
        if self.model is not None:
            self.sink: GradSink = self.model.attach_grad_sink(bucket_mb=bucket_mb)
            have = {id(p) for b in self.sink.buckets for p in b.params}
            if have != {id(p) for p in trainable}:
                raise ValueError("FusedAdamW(model=...) must be given exactly the model's trainable parameters")
        else:
            self.sink = GradSink(trainable, bucket_mb)
        self._steps = 0
        self._bf16 = self.model is not None and ops.torch_dtype(self._model_dtype()) == torch.bfloat16
        with torch.no_grad():
            for b in self.sink.buckets:
                dev = b.flat.device
                b.flat_p = torch.zeros(b.numel, dtype=torch.float32, device=dev)
                for p, off in zip(b.params, b.offsets):
                    view = b.flat_p[off:off + p.numel()].view_as(p)
                    view.copy_(p.detach())
                    p.data = view                                   # the Parameter now lives in the flat buffer
                b.flat_m = torch.zeros_like(b.flat_p)
                b.flat_v = torch.zeros_like(b.flat_p)
                b.flat_shadow = b.flat_p.to(torch.bfloat16) if self._bf16 else None
        self._shadow_entries: List[tuple] = []
        self._desc = None
        if self._bf16:
            self._build_shadow_table()
            self._refresh_shadows()

    # ------------------------------------------------------------------------------------------
    def _model_dtype(self) -> int:
        from .model import _PRECISIONS
        return _PRECISIONS[self.model.precision][2]

    def _slot(self, p):
        bi, pi = self.sink._where[id(p)]
        b = self.sink.buckets[bi]
        return b, b.offsets[pi]

    def _build_shadow_table(self) -> None:
        """One entry per weight the tensor-core GEMMs read: (cache key, weight, gamma, straight shadow view, wt)."""
        m = self.model
        items = [("patch_embed", m.patch_embed.weight, None, False)]
        for i, blk in enumerate(m.blocks):
            n1w, n1b, qkvw, qkvb, pw, pb, g1, n2w, n2b, f1w, f1b, f2w, f2b, g2 = blk.tensors()
            items += [((i, "qkv"), qkvw, None, True), ((i, "proj"), pw, g1, True),
                      ((i, "fc1"), f1w, None, True), ((i, "fc2"), f2w, g2, True)]
        descs, tiles = [], 0
        for key, w, gamma, need_t in items:
            b, off = self._slot(w)
            R = w.shape[0]
            C = w.numel() // R
            w_sh = b.flat_shadow[off:off + w.numel()].view(R, C)
            wt_sh = torch.empty((C, R), dtype=torch.bfloat16, device=w.device) if need_t else None
            self._shadow_entries.append((key, w, gamma, w_sh, wt_sh))
            if need_t:
                tx, ty = (C + 31) // 32, (R + 31) // 32
                descs.append(L.ShadowDesc(w.data_ptr(), 0 if gamma is None else gamma.data_ptr(), wt_sh.data_ptr(),
                                          R, C, tiles, tx))
                tiles += tx * ty
        arr = (L.ShadowDesc * len(descs))(*descs)
        raw = bytes(memoryview(arr))
        dev = self.sink.buckets[0].flat.device
        self._desc = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._desc_count, self._desc_tiles = len(descs), tiles

    def _refresh_shadows(self) -> None:
        ops.shadow_t_multi(self._desc, self._desc_count, self._desc_tiles, L.BF16)
        sh = self.model._shadows
        for key, w, gamma, w_sh, wt_sh in self._shadow_entries:
            sh.adopt(key, w, gamma, L.BF16, w_sh, wt_sh)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        self._steps += 1
        self.sink.attach_grads()
        with torch.cuda.device(self.sink.buckets[0].flat.device):
            for b in self.sink.buckets:
                ops.adamw(b.flat_p, b.flat, b.flat_m, b.flat_v, float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                          float(g["weight_decay"]), self._steps, 1.0, shadow=b.flat_shadow)
            _bump_versions([p for b in self.sink.buckets for p in b.params])
            if self._bf16:
                self._refresh_shadows()
        self.sink.consumed = True
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        """One memset per bucket.  ``set_to_none`` is accepted for interface compatibility; gradients stay views of
        the flat buffers (that is what the backward kernels and the all-reduce write / read)."""
        self.sink.zero()
        self.sink.attach_grads()

    # ---- state (torch.optim layout: state[p] = {"step", "exp_avg", "exp_avg_sq"}) -----------------------------
    def _publish_state(self) -> None:
        for b in self.sink.buckets:
            for p, off in zip(b.params, b.offsets):
                self.state[p] = {"step": torch.tensor(float(self._steps)),
                                 "exp_avg": b.flat_m[off:off + p.numel()].view_as(p),
                                 "exp_avg_sq": b.flat_v[off:off + p.numel()].view_as(p)}

    def state_dict(self):
        self._publish_state()
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for b in self.sink.buckets:
                for p, off in zip(b.params, b.offsets):
                    st = self.state.get(p)
                    if not st:
                        continue
                    b.flat_m[off:off + p.numel()].view_as(p).copy_(st["exp_avg"])
                    b.flat_v[off:off + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                    self._steps = max(self._steps, int(float(st["step"])))
        self._publish_state()

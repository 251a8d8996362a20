"""Loss and running metrics on the device.

The reference loop computes ``CrossEntropyLoss(weight, label_smoothing)`` (train.py:167-170,225) and then
synchronises with the host three times per step to keep running sums and per-sample probabilities
(train.py:229-235; ``evaluate`` :77-105 does the same per validation batch).  Here one launch per step
(``tvit_ce_loss``) produces the loss, its gradient w.r.t. the logits and -- on the device -- the running
``loss * B`` / correct / count sums plus the positive-class probabilities and labels the epoch-end AUC needs.
``DeviceMetrics.compute()`` is the only host synchronisation: once per epoch.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


def roc_auc(labels: np.ndarray, scores: np.ndarray) -> float:
    """Area under the ROC curve by the rank-sum (Mann-Whitney) statistic with average ranks for ties; equals
    sklearn.metrics.roc_auc_score for binary labels (train.py:7,100,239).  NaN when only one class is present."""
    labels = np.asarray(labels).astype(bool)
    scores = np.asarray(scores, dtype=np.float64)
    n_pos, n_neg = int(labels.sum()), int((~labels).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = np.argsort(scores, kind="mergesort")
    s = scores[order]
    ranks = np.empty(len(s), dtype=np.float64)
    bounds = np.flatnonzero(np.r_[True, s[1:] != s[:-1], True])        # tie groups share their average rank
    for a, b in zip(bounds[:-1], bounds[1:]):
        ranks[a:b] = 0.5 * (a + b - 1) + 1.0
    r = np.empty_like(ranks)
    r[order] = ranks
    return float((r[labels].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


class DeviceMetrics:
    """Running loss / accuracy / AUC inputs accumulated on the GPU by the loss kernel."""

    def __init__(self, device, capacity: int = 4096):
        self.device = torch.device(device)
        self.acc = torch.zeros(3, dtype=torch.float32, device=self.device)      # loss * B, correct, count
        self.probs = torch.empty(capacity, dtype=torch.float32, device=self.device)
        self.labels = torch.empty(capacity, dtype=torch.float32, device=self.device)
        self.count = 0                                                            # host-side mirror of acc[2]

    def reset(self) -> None:
        self.acc.zero_()
        self.count = 0

    def _reserve(self, n: int):
        need = self.count + n
        if need > self.probs.numel():
            cap = max(need, 2 * self.probs.numel())
            for name in ("probs", "labels"):
                new = torch.empty(cap, dtype=torch.float32, device=self.device)
                new[:self.count].copy_(getattr(self, name)[:self.count])
                setattr(self, name, new)
        off = self.count
        self.count = need
        return self.probs[off:need], self.labels[off:need]

    @torch.no_grad()
    def update(self, logits: torch.Tensor, labels: torch.Tensor, class_weight: Optional[torch.Tensor] = None,
               label_smoothing: float = 0.0) -> None:
        """Evaluation-side update (no gradient): the per-batch body of ``evaluate`` (train.py:88-99)."""
        p, l = self._reserve(logits.shape[0])
        ops.ce_loss(logits.detach().float().contiguous(), labels.contiguous(), class_weight, label_smoothing, None,
                    None, self.acc, p, l)

    def compute(self) -> Dict[str, float]:
        """ONE device->host transfer: {"loss", "acc", "auc", "count"} exactly as train.py:237-241 / :100-105."""
        n = self.count
        host = torch.cat([self.acc, self.probs[:n], self.labels[:n]]).cpu().numpy()
        loss_sum, correct, total = (float(v) for v in host[:3])
        probs, labels = host[3:3 + n], host[3 + n:3 + 2 * n]
        return {"loss": loss_sum / max(total, 1.0), "acc": correct / max(total, 1.0),
                "auc": roc_auc(labels > 0.5, probs), "count": int(total)}


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, label_smoothing, metrics):
        logits = logits.float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        macc = p = l = None
        if metrics is not None:
            p, l = metrics._reserve(logits.shape[0])
            macc = metrics.acc
        ops.ce_loss(logits, labels.contiguous(), weight, label_smoothing, loss, dlogits, macc, p, l)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None, None


class CrossEntropyLoss(nn.Module):
    """``torch.nn.CrossEntropyLoss(weight=..., label_smoothing=...)`` (mean reduction) as one fused
    forward+backward launch; pass ``metrics=DeviceMetrics(...)`` to also accumulate the loop's running metrics."""

    def __init__(self, weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                 metrics: Optional[DeviceMetrics] = None):
        super().__init__()
        self.register_buffer("weight", None if weight is None else weight.detach().float().contiguous())
        self.label_smoothing = float(label_smoothing)
        self.metrics = metrics

    def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if not logits.is_cuda:
            raise RuntimeError("neural_vit_b200.CrossEntropyLoss runs on CUDA tensors only (no CPU fallback)")
        with torch.cuda.device(logits.device):
            if torch.is_grad_enabled() and logits.requires_grad:
                return _CEFn.apply(logits, labels, self.weight, self.label_smoothing, self.metrics)
            loss = torch.empty((), dtype=torch.float32, device=logits.device)
            macc = p = l = None
            if self.metrics is not None:
                p, l = self.metrics._reserve(logits.shape[0])
                macc = self.metrics.acc
            ops.ce_loss(logits.detach().float().contiguous(), labels.contiguous(), self.weight, self.label_smoothing,
                        loss, None, macc, p, l)
            return loss

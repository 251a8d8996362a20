"""Checkpoint I/O in the reference's layout.

The reference only ever *writes* checkpoints -- ``{"model_state": model.state_dict(), "config": asdict(model.config)}``
(train.py:265-275, 290-295) -- and never loads one.  ``save_checkpoint`` writes that exact layout from the drop-in
model and ``load_checkpoint`` restores a model (drop-in or written by the reference itself) from it, so training can
resume and reference checkpoints can be evaluated on the B200 path.
"""
from __future__ import annotations

from dataclasses import asdict
from typing import Optional

import torch

from .model import Temporal3DViT, Temporal3DViTConfig


def save_checkpoint(model: Temporal3DViT, path: str) -> None:
    torch.save({"model_state": model.state_dict(), "config": asdict(model.config)}, path)


def load_checkpoint(path: str, device: Optional[str] = None, precision: Optional[str] = None) -> Temporal3DViT:
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(ckpt, dict) or "model_state" not in ckpt or "config" not in ckpt:
        raise ValueError(f"{path} is not a Temporal 3D ViT checkpoint ({{'model_state', 'config'}} expected)")
    known = Temporal3DViTConfig.__dataclass_fields__.keys()
    unknown = set(ckpt["config"]) - set(known)
    if unknown:
        raise ValueError(f"checkpoint config has unknown fields {sorted(unknown)}")
    model = Temporal3DViT(Temporal3DViTConfig(**ckpt["config"]), precision=precision)
    model.load_state_dict(ckpt["model_state"], strict=True)
    if device is not None:
        model.to(device)
    return model

"""neural_vit_b200 -- B200-native (sm_100a) implementation of the Temporal 3D ViT training hot path.

Public names mirror the reference's ``temporal_vit.models.model``.
"""
from .model import CONFIGS, Temporal3DViT, Temporal3DViTConfig  # noqa: F401

__all__ = ["CONFIGS", "Temporal3DViT", "Temporal3DViTConfig"]

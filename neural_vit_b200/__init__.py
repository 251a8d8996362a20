"""neural_vit_b200 -- B200-native (sm_100a) implementation of the Temporal 3D ViT training hot path.

Public names mirror the reference's ``temporal_vit.models.model``; the pieces around the path (SURVEY.md section 8
rows e and f) are exported next to them.
"""
from .model import CONFIGS, Temporal3DViT, Temporal3DViTConfig  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .loss import CrossEntropyLoss, DeviceMetrics, roc_auc  # noqa: F401
from .data import DevicePrefetcher, RankShardSampler  # noqa: F401
from .checkpoint import load_checkpoint, save_checkpoint  # noqa: F401
from .ddp import BucketedAllReduce  # noqa: F401

__all__ = ["CONFIGS", "Temporal3DViT", "Temporal3DViTConfig", "FusedAdamW", "CrossEntropyLoss", "DeviceMetrics",
           "roc_auc", "DevicePrefetcher", "RankShardSampler", "load_checkpoint", "save_checkpoint",
           "BucketedAllReduce"]

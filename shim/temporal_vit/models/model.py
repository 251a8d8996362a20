"""Import-path shim: put `<repo>/shim` ahead of the reference on PYTHONPATH and the reference's
own training loop (`from temporal_vit.models.model import CONFIGS, Temporal3DViT, Temporal3DViTConfig`,
train.py:13, train_hptune.py:29) picks up the B200-native model with no source change.
See INTEGRATION.md."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from neural_vit_b200.model import CONFIGS, Temporal3DViT, Temporal3DViTConfig  # noqa: E402,F401

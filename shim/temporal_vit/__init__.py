"""Import-path shim.

This directory is put AHEAD of the reference on ``sys.path``.  ``pkgutil.extend_path`` makes the package span both
locations, so every reference sub-package (``temporal_vit.training``, ``temporal_vit.data`` ...) still resolves to
the reference's own files and only ``temporal_vit.models.model`` -- which exists here as well and is found first --
is replaced by the B200-native implementation.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

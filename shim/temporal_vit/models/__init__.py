"""Spans the shim's and the reference's ``temporal_vit/models`` directories; ``model.py`` of the shim wins."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

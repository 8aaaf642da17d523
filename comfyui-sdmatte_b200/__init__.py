"""B200-native drop-in for the `Apply SDMatte` ComfyUI node (reference: /root/reference/__init__.py:1-6)."""
from . import engine  # noqa: F401

try:  # the node module needs torch/torchvision only; ComfyUI modules are optional (stubbed when absent)
    from .sdmatte_nodes import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS
except Exception as _e:  # pragma: no cover - surfaced when the node is actually used
    NODE_CLASS_MAPPINGS = {}
    NODE_DISPLAY_NAME_MAPPINGS = {}
    _NODE_IMPORT_ERROR = _e

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS", "engine"]

"""Minimal `mmcv` namespace exposing only `mmcv.ops` (the native-operator part STMask's hot path uses)."""
from . import ops  # noqa: F401

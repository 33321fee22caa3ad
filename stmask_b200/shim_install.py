"""Make the reference's three operator imports resolve to this library WITHOUT shadowing anything else:

    import stmask_b200; stmask_b200.install_shims()
    from dcn_v2 import DCN                               # backbone.py:5
    from mmcv.ops import DeformConv2d, roi_align         # Featurealign.py:3, track_to_segment_head.py:6
    from spatial_correlation_sampler import spatial_correlation_sample   # track_to_segment_head.py:4

`mmcv` is special: the reference also uses its image / io / runner / parallel helpers, so a real mmcv (when
one is installed) stays in charge and only the hot-path names of `mmcv.ops` are replaced; without one a
minimal `mmcv` namespace with just `mmcv.ops` is registered."""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import sys
import types

HOT_PATH_OPS = ("DeformConv2d", "DeformConv2dPack", "ModulatedDeformConv2d", "ModulatedDeformConv2dPack",
                "deform_conv2d", "modulated_deform_conv2d", "roi_align")


def overlay_mmcv_ops(mmcv_module: types.ModuleType, real_first: bool) -> types.ModuleType:
    """Return the module to serve as `mmcv.ops`: the real one (if it imports — mmcv-lite has none, mmcv-full's
    needs its compiled extension) with the hot-path operators overridden, else this library's operators alone."""
    from .compat import mmcv_ops as ours
    target = None
    if real_first:
        saved = sys.modules.pop("mmcv.ops", None)
        try:
            search = [p for p in getattr(mmcv_module, "__path__", [])]
            spec = importlib.machinery.PathFinder.find_spec("ops", search[:-1] if len(search) > 1 else search)
            if spec is not None and spec.loader is not None:
                real = importlib.util.module_from_spec(spec)
                real.__name__ = "mmcv.ops"
                sys.modules["mmcv.ops"] = real
                spec.loader.exec_module(real)
                target = real
        except Exception:              # real mmcv.ops not importable (no compiled _ext): serve ours alone
            target = None
            if saved is not None:
                sys.modules["mmcv.ops"] = saved
    if target is None:
        target = types.ModuleType("mmcv.ops")
        target.__doc__ = ours.__doc__
        target.__all__ = list(ours.__all__)
    for name in HOT_PATH_OPS:
        setattr(target, name, getattr(ours, name))
    target.__stmask_b200__ = True
    sys.modules["mmcv.ops"] = target
    mmcv_module.ops = target
    return target


def install_shims() -> None:
    """Register `dcn_v2`, `spatial_correlation_sampler` and the hot-path part of `mmcv.ops` in `sys.modules`."""
    from .compat import dcn_v2, spatial_correlation_sampler
    sys.modules["dcn_v2"] = dcn_v2
    sys.modules["spatial_correlation_sampler"] = spatial_correlation_sampler
    mm = sys.modules.get("mmcv")
    real = False
    if mm is None:
        try:
            mm = importlib.import_module("mmcv")
            real = not getattr(getattr(mm, "ops", None), "__stmask_b200__", False)
        except ImportError:
            mm = types.ModuleType("mmcv")
            mm.__path__ = []
            sys.modules["mmcv"] = mm
    else:
        real = bool(getattr(mm, "__path__", None))
    if getattr(getattr(mm, "ops", None), "__stmask_b200__", False):
        return
    overlay_mmcv_ops(mm, real_first=real)

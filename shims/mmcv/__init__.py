"""`import mmcv` with <repo>/shims ahead on PYTHONPATH.

This directory must NOT shadow a real mmcv: the reference uses far more of it than the three native
operators on the hot path (`mmcv.imread`, `mmcv.dump`, `mmcv.is_str`, `mmcv.parallel.DataContainer`,
`mmcv.runner.*`, `mmcv.imresize` ... in datasets/, layers/eval_utils.py, layers/box_utils.py).  So this
module FALLS THROUGH: if another `mmcv` package is importable further down `sys.path`, it is executed
under this module's name (its `__path__` first, so `mmcv.runner`, `mmcv.parallel`, ... resolve to the real
package) and only the hot-path operators of `mmcv.ops` are replaced.  Without a real mmcv this is a
namespace holding `mmcv.ops` alone.  `stmask_b200.install_shims()` does the same without PYTHONPATH.
"""
import importlib.machinery
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _real_spec():
    for entry in sys.path:
        root = os.path.abspath(entry or os.getcwd())
        if root == os.path.dirname(_here):
            continue
        spec = importlib.machinery.PathFinder.find_spec("mmcv", [root])
        if spec is not None and spec.origin and os.path.dirname(os.path.abspath(spec.origin)) != _here \
                and spec.submodule_search_locations:
            return spec
    return None


_spec = _real_spec()
if _spec is not None:
    # become the real package: its search path first (sub-packages resolve there), ours last (for .ops)
    __path__ = list(_spec.submodule_search_locations) + [_here]     # noqa: F811
    __file__ = _spec.origin
    with open(_spec.origin, "rb") as _f:
        exec(compile(_f.read(), _spec.origin, "exec"), globals())

from stmask_b200.shim_install import overlay_mmcv_ops as _overlay  # noqa: E402

ops = _overlay(sys.modules[__name__], real_first=_spec is not None)

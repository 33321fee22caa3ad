"""CPU oracle for the STMask hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the CPU
baseline.  Nothing under ``stmask_b200/`` imports it (tests/test_boundary.py checks).

Parity pin status: *unpinned by the reference's own tests* (it has none).
The C restatement in ``stm_oracle.c`` is pinned against torchvision's CPU
``deform_conv2d`` and against the reference's Python call sites run here
(``oracle/make_golden.py`` -> ``tests/golden/``).
"""
from .oracle import (  # noqa: F401
    build,
    correlate,
    correlation,
    deform_conv2d,
    fcb_ali_offsets,
    feature_align,
    lib_path,
    np_correlation,
    np_deform_conv2d,
    num_threads,
    out_size,
    roi_align,
)

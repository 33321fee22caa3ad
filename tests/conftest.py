import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SHIMS = os.path.join(ROOT, "shims")       # `import dcn_v2`, `import mmcv.ops`, `import spatial_correlation_sampler`
if SHIMS not in sys.path:
    sys.path.insert(0, SHIMS)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_meta(z, key="meta"):
    return json.loads(bytes(z[key]).decode())


def rel_err(a, b):
    """max |a-b| / max |b| — the 'relative error on features' of BASELINE.json's north star."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")

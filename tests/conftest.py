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
    # torch-side fp32 reference math in the tests (F.conv2d, matmul) must be real fp32, whatever the test order
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def detections_after_fast_nms(logits, boxes, centerness, conf_thresh=0.05, nms_thresh=0.5, top_k=200):
    """Restatement (test infrastructure) of what the reference does with a head's class logits at eval:
    softmax, candidate filter `max_c>0 conf > eval_conf_thresh` (layers/functions/TF_utils.py:68-74), then
    Detect_TF.cc_fast_nms (layers/functions/detection_TF.py:85-134): score = max non-background class prob
    x centerness, sort descending, top_k, IoU = triu(jaccard, 1), keep a box iff no higher-scoring box overlaps
    it by more than nms_thresh; classes += 1.  One prior per pixel.  Returns (prior index, class, score)."""
    import torch
    logits = torch.as_tensor(logits, dtype=torch.float32)
    boxes = torch.as_tensor(boxes, dtype=torch.float32)
    centerness = torch.as_tensor(centerness, dtype=torch.float32)
    c, h, w = logits.shape[-3:]
    conf = torch.softmax(logits.reshape(c, h * w).t(), -1)                # [priors, classes]
    conf_t = conf.t().contiguous()
    keep = torch.max(conf_t[1:, :], dim=0)[0] > conf_thresh
    prior = torch.nonzero(keep).view(-1)
    scores, classes = conf_t[1:, keep].max(dim=0)
    scores = scores * centerness[keep]
    b = boxes[keep]
    _, idx = scores.sort(0, descending=True)
    idx = idx[:top_k]
    bi = b[idx]
    lt = torch.max(bi[:, None, :2], bi[None, :, :2])
    rb = torch.min(bi[:, None, 2:], bi[None, :, 2:])
    inter = (rb - lt).clamp(min=0).prod(-1)
    area = (bi[:, 2] - bi[:, 0]) * (bi[:, 3] - bi[:, 1])
    iou = torch.triu(inter / (area[:, None] + area[None, :] - inter), diagonal=1)
    out = idx[iou.max(dim=0)[0] <= nms_thresh]
    return prior[out].numpy(), (classes[out] + 1).numpy(), scores[out].numpy()

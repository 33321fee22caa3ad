"""On-device multi-GPU correctness (needs >= 2 GPUs; skipped on a 1-GPU box): frame-sharded temporal fusion with the
one-frame halos exchanged by NCCL send/recv must be BIT-IDENTICAL to the single-GPU result.  Runs bench.py's own
self-check (`multi_gpu_check`) under torch.distributed.run, i.e. the exact code path the scaling runs time."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["frame", "clip"])
def test_sharded_temporal_fusion_equals_single_gpu(mode):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3", "--clips", "6",
           "--frames-per-clip", "7", "--sharding", mode, "--no-e2e", "--no-cpu-baseline", "--no-extras", "--multi-gpu-check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    chk = line["multi_gpu_check"]
    assert chk["tf_concat_bit_identical"] is True and chk["mismatching_pairs"] == 0, chk
    assert chk["pairs_checked"] == 6 * 6
    assert chk["halos_per_rank"] == (6 if mode == "frame" else 0), chk          # 7 frames cut in two: one boundary per clip
    assert line["n_gpus"] == 2 and line["config"]["sharding"] == mode

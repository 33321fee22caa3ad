"""Host (launch) time vs device time of one hot-path step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stmask_b200 import sharding
from stmask_b200.hotpath import HotPath, HotPathConfig
dev = torch.device("cuda", 0)
hp = HotPath(HotPathConfig(), dev)
plan = sharding.make_plan(2, 36, 1, "clip")
inp = hp.make_inputs(72, dev)
for _ in range(3):
    hp(inp, plan, 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    hp(inp, plan, 0)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step {1e3 * (t1 - t0) / 10:.3f} ms; wall per step incl. drain {1e3 * (t2 - t0) / 10:.3f} ms")

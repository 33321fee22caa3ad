"""Per-phase globaltimer trace (ns) of the tcgen05 correlation kernel: first 4 tiles of every CTA.
    python tools/corr_trace.py [pairs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stmask_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 71
from stmask_b200 import sharding
x = torch.randn(n + 1, 256, 24, 40, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
t = torch.randn_like(x)
ri, ni = sharding.pair_index_tensors(sharding.make_plan(1, n + 1, 1), 0, x.device)
buf = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
run = lambda: ops.correlation_pairs(x, ri, ni, 11, 1, scale=1 / 256, relu=True, feats=t, feat_channel_offset=128)
run(); torch.cuda.synchronize()
os.environ["STM_DEBUG_BUF"] = hex(buf.data_ptr())
run(); torch.cuda.synchronize()
t = buf.view(148, 64).cpu()
t0 = int(t[:, 59].min())
ev = ["tma first", "tma last", "mma first full", "mma last full", "epi start", "phase0 done", "tmem_full seen", "phase1 done", "phase2 done"]
for cta in (0, 73, 147):
    print(f"CTA {cta}: body start +{int(t[cta, 59]) - t0} ns, teardown +{int(t[cta, 58]) - t0} ns")
    for k in range(4):
        row = [int(t[cta, k * 12 + e]) for e in range(9)]
        if row[4] == 0:
            continue
        print(f"   tile {k}: " + ", ".join(f"{ev[e]}={row[e] - t0}" for e in range(9)))
d = t.clone().float()
d[t == 0] = float("nan")
print("median over CTAs (ns from first body start):")
for k in range(4):
    print(f"   tile {k}: " + ", ".join(f"{ev[e]}={float(torch.nanmedian(d[:, k * 12 + e])) - t0:.0f}" for e in range(9)))
print(f"   teardown median {float(torch.nanmedian(d[:, 58])) - t0:.0f} max {float(d[:, 58].max()) - t0:.0f}")

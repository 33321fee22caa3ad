"""Per-phase clock64 trace of the tcgen05 correlation kernel (first tile of every CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stmask_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 71
fused = (sys.argv[2] != "plain") if len(sys.argv) > 2 else True
x1 = torch.randn(n, 256, 24, 40, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
x2, t1, t2 = torch.randn_like(x1), torch.randn_like(x1), torch.randn_like(x1)
buf = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
def run():
    if fused:
        return ops.correlation(x1, x2, 11, 1, scale=1 / 256, relu=True, feats=(t1, t2), channels_last=True)
    return ops.correlation(x1, x2, 11, 1, channels_last=False)
run(); torch.cuda.synchronize()
os.environ["STM_DEBUG_BUF"] = hex(buf.data_ptr())
run(); torch.cuda.synchronize()
t = buf.view(148, 32).cpu()
names = {1: "tma c0 issued", 2: "tma c1 issued", 3: "tma c2 issued", 4: "tma c3 issued", 5: "mma sees c0", 6: "mma sees c1", 7: "mma sees c2",
         8: "mma sees c3", 9: "mma committed tile", 10: "epi sees tmem_full", 11: "phase1 done", 12: "phase2/3 done (tile 0)", 13: "tile 1 done", 14: "teardown"}
for cta in (0, 73, 147):
    base = int(t[cta, 0])
    print(f"CTA {cta}:", ", ".join(f"{names[i]}={int(t[cta, i]) - base}" for i in range(1, 15) if int(t[cta, i]) > 0))
d = (t[:, 1:15] - t[:, :1]).float()
d[t[:, 1:15] == 0] = float("nan")
print("median over CTAs:", {names[i + 1]: float(torch.nanmedian(d[:, i])) for i in range(14)})

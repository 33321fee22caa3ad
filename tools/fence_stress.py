"""Stress check of the consumer-side proxy fence (STM_DCN_CFENCE=1, default) against the producer-side fence:
the full-size FCB 3x5 launch repeated many times must reproduce the producer-fence result bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stmask_b200 import ops
from stmask_b200.hotpath import fpn_level_sizes
torch.manual_seed(0)
dev = "cuda"
F = 72
spec = ops.ConvSpec(256, 256, (3, 5), 1, (1, 2))
w = (torch.randn(256, 256, 3, 5, device=dev) / (256 * 15) ** 0.5).bfloat16()
wp = ops.pack_weight(w, spec, torch.bfloat16)
lv = fpn_level_sizes()
xs = [torch.randn(F, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for h, ww in lv]
offs = [torch.randn(F, 30, h, ww, device=dev) for h, ww in lv]
os.environ["STM_DCN_CFENCE"] = "0"
ref = [t.clone() for t in ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True)]
os.environ["STM_DCN_CFENCE"] = "1"
bad = 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for i in range(n):
    out = ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True)
    if not all(torch.equal(a, b) for a, b in zip(out, ref)):
        bad += 1
torch.cuda.synchronize()
print(f"consumer-side fence: {n} launches, {bad} differ from the producer-side-fence result")
sys.exit(1 if bad else 0)

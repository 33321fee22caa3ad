import sys, torch
from stmask_b200 import ops, _lib as L
dev='cuda'
hint = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
spec = ops.ConvSpec(256, 32, 3, 1, 1)
x = torch.randn(B, 24, 40, 256, device=dev).bfloat16().permute(0,3,1,2)
w = (torch.randn(32,256,3,3, device=dev)/48).bfloat16()
wp = ops.pack_weight(w, spec, torch.bfloat16)
print(ops.deform_conv2d_variant([tuple(x.shape)], spec, torch.bfloat16, zero_offset=True, hint=hint))
y = ops.deform_conv2d_multi([x],[None],None,wp,None,spec,out_f32=True, hint=hint)[0]
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x.float(), w.float(), padding=1)
print((y-ref).abs().max())

"""GPU bring-up helper: tcgen05 kernels vs the (oracle-verified) SIMT kernels, one case per
subprocess so that a trap in one case does not take the others down.
    python tools/tc_debug.py dcn|corr [case_index]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DCN_CASES = [
    # B, Cin, Cout, H, W, kh, kw, stride, dg, mask, bias, zero_offsets
    (1, 64, 64, 8, 16, 1, 1, 1, 1, False, False, True),      # 1 k-block, pure GEMM check
    (1, 64, 64, 8, 16, 3, 3, 1, 1, False, False, True),
    (1, 64, 128, 8, 16, 3, 3, 1, 1, False, False, False),
    (2, 256, 256, 12, 20, 3, 3, 1, 1, False, False, False),
    (2, 256, 256, 12, 20, 3, 5, 1, 1, False, False, False),
    (2, 256, 256, 24, 40, 5, 3, 1, 4, False, False, False),
    (2, 128, 128, 24, 40, 3, 3, 2, 1, True, True, False),
    (1, 512, 512, 12, 20, 3, 3, 1, 1, True, True, False),
    (8, 256, 256, 48, 80, 3, 5, 1, 1, False, False, False),   # big: M_TILES = 2 path
    (3, 256, 48, 23, 40, 3, 3, 1, 1, False, True, False),     # odd N tile
]


def run_dcn(i):
    import torch
    from stmask_b200 import ops
    B, Cin, Cout, H, W, kh, kw, s, dg, um, ub, zero = DCN_CASES[i]
    torch.manual_seed(i)
    dev = "cuda"
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    spec = ops.ConvSpec(Cin, Cout, (kh, kw), s, pad, 1, 1, dg)
    Ho, Wo = spec.out_hw(H, W)
    x = torch.randn(B, Cin, H, W, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, kh, kw, device=dev) / (Cin * kh * kw) ** 0.5).bfloat16()
    off = torch.zeros(B, dg * 2 * kh * kw, Ho, Wo, device=dev) if zero else torch.randn(B, dg * 2 * kh * kw, Ho, Wo, device=dev) * 2
    mask = torch.rand(B, dg * kh * kw, Ho, Wo, device=dev) if um else None
    bias = torch.randn(Cout, device=dev) if ub else None
    ref = ops.deform_conv2d(x, off, w, bias, mask, s, pad, 1, 1, dg, backend="simt").float()
    torch.cuda.synchronize()
    got = ops.deform_conv2d(x, off, w, bias, mask, s, pad, 1, 1, dg, backend="tcgen05").float()
    torch.cuda.synchronize()
    err = (got - ref).abs()
    rel = float(err.max() / ref.abs().max())
    res = {"case": DCN_CASES[i], "rel_err": rel, "ok": rel < 1e-2}
    if rel >= 1e-2:
        bad = (err > 0.02 * ref.abs().max())
        res["bad_frac"] = float(bad.float().mean())
        idx = bad.nonzero()[:6].tolist()
        res["first_bad(b,c,h,w)"] = idx
        res["got/ref"] = [(float(got[tuple(j)]), float(ref[tuple(j)])) for j in idx[:4]]
        res["bad_per_channel_head"] = bad.float().mean(dim=(0, 2, 3))[:16].tolist()
        res["bad_per_row_head"] = bad.float().mean(dim=(0, 1, 3))[:8].tolist()
    print(json.dumps(res))


CORR_CASES = [((2, 256, 24, 40), 11, 1), ((1, 256, 48, 80), 11, 1), ((2, 256, 24, 40), 11, 2), ((3, 256, 3, 5), 11, 1),
              ((2, 256, 6, 10), 11, 2), ((2, 256, 23, 40), 11, 1), ((2, 64, 12, 20), 5, 1), ((1, 512, 12, 20), 11, 1),
              ((8, 256, 48, 80), 11, 1)]


def run_corr(i):
    import torch
    from stmask_b200 import ops
    shape, P, d = CORR_CASES[i]
    torch.manual_seed(i)
    x1 = torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    x2 = torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    res = {"case": CORR_CASES[i]}
    for cl in (False, True):
        ref = ops.correlation(x1, x2, P, d, scale=1 / shape[1], leaky_slope=0.1, backend="simt", out_dtype=torch.float32, channels_last=cl)
        torch.cuda.synchronize()
        got = ops.correlation(x1, x2, P, d, scale=1 / shape[1], leaky_slope=0.1, backend="tcgen05", out_dtype=torch.float32, channels_last=cl)
        torch.cuda.synchronize()
        err = (got - ref).abs()
        rel = float(err.max() / ref.abs().max())
        res[f"rel_err_cl{int(cl)}"] = rel
        if rel >= 1e-2:
            bad = err > 0.02 * ref.abs().max()
            res["bad_frac"] = float(bad.float().mean())
            res["first_bad(b,k,y,x)"] = bad.nonzero()[:8].tolist()
            res["bad_per_k_head"] = bad.float().mean(dim=(0, 2, 3))[:24].tolist()
    print(json.dumps(res))


if __name__ == "__main__":
    kind = sys.argv[1]
    cases = DCN_CASES if kind == "dcn" else CORR_CASES
    if len(sys.argv) > 2:
        (run_dcn if kind == "dcn" else run_corr)(int(sys.argv[2]))
    else:
        for i in range(len(cases)):
            try:
                r = subprocess.run([sys.executable, __file__, kind, str(i)], capture_output=True, text=True, timeout=120)
                out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                print(f"[{kind} {i}] rc={r.returncode} {out}")
                if r.returncode != 0:
                    print("   stderr:", r.stderr.strip().splitlines()[-3:])
            except subprocess.TimeoutExpired:
                print(f"[{kind} {i}] TIMEOUT {cases[i]}")

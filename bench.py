#!/usr/bin/env python
"""bench.py — frames/sec of the STMask R101-DCN-FPN FCA+FCB(ada)+TF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path (stmask_b200/hotpath.py: 11 backbone DCNv2 layers, FCB over P3..P7 for
the 3 anchor kernels, temporal-fusion correlation+concat) over this rank's frames.  Workload: BASELINE.json
configs[3] — 36-frame 360x640 clips (padded 384x640), bf16 — two clips (72 frames) per GPU, weak scaling;
for N > 1 every clip is cut frame-wise across the ranks, so each rank exchanges one-frame feature halos
(NCCL send/recv) inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS_PER_GPU = 2
FRAMES_PER_CLIP = 36
METRIC = "frames/sec @360x640 R101 FCA+FCB+TF"


def _traffic(key, count_key, count):
    """Measured DRAM bytes per launch of the dominant kernels (one ncu --set full capture each, profiles/)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[key]
        return t["bytes"] if t[count_key] == count else None
    except (OSError, KeyError, ValueError):
        return None


def _time_launches(fn, n):
    """Average device time of one launch: n launches queued back to back between ONE pair of CUDA events on the
    launching (current) stream, so host launch latency is not part of a sub-100-us kernel's figure.  The
    operands of consecutive launches (> 126 MB read + written) do not fit L2."""
    import torch
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every few ms (the timed
    region is tens of ms long; nvidia-smi's loop is too coarse for it).  Falls back to `nvidia-smi -lms`."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.stop = threading.Event()
        self.thread = self.proc = None
        self.src = "nvml"

    def _nvml_loop(self, nv, h):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", None) or nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", None) or nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", None) or nv.nvmlClocksThrottleReasonSwPowerCap))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = get_reasons(h)
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            self.stop.wait(self.period)

    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1]))
                self.max_mhz = float(r[2])
                for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id
            dom = torch.cuda.get_device_properties(self.index).pci_domain_id
            dev_id = torch.cuda.get_device_properties(self.index).pci_device_id
            h = nv.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev_id:02x}.0")
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.src = "nvidia-smi"
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                              "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._smi_loop, daemon=True)
                self.thread.start()
            except OSError:
                self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        out = {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.src}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


# ----------------------------------------------------------------------------------------------
# CPU leg: the reference's CPU path for the same operators on the box's host cores
# ----------------------------------------------------------------------------------------------
def cpu_hot_path_frame(hp_cfg, seed=0):
    """Build the closure that runs ONE frame (+ one TF pair) of the hot path on the CPU: DCN through
    torchvision.ops.deform_conv2d (the CPU deformable conv the north star names; mmcv-full 1.1.2 / dcn_v2
    are not installable), correlation through the oracle's OpenMP C port, fp32, all host threads."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from torchvision.ops import deform_conv2d as tv_dcn

    import oracle
    from stmask_b200 import backbone_dcn
    from stmask_b200.hotpath import CORR_LEVEL, FPN_CHANNELS, HEAD_KERNELS, fpn_level_sizes

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    shapes = backbone_dcn.dcn_layer_shapes(*hp_cfg.resnet_args, hp_cfg.height, hp_cfg.width)
    levels = fpn_level_sizes(hp_cfg.height, hp_cfg.width)
    dcn = []
    for s in shapes:
        c = s.channels
        dcn.append((torch.randn(1, c, s.in_h, s.in_w, generator=g), torch.randn(c, c, 3, 3, generator=g) / (9 * c) ** 0.5,
                    torch.randn(c, generator=g) * 0.1, torch.randn(27, c, 3, 3, generator=g) * 0.02, torch.randn(27, generator=g) * 0.5, s.stride))
    fcb = []
    for kh, kw in HEAD_KERNELS:
        w = torch.randn(FPN_CHANNELS, FPN_CHANNELS, kh, kw, generator=g) / (FPN_CHANNELS * kh * kw) ** 0.5
        wo = torch.randn(2 * kh * kw, 4, 1, 1, generator=g) * 0.5
        fcb.append((w, wo, ((kh - 1) // 2, (kw - 1) // 2)))
    xs = [torch.randn(1, FPN_CHANNELS, h, w, generator=g) for h, w in levels]
    boxes = [torch.randn(1, 4, h, w, generator=g) for h, w in levels]
    h, w = levels[CORR_LEVEL]
    f1, f2 = (np.random.default_rng(seed).standard_normal((1, FPN_CHANNELS, h, w)).astype(np.float32) for _ in range(2))
    t1, t2 = torch.randn(1, FPN_CHANNELS, h, w, generator=g), torch.randn(1, FPN_CHANNELS, h, w, generator=g)

    def run():
        with torch.no_grad():
            for x, wt, b, wc, bc, st in dcn:
                out = F.conv2d(x, wc, bc, stride=st, padding=1)
                tv_dcn(x, out[:, :18], wt, b, (st, st), (1, 1), (1, 1), torch.sigmoid(out[:, 18:]))
            if hp_cfg.fcb:
                for wt, wo, pad in fcb:
                    for x, bx in zip(xs, boxes):
                        torch.relu_(tv_dcn(x, F.conv2d(bx, wo), wt, None, (1, 1), pad))
            if hp_cfg.temporal_fusion:
                corr = oracle.correlate(f1, f2, 11, 1, accum64=False)
                torch.relu_(torch.cat([torch.from_numpy(corr), t1, t2], 1))
    return run


def cpu_baseline(hp_cfg, budget_s=20.0):
    run = cpu_hot_path_frame(hp_cfg)
    run()                                   # warm-up (thread pools, page faults)
    t0 = time.perf_counter()
    n = 0
    while True:
        run()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 8:
            break
    return {"value": n / dt, "unit": "frames/sec", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{n} frame(s) (+1 TF pair each) of the same hot path, fp32: DCN = torchvision.ops.deform_conv2d CPU "
                      f"(the reference-side CPU deformable conv named by the north star), correlation = oracle C port (OpenMP); "
                      f"{dt:.1f} s of CPU work"}


def run_reference(args, hp_cfg, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run = cpu_hot_path_frame(hp_cfg)
    for _ in range(max(args.warmup, 1)):
        run()
    steps = max(args.steps, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    v = steps / dt
    cores = os.cpu_count() or 1
    sample = "each step = 1 frame (+1 TF pair) of the hot path on the host cores, fp32: torchvision.ops.deform_conv2d CPU + oracle C correlation (OpenMP)"
    print(file=real_stdout, flush=True, *[json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/sec", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{hp_cfg.name()} hot path, 360x640 (padded 384x640), CPU reference path"},
        "cpu_baseline": {"value": v, "unit": "frames/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })])


# ----------------------------------------------------------------------------------------------
def _claim_stdout():
    """Libraries (NCCL prints its version banner) may write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backbone", default="r101", choices=["r50", "r101"])
    ap.add_argument("--fcb", default="ada", choices=["ada", "ali", "none"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--backend", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--clips-per-gpu", type=int, default=CLIPS_PER_GPU)
    ap.add_argument("--frames-per-clip", type=int, default=FRAMES_PER_CLIP)
    ap.add_argument("--sharding", default="auto", choices=["auto", "clip", "frame"],
                    help="auto = frame-wise cuts (halo exchange in the timed region) for N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=12, help="frames per chunk of the streamed end-to-end path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import torch
    from stmask_b200.hotpath import HotPath, HotPathConfig

    hp_cfg = HotPathConfig(backbone=args.backbone, fcb=None if args.fcb == "none" else args.fcb,
                           dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32, backend=args.backend)
    if args.impl == "reference":
        run_reference(args, hp_cfg, real_stdout)
        return

    import torch.distributed as dist
    from stmask_b200 import _lib, ops, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if _lib.lib().stm_device_supported(local_rank) != 1:
        raise SystemExit("stmask_b200 needs an sm_100 (B200) device; there is no fallback")

    n_clips = args.clips_per_gpu * world
    mode = args.sharding if args.sharding != "auto" else ("frame" if world > 1 else "clip")
    plan = sharding.make_plan(n_clips, args.frames_per_clip, world, mode)
    n_local = plan.local_frames(rank)
    total_frames = n_clips * args.frames_per_clip
    hp = HotPath(hp_cfg, dev, seed=0)
    inp = hp.make_inputs(n_local, dev, seed=rank)
    in_bytes = sum(t.numel() * t.element_size() for t in inp.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    full_plan = plan if world > 1 else sharding.make_plan(args.clips_per_gpu, args.frames_per_clip, 1, "clip")

    def step(x):
        o = hp(x, full_plan, rank)
        if isinstance(o.get("tf.concat"), list):        # one tensor per whole clip
            for i, t_ in enumerate(o.pop("tf.concat")):
                o[f"tf.concat{i}"] = t_
        return o

    for _ in range(args.warmup):
        out = step(inp)
    out_bytes = sum(t.numel() * t.element_size() for t in out.values())
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()              # a cyclic-GC pause inside the timed region would stall this rank and, through the halos, all others
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        marks = []
        for _ in range(args.steps):
            out = step(inp)
            if os.environ.get("STM_BENCH_TRACE"):
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        e1.record()
        barrier()
    gc.enable()
    ms = e0.elapsed_time(e1)
    if marks:
        print(f"rank {rank} per-step ms:", " ".join(f"{a.elapsed_time(b):.2f}" for a, b in zip([e0] + marks[:-1], marks)), file=sys.stderr)
    launches = _lib.launch_count() - n0
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = total_frames * args.steps / (ms / 1e3)

    # ---------------- per-kernel rooflines (dominant kernel: the FCB deformable conv; plus correlation) ----------
    peaks = _peaks()
    roof = corr_roof = None
    reps = max(5, args.steps)
    if hp_cfg.fcb:
        m = hp.fcb[1]                                   # 3x5 kernel, all five levels, one launch
        xs = [inp[f"fcb.x{l}"] for l in range(5)]
        offs = [m.offsets(inp[f"fcb.box{l}.1"]) for l in range(5)]
        spec = m.conv_adaption.spec()
        wp = m.conv_adaption._cache.weight(m.conv_adaption.weight, spec, xs[0].dtype)
        outs = ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True, backend=args.backend)
        k_ms = _time_launches(lambda: ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True, backend=args.backend, outs=outs), reps)
        px = sum(h * w for h, w in hp.level_sizes)
        flops = 2.0 * n_local * px * 256 * 256 * 15
        ach = flops / (k_ms / 1e3) / 1e12
        be = ops.deform_conv2d_backend(tuple(xs[0].shape), spec, xs[0].dtype, args.backend)
        roof = {"kernel": f"deform_conv2d[{be}] FCB 3x5 256->256, P3..P7, {n_local} frames, one launch", "bound": "tensor",
                "achieved": ach, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_burst"],
                "traffic": _traffic("dcn_fcb35", "frames", n_local), "ms_per_launch": k_ms, "flops_per_launch": flops,
                "peak_source": peaks["src"] + ", burst (kernel timed alone)"}
    if hp_cfg.temporal_fusion:
        one_clip = sharding.make_plan(1, n_local, 1)           # n_local - 1 pairs, read in place through index arrays
        fr = inp["tf.fpn"][1:]
        k_ms = _time_launches(lambda: hp._tf_pairs(inp["tf.fpn"], inp["tf.t2s"], one_clip, 0, None), 4 * reps)
        es = 2 if hp_cfg.dtype == torch.bfloat16 else 4
        npx = fr.shape[0] * fr.shape[2] * fr.shape[3]
        # H*W*(2C + P^2) bytes, plus the 2*Ct WRITTEN concat bytes because this kernel copies them;
        # the 2*Ct feature bytes it also has to READ and the 7 zero pad channels of the padded layout it writes are not counted
        nbytes = npx * (2 * 256 + 121 + 2 * 256) * es
        tr_bytes = _traffic("corr_fused", "pairs", int(fr.shape[0]))
        ach = nbytes / (k_ms / 1e3) / 1e9
        corr_roof = {"kernel": f"correlation+concat[{ops.correlation_backend(tuple(fr.shape), fr.dtype, 11, 1, args.backend)}] "
                               f"P=11 C=256 24x40, {fr.shape[0]} frame pairs, one launch", "bound": "hbm", "achieved": ach,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": tr_bytes, "ms_per_launch": k_ms,
                     "bytes_per_launch": nbytes,
                     "note": "pairs are read in place from the frame batch: a frame is x2 of one pair and x1 of the next but comes from "
                             "DRAM once, so the measured traffic is below the per-pair algorithmic bytes",
                     "peak_source": peaks["src"]}

    # ---------------- end to end: pinned host inputs -> device -> hot path -> host results -----------------------
    e2e = None
    if not args.no_e2e:
        from stmask_b200.hotpath import StreamedIO
        host_in = {k: v.cpu().pin_memory() for k, v in inp.items()}
        del inp, out
        torch.cuda.empty_cache()
        io = StreamedIO(dev, chunk_frames=args.e2e_chunk)
        # result buffers: shapes from one (untimed) device-resident step
        d_in = {k: v.to(dev) for k, v in host_in.items()}
        ref_out = dict(hp._frames_only({k: v for k, v in d_in.items() if not k.startswith("tf.")}))
        if hp_cfg.temporal_fusion:
            ref_out.update(hp._tf_only({k: v for k, v in d_in.items() if k.startswith("tf.")}, plan, rank, None))
        host_out = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in ref_out.items()}
        e2e_out_bytes = sum(t.numel() * t.element_size() for t in host_out.values())
        del d_in, ref_out
        torch.cuda.empty_cache()

        def e2e_step():
            hp.forward_streamed(host_in, host_out, io, plan, rank)

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": total_frames * e2e_steps / float(t_e.item()), "unit": "frames/sec",
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": e2e_out_bytes, "steps": e2e_steps,
               "ms_per_step": 1e3 * float(t_e.item()) / e2e_steps,
               "note": f"HotPath.forward_streamed: pinned HOST inputs -> device -> hot path -> pinned HOST results, every step; "
                       f"{args.e2e_chunk}-frame chunks, H2D / kernels / D2H overlapped on three streams; per rank per step bytes"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(hp_cfg)

    if rank == 0:
        fl = hp.flops_per_frame()
        line = {
            "metric": METRIC, "value": value, "unit": "frames/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if hp_cfg.dtype == torch.bfloat16 else "f32", "data": "synthetic",
            "config": {"workload": f"{hp_cfg.name()} hot path (BASELINE.json configs[3]): {args.clips_per_gpu} clips x "
                                   f"{args.frames_per_clip} frames per GPU, 360x640 padded to 384x640, random weights, non-zero offsets",
                       "frames_per_step": total_frames, "sharding": plan.mode, "halos_per_rank": len(plan.recv_halos(min(1, world - 1))),
                       "l2": f"inputs+outputs per step = {(in_bytes + out_bytes) / 1e6:.0f} MB per GPU > 126 MB L2 (no explicit flush needed)",
                       "dcn_gflop_per_frame": (fl["backbone_dcn"] + fl["fcb"]) / 1e9, "backend": args.backend},
            "roofline": roof, "roofline_correlation": corr_roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — frames/sec of the STMask R101-DCN-FPN FCA+FCB(ada)+TF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c5|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path (stmask_b200/hotpath.py: 11 backbone DCNv2 layers with their offset/mask
predictor, FCB over P3..P7 for the 3 anchor kernels, temporal-fusion correlation+concat) over ALL frames of the
workload.  Default workload = BASELINE.json configs[4], the configuration the metric's 1/2/4/8-GPU curve is
quoted on: 64 clips x 16 frames (1024 frames) of 360x640 (padded 384x640), bf16, STRONG scaling — the same
1024 frames are sharded over the N ranks (it fits one GPU: 31 GB).  For N > 1 every clip is cut frame-wise
across the ranks ("frame" sharding: 64 one-frame feature halos per rank at N = 8, exchanged by NCCL send/recv
inside the timed region); the same run also times whole-clip sharding (no halo) and reports it as
`clip_sharding`.  `--workload c3` is configs[3] (2 x 36-frame clips per GPU, weak scaling, the round-1 line).
Prints ONE JSON line on rank 0.
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

WORKLOADS = {
    # name: (clips, frames per clip, scaling, clips are per GPU?)
    "c5": (64, 16, "strong", False),      # BASELINE.json configs[4]
    "c3": (2, 36, "weak", True),          # BASELINE.json configs[3], two clips per GPU
}
METRIC = "frames/sec @360x640 R101 FCA+FCB+TF"


def _traffic(key, count_key, count):
    """Measured DRAM bytes per launch of the dominant kernels (one ncu --set full capture each, profiles/),
    scaled linearly to this run's unit count (DRAM traffic of both kernels is proportional to frames / pairs)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))[key]
            return t["bytes"] * count / t[count_key]
        except (OSError, KeyError, ValueError, ZeroDivisionError):
            continue
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


def _time_graph(fn, n):
    """Average device time of one small launch, replayed n times from a CUDA graph: the operator's Python / ctypes front end
    (allocations, descriptor structs, tensor-map encoding: tens of microseconds) is paid once at capture and the GPU never
    waits for the host — what a serving loop that captures its step would see.  Falls back to `_time_launches`."""
    import torch
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        # these launches move less than the 126 MB L2 holds: write a 256 MB buffer between replays so that every timed replay
        # starts with its operands in HBM, and bracket each replay with its own pair of events
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            flush.zero_()
            a.record()
            g.replay()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2], "cuda graph replay, L2 flushed before every replay, median"
    except Exception:                                        # capture refused (should not happen: the ABI is capturable)
        torch.cuda.synchronize()
        return _time_launches(fn, n), "back-to-back launches"


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


def _bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore the first-touch placement of the pinned host slabs it allocates) to the CPUs of
    the NUMA node the GPU hangs off.  Best effort; returns a description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa node {node} ({len(allowed)} cpus)"
        return f"numa node {node} (not in the allowed cpu set)"
    except Exception as e:       # noqa: BLE001 — purely advisory
        return f"unavailable ({type(e).__name__})"


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--backbone", default="r101", choices=["r50", "r101"])
    ap.add_argument("--fcb", default="ada", choices=["ada", "ali", "none"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--backend", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--clips", type=int, default=0, help="override the workload's clip count")
    ap.add_argument("--frames-per-clip", type=int, default=0)
    ap.add_argument("--sharding", default="auto", choices=["auto", "clip", "frame"],
                    help="auto = frame-wise cuts (halo exchange in the timed region) for N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained run, the per-layer table and the sweeps")
    ap.add_argument("--multi-gpu-check", action="store_true", help="run the on-device multi-GPU correctness check even with --no-extras")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="frames per chunk of the streamed end-to-end path")
    ap.add_argument("--sustain-s", type=float, default=2.5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import torch
    from stmask_b200.hotpath import HotPath, HotPathConfig
    torch.set_grad_enabled(False)            # inference benchmark: the operators are forward-only and say so loudly

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
    numa = _bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if _lib.lib().stm_device_supported(local_rank) != 1:
        raise SystemExit("stmask_b200 needs an sm_100 (B200) device; there is no fallback")

    wl_clips, wl_fpc, scaling, per_gpu = WORKLOADS[args.workload]
    n_clips = (args.clips or wl_clips) * (world if per_gpu else 1)
    fpc = args.frames_per_clip or wl_fpc
    mode = args.sharding if args.sharding != "auto" else ("frame" if world > 1 else "clip")
    plan = sharding.make_plan(n_clips, fpc, world, mode)
    n_local = plan.local_frames(rank)
    total_frames = n_clips * fpc
    hp = HotPath(hp_cfg, dev, seed=0)
    inp = hp.make_inputs(n_local, dev, seed=rank, on_device=True)
    in_bytes = sum(t.numel() * t.element_size() for t in inp.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, sample_clocks=True):
        """W untimed + K timed steps bracketed by barrier + synchronize; device time (CUDA events), max over ranks."""
        for _ in range(warmup):
            step_fn()
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import gc
        gc.collect()
        gc.disable()          # a cyclic-GC pause inside the timed region would stall this rank and, through the halos, all others
        with ClockSampler(local_rank) as clocks:
            barrier()
            e0.record()
            for _ in range(steps):
                step_fn()
            e1.record()
            barrier()
        gc.enable()
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()), _lib.launch_count() - n0, clocks.summary()

    out_holder = {}

    def step():
        out_holder["o"] = hp(inp, plan, rank)

    ms, launches, clocks = timed(step, args.steps, args.warmup)
    out = out_holder["o"]
    out_bytes = sum(t.numel() * t.element_size() for t in out.values())
    value = total_frames * args.steps / (ms / 1e3)

    extras = {}
    peaks = _peaks()
    if not args.no_extras:
        # ---- the other sharding of the same workload (whole clips per rank: no halo) ----
        if world > 1:
            other = "clip" if mode == "frame" else "frame"
            plan2 = sharding.make_plan(n_clips, fpc, world, other)
            if plan2.local_frames(rank) == n_local:
                ms2, _, _ = timed(lambda: out_holder.__setitem__("o", hp(inp, plan2, rank)), max(3, args.steps // 2), 2)
                extras[f"{other}_sharding"] = {"value": total_frames * max(3, args.steps // 2) / (ms2 / 1e3), "unit": "frames/sec",
                                               "ms_per_step": ms2 / max(3, args.steps // 2),
                                               "halos_per_rank": max(len(plan2.recv_halos(r)) for r in range(world))}
        # ---- sustained: the same step for >= sustain-s seconds, clocks and power sampled throughout ----
        n_sus = max(args.steps, int(args.sustain_s / max(ms / args.steps / 1e3, 1e-6)) + 1)
        ms_s, _, clk_s = timed(step, n_sus, 0)
        extras["sustained"] = {"value": total_frames * n_sus / (ms_s / 1e3), "unit": "frames/sec", "steps": n_sus,
                               "seconds": ms_s / 1e3, "ms_per_step": ms_s / n_sus, "clocks": clk_s}

    # ---------------- per-kernel rooflines (dominant kernel: the FCB deformable conv; plus correlation) ----------
    roof = corr_roof = None
    reps = max(5, min(args.steps, 20))
    fpn_shapes = None
    if hp_cfg.fcb:
        m = hp.fcb[1]                                   # 3x5 kernel, all five levels, ONE launch (offsets derived in the kernel)
        xs = [inp[f"fcb.x{l}"] for l in range(5)]
        boxes = [inp[f"fcb.box{l}.1"] for l in range(5)]
        spec = m.conv_adaption.spec()
        outs = m.calibrate_levels(xs, boxes)
        fused = bool(m._fused)
        with ClockSampler(local_rank) as kclk:
            k_ms = _time_launches(lambda: m.calibrate_levels(xs, boxes, outs=outs), reps)
        kc = kclk.summary()
        px = sum(h * w for h, w in hp.level_sizes)
        flops = 2.0 * n_local * px * 256 * 256 * 15
        ach = flops / (k_ms / 1e3) / 1e12
        fpn_shapes = [tuple(x.shape) for x in xs]
        variant = ops.deform_conv2d_variant(fpn_shapes, spec, xs[0].dtype, args.backend, fcb=fused)
        # Which measured peak applies: MEASURED_PEAKS.json gives the cuBLAS bf16 rate as a burst (a kernel timed alone at
        # full clocks) and sustained under the 1 kW power cap.  A launch over all this rank's frames runs for many ms and,
        # repeated back to back, drives the GPU into the power cap (sw_power_cap, SM clock well below max): that is the
        # sustained regime.  The same launch over 74 frames (~1 ms) stays at full clocks: reported as `burst_probe`.
        capped = "sw_power_cap" in (kc.get("reasons") or []) or (kc.get("sm_mhz") and kc.get("sm_max_mhz") and kc["sm_mhz"] < 0.95 * kc["sm_max_mhz"])
        peak = peaks["bf16_sustained"] if capped else peaks["bf16_burst"]
        roof = {"kernel": f"deform_conv2d FCB 3x5 256->256, P3..P7, {n_local} frames, one launch [{variant}]", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": _traffic("dcn_fcb35", "frames", n_local), "ms_per_launch": k_ms, "flops_per_launch": flops,
                "regime": "sustained (power-capped during the timed launches)" if capped else "burst (full clocks during the timed launches)",
                "clocks_during_launches": kc, "frac_of_burst_peak": ach / peaks["bf16_burst"],
                "frac_of_sustained_peak": ach / peaks["bf16_sustained"],
                "peak_source": peaks["src"] + (", sustained" if capped else ", burst")}
        # 74 frames: 2958 CTAs = 9.99 waves of the 296 CTAs resident on 148 SMs (72 frames leave a 0.72-full last wave)
        nb = min(74, n_local)
        if nb < n_local:
            xb, ob, yb = [x[:nb] for x in xs], [o[:nb] for o in boxes], [y[:nb] for y in outs]
            time.sleep(0.5)                        # let the clocks recover from the capped stretch above
            with ClockSampler(local_rank) as bclk:
                b_ms = _time_launches(lambda: m.calibrate_levels(xb, ob, outs=yb), 20)
            bfl = 2.0 * nb * px * 256 * 256 * 15
            roof["burst_probe"] = {"frames": nb, "ms_per_launch": b_ms, "achieved": bfl / (b_ms / 1e3) / 1e12, "peak": peaks["bf16_burst"],
                                   "frac": bfl / (b_ms / 1e3) / 1e12 / peaks["bf16_burst"], "clocks": bclk.summary(),
                                   "variant": ops.deform_conv2d_variant([(nb,) + tuple(x.shape[1:]) for x in xs], spec, xs[0].dtype, args.backend, fcb=fused)}
        del outs
    if hp_cfg.temporal_fusion:
        # the step's own temporal-fusion launch: this rank's (t-1, t) pairs read in place through index arrays (for N > 1 the
        # pairs whose reference frame is a received halo are left out here: local clips of the local frames only)
        tf_plan = plan if world == 1 else sharding.make_plan(1, n_local, 1)
        n_pairs = tf_plan.local_pairs(0)
        fr = inp["tf.fpn"][:n_pairs]
        k_ms = _time_launches(lambda: hp._tf_pairs(inp["tf.fpn"], inp["tf.t2s"], tf_plan, 0, None), 4 * reps)
        es = 2 if hp_cfg.dtype == torch.bfloat16 else 4
        npx = fr.shape[0] * fr.shape[2] * fr.shape[3]
        # H*W*(2C + P^2) bytes, plus the 2*Ct WRITTEN concat bytes because this kernel copies them;
        # the 2*Ct feature bytes it also has to READ and the 7 zero pad channels of the padded layout it writes are not counted
        nbytes = npx * (2 * 256 + 121 + 2 * 256) * es
        ach = nbytes / (k_ms / 1e3) / 1e9
        corr_roof = {"kernel": f"correlation+concat[{ops.correlation_backend(tuple(fr.shape), fr.dtype, 11, 1, args.backend)}] "
                               f"P=11 C=256 24x40, {fr.shape[0]} frame pairs, one launch", "bound": "hbm", "achieved": ach,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                     "traffic": _traffic("corr_fused", "pairs", int(fr.shape[0])), "ms_per_launch": k_ms,
                     "bytes_per_launch": nbytes, "frac_plain_8d_bytes": npx * (2 * 256 + 121) * es / (k_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                     "note": "pairs are read in place from the frame batch: a frame is x2 of one pair and x1 of the next but comes from "
                             "DRAM once, so the measured traffic is below the per-pair algorithmic bytes",
                     "peak_source": peaks["src"]}

    # ---------------- on-device multi-GPU correctness: rank-sharded temporal fusion (halos over NCCL) == one GPU ----------
    if world > 1 and hp_cfg.temporal_fusion and (not args.no_extras or args.multi_gpu_check):
        extras["multi_gpu_check"] = multi_gpu_check(hp, inp, plan, rank, world, dev, n_clips, fpc)

    if not args.no_extras and rank == 0:
        extras["roofline_layers"] = layer_table(hp, inp, n_local, peaks, reps, args.backend)
        if world == 1:
            extras["roofline_sweep"] = operator_sweep(dev, peaks, reps)

    # ---------------- end to end: pinned host slabs -> device -> hot path -> pinned host slabs --------------------
    e2e = None
    if not args.no_e2e:
        from stmask_b200.hotpath import StreamedIO
        io = StreamedIO(hp, dev, n_local, chunk_frames=args.e2e_chunk)
        host = io.host_buffers(plan.local_pairs(rank))
        h_in, h_out, h_tin, h_tout = host
        # fill the pinned input slabs once from the synthetic device tensors (this is the "data loader" side)
        for ci in range(io.n_chunks):
            a, b = ci * io.chunk, min(n_local, (ci + 1) * io.chunk)
            for k, v in io.lin.views(h_in[ci], b - a).items():
                v.copy_(inp[k][a:b])
        for k, v in io.tin.views(h_tin).items():
            v.copy_(inp[k])
        torch.cuda.synchronize()
        del inp, out, out_holder
        torch.cuda.empty_cache()
        e2e_in = sum(t.numel() for t in h_in) + h_tin.numel()
        e2e_out = sum(t.numel() for t in h_out) + (h_tout.numel() if h_tout is not None else 0)

        def e2e_step():
            hp.forward_streamed(host, io, plan, rank)

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
        dt = float(t_e.item())
        # the host-copy ceiling of this box for the same slabs: H2D and D2H of every chunk concurrently, no kernels
        d_a = torch.empty_like(h_in[0], device=dev)
        d_b = torch.empty_like(h_out[0], device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        barrier()
        t0 = time.perf_counter()
        for ci in range(io.n_chunks):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in[ci], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[ci].copy_(d_b, non_blocking=True)
        barrier()
        dc = time.perf_counter() - t0
        t_c = torch.tensor([dc], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        dc = float(t_c.item())
        copy_bytes = sum(t.numel() for t in h_in) + sum(t.numel() for t in h_out)
        e2e = {"value": total_frames * e2e_steps / dt, "unit": "frames/sec",
               "h2d_bytes_per_step": e2e_in, "d2h_bytes_per_step": e2e_out, "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "host_copy_ceiling": {"frames_per_sec": total_frames / dc, "gb_per_s_each_way_per_gpu": copy_bytes / 2 / dc / 1e9,
                                     "how": "the same pinned slabs copied H2D and D2H concurrently with no kernels, max over ranks"},
               "frac_of_host_copy_ceiling": (total_frames * e2e_steps / dt) / (total_frames / dc),
               "host_placement": numa,
               "note": f"HotPath.forward_streamed: pinned HOST slabs -> device -> hot path -> pinned HOST slabs, every step; "
                       f"{args.e2e_chunk}-frame chunks, ONE copy per chunk and direction, H2D / kernels / D2H overlapped on three "
                       f"streams with double-buffered device slabs; bytes are per rank per step.  The boundary ships the path's "
                       f"intermediate activations (DCN inputs, FPN levels), which a full pipeline would keep on the device: "
                       f"this number is bound by the host link, not by the kernels"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(hp_cfg)

    if rank == 0:
        fl = hp.flops_per_frame()
        line = {
            "metric": METRIC, "value": value, "unit": "frames/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if hp_cfg.dtype == torch.bfloat16 else "f32", "data": "synthetic",
            "config": {"workload": f"{hp_cfg.name()} HOT PATH ONLY (11 DCNv2 layers + offset predictors, FCB 3 kernels x P3..P7, TF "
                                   f"correlation+concat on synthetic activations; FPN / heads / NMS are not part of it) — BASELINE.json "
                                   f"configs[{4 if args.workload == 'c5' else 3}]: {n_clips} clips x {fpc} frames, 360x640 padded to 384x640, "
                                   f"random weights, non-zero offsets",
                       "frames_per_step": total_frames, "frames_per_gpu": n_local, "sharding": plan.mode,
                       "halos_per_rank": max(len(plan.recv_halos(r)) for r in range(world)),
                       "l2": f"inputs+outputs per step = {(in_bytes + out_bytes) / 1e6:.0f} MB per GPU > 126 MB L2 (no explicit flush needed)",
                       "dcn_gflop_per_frame": (fl["backbone_dcn"] + fl["fcb"]) / 1e9, "backend": args.backend},
            "roofline": roof, "roofline_correlation": corr_roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_check(hp, inp, plan, rank, world, dev, n_clips, fpc):
    """Every rank computes the temporal-fusion concat of ITS frame pairs with the halos it received over NCCL (the
    step's own code path); the features are all-gathered, rank 0 recomputes ALL pairs of all clips on one GPU without any
    sharding and compares the gathered per-rank results with it bit for bit."""
    import torch
    import torch.distributed as dist
    from stmask_b200 import sharding
    mine = hp(inp, plan, rank)["tf.concat"].contiguous(memory_format=torch.channels_last)
    counts = [plan.local_frames(r) for r in range(world)]
    pairs = [plan.local_pairs(r) for r in range(world)]

    def gather(t, ns):
        nhwc = t.permute(0, 2, 3, 1).contiguous()
        bufs = [nhwc.new_empty((n,) + tuple(nhwc.shape[1:])) for n in ns]
        dist.all_gather(bufs, nhwc) if len(set(ns)) == 1 else _all_gather_ragged(bufs, nhwc, rank, world)
        return bufs
    fpn, t2s, outs = gather(inp["tf.fpn"], counts), gather(inp["tf.t2s"], counts), gather(mine, pairs)
    res = None
    if rank == 0:
        h, w, c = fpn[0].shape[1:]
        full_f = fpn[0].new_empty((n_clips * fpc, h, w, c))
        full_t = torch.empty_like(full_f)
        for r in range(world):
            for i, (clip, f) in enumerate(sharding.frame_order(plan, r)):
                full_f[clip * fpc + f], full_t[clip * fpc + f] = fpn[r][i], t2s[r][i]
        one = sharding.make_plan(n_clips, fpc, 1, "clip")
        ref = hp._tf_pairs(full_f.permute(0, 3, 1, 2), full_t.permute(0, 3, 1, 2), one, 0, None).permute(0, 2, 3, 1)
        pos = {cf: i for i, cf in enumerate(sharding.pair_frames(one, 0))}
        bad = 0
        for r in range(world):
            idx = torch.tensor([pos[cf] for cf in sharding.pair_frames(plan, r)], device=dev, dtype=torch.long)
            if idx.numel():
                bad += int((ref.index_select(0, idx) != outs[r]).any(dim=(1, 2, 3)).sum())
        res = {"tf_concat_bit_identical": bad == 0, "pairs_checked": int(sum(pairs)), "mismatching_pairs": bad,
               "halos_per_rank": max(len(plan.recv_halos(r)) for r in range(world)), "sharding": plan.mode,
               "how": "per-rank pair-indexed correlation+concat with NCCL halos vs. the unsharded single-GPU launch on all-gathered features"}
    torch.cuda.synchronize()
    return res


def _all_gather_ragged(bufs, mine, rank, world):
    import torch.distributed as dist
    for r in range(world):
        if r == rank:
            bufs[r].copy_(mine)
        dist.broadcast(bufs[r], src=r)


def layer_table(hp, inp, n_local, peaks, reps, backend):
    """Per-layer roofline: every distinct backbone DCNv2 shape (deformable conv and its plain-conv predictor) and the
    three FCB kernels, each timed alone (burst peak)."""
    import torch
    from stmask_b200 import ops
    rows, seen = [], set()
    for i, (s, m) in enumerate(zip(hp.dcn_shapes, hp.backbone_dcn)):
        key = (s.channels, s.in_h, s.in_w, s.stride)
        if key in seen:
            continue
        seen.add(key)
        x = inp[f"dcn{i}.x"]
        com = m.conv_offset_mask
        with torch.no_grad():
            om = m._predictor([x], com.weight, com.bias, com.stride, com.padding, com.dilation, out_f32=True, out_planar=True)[0]
            spec = m._spec()
            wp, bf = m._cache.weight(m.weight, spec, x.dtype), m._cache.bias(m.bias)
            y = ops.deform_conv2d_multi([x], [om[:, :18]], [om[:, 18:27]], wp, bf, spec, mask_sigmoid=True, backend=backend)
            t = _time_launches(lambda: ops.deform_conv2d_multi([x], [om[:, :18]], [om[:, 18:27]], wp, bf, spec, mask_sigmoid=True,
                                                               backend=backend, outs=y), reps)
            tp = _time_launches(lambda: m._predictor([x], com.weight, com.bias, com.stride, com.padding, com.dilation, out_f32=True, out_planar=True), reps)
        fl = float(n_local) * s.flops_per_frame
        n_same = sum(1 for q in hp.dcn_shapes if (q.channels, q.in_h, q.in_w, q.stride) == key)
        # the offset / mask-logit predictor (a regular 3x3 conv to 27 -> 32 fp32 channels) is judged on HBM bytes: it reads x once
        # and writes its fp32 output once
        pbytes = float(x.numel() * x.element_size() + om.numel() * om.element_size())
        rows.append({"layer": f"backbone DCNv2 C={s.channels} {s.in_h}x{s.in_w} s{s.stride} (x{n_same})", "ms": t, "tflops": fl / t / 1e9,
                     "frac": fl / t / 1e9 / peaks["bf16_burst"], "predictor_ms": tp, "predictor_gbs": pbytes / tp / 1e6,
                     "predictor_frac_hbm": pbytes / tp / 1e6 / peaks["hbm_gbs"],
                     "variant": ops.deform_conv2d_variant([tuple(x.shape)], spec, x.dtype, backend),
                     "predictor_variant": ops.deform_conv2d_variant([tuple(x.shape)], ops.ConvSpec(s.channels, 32, 3, s.stride, 1), x.dtype, backend,
                                                                    zero_offset=True)})
    px = sum(h * w for h, w in hp.level_sizes)
    for k, m in enumerate(hp.fcb):
        kh, kw = m.kernel_size
        xs = [inp[f"fcb.x{l}"] for l in range(5)]
        boxes = [inp[f"fcb.box{l}.{k}"] for l in range(5)]
        with torch.no_grad():
            outs = m.calibrate_levels(xs, boxes)
            t = _time_launches(lambda: m.calibrate_levels(xs, boxes, outs=outs), reps)
        fl = 2.0 * n_local * px * 256 * 256 * kh * kw
        rows.append({"layer": f"FCB {kh}x{kw} 256->256 P3..P7 (one launch, offsets {'derived in the kernel' if m._fused else 'from a separate kernel'})",
                     "ms": t, "tflops": fl / t / 1e9, "frac": fl / t / 1e9 / peaks["bf16_burst"]})
        del outs
    return {"peak_tflops": peaks["bf16_burst"], "peak_source": peaks["src"] + ", burst", "rows": rows}


def operator_sweep(dev, peaks, reps):
    """BASELINE.json configs[1]: batch 8 over P3..P7 — deformable conv 3x3 / 3x5 with deform_groups 1 / 4 (one grouped
    launch each) and the plain 121-channel correlation cost volume with dilation_patch 1 / 2."""
    import torch
    from stmask_b200 import ops
    from stmask_b200.hotpath import fpn_level_sizes
    lv = fpn_level_sizes()
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn((8, h, w, 256), generator=g, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2) for h, w in lv]
    px = sum(h * w for h, w in lv)
    rows = []
    for (kh, kw) in ((3, 3), (3, 5)):
        for dg in (1, 4):
            spec = ops.ConvSpec(256, 256, (kh, kw), 1, ((kh - 1) // 2, (kw - 1) // 2), 1, 1, dg)
            w = (torch.randn(256, 256, kh, kw, generator=g, device=dev) / (256 * kh * kw) ** 0.5).bfloat16()
            wp = ops.pack_weight(w, spec, torch.bfloat16)
            offs = [torch.randn((8, dg * 2 * kh * kw, h, ww), generator=g, device=dev) * 2.0 for h, ww in lv]
            masks = [torch.rand((8, dg * kh * kw, h, ww), generator=g, device=dev) for h, ww in lv]
            outs = ops.deform_conv2d_multi(xs, offs, masks, wp, None, spec)
            t, how = _time_graph(lambda: ops.deform_conv2d_multi(xs, offs, masks, wp, None, spec, outs=outs), 4 * reps)
            fl = 2.0 * 8 * px * 256 * 256 * kh * kw
            rows.append({"op": f"modulated deform conv {kh}x{kw} dg={dg}, batch 8, P3..P7, one launch", "ms": t, "tflops": fl / t / 1e9,
                         "frac": fl / t / 1e9 / peaks["bf16_burst"], "bound": "tensor", "timing": how})
    x2 = [torch.randn((8, h, w, 256), generator=g, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2) for h, w in lv]
    for d in (1, 2):
        fn = lambda: ops.correlation_multi(xs, x2, 11, d) if hasattr(ops, "correlation_multi") else \
            [ops.correlation(a, b, 11, d, channels_last=True) for a, b in zip(xs, x2)]
        fn()
        t, how = _time_graph(fn, 4 * reps)
        nb = 8.0 * px * (2 * 256 + 121) * 2
        rows.append({"op": f"correlation P=11 d={d}, batch 8, P3..P7, NHWC cost volume, "
                           f"{'one grouped launch' if hasattr(ops, 'correlation_multi') else 'one launch per level'}",
                     "ms": t, "gbs": nb / t / 1e6, "frac": nb / t / 1e6 / peaks["hbm_gbs"], "bound": "hbm", "timing": how})
    return {"rows": rows, "peaks": {"bf16_tflops": peaks["bf16_burst"], "hbm_gbs": peaks["hbm_gbs"]}}


if __name__ == "__main__":
    main()

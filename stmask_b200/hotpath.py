"""The STMask feature-calibration + temporal-fusion hot path as ONE schedulable unit.

A "step" = one pass over a batch of frames of every operator SURVEY.md §8(a) puts on the path:

  a1/a2  backbone DCNv2 `conv2` of every DCN bottleneck (R50: 7, R101: 11 layers; reference
         backbone.py:20-26,45,105-138), through the drop-in `dcn_v2.DCN` module (offset/mask
         predictor + modulated deformable conv);
  a3/a4  FCB box-guided calibration: for each of the 3 anchor-shaped kernels (3x3, 3x5, 5x3;
         config.py:657-659) offsets from the regressed box deltas (ada: 1x1 conv, ali: closed form)
         and relu(DeformConv2d(conf_x, offset)) over the five FPN levels P3..P7 in one grouped
         launch (Featurealign.py:42-72, prediction_head_FC.py:157-167, shared head STMask.py:91-92);
  a5/a6  temporal fusion: correlate(fpn_{t-1}, fpn_t) / C, concat with T2S features, ReLU — one
         kernel — for every consecutive frame pair of a clip (track_to_segment_head.py:40-62,
         TF_utils.py:28-31), with the previous frame coming from a neighbour rank's halo at a
         shard boundary (sharding.py).

Everything between these operators (1x1/3x3 dense convs, BN, FPN, the prediction convs) is
out of scope and is replaced by synthetic activations of the right shape.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import backbone_dcn, ops, sharding
from .compat.dcn_v2 import DCN
from .feature_align import FeatureAlign

FPN_STRIDES = (8, 16, 32, 64, 128)
HEAD_KERNELS = ((3, 3), (3, 5), (5, 3))         # cfg.head_layer_params kernel sizes (config.py:657-659)
FPN_CHANNELS = 256                               # cfg.fpn.num_features (config.py:364)
CORR_PATCH = 11                                  # cfg.correlation_patch_size (config.py:690)
CORR_LEVEL = 1                                   # cfg.correlation_selected_layer (config.py:691)
NUM_CLASSES = 41


def fpn_level_sizes(height: int = 384, width: int = 640) -> List[Tuple[int, int]]:
    """P3..P7 sizes: C3/C4/C5 strides 8/16/32, then two stride-2 3x3 convs (FPN.py:43-59)."""
    sizes = []
    h, w = height, width
    for _ in range(3):
        h, w = (h + 1) // 2, (w + 1) // 2
    for lvl in range(5):
        if lvl > 0:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        sizes.append((h, w))
    return sizes


@dataclass
class HotPathConfig:
    backbone: str = "r101"             # "r50" | "r101"
    fcb: Optional[str] = "ada"         # "ada" | "ali" | None (FCA only)
    temporal_fusion: bool = True
    height: int = 384                  # 360 padded to a multiple of 32 (config.py:118-120)
    width: int = 640
    dtype: torch.dtype = torch.bfloat16
    backend: str = "auto"

    @property
    def resnet_args(self):
        return backbone_dcn.RESNET101_DCN if self.backbone == "r101" else backbone_dcn.RESNET50_DCN

    def name(self) -> str:
        parts = ["FCA"]
        if self.fcb:
            parts.append(f"FCB({self.fcb})")
        if self.temporal_fusion:
            parts.append("TF")
        return f"{'R101' if self.backbone == 'r101' else 'R50'}-DCN-FPN " + "+".join(parts)


class SlabLayout:
    """Byte layout of ONE pinned-host / device slab holding, for a chunk of `n` frames, every tensor of `shapes`
    back to back (each as NHWC-contiguous [n, H, W, C]): one host<->device copy per chunk and direction instead of
    one per tensor (r1: ~60 small copies per chunk reached 34 GB/s; one slab reaches the link's rate)."""

    def __init__(self, shapes: Dict[str, Tuple[Tuple[int, int, int], torch.dtype]], n: int):
        self.n = n
        self.fields = []                 # (key, (C, H, W), dtype, byte offset, byte length)
        off = 0
        for k, ((c, h, w), dt) in shapes.items():
            nb = n * c * h * w * torch.empty((), dtype=dt).element_size()
            self.fields.append((k, (c, h, w), dt, off, nb))
            off += (nb + 255) // 256 * 256
        self.nbytes = off

    def views(self, slab: torch.Tensor, n: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """NCHW-shaped channels-last views [n, C, H, W] into a uint8 slab (first n frames of every field)."""
        n = self.n if n is None else n
        out = {}
        for k, (c, h, w), dt, off, nb in self.fields:
            es = torch.empty((), dtype=dt).element_size()
            t = slab[off:off + n * c * h * w * es].view(dt).view(n, h, w, c).permute(0, 3, 1, 2)
            out[k] = t
        return out


class StreamedIO:
    """Streams, slab layouts and double-buffered device staging of `HotPath.forward_streamed`."""

    def __init__(self, hp: "HotPath", device, n_frames: int, chunk_frames: int = 16):
        self.device = device
        self.chunk = chunk_frames
        self.n = n_frames
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        fin, fout, tin, tout = hp.io_shapes()
        self.lin, self.lout = SlabLayout(fin, chunk_frames), SlabLayout(fout, chunk_frames)
        self.tin, self.tout = SlabLayout(tin, n_frames), None
        self.tout_shapes = tout
        self.d_in = [torch.empty(self.lin.nbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self.d_out = [torch.empty(self.lout.nbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self.d_tin = torch.empty(max(self.tin.nbytes, 1), dtype=torch.uint8, device=device)
        self.free_in = [torch.cuda.Event() for _ in range(2)]      # compute finished reading d_in[i]
        self.free_out = [torch.cuda.Event() for _ in range(2)]     # D2H finished reading d_out[i]
        self.n_chunks = (n_frames + chunk_frames - 1) // chunk_frames

    def host_buffers(self, n_pairs: int):
        """Pinned host slabs: inputs / outputs per chunk, temporal-fusion inputs (all frames) and result."""
        pin = lambda nb: torch.empty(max(nb, 1), dtype=torch.uint8).pin_memory()
        h_in = [pin(self.lin.nbytes) for _ in range(self.n_chunks)]
        h_out = [pin(self.lout.nbytes) for _ in range(self.n_chunks)]
        h_tin = pin(self.tin.nbytes)
        self.tout = SlabLayout(self.tout_shapes, n_pairs) if self.tout_shapes and n_pairs > 0 else None
        h_tout = pin(self.tout.nbytes) if self.tout is not None else None
        return h_in, h_out, h_tin, h_tout


class HotPath(torch.nn.Module):
    """Holds the hot path's weights (random, seeded; offset predictors NON-zero) and runs it."""

    def __init__(self, cfg: HotPathConfig, device, seed: int = 0):
        super().__init__()
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        self.dcn_shapes = backbone_dcn.dcn_layer_shapes(*cfg.resnet_args, cfg.height, cfg.width)
        self.level_sizes = fpn_level_sizes(cfg.height, cfg.width)
        self.backbone_dcn = torch.nn.ModuleList()
        for s in self.dcn_shapes:
            m = DCN(s.channels, s.channels, kernel_size=3, stride=s.stride, padding=1, dilation=1, deformable_groups=1)
            k = s.channels * 9
            with torch.no_grad():
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) / k ** 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.conv_offset_mask.weight.copy_(torch.randn(m.conv_offset_mask.weight.shape, generator=g) * 0.05 / 3)
                m.conv_offset_mask.bias.copy_(torch.randn(m.conv_offset_mask.bias.shape, generator=g) * 0.5)
            self.backbone_dcn.append(m)
        self.fcb = torch.nn.ModuleList()
        if cfg.fcb:
            for ks in HEAD_KERNELS:
                m = FeatureAlign(FPN_CHANNELS, NUM_CLASSES, kernel_size=ks, deformable_groups=1,
                                 use_pred_offset=(cfg.fcb == "ada"))
                with torch.no_grad():
                    m.conv_adaption.weight.copy_(torch.randn(m.conv_adaption.weight.shape, generator=g) /
                                                 (FPN_CHANNELS * ks[0] * ks[1]) ** 0.5)
                    if cfg.fcb == "ada":
                        m.conv_offset.weight.copy_(torch.randn(m.conv_offset.weight.shape, generator=g) * 0.5)
                self.fcb.append(m)
        self.to(device=device)
        # activations / weights of the operators are cfg.dtype; offset predictors' tiny weights stay fp32
        for m in self.backbone_dcn:
            m.to(cfg.dtype)
        for m in self.fcb:
            m.conv_adaption.to(cfg.dtype)

    @property
    def _halo_buffers(self) -> dict:
        b = getattr(self, "_halo_bufs", None)
        if b is None:
            b = self._halo_bufs = {}
        return b

    def _comm_stream(self, device) -> "torch.cuda.Stream":
        st = getattr(self, "_comm", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device, priority=-1)
            self._comm = st
        return st

    # ---------------------------------------------------------------- synthetic activations
    def make_inputs(self, n_frames: int, device, seed: int = 0, pinned_host: bool = False,
                    on_device: bool = False) -> Dict[str, torch.Tensor]:
        """Synthetic N(0,1) activations of the shapes the reference produces for `n_frames` frames
        (NHWC memory, cfg.dtype); box deltas N(0,1) fp32.  `on_device`: draw them with the device's generator
        (the 1024-frame workload is 15 GB of inputs; the host generator would take minutes)."""
        g = torch.Generator(device=device if on_device else "cpu").manual_seed(1000 + seed)
        dt = self.cfg.dtype

        def mk(shape, dtype):
            if on_device:
                n, c, h, w = shape
                return torch.randn((n, h, w, c), generator=g, dtype=dtype, device=device).permute(0, 3, 1, 2)
            t = torch.randn(shape, generator=g, dtype=torch.float32).to(dtype).contiguous(memory_format=torch.channels_last)
            if pinned_host:
                return t.pin_memory()
            return t.to(device)

        inp: Dict[str, torch.Tensor] = {}
        for i, s in enumerate(self.dcn_shapes):
            inp[f"dcn{i}.x"] = mk((n_frames, s.channels, s.in_h, s.in_w), dt)
        if self.cfg.fcb:
            for l, (h, w) in enumerate(self.level_sizes):
                inp[f"fcb.x{l}"] = mk((n_frames, FPN_CHANNELS, h, w), dt)
                for k in range(len(HEAD_KERNELS)):
                    inp[f"fcb.box{l}.{k}"] = mk((n_frames, 4, h, w), torch.float32)
        if self.cfg.temporal_fusion:
            h, w = self.level_sizes[CORR_LEVEL]
            inp["tf.fpn"] = mk((n_frames, FPN_CHANNELS, h, w), dt)
            inp["tf.t2s"] = mk((n_frames, FPN_CHANNELS, h, w), dt)
        return inp

    # ---------------------------------------------------------------- the step
    @torch.no_grad()
    def forward(self, inp: Dict[str, torch.Tensor], plan: Optional[sharding.ShardPlan] = None, rank: int = 0,
                group=None) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        halo = None
        comm = None
        if self.cfg.temporal_fusion and plan is not None and plan.world_size > 1:
            # Post the halo exchange first, on a high-priority SIDE stream: the NCCL send/recv completes only when
            # the neighbour rank has posted its side, and on the compute stream that wait would put all ranks in
            # lock step every step.  The compute stream joins the side stream right before the correlation, so the
            # NVLink transfer (and any rank skew up to the DCN work's duration) hides behind the DCN kernels.
            main = torch.cuda.current_stream()
            if inp["tf.fpn"].is_cuda:
                comm = self._comm_stream(inp["tf.fpn"].device)
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    # persistent message / receive buffers: the side stream waited for the previous step's
                    # correlation (comm.wait_stream above), so reusing them is safe and nothing is allocated here
                    halo = sharding.exchange_halo(plan, rank, [inp["tf.fpn"], inp["tf.t2s"]], group, self._halo_buffers)
            else:
                halo = sharding.exchange_halo(plan, rank, [inp["tf.fpn"], inp["tf.t2s"]], group)
        for i, m in enumerate(self.backbone_dcn):
            out[f"dcn{i}.y"] = m(inp[f"dcn{i}.x"])
        for k, m in enumerate(self.fcb):
            xs = [inp[f"fcb.x{l}"] for l in range(len(self.level_sizes))]
            boxes = [inp[f"fcb.box{l}.{k}"] for l in range(len(self.level_sizes))]
            for l, y in enumerate(m.calibrate_levels(xs, boxes)):
                out[f"fcb.y{l}.{k}"] = y
        if self.cfg.temporal_fusion:
            n = inp["tf.fpn"].shape[0]
            if comm is not None:
                main.wait_stream(comm)
            if plan is None:
                plan = sharding.make_plan(1, n, 1, "clip")
            tf = self._tf_pairs(inp["tf.fpn"], inp["tf.t2s"], plan, rank, halo)
            if tf is not None:
                out["tf.concat"] = tf
        return out

    # ---------------------------------------------------------------- end to end from host memory
    def io_shapes(self):
        """Per-frame (C, H, W) / dtype of every per-frame input and output, and of the temporal-fusion inputs / result."""
        dt = self.cfg.dtype
        fin, fout, tin, tout = {}, {}, {}, {}
        for i, s in enumerate(self.dcn_shapes):
            fin[f"dcn{i}.x"] = ((s.channels, s.in_h, s.in_w), dt)
            fout[f"dcn{i}.y"] = ((s.channels, s.out_h, s.out_w), dt)
        if self.cfg.fcb:
            for l, (h, w) in enumerate(self.level_sizes):
                fin[f"fcb.x{l}"] = ((FPN_CHANNELS, h, w), dt)
                for k in range(len(HEAD_KERNELS)):
                    fin[f"fcb.box{l}.{k}"] = ((4, h, w), torch.float32)
                    fout[f"fcb.y{l}.{k}"] = ((FPN_CHANNELS, h, w), dt)
        if self.cfg.temporal_fusion:
            h, w = self.level_sizes[CORR_LEVEL]
            tin["tf.fpn"] = ((FPN_CHANNELS, h, w), dt)
            tin["tf.t2s"] = ((FPN_CHANNELS, h, w), dt)
            from .temporal_fusion import padded_corr_channels
            tout["tf.concat"] = ((padded_corr_channels(CORR_PATCH) + 2 * FPN_CHANNELS, h, w), dt)
        return fin, fout, tin, tout

    @torch.no_grad()
    def forward_streamed(self, host, io: "StreamedIO", plan: Optional[sharding.ShardPlan] = None, rank: int = 0, group=None) -> None:
        """The same step as `forward`, fed from PINNED HOST slabs and delivering every result into pinned host
        slabs.  `host` = (h_in, h_out, h_tin, h_tout) from `io.host_buffers()`: per chunk of `io.chunk` frames ONE
        input slab and ONE output slab (`io.lin` / `io.lout` give the tensor views), plus one slab with the
        temporal-fusion inputs of all frames and one for its result.  Per chunk: one H2D copy, the kernels (writing
        straight into the output slab's views), one D2H copy; the three run on three streams with double-buffered
        device slabs, so both PCIe directions and the compute overlap.  Temporal fusion goes first (its inputs are
        small and every later chunk is independent of it)."""
        h_in, h_out, h_tin, h_tout = host
        keep = []
        if self.cfg.temporal_fusion and io.tin.nbytes:
            ev_in, ev_done = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(io.s_in):
                io.d_tin.copy_(h_tin, non_blocking=True)
                ev_in.record(io.s_in)
            with torch.cuda.stream(io.s_comp):
                io.s_comp.wait_event(ev_in)
                sub = self._tf_only(io.tin.views(io.d_tin), plan, rank, group)
                ev_done.record(io.s_comp)
            if sub:
                t = sub["tf.concat"]
                keep.append(t)
                with torch.cuda.stream(io.s_out):
                    io.s_out.wait_event(ev_done)
                    t.record_stream(io.s_out)
                    io.tout.views(h_tout)["tf.concat"].copy_(t, non_blocking=True)
        for ci in range(io.n_chunks):
            a = ci * io.chunk
            nf = min(io.n, a + io.chunk) - a
            buf = ci & 1
            ev_in, ev_done = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(io.s_in):
                io.s_in.wait_event(io.free_in[buf])              # the kernels of chunk ci-2 have read this device slab
                io.d_in[buf].copy_(h_in[ci], non_blocking=True)
                ev_in.record(io.s_in)
            with torch.cuda.stream(io.s_comp):
                io.s_comp.wait_event(ev_in)
                io.s_comp.wait_event(io.free_out[buf])           # the D2H of chunk ci-2 has drained this output slab
                self._frames_only(io.lin.views(io.d_in[buf], nf), outs=io.lout.views(io.d_out[buf], nf))
                io.free_in[buf].record(io.s_comp)
                ev_done.record(io.s_comp)
            with torch.cuda.stream(io.s_out):
                io.s_out.wait_event(ev_done)
                h_out[ci].copy_(io.d_out[buf], non_blocking=True)
                io.free_out[buf].record(io.s_out)
        io.s_out.synchronize()
        io.s_comp.synchronize()

    @torch.no_grad()
    def _frames_only(self, inp, outs: Optional[Dict[str, torch.Tensor]] = None):
        """Backbone DCN + FCB of a batch of frames (per-frame independent operators).  `outs`: preallocated
        channels-last result tensors (views of an output slab) the kernels write into directly."""
        out = {}
        for i, m in enumerate(self.backbone_dcn):
            out[f"dcn{i}.y"] = m(inp[f"dcn{i}.x"], out=None if outs is None else outs[f"dcn{i}.y"])
        nl = len(self.level_sizes)
        for k, m in enumerate(self.fcb):
            xs = [inp[f"fcb.x{l}"] for l in range(nl)]
            boxes = [inp[f"fcb.box{l}.{k}"] for l in range(nl)]
            ys = m.calibrate_levels(xs, boxes, outs=None if outs is None else [outs[f"fcb.y{l}.{k}"] for l in range(nl)])
            for l, y in enumerate(ys):
                out[f"fcb.y{l}.{k}"] = y
        return out

    @torch.no_grad()
    def _tf_only(self, inp, plan, rank, group):
        """Temporal fusion of every local frame pair (halo exchange included); ONE [pairs, 633, H, W] result."""
        n = inp["tf.fpn"].shape[0]
        if plan is None:
            plan = sharding.make_plan(1, n, 1, "clip")
        halo = None
        if plan.world_size > 1:
            halo = sharding.exchange_halo(plan, rank, [inp["tf.fpn"], inp["tf.t2s"]], group)
        tf = self._tf_pairs(inp["tf.fpn"], inp["tf.t2s"], plan, rank, halo)
        return {} if tf is None else {"tf.concat": tf}

    def _tf_pairs(self, fpn, t2s, plan, rank, halo):
        """relu(cat[correlate(fpn[t-1], fpn[t]) / C, t2s[t-1], t2s[t]]) for every (t-1, t) pair whose frame t is local:
        ONE launch that reads the pairs straight out of the frame batch (+ the received halo frames) through index
        arrays — no gathered copies.  bf16 runs the pair-indexed tcgen05 kernel; fp32 gathers first."""
        if fpn.dtype == torch.bfloat16 and fpn.is_cuda:
            ref_idx, next_idx = sharding.pair_index_tensors(plan, rank, fpn.device)
            if next_idx.numel() == 0:
                return None
            from .temporal_fusion import padded_corr_channels
            return ops.correlation_pairs(fpn, ref_idx, next_idx, CORR_PATCH, 1, scale=1.0 / fpn.shape[1], relu=True, feats=t2s,
                                         halo=halo[0] if halo else None, feats_halo=halo[1] if halo else None,
                                         feat_channel_offset=padded_corr_channels(CORR_PATCH))
        fpn_ref, fpn_next = sharding.temporal_pairs(plan, rank, fpn, halo[0] if halo else None)
        t2s_ref, t2s_next = sharding.temporal_pairs(plan, rank, t2s, halo[1] if halo else None)
        if fpn_next.shape[0] == 0:
            return None
        return self.temporal_fusion(fpn_ref, fpn_next, t2s_ref, t2s_next)

    def temporal_fusion(self, fpn_ref, fpn_next, t2s_ref, t2s_next):
        from .temporal_fusion import correlate_concat
        return correlate_concat(fpn_ref, fpn_next, t2s_ref, t2s_next, CORR_PATCH, 1, channels_last=True, padded=True)

    # ---------------------------------------------------------------- accounting
    def flops_per_frame(self) -> Dict[str, float]:
        """Algorithmic DCN flops per frame: 2*Ho*Wo*Cout*Cin*kh*kw, gather not counted."""
        px = sum(h * w for h, w in self.level_sizes)
        return {
            "backbone_dcn": float(sum(s.flops_per_frame for s in self.dcn_shapes)),
            "fcb": float(sum(2 * px * FPN_CHANNELS * FPN_CHANNELS * kh * kw for kh, kw in HEAD_KERNELS)) if self.cfg.fcb else 0.0,
        }

    def corr_bytes_per_pair(self) -> float:
        """Algorithmic correlation bytes per frame pair: H*W*(2*C + P*P)*sizeof."""
        if not self.cfg.temporal_fusion:
            return 0.0
        h, w = self.level_sizes[CORR_LEVEL]
        return float(h * w * (2 * FPN_CHANNELS + CORR_PATCH * CORR_PATCH) * (2 if self.cfg.dtype == torch.bfloat16 else 4))

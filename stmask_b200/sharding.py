"""Clip / frame sharding of the hot path over the GPUs of one box, and the one-frame feature
halo that temporal fusion needs at shard boundaries.

The reference is single-GPU, batch 1, strictly sequential (eval.py:590-605, track_TF.py:43-54);
there is nothing to port.  Backbone DCN and FCB are per-frame independent; the correlation for
frame t reads frame t-1's `{fpn_feat, T2S_feat}` (TF_utils.py:22-31).  So the only exchange is:
rank r sends the features of its LAST local frame of a clip to the rank that owns the NEXT frame
of that clip.  Point-to-point `isend/irecv` (NCCL send/recv over NVLink on GPUs, gloo in the CPU
tests), all boundaries of a step batched into one message per neighbour; no all-reduce.

One process per GPU; `torch.distributed` is plumbing.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Segment:
    """A run of consecutive frames of one clip owned by one rank."""
    clip: int
    start: int       # first frame (inclusive), index within the clip
    stop: int        # last frame (exclusive)
    offset: int      # position of `start` in the rank's local frame batch

    @property
    def length(self) -> int:
        return self.stop - self.start


@dataclass(frozen=True)
class Halo:
    """Rank `src` owns frame `frame-1` of `clip`, rank `dst` owns `frame`."""
    clip: int
    frame: int
    src: int
    dst: int


@dataclass
class ShardPlan:
    world_size: int
    n_clips: int
    frames_per_clip: int
    mode: str
    segments: List[List[Segment]]         # per rank
    halos: List[Halo]

    def local_frames(self, rank: int) -> int:
        return sum(s.length for s in self.segments[rank])

    def recv_halos(self, rank: int) -> List[Halo]:
        return [h for h in self.halos if h.dst == rank]

    def send_halos(self, rank: int) -> List[Halo]:
        return [h for h in self.halos if h.src == rank]

    def local_pairs(self, rank: int) -> int:
        """(t-1, t) frame pairs whose frame t is local: every local frame that is not a clip start."""
        return sum(s.length - (1 if s.start == 0 else 0) for s in self.segments[rank])


def make_plan(n_clips: int, frames_per_clip: int, world_size: int, mode: str = "clip") -> ShardPlan:
    """`clip`:  flatten (clip, frame) and cut into `world_size` contiguous ranges (whole clips per rank
                when n_clips % world_size == 0  =>  no halo at all).
       `frame`: every clip's frames are cut into `world_size` contiguous ranges (rank r owns the r-th
                range of EVERY clip  =>  world_size-1 boundaries per clip); this is the layout that
                exercises the halo exchange."""
    if n_clips < 0 or frames_per_clip < 1 or world_size < 1:
        raise ValueError("bad plan arguments")
    if mode not in ("clip", "frame"):
        raise ValueError(f"unknown sharding mode {mode!r}")
    segs: List[List[Segment]] = [[] for _ in range(world_size)]
    owner: Dict[Tuple[int, int], int] = {}

    def add(rank: int, clip: int, a: int, b: int):
        if b <= a:
            return
        off = sum(s.length for s in segs[rank])
        segs[rank].append(Segment(clip, a, b, off))
        for f in range(a, b):
            owner[(clip, f)] = rank

    if mode == "clip":
        total = n_clips * frames_per_clip
        for r in range(world_size):
            lo, hi = r * total // world_size, (r + 1) * total // world_size
            f = lo
            while f < hi:
                clip, a = divmod(f, frames_per_clip)
                b = min(frames_per_clip, a + (hi - f))
                add(r, clip, a, b)
                f += b - a
    else:
        # frames_per_clip = base * world_size + extra: `extra` ranks take base + 1 frames of a clip.  Which ranks do
        # rotates from clip to clip, so the per-rank totals stay balanced (36 frames over 8 ranks: 4/5 alternate and
        # every rank ends up with 9 frames per two clips instead of 8 vs 10).
        base, extra = divmod(frames_per_clip, world_size)
        for clip in range(n_clips):
            first = (clip * extra) % world_size
            a = 0
            for r in range(world_size):
                b = a + base + (1 if (r - first) % world_size < extra else 0)
                add(r, clip, a, b)
                a = b
    halos = []
    for r in range(world_size):
        for s in segs[r]:
            if s.start > 0:
                src = owner[(s.clip, s.start - 1)]
                if src != r:
                    halos.append(Halo(s.clip, s.start, src, r))
    return ShardPlan(world_size, n_clips, frames_per_clip, mode, segs, halos)


_SEND_ROWS_CACHE: Dict[Tuple, torch.Tensor] = {}


def _send_rows(plan: ShardPlan, rank: int, dst: int, rows: List[int], device) -> torch.Tensor:
    key = (plan.world_size, plan.n_clips, plan.frames_per_clip, plan.mode, rank, dst, str(device))
    t = _SEND_ROWS_CACHE.get(key)
    if t is None:
        t = torch.tensor(rows, dtype=torch.long, device=device)
        _SEND_ROWS_CACHE[key] = t
    return t


def exchange_halo(plan: ShardPlan, rank: int, feats: Sequence[torch.Tensor], group=None,
                  buffers: Optional[dict] = None) -> List[Optional[torch.Tensor]]:
    """Send the last-frame features of every local segment that another rank continues, receive the
    halos this rank needs.  `buffers` (a dict the caller keeps between steps) makes the message and receive
    buffers persistent: no allocation inside a step (a cudaMalloc in the caching allocator is a multi-ms,
    device-synchronising stall that the halo dependency then propagates to every rank).  `feats` = per-frame feature tensors of the local batch, each
    [n_local, C, H, W] (e.g. fpn_feat and T2S_feat).

    Per neighbour and feature tensor ONE message: the boundary frames gathered by one index_select on the NHWC
    view (`[n_boundaries, H, W, C]`, the kernels' layout, whatever the local memory format is), received
    straight into the buffer the correlation kernel reads.  All isend/irecv of a step go out in one batch
    (few large messages: many small ones cost NCCL far more than the gather).

    Returns, for each f in feats, a channels-last tensor [n_recv, C, H, W] ordered like plan.recv_halos(rank)
    (None when this rank receives nothing)."""
    sends = plan.send_halos(rank)
    recvs = plan.recv_halos(rank)
    if plan.world_size == 1 or (not sends and not recvs):
        return [None for _ in feats]
    seg_of = {(s.clip, s.stop): s for s in plan.segments[rank]}
    ops_, keep = [], []
    by_dst: Dict[int, List[Halo]] = {}
    for h in sends:
        by_dst.setdefault(h.dst, []).append(h)
    for dst, hs in sorted(by_dst.items()):
        rows = [seg_of[(h.clip, h.frame)].offset + seg_of[(h.clip, h.frame)].length - 1 for h in hs]
        idx_t = _send_rows(plan, rank, dst, rows, feats[0].device)
        for j, f in enumerate(feats):
            nhwc = f.permute(0, 2, 3, 1)
            if buffers is not None:
                msg = buffers.get(("send", dst, j))
                if msg is None or msg.shape != (len(rows),) + tuple(nhwc.shape[1:]) or msg.dtype != f.dtype or msg.device != f.device:
                    msg = buffers[("send", dst, j)] = f.new_empty((len(rows),) + tuple(nhwc.shape[1:]))
                torch.index_select(nhwc, 0, idx_t, out=msg)
            else:
                msg = nhwc.index_select(0, idx_t)                     # contiguous [n, H, W, C]
            keep.append(msg)
            ops_.append(dist.P2POp(dist.isend, msg, dst, group))
    by_src: Dict[int, List[Halo]] = {}
    for h in recvs:
        by_src.setdefault(h.src, []).append(h)
    bufs: Dict[int, List[torch.Tensor]] = {}
    for src, hs in sorted(by_src.items()):
        bufs[src] = []
        for j, f in enumerate(feats):
            shape = (len(hs),) + tuple(f.shape[2:]) + (f.shape[1],)
            b = buffers.get(("recv", src, j)) if buffers is not None else None
            if b is None or b.shape != shape or b.dtype != f.dtype or b.device != f.device:
                b = f.new_empty(shape)
                if buffers is not None:
                    buffers[("recv", src, j)] = b
            bufs[src].append(b)
        for b in bufs[src]:
            ops_.append(dist.P2POp(dist.irecv, b, src, group))
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    if not recvs:
        return [None for _ in feats]
    out = []
    for j in range(len(feats)):
        if len(by_src) == 1:
            nhwc = bufs[next(iter(by_src))][j]                         # already in plan.recv_halos order
        else:
            pos = {}
            for src, hs in by_src.items():
                for i, h in enumerate(hs):
                    pos[(h.clip, h.frame)] = (src, i)
            nhwc = torch.stack([bufs[pos[(h.clip, h.frame)][0]][j][pos[(h.clip, h.frame)][1]] for h in recvs], 0)
        out.append(nhwc.permute(0, 3, 1, 2))                          # NCHW-shaped view of NHWC memory (channels-last)
    return out


def temporal_pairs(plan: ShardPlan, rank: int, feat: torch.Tensor, halo: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ref, next) batches for the correlation of this rank: next = every local frame that is not a
    clip start; ref = the frame before it — local, or the received halo at a shard boundary.
    Mirrors the train-time batched call (STMask.py:289-297: ref = frames[::2], next = frames[1::2])
    generalised to clips of any length.  The gather runs on the NHWC view, so channels-last
    features stay channels-last (no re-layout before the correlation kernel)."""
    ref_idx, next_idx = pair_indices(plan, rank)
    src = feat if halo is None else torch.cat([feat, halo.to(feat.dtype)], 0)
    dev = feat.device
    ri = torch.as_tensor(ref_idx, device=dev, dtype=torch.long)
    ni = torch.as_tensor(next_idx, device=dev, dtype=torch.long)
    if feat.dim() == 4 and feat.stride(1) == 1:
        return (src.permute(0, 2, 3, 1).index_select(0, ri).permute(0, 3, 1, 2),
                feat.permute(0, 2, 3, 1).index_select(0, ni).permute(0, 3, 1, 2))
    return src.index_select(0, ri), feat.index_select(0, ni)


def pair_indices(plan: ShardPlan, rank: int) -> Tuple[List[int], List[int]]:
    """Row indices (into [local frames ++ received halos]) of the reference and the next frame of every pair."""
    recvs = plan.recv_halos(rank)
    halo_row = {(h.clip, h.frame): i for i, h in enumerate(recvs)}
    ref_idx, next_idx = [], []
    n_local = plan.local_frames(rank)
    for s in plan.segments[rank]:
        for f in range(s.start, s.stop):
            if f == 0:
                continue
            loc = s.offset + (f - s.start)
            next_idx.append(loc)
            ref_idx.append(loc - 1 if f > s.start else n_local + halo_row[(s.clip, f)])
    return ref_idx, next_idx


def pair_frames(plan: ShardPlan, rank: int) -> List[Tuple[int, int]]:
    """(clip, frame t) of every (t-1, t) pair of this rank, in the order `pair_indices` / the correlation output use."""
    return [(s.clip, f) for s in plan.segments[rank] for f in range(s.start, s.stop) if f > 0]


def frame_order(plan: ShardPlan, rank: int) -> List[Tuple[int, int]]:
    """(clip, frame) of every local frame of this rank, in local batch order."""
    return [(s.clip, f) for s in plan.segments[rank] for f in range(s.start, s.stop)]


_PAIR_INDEX_CACHE: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def pair_index_tensors(plan: ShardPlan, rank: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """`pair_indices` as int32 device tensors (what `ops.correlation_pairs` takes), cached per plan / rank / device so
    a step does no host-to-device copy for them."""
    key = (plan.world_size, plan.n_clips, plan.frames_per_clip, plan.mode, rank, str(device))
    hit = _PAIR_INDEX_CACHE.get(key)
    if hit is None:
        ref_idx, next_idx = pair_indices(plan, rank)
        hit = (torch.tensor(ref_idx, dtype=torch.int32, device=device), torch.tensor(next_idx, dtype=torch.int32, device=device))
        _PAIR_INDEX_CACHE[key] = hit
    return hit


def pair_slices(plan: ShardPlan, rank: int) -> Optional[List[Tuple[int, int]]]:
    """When no pair of this rank crosses a shard boundary, the pairs of a segment are simply
    (frames[a:b-1], frames[a+1:b]): return the [a, b) ranges so the caller can use VIEWS (zero copies).
    Returns None when a halo is involved."""
    if plan.recv_halos(rank):
        return None
    return [(s.offset, s.offset + s.length) for s in plan.segments[rank] if s.length > 1]

"""CPU: clip/frame partitioning and the one-frame halo exchange (world_size 2 over gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stmask_b200.sharding import exchange_halo, make_plan, temporal_pairs


def test_clip_mode_whole_clips_need_no_halo():
    for g in (1, 2, 4, 8):
        plan = make_plan(64, 16, g, "clip")
        assert plan.halos == []
        assert [plan.local_frames(r) for r in range(g)] == [1024 // g] * g
        assert sum(plan.local_pairs(r) for r in range(g)) == 64 * 15


def test_clip_mode_uneven_split_creates_boundaries():
    plan = make_plan(3, 5, 2, "clip")          # 15 frames -> 7 + 8; clip 1 is cut between frame 1 and 2
    assert [plan.local_frames(r) for r in range(2)] == [7, 8]
    assert [(h.clip, h.frame, h.src, h.dst) for h in plan.halos] == [(1, 2, 0, 1)]
    assert sum(plan.local_pairs(r) for r in range(2)) == 3 * 4


def test_frame_mode_boundaries():
    plan = make_plan(64, 16, 8, "frame")
    assert len(plan.halos) == 64 * 7
    assert all(h.dst == h.src + 1 for h in plan.halos)
    assert [plan.local_frames(r) for r in range(8)] == [128] * 8
    assert sum(plan.local_pairs(r) for r in range(8)) == 64 * 15
    assert len(plan.recv_halos(0)) == 0 and len(plan.recv_halos(3)) == 64 and len(plan.send_halos(7)) == 0
    with pytest.raises(ValueError):
        make_plan(1, 1, 1, "bogus")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, n_clips, fpc, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = make_plan(n_clips, fpc, world, mode)
        C1, C2, H, W = 6, 4, 3, 5
        # global features: value encodes (clip, frame, channel) so misrouted halos are detectable
        g = torch.Generator().manual_seed(0)
        full1 = torch.randn(n_clips, fpc, C1, H, W, generator=g)
        full2 = torch.randn(n_clips, fpc, C2, H, W, generator=g)
        loc1 = torch.cat([full1[s.clip, s.start:s.stop] for s in plan.segments[rank]], 0)
        loc2 = torch.cat([full2[s.clip, s.start:s.stop] for s in plan.segments[rank]], 0)
        if rank == 1:   # exercise the channels-last message path on one rank's send side too
            loc1 = loc1.contiguous(memory_format=torch.channels_last)
            loc2 = loc2.contiguous(memory_format=torch.channels_last)
        bufs = {}                                  # persistent message / receive buffers: second step must reuse them
        exchange_halo(plan, rank, [torch.zeros_like(loc1), torch.zeros_like(loc2)], buffers=bufs)
        ptrs = sorted((k, v.data_ptr()) for k, v in bufs.items())
        h1, h2 = exchange_halo(plan, rank, [loc1, loc2], buffers=bufs)
        assert ptrs == sorted((k, v.data_ptr()) for k, v in bufs.items())
        ref1, nxt1 = temporal_pairs(plan, rank, loc1, h1)
        ref2, nxt2 = temporal_pairs(plan, rank, loc2, h2)
        # expected pairs straight from the global tensors
        e_ref1, e_nxt1, e_ref2 = [], [], []
        for s in plan.segments[rank]:
            for f in range(s.start, s.stop):
                if f > 0:
                    e_ref1.append(full1[s.clip, f - 1]); e_nxt1.append(full1[s.clip, f]); e_ref2.append(full2[s.clip, f - 1])
        ok = (torch.equal(ref1, torch.stack(e_ref1)) and torch.equal(nxt1, torch.stack(e_nxt1))
              and torch.equal(ref2, torch.stack(e_ref2)) and ref1.shape[0] == plan.local_pairs(rank)
              and nxt2.shape[0] == plan.local_pairs(rank))
        q.put((rank, bool(ok), len(plan.recv_halos(rank))))
    finally:
        dist.destroy_process_group()


def test_frame_mode_rotates_the_odd_frames_so_ranks_stay_balanced():
    plan = make_plan(16, 36, 8, "frame")       # 36 frames over 8 ranks: 4 or 5 per clip
    assert [plan.local_frames(r) for r in range(8)] == [72] * 8
    assert len(plan.halos) == 16 * 7
    for clip in range(16):                      # every clip is covered exactly once, in order
        segs = sorted((s.start, s.stop) for r in range(8) for s in plan.segments[r] if s.clip == clip)
        assert segs[0][0] == 0 and segs[-1][1] == 36 and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
    plan = make_plan(3, 5, 4, "frame")          # fewer frames than ranks * 2: some ranks skip a clip's remainder
    assert sum(plan.local_frames(r) for r in range(4)) == 15
    assert sum(plan.local_pairs(r) for r in range(4)) == 3 * 4


@pytest.mark.parametrize("mode,n_clips,fpc", [("frame", 3, 8), ("frame", 3, 7), ("clip", 3, 5), ("clip", 4, 4)])
def test_halo_exchange_world_size_2_gloo(mode, n_clips, fpc):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, n_clips, fpc, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    n_halo = {"frame": n_clips, "clip": 1 if (n_clips, fpc) == (3, 5) else 0}[mode]
    assert res[1][2] == n_halo and res[0][2] == 0


def test_padded_concat_layout_helpers_cpu():
    """padded layout <-> reference layout: unpad_concat is the exact inverse, and a consumer conv with
    pad_concat_weight gives the same result on the padded tensor as the original weight on the reference tensor."""
    import torch.nn.functional as F
    from stmask_b200.temporal_fusion import pad_concat_weight, padded_corr_channels, unpad_concat
    assert padded_corr_channels(11) == 128 and padded_corr_channels(5) == 32 and padded_corr_channels(3) == 16
    torch.manual_seed(0)
    ref = torch.randn(2, 121 + 2 * 16, 7, 7)
    padded = torch.zeros(2, 128 + 2 * 16, 7, 7)
    padded[:, :121], padded[:, 128:] = ref[:, :121], ref[:, 121:]
    assert torch.equal(unpad_concat(padded), ref)
    w = torch.randn(8, 121 + 2 * 16, 3, 3)
    assert torch.allclose(F.conv2d(padded, pad_concat_weight(w), padding=1), F.conv2d(ref, w, padding=1), atol=1e-5)
    with pytest.raises(ValueError):
        pad_concat_weight(torch.zeros(8, 100, 3, 3))


def test_pair_index_tensors_match_pair_indices():
    from stmask_b200.sharding import pair_index_tensors, pair_indices
    for world, mode in ((1, "clip"), (2, "frame"), (4, "frame"), (2, "clip")):
        plan = make_plan(3, 7, world, mode)
        for r in range(world):
            ri, ni = pair_index_tensors(plan, r, "cpu")
            ref, nxt = pair_indices(plan, r)
            assert ri.dtype == torch.int32 and ri.tolist() == ref and ni.tolist() == nxt
            assert len(nxt) == plan.local_pairs(r)
            n_local = plan.local_frames(r)
            assert all(i < n_local for i in nxt) and sum(i >= n_local for i in ref) == len(plan.recv_halos(r))


def test_plans_cover_every_frame_once_property():
    """Any (clips, frames per clip, world, mode): every frame has exactly one owner, segments of a clip are in rank
    order and contiguous, halos sit exactly at the cuts between different owners, and the pair indices address the
    previous frame of the same clip (locally, or through the matching received halo)."""
    from hypothesis import given, settings, strategies as st
    from stmask_b200.sharding import pair_indices

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 9), st.integers(1, 40), st.integers(1, 8), st.sampled_from(["clip", "frame"]))
    def check(n_clips, fpc, world, mode):
        plan = make_plan(n_clips, fpc, world, mode)
        owner = {}
        for r in range(world):
            off = 0
            for s in plan.segments[r]:
                assert s.offset == off and s.length > 0
                off += s.length
                for f in range(s.start, s.stop):
                    assert (s.clip, f) not in owner
                    owner[(s.clip, f)] = (r, s.offset + f - s.start)
            assert off == plan.local_frames(r)
        assert len(owner) == n_clips * fpc
        cuts = {(c, f) for c in range(n_clips) for f in range(1, fpc) if owner[(c, f)][0] != owner[(c, f - 1)][0]}
        assert {(h.clip, h.frame) for h in plan.halos} == cuts
        assert all(owner[(h.clip, h.frame)][0] == h.dst and owner[(h.clip, h.frame - 1)][0] == h.src for h in plan.halos)
        if mode == "frame" and n_clips % world == 0:
            counts = [plan.local_frames(r) for r in range(world)]
            assert max(counts) - min(counts) <= max(1, n_clips // world) if fpc % world else max(counts) == min(counts)
        for r in range(world):
            ref, nxt = pair_indices(plan, r)
            n_local = plan.local_frames(r)
            recv = plan.recv_halos(r)
            local = {v[1]: k for k, v in owner.items() if v[0] == r}
            for a, b in zip(ref, nxt):
                clip, f = local[b]
                assert f > 0
                if a < n_local:
                    assert local[a] == (clip, f - 1)
                else:
                    assert (recv[a - n_local].clip, recv[a - n_local].frame) == (clip, f)
            assert len(nxt) == plan.local_pairs(r) == sum(1 for (c, f), v in owner.items() if v[0] == r and f > 0)

    check()

"""GPU parity of the device-side tracker (csrc/track.cu behind stm_track_update_fwd, host mirror stmask_b200/tracker.py)
against the reference's own Track_TF.track (tests/golden/tracker.npz, track_TF.py:52-181) and the numpy restatement
oracle/track_oracle.py.  Decisions (matches, ids, keep flags, tracked counters) must be identical; features <= 1e-5."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _pad_dets(frames, max_det, K, E, device):
    """frames: list (one per clip) of dicts of numpy arrays (or None = no detections) -> padded device tensors."""
    C = len(frames)
    out = {"count": torch.zeros(C, dtype=torch.int32), "box": torch.zeros(C, max_det, 4), "score": torch.zeros(C, max_det),
           "cls": torch.zeros(C, max_det, dtype=torch.int32), "coeff": torch.zeros(C, max_det, K), "track": torch.zeros(C, max_det, E),
           "centerness": torch.zeros(C, max_det)}
    for c, d in enumerate(frames):
        if d is None:
            continue
        n = d["box"].shape[0]
        out["count"][c] = n
        for k in ("box", "score", "coeff", "track", "centerness"):
            out[k][c, :n] = torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float32))
        out["cls"][c, :n] = torch.from_numpy(d["cls"].astype(np.int32))
    return {k: v.to(device) for k, v in out.items()}


def _golden_frame(z, f):
    return {"box": z[f"f{f}.det.box"], "score": z[f"f{f}.det.score"], "cls": z[f"f{f}.det.class"].astype(np.int32),
            "coeff": z[f"f{f}.det.mask_coeff"], "track": z[f"f{f}.det.track"], "centerness": z[f"f{f}.det.centerness"]}


def _check_state(trk, clip, want, tag):
    st = {k: v[clip].cpu().numpy() for k, v in trk.state.items() if k != "n_obj"}
    n = int(trk.state["n_obj"][clip])
    assert n == want["box"].shape[0], (tag, n, want["box"].shape[0])
    for k, wk in (("box", "box"), ("score", "score"), ("coeff", "mask_coeff"), ("track", "track"), ("centerness", "centerness"), ("mask", "mask")):
        assert np.abs(st[k][:n] - want[wk]).max() <= 1e-5, (tag, k)
    assert np.array_equal(st["cls"][:n], want["class"].astype(np.int32)), tag
    assert np.array_equal(st["tracked"][:n], want["tracked_mask"].astype(np.int32)), tag


def test_device_tracker_replays_the_reference_clip(cuda_device):
    """Three clips in one batch: the golden clip, the golden clip one frame late (so the two are at different points of
    the state machine in every launch), and the golden clip again behind a frame with no detections."""
    from stmask_b200.tracker import DeviceTracker
    z = load_golden("tracker.npz")
    F = int(z["n_frames"])
    K, E = z["f0.det.mask_coeff"].shape[1], z["f0.det.track"].shape[1]
    H, W = z["f0.proto"].shape[:2]
    cap, max_det = 12, 7
    trk = DeviceTracker(3, cap, K, E, (H, W), cuda_device, match_coeff=z["match_coeff"], conf_thresh=float(z["conf_thresh"]))
    delay = [0, 1, 1]
    for step in range(F + 1):
        fs = [step - d for d in delay]                          # golden frame index per clip (or out of range)
        live = [0 <= f < F for f in fs]
        dets = _pad_dets([_golden_frame(z, f) if ok else None for f, ok in zip(fs, live)], max_det, K, E, cuda_device)
        proto = torch.stack([torch.from_numpy(z[f"f{f}.proto"] if ok else z["f0.proto"]) for f, ok in zip(fs, live)]).to(cuda_device)
        first = torch.tensor([bool(z[f"f{f}.is_first"]) if ok else True for f, ok in zip(fs, live)])
        # CandidateShift's write-back with the recorded TemporalNet outputs
        loc = torch.zeros(3, cap, 4)
        dco = torch.zeros(3, cap, K)
        shifted = [False] * 3
        for c, (f, ok) in enumerate(zip(fs, live)):
            if ok and f"f{f}.shift.loc" in z.files:
                n = z[f"f{f}.shift.loc"].shape[0]
                loc[c, :n], dco[c, :n] = torch.from_numpy(z[f"f{f}.shift.loc"]), torch.from_numpy(z[f"f{f}.shift.coeff"])
                shifted[c] = True
        trk.apply_shift(loc.to(cuda_device), dco.to(cuda_device), proto)
        for c, (f, ok) in enumerate(zip(fs, live)):
            if shifted[c]:
                n = z[f"f{f}.shifted.box"].shape[0]
                assert np.abs(trk.state["box"][c, :n].cpu().numpy() - z[f"f{f}.shifted.box"]).max() <= 1e-6
                assert np.abs(trk.state["mask"][c, :n].cpu().numpy() - z[f"f{f}.shifted.mask"]).max() <= 1e-5
        det_slot, keep = trk.step(dets, proto, first)
        for c, (f, ok) in enumerate(zip(fs, live)):
            if not ok:
                continue
            want = {k: z[f"f{f}.state.{k}"] for k in ("box", "score", "class", "mask_coeff", "track", "centerness", "tracked_mask", "mask")}
            _check_state(trk, c, want, (step, c))
            assert np.array_equal(np.nonzero(keep[c].cpu().numpy())[0], z[f"f{f}.out.box_ids"]), (step, c)


def test_device_tracker_vs_oracle_many_objects(cuda_device):
    """Random crowded clips (up to 40 detections against up to 48 tracked objects, 128-d embeddings, a state that fills up)
    against oracle/track_oracle.py with the kernel's capacity rule."""
    from oracle import track_oracle as T
    from stmask_b200.tracker import DeviceTracker
    rng = np.random.default_rng(11)
    C, cap, max_det, K, E, H, W = 4, 48, 40, 32, 128, 24, 40
    coeffs = (0.0, 1.0, 2.0, 0.5)
    trk = DeviceTracker(C, cap, K, E, (H, W), cuda_device, match_coeff=coeffs, conf_thresh=0.3)
    n_ident = 60
    emb = rng.standard_normal((n_ident, E)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    ctr = (0.1 + 0.8 * rng.random((n_ident, 2))).astype(np.float32)
    wh = (0.08 + 0.25 * rng.random((n_ident, 2))).astype(np.float32)
    icoeff = rng.standard_normal((n_ident, K)).astype(np.float32)
    icls = rng.integers(1, 41, n_ident).astype(np.int32)
    states = [None] * C
    for f in range(6):
        frames, protos, firsts = [], [], []
        for c in range(C):
            n = int(rng.integers(0, max_det + 1)) if not (f == 3 and c == 1) else 0
            ids = rng.integers(0, min(n_ident, 20 + 10 * f), n)
            cc = ctr[ids] + 0.01 * rng.standard_normal((n, 2)).astype(np.float32)
            ww = wh[ids] * (1 + 0.05 * rng.standard_normal((n, 2)).astype(np.float32))
            tr = emb[ids] + 0.1 * rng.standard_normal((n, E)).astype(np.float32)
            tr /= np.maximum(np.linalg.norm(tr, axis=1, keepdims=True), 1e-6)
            frames.append({"box": np.concatenate([cc - ww / 2, cc + ww / 2], 1).astype(np.float32),
                           "score": (0.1 + 0.85 * rng.random(n)).astype(np.float32), "cls": icls[ids],
                           "coeff": (icoeff[ids] + 0.1 * rng.standard_normal((n, K))).astype(np.float32),
                           "track": tr.astype(np.float32), "centerness": rng.random(n).astype(np.float32)})
            protos.append(np.maximum(rng.standard_normal((H, W, K)), 0).astype(np.float32))
            firsts.append(f == 0 or (f == 4 and c == 2))
        proto = torch.from_numpy(np.stack(protos)).to(cuda_device)
        loc = (0.3 * rng.standard_normal((C, cap, 4))).astype(np.float32)
        dco = (0.05 * rng.standard_normal((C, cap, K))).astype(np.float32)
        trk.apply_shift(torch.from_numpy(loc).to(cuda_device), torch.from_numpy(dco).to(cuda_device), proto)
        det_slot, keep = trk.step(_pad_dets(frames, max_det, K, E, cuda_device), proto, torch.tensor(firsts))
        det_slot, keep = det_slot.cpu().numpy(), keep.cpu().numpy()
        for c in range(C):
            st = None if firsts[c] else states[c]
            if st is not None:
                n = st["box"].shape[0]
                st = T.apply_shift(st, loc[c, :n], dco[c, :n], protos[c])
            d = dict(frames[c])
            d["mask"] = T.generate_mask(protos[c], d["coeff"], d["box"])
            st, want_slot, want_keep = T.track_update(st, d, firsts[c], coeffs, conf_thresh=0.3, capacity=cap)
            states[c] = st
            n = 0 if st is None else st["box"].shape[0]
            assert int(trk.state["n_obj"][c]) == n, (f, c)
            nd = d["box"].shape[0]
            assert np.array_equal(det_slot[c, :nd], want_slot), (f, c)
            assert np.array_equal(np.nonzero(keep[c])[0], np.nonzero(want_keep)[0]), (f, c)
            if n:
                for k in ("box", "score", "coeff", "track", "centerness", "mask"):
                    assert np.abs(trk.state[k][c, :n].cpu().numpy() - st[k]).max() <= 2e-5, (f, c, k)
                assert np.array_equal(trk.state["cls"][c, :n].cpu().numpy(), st["cls"])
                assert np.array_equal(trk.state["tracked"][c, :n].cpu().numpy(), st["tracked"])
    assert max(s["box"].shape[0] for s in states if s is not None) == cap          # the capacity rule was exercised


def test_track_update_argument_checks(cuda_device):
    from stmask_b200 import ops
    from stmask_b200.tracker import DeviceTracker
    trk = DeviceTracker(1, 4, 8, 8, (6, 10), cuda_device)
    dets = _pad_dets([None], 3, 8, 8, cuda_device)
    dets["mask_bits"] = torch.zeros(1, 3, 2, dtype=torch.int32, device=cuda_device)
    with pytest.raises(ValueError):
        ops.track_update(trk.state, dets, torch.zeros(1, 3, 5, device=cuda_device), None, match_coeff=(0, 1, 2, 0))
    with pytest.raises(ValueError):
        ops.track_update(trk.state, dets, torch.zeros(1, 3, 4, device=cuda_device), None, match_coeff=(0, 1, 2))


def test_clip_pipeline_glue_vs_oracle(cuda_device):
    """stmask_b200/clip_pipeline.py: NMS -> gathers -> CandidateShift (correlation, RoIAlign, TemporalNet) -> tracker for two
    clips over three frames.  The kernels are pinned elsewhere; here the GLUE is: the gathered detection rows must be the head's
    rows at the NMS indices, and the tracker state must equal the numpy restatement replayed on the pipeline's own detections and shifts."""
    from oracle import track_oracle as T
    from stmask_b200.clip_pipeline import ClipPipeline
    from stmask_b200.temporal_net import TemporalNet
    g = torch.Generator(device=cuda_device).manual_seed(9)
    rnd = lambda *s: torch.randn(*s, generator=g, device=cuda_device)
    C, cap, K, E, P, H, W = 2, 24, 32, 16, 600, 24, 40
    net = TemporalNet(633).to(cuda_device).to(torch.bfloat16)
    pipe = ClipPipeline(net, C, cap, K, E, (H, W), cuda_device, top_k=12, conf_thresh=0.2, nms_thresh=0.5)
    priors = torch.cat([torch.rand(P, 2, generator=g, device=cuda_device) * 0.8 + 0.1, torch.rand(P, 2, generator=g, device=cuda_device) * 0.2 + 0.05], 1)
    states = [None] * C
    for f in range(3):
        conf = rnd(C, P, 41) * 2.0
        preds = {"conf": conf, "loc": rnd(C, P, 4) * 0.2, "centerness": torch.rand(C, P, 1, generator=g, device=cuda_device),
                 "mask_coeff": rnd(C, P, K), "track": torch.nn.functional.normalize(rnd(C, P, E), dim=-1), "priors": priors[None]}
        fpn = rnd(C, 256, H, W).bfloat16().contiguous(memory_format=torch.channels_last)
        t2s = rnd(C, 256, H, W).bfloat16().contiguous(memory_format=torch.channels_last)
        proto = torch.relu(rnd(C, H, W, K))
        first = torch.tensor([f == 0, f == 0])
        out = pipe.step(preds, fpn, t2s, proto, first)
        cnt = out["count"].cpu().numpy()
        assert cnt.min() > 0
        for c in range(C):
            n = int(cnt[c])
            idx = out["index"][c, :n].long()
            assert torch.equal(out["coeff"][c, :n], preds["mask_coeff"][c, idx]) and torch.equal(out["track"][c, :n], preds["track"][c, idx])
            assert torch.equal(out["centerness"][c, :n], preds["centerness"][c, idx, 0])
            det = {"box": out["box"][c, :n].cpu().numpy(), "score": out["score"][c, :n].cpu().numpy(), "cls": out["cls"][c, :n].cpu().numpy(),
                   "coeff": out["coeff"][c, :n].cpu().numpy(), "track": out["track"][c, :n].cpu().numpy(),
                   "centerness": out["centerness"][c, :n].cpu().numpy()}
            pr = proto[c].cpu().numpy()
            det["mask"] = T.generate_mask(pr, det["coeff"], det["box"])
            st = states[c]
            if st is not None and f > 0:
                m = st["box"].shape[0]
                st = T.apply_shift(st, out["loc_shift"][c, :m].float().cpu().numpy(), out["coeff_shift"][c, :m].float().cpu().numpy(), pr)
            states[c], want_slot, want_keep = T.track_update(st, det, f == 0, pipe.tracker.match_coeff, conf_thresh=0.2, capacity=cap)
            m = states[c]["box"].shape[0]
            assert int(pipe.tracker.state["n_obj"][c]) == m
            assert np.array_equal(out["det_slot"][c, :n].cpu().numpy(), want_slot)
            assert np.array_equal(np.nonzero(out["keep"][c].cpu().numpy())[0], np.nonzero(want_keep)[0])
            assert np.abs(pipe.tracker.state["box"][c, :m].cpu().numpy() - states[c]["box"]).max() <= 1e-4
            assert np.abs(pipe.tracker.state["coeff"][c, :m].cpu().numpy() - states[c]["coeff"]).max() <= 1e-4

"""GPU: the cascade and the whole drop-in against the oracle / committed goldens (BASELINE.json tolerances:
same face count per frame, box IoU >= 0.95, embedding cosine >= 0.999, identical flagged set outside a 1e-3 band)."""
import importlib.util
import os

import cv2
import numpy as np
import pytest
import torch

import helpers as H
from oracle.reference_run import reference_run_frames
from truely_b200 import model as M
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _match_boxes(got, ref):
    """every reference box has a distinct detected box with IoU >= 0.95"""
    used = set()
    for r in ref:
        best, bi = 0.0, -1
        for j, g in enumerate(got):
            if j in used:
                continue
            v = H.box_iou(r, g)
            if v > best:
                best, bi = v, j
        assert best >= 0.95, f"reference box {r} best IoU {best}"
        used.add(bi)


def test_detect_matches_oracle_per_stage(analyzer):
    """Face count, boxes and the candidate counts entering R-Net / O-Net, single and multi face, two resolutions."""
    mt = H.oracle_mtcnn()
    cases = [SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=0.5, seed=3),
             SyntheticClip(540, 960, 60, 64, n_faces=(3, 3), face_h=(60.0, 160.0), seed=5),
             SyntheticClip(240, 320, 30, 64, n_faces=(0, 0), seed=8)]          # no face at all
    for clip in cases:
        frames = [clip.frame(i) for i in (0, 7, 19, 33)]
        res = analyzer.process_frames(np.stack(frames), detail=True)
        for k, f in enumerate(frames):
            tr = {}
            boxes, probs = mt.detect(f, trace=tr)
            n_ref = 0 if boxes is None else len(boxes)
            assert res.nfaces[k] == n_ref, f"frame {k}: {res.nfaces[k]} faces vs oracle {n_ref}"
            assert res.counts[k, 1] == len(tr["s1_boxes"]), "R-Net input count"
            assert res.counts[k, 2] == len(tr["s2_boxes"]), "O-Net input count"
            if n_ref:
                _match_boxes(res.boxes[k, :n_ref, :4], boxes)
                # largest-first order and scores
                assert H.box_iou(res.boxes[k, 0, :4], boxes[0]) >= 0.95
                assert abs(res.boxes[k, 0, 4] - probs[0]) < 1e-3


def test_detect_stress_low_thresholds(analyzer):
    """Hundreds of candidates per frame (thresholds lowered in both paths): exercises NMS blocks, ties and capacity."""
    import copy
    mt2 = copy.deepcopy(H.oracle_mtcnn())
    mt2.thresholds = [0.3, 0.4, 0.5]
    a3 = _analyzer_with_thresholds((0.3, 0.4, 0.5))
    clip = SyntheticClip(540, 960, 60, 64, n_faces=(4, 4), face_h=(50.0, 200.0), seed=15)
    frames = [clip.frame(i) for i in (2, 30)]
    res = a3.process_frames(np.stack(frames), detail=True)
    for k, f in enumerate(frames):
        tr = {}
        boxes, _ = mt2.detect(f, trace=tr)
        n_ref = 0 if boxes is None else len(boxes)
        assert res.counts[k, 1] == len(tr["s1_boxes"])
        assert res.counts[k, 2] == len(tr["s2_boxes"])
        assert res.nfaces[k] == n_ref
        if n_ref:
            _match_boxes(res.boxes[k, :n_ref, :4], boxes)
    a3.close()


def _analyzer_with_thresholds(thr):
    """Analyzer whose trl_config_t carries other MTCNN thresholds (the reference's are fixed literals)."""
    import ctypes as C
    from truely_b200 import _lib as L
    from truely_b200 import weights as W
    from truely_b200.model import Analyzer
    an = Analyzer.__new__(Analyzer)
    an.torch = torch
    an.lib = L.load()
    an.device = 0
    cfg = L.Config()
    an.lib.trl_default_config(C.byref(cfg))
    cfg.thresholds[0], cfg.thresholds[1], cfg.thresholds[2] = thr
    cfg.facenet_impl = 1
    mt, _ = W.load_mtcnn_state()
    fn, _ = W.load_facenet_state()
    blobs = [W.pack_mtcnn(mt, "pnet"), W.pack_mtcnn(mt, "rnet"), W.pack_mtcnn(mt, "onet"), W.pack_facenet(fn)]
    w = L.Weights()
    fp = C.POINTER(C.c_float)
    w.h_pnet, w.pnet_len = blobs[0].ctypes.data_as(fp), blobs[0].size
    w.h_rnet, w.rnet_len = blobs[1].ctypes.data_as(fp), blobs[1].size
    w.h_onet, w.onet_len = blobs[2].ctypes.data_as(fp), blobs[2].size
    w.h_facenet, w.facenet_len = blobs[3].ctypes.data_as(fp), blobs[3].size
    ctx = C.c_void_p()
    rc = an.lib.trl_create(0, C.byref(w), C.byref(cfg), C.byref(ctx))
    assert rc == 0, an.lib.trl_last_error(None)
    an.ctx, an.cfg = ctx, cfg
    an.stream = torch.cuda.Stream(device=0)
    an.box_cap, an.crop_size = cfg.box_cap_frame, 80
    return an


def test_capacity_overflow_is_reported_not_truncated():
    import ctypes as C
    from truely_b200 import _lib as L
    from truely_b200.model import Analyzer
    an = Analyzer(device=0, facenet_impl=1, cand_cap_scale=2, cand_cap_frame=2, box_cap_frame=2)
    clip = SyntheticClip(540, 960, 60, 8, n_faces=(4, 4), face_h=(50.0, 200.0), seed=15)
    with pytest.raises(L.TrlError) as e:
        an.process_frames(np.stack([clip.frame(0)]), detail=True)
    assert e.value.code == L.TRL_E_CAPACITY
    an.close()


def test_uncapped_cascade_matches_oracle_beyond_the_smem_nms_limit():
    """upstream detect_face has no candidate cap.  With the P-Net threshold lowered to 0.01 every cell
    of the finest level is a candidate (6256 > 2048 in one NMS group): the default capacities report TRL_E_CAPACITY, the
    raised ones (trl_set_capacity -> global-memory NMS) reproduce the oracle's lists exactly."""
    import copy
    from truely_b200 import _lib as L
    thr = (0.01, 0.7, 0.7)
    mt2 = copy.deepcopy(H.oracle_mtcnn())
    mt2.thresholds = list(thr)
    an = _analyzer_with_thresholds(thr)
    clip = SyntheticClip(240, 320, 30, 8, n_faces=(1, 1), face_h=(70.0, 110.0), seed=31)
    frame = clip.frame(3)
    with pytest.raises(L.TrlError) as e:
        an.process_frames(frame[None], detail=True)
    assert e.value.code == L.TRL_E_CAPACITY
    an.set_capacity(*M.Analyzer.BIG_CAPS)
    res = an.process_frames(frame[None], detail=True)
    tr = {}
    boxes, _ = mt2.detect(frame, trace=tr)
    n_ref = 0 if boxes is None else len(boxes)
    assert len(tr["s1_boxes"]) > 1024, "the case must exceed the fast path's per-frame capacity to mean anything"
    assert res.counts[0, 1] == len(tr["s1_boxes"]), "R-Net input count"
    assert res.counts[0, 2] == len(tr["s2_boxes"]), "O-Net input count"
    assert res.nfaces[0] == n_ref
    if n_ref:
        _match_boxes(res.boxes[0, :n_ref, :4], boxes)
    an.close()


def test_stream_retries_capacity_overflow_instead_of_aborting(analyzer):
    """analyze_stream on a context whose capacities are far too small: every chunk overflows, is redone through the
    uncapped path (frame by frame, halo re-chained) and the trace equals the one of the normally sized context."""
    from truely_b200.model import Analyzer
    clip = SyntheticClip(360, 640, 30, 96, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=1.0, seed=12)
    frames = [f for f in clip]
    ref = M.analyze_stream(iter(frames), 30, 640, 360, analyzer=analyzer, chunk=5, keep_emb=True)
    small = Analyzer(device=0, cand_cap_scale=2, cand_cap_frame=2, box_cap_frame=2)
    got = M.analyze_stream(iter(frames), 30, 640, 360, analyzer=small, chunk=5, keep_emb=True)
    assert got.timings["capacity_retries"] >= 1 and ref.timings["capacity_retries"] == 0
    assert (small.cfg.cand_cap_scale, small.cfg.cand_cap_frame, small.cfg.box_cap_frame) == (2, 2, 2)    # restored
    assert got.valid == ref.valid and got.nfaces == ref.nfaces and sum(ref.valid) > 15
    assert all(np.array_equal(a, b) for a, b in zip(got.box, ref.box))
    assert all(np.array_equal(a, b) for a, b in zip(got.emb, ref.emb))
    assert got.sim == ref.sim and got.flagged == ref.flagged and got.score == ref.score
    small.close()


def test_golden_run_flagged_set_and_score(analyzer):
    """The whole hot loop on the golden clip vs the committed oracle trace."""
    g = np.load(os.path.join(GOLD, "reference_run.npz"))
    clip = _mg().golden_clip()
    tr = M.analyze_stream(iter(clip), clip.fps, clip.width, clip.height, writer=None, analyzer=analyzer, chunk=16, keep_emb=True)
    assert tr.frame_count == int(g["frame_count"]) and tr.frame_index == list(g["frame_index"])
    assert tr.nfaces == list(g["n_faces"])
    assert tr.valid == [bool(v) for v in g["embedded"]]
    for k in range(len(tr.frame_index)):
        if not tr.valid[k]:
            continue
        assert np.abs(tr.box[k] - g["box"][k]).max() <= 1, f"frame {k} box"
        # the crop may differ by one pixel when a float box lands on the other side of an integer; cosine bar still holds
        assert H.cosine(tr.emb[k], g["emb"][k]) >= 0.999, f"frame {k} embedding"
        if not np.isnan(g["sim"][k]):
            assert abs(tr.sim[k] - g["sim"][k]) < 1e-3, f"frame {k} sim {tr.sim[k]} vs {g['sim'][k]}"
    ref_sim = [None if np.isnan(v) else float(v) for v in g["sim"]]
    H.assert_flags_match_outside_band(tr.valid, tr.sim, tr.flagged, tr.score, tr.frame_count, tr.stride, clip.fps,
                                      [bool(v) for v in g["embedded"]], ref_sim, [bool(v) for v in g["flagged"]], int(g["score"]))


def test_run_dropin_on_encoded_clip_matches_oracle(analyzer, tmp_path):
    """run(video_path_one, video_path_two) on an mp4 (decode path) vs the oracle loop on the same decoded frames."""
    clip = SyntheticClip(240, 320, 30, 140, n_faces=(1, 1), face_h=(70.0, 110.0), jitter=1.6, seed=31)
    src = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 30, (320, 240))
    assert wr.isOpened()
    for f in clip:
        wr.write(f)
    wr.release()
    dst = str(tmp_path / "clip_output.mp4")
    M._ANALYZER = analyzer
    score = M.run(src, dst)
    assert isinstance(score, int) and 0 <= score <= 100
    assert os.path.exists(dst) and os.path.getsize(dst) > 0            # server/server.py:612-627 rejects a missing file
    cap = cv2.VideoCapture(dst)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 140
    cap.release()
    # oracle on the same decoded frames
    cap = cv2.VideoCapture(src)
    frames = []
    while True:
        ok_, f = cap.read()
        if not ok_:
            break
        frames.append(f)
    cap.release()
    ref = reference_run_frames(iter([f.copy() for f in frames]), 30, 320, 240, H.oracle_mtcnn(), H.oracle_facenet())
    tr = M.analyze_stream(iter(frames), 30, 320, 240, analyzer=analyzer, keep_emb=True)
    assert tr.frame_count == ref.frame_count
    for k, f in enumerate(ref.frames):
        assert tr.nfaces[k] == f.n_faces
        assert tr.valid[k] == f.embedded
        if f.embedded:
            assert H.cosine(tr.emb[k], f.emb) >= 0.999
            if f.sim is not None:
                assert abs(tr.sim[k] - f.sim) < 1e-3
    H.assert_flags_match_outside_band(tr.valid, tr.sim, tr.flagged, tr.score, tr.frame_count, tr.stride, 30,
                                      [f.embedded for f in ref.frames], [f.sim for f in ref.frames],
                                      [f.flagged for f in ref.frames], ref.score)
    assert score == tr.score


def test_fused_process_equals_staged_calls(analyzer):
    clip = SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=1.0, seed=77)
    frames = np.stack([clip.frame(i) for i in range(0, 40, 4)])
    a = analyzer.process_frames(frames, detail=True)
    b = analyzer.process_frames(frames, detail=False)
    assert np.array_equal(a.nfaces, b.nfaces) and np.array_equal(a.box, b.box) and np.array_equal(a.valid, b.valid)
    assert np.array_equal(a.emb, b.emb) and np.array_equal(a.below, b.below)
    assert np.allclose(a.sim, b.sim, equal_nan=True)


def test_chunking_and_host_buffer_path_do_not_change_results(analyzer):
    """size-independent properties at the bench's frame size: the cascade is per-frame independent, so any chunking of
    the batch and the pinned-host (H2D overlapped) path must give bit-identical boxes, crops, embeddings and flags"""
    an = analyzer
    clip = SyntheticClip(720, 1280, 30, 240, n_faces=(1, 1), seed=3)
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()[:40]])
    d = torch.from_numpy(frames).cuda()
    out0 = an.analyze_resident(d, chunk=40, pipeline=False)
    torch.cuda.synchronize()                       # the analyzer works on its own stream
    ref = {k: out0[k][:40].clone() for k in ("box", "valid", "emb", "sim", "below", "has_sim", "crops")}
    torch.cuda.synchronize()
    pinned = torch.from_numpy(frames).pin_memory()
    stage = torch.empty((3, 16, 720, 1280, 3), dtype=torch.uint8, device="cuda")
    # pipeline=True (the default): the tail of chunk k runs on the library's internal stream under the pyramid of chunk
    # k+1 (trl_detect_align_async); it must not change a bit either, with ragged last chunks and growing chunk sizes
    for kwargs in (dict(chunk=7, pipeline=False), dict(chunk=7), dict(chunk=13), dict(chunk=40),
                   dict(chunk=16, h2d=True, dev_frames=stage, pipeline=False), dict(chunk=16, h2d=True, dev_frames=stage),
                   dict(chunk=9, h2d=True, dev_frames=stage[:2, :9])):
        src = pinned if kwargs.get("h2d") else d
        out = an.analyze_resident(src, **kwargs)
        an.stream.synchronize()
        torch.cuda.synchronize()
        for k, v in ref.items():
            a, b = out[k][:40], v[:40]
            if a.is_floating_point():          # sim of a frame without a predecessor is NaN by design
                a, b = torch.nan_to_num(a, nan=-2.0), torch.nan_to_num(b, nan=-2.0)
            assert torch.equal(a, b), f"{k} differs with {list(kwargs)}"
    assert int(ref["valid"][:40].sum()) > 30          # the property is not vacuous: faces were found


def test_1080p_multi_face_matches_oracle(analyzer):
    """BASELINE.json configs[3] shape: 1080p, 4-8 faces per frame, 12 pyramid scales -- same face count, IoU >= 0.95"""
    an = analyzer
    clip = SyntheticClip(1080, 1920, 60, 64, n_faces=(4, 8), face_h=(60.0, 300.0), seed=9)
    frames = np.stack([clip.frame(i) for i in (0, 24, 48)])
    res = an.process_frames(frames, detail=True)
    mt = H.oracle_mtcnn()
    for i, f in enumerate(frames):
        boxes, _ = mt.detect(f)
        n_ref = 0 if boxes is None else len(boxes)
        assert int(res.counts[i, 3]) == n_ref, f"frame {i}: {int(res.counts[i, 3])} faces, oracle {n_ref}"
        if n_ref:
            _match_boxes(res.boxes[i, :n_ref, :4], boxes)


def test_odd_frame_shapes_match_oracle(analyzer):
    """Frame widths whose rows are neither 16- nor 4-byte aligned (3 W % 4 != 0) take the byte-aligned pyramid path and
    the cp.async P-Net staging; the cascade must still agree with the oracle (the reference accepts any frame size)."""
    mt = H.oracle_mtcnn()
    for (h, w, seed) in ((363, 641, 21), (301, 403, 22)):
        clip = SyntheticClip(h, w, 30, 40, n_faces=(1, 2), face_h=(70.0, 120.0), seed=seed)
        frames = [clip.frame(i) for i in (0, 9, 23)]
        res = analyzer.process_frames(np.stack(frames), detail=True)
        for k, f in enumerate(frames):
            boxes, _ = mt.detect(f)
            n_ref = 0 if boxes is None else len(boxes)
            assert res.nfaces[k] == n_ref, f"{h}x{w} frame {k}: {res.nfaces[k]} faces vs oracle {n_ref}"
            if n_ref:
                _match_boxes(res.boxes[k, :n_ref, :4], boxes)


def test_faceless_and_tiny_frames(analyzer):
    """mtcnn.detect returns (None, [None]) on frames without a face; the reference then skips the frame
    (server/model.py:48): no embedding, no comparison, the run-length counter does not move, score 0.  On frames too small
    for a single pyramid level (min side * 0.6 < 12) upstream detect_face raises (torch.cat of an empty list), and so does
    the oracle; the library reports TRL_E_INVALID instead of launching anything."""
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, size=(6, 360, 640, 3), dtype=np.uint8)
    res = analyzer.process_frames(noise)
    assert not res.valid.any() and (res.nfaces == 0).all()
    d = torch.from_numpy(noise).cuda()
    out = analyzer.analyze_resident(d, chunk=4)
    analyzer.stream.synchronize()
    assert int(out["valid"][:6].sum()) == 0 and int(out["has_sim"][:6].sum()) == 0 and int(out["below"][:6].sum()) == 0
    score, flagged, _ = M.score_from_flags(out["valid"][:6].cpu().numpy(), out["has_sim"][:6].cpu().numpy(),
                                           out["below"][:6].cpu().numpy(), 24, 30, 4)
    assert score == 0 and not any(flagged)
    tiny = rng.integers(0, 256, size=(3, 18, 31, 3), dtype=np.uint8)       # 18 * 0.6 = 10.8 < 12: zero pyramid levels
    with pytest.raises((RuntimeError, ValueError)):
        H.oracle_mtcnn().detect(tiny[0])
    from truely_b200._lib import TrlError
    with pytest.raises(TrlError) as ei:
        analyzer.process_frames(tiny)
    assert ei.value.code == -1 and "too small" in str(ei.value)
    res = analyzer.process_frames(noise[:2])                               # the context is still usable afterwards
    assert (res.nfaces == 0).all()


def test_full_size_run_is_deterministic_and_schedule_independent(analyzer):
    """BASELINE.json configs[1] at full size (450 processed 720p frames): two runs give identical bits, and neither the
    number of frames per cascade call nor the two-stream schedule changes boxes, crops, embeddings or flags."""
    an = analyzer
    clip = SyntheticClip(720, 1280, 30, 1800, n_faces=(1, 1), jitter=1.2, seed=0)
    idx = clip.processed_indices()
    assert len(idx) == 450
    d = torch.empty((450, 720, 1280, 3), dtype=torch.uint8, device="cuda")
    for k, i in enumerate(idx):
        d[k].copy_(torch.from_numpy(clip.frame(i)))
    keys = ("box", "valid", "crops", "emb", "sim", "below", "has_sim")

    def run(**kw):
        out = an.analyze_resident(d, **kw)
        an.stream.synchronize()
        torch.cuda.synchronize()
        return {k: torch.nan_to_num(out[k][:450].clone().float(), nan=-2.0) for k in keys}

    ref = run(chunk=90, pipeline=False)
    for kw in (dict(chunk=90, pipeline=False), dict(chunk=90), dict(chunk=225), dict(chunk=450)):
        got = run(**kw)
        for k in keys:
            assert torch.equal(got[k], ref[k]), f"{k} differs with {kw}"
    assert int(ref["valid"].sum()) >= 440 and int(ref["below"].sum()) > 0       # faces found, comparisons on both sides of 0.99


def test_host_staging_buffers(analyzer):
    """trl_host_alloc / trl_host_free: page-locked (write-combined) staging memory is a valid source of asynchronous copies
    and the whole host-buffer path gives the same bits from it as from torch's pinned memory."""
    from truely_b200.model import staging_empty
    clip = SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(90.0, 130.0), seed=4)
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()[:12]])
    stage = torch.empty((2, 5, 360, 640, 3), dtype=torch.uint8, device="cuda")
    outs = []
    for wc in (None, True, False):
        if wc is None:
            src = torch.from_numpy(frames).pin_memory()
        else:
            src = staging_empty(torch, frames.shape, write_combined=wc)
            src.copy_(torch.from_numpy(frames))
            assert src.is_pinned()
        out = analyzer.analyze_resident(src, chunk=5, h2d=True, dev_frames=stage)
        analyzer.stream.synchronize()
        torch.cuda.synchronize()
        outs.append({k: torch.nan_to_num(out[k][:12].clone().float(), nan=-2.0) for k in ("box", "valid", "emb", "sim", "below")})
        del src
    for o in outs[1:]:
        for k, v in outs[0].items():
            assert torch.equal(o[k], v), k
    assert int(outs[0]["valid"].sum()) >= 10
    import ctypes as C
    lib = analyzer.lib
    p = C.c_void_p()
    assert lib.trl_host_alloc(0, 1, C.byref(p)) == -1                       # TRL_E_INVALID
    assert lib.trl_host_alloc(1 << 20, 1, C.byref(p)) == 0 and p.value
    assert lib.trl_host_free(p) == 0


BUNDLED = os.path.join(GOLD, "bundled_veo3_360p.mp4")


def test_bundled_clip_matches_oracle_golden(analyzer):
    """BASELINE.json configs[0]: the reference's bundled clip (640x360 h264, 960 frames -> 240 processed) through
    run_trace(), against the oracle's committed trace of the same OpenCV decode (tests/golden/make_bundled_golden.py).
    North-star tolerances: same face count per frame, box within one pixel of the oracle's truncated box (IoU >= 0.95),
    embedding cosine >= 0.999, sims within 1e-3, identical flags and score outside the 1e-3 band."""
    g = np.load(os.path.join(GOLD, "bundled_clip_run.npz"))
    tr = M.run_trace(BUNDLED, None, analyzer=analyzer, keep_emb=True)
    assert tr.frame_count == int(g["frame_count"]) == 960 and tr.stride == 4
    assert tr.frame_index == list(g["frame_index"]) and len(tr.frame_index) == 240
    assert tr.nfaces == list(g["n_faces"]), "face count per frame"
    assert tr.valid == [bool(v) for v in g["embedded"]]
    assert sum(tr.valid) == 60 and max(tr.nfaces) == 4                      # the clip exercises multi-candidate frames
    for k in range(240):
        if not tr.valid[k]:
            continue
        assert H.box_iou(tr.box[k].astype(np.float64), g["box"][k].astype(np.float64)) >= 0.95, f"frame {k} box"
        assert np.abs(tr.box[k] - g["box"][k]).max() <= 1, f"frame {k} box"
        assert H.cosine(tr.emb[k], g["emb"][k]) >= 0.999, f"frame {k} embedding cosine"
        if not np.isnan(g["sim"][k]):
            assert abs(tr.sim[k] - float(g["sim"][k])) < 1e-3, f"frame {k} sim {tr.sim[k]} vs {g['sim'][k]}"
    ref_sim = [None if np.isnan(v) else float(v) for v in g["sim"]]
    H.assert_flags_match_outside_band(tr.valid, tr.sim, tr.flagged, tr.score, tr.frame_count, tr.stride, 30,
                                      [bool(v) for v in g["embedded"]], ref_sim, [bool(v) for v in g["flagged"]], int(g["score"]))


class _FrameSink:
    """stands in for cv2.VideoWriter: keeps the annotated frames as they would be handed to the encoder"""

    def __init__(self):
        self.frames = []

    def write(self, f):
        self.frames.append(f.copy())


def test_annotated_frames_equal_the_oracle_loop(analyzer):
    """a12 (server/model.py:66-77): every frame is written, in order, and the processed ones carry the reference's
    overlay.  The stream handed to the writer is compared pixel for pixel with reference_run_frames(annotate=True) on the
    same frames (a jittery clip long enough to pass the 15-frame run: both overlay kinds occur)."""
    clip = SyntheticClip(240, 320, 30, 140, n_faces=(1, 1), face_h=(70.0, 110.0), jitter=1.6, seed=31)
    frames = [f for f in clip]
    ref_sink, got_sink = _FrameSink(), _FrameSink()
    ref = reference_run_frames(iter([f.copy() for f in frames]), 30, 320, 240, H.oracle_mtcnn(), H.oracle_facenet(),
                               writer=ref_sink, annotate=True)
    tr = M.analyze_stream(iter([f.copy() for f in frames]), 30, 320, 240, writer=got_sink, analyzer=analyzer, chunk=8)
    assert len(got_sink.frames) == len(ref_sink.frames) == 140
    assert sum(f.flagged for f in ref.frames) > 0 and sum(f.sim is not None and not f.flagged for f in ref.frames) > 0
    same_box = 0
    for k, f in enumerate(ref.frames):
        i = f.frame_index
        if f.embedded and np.array_equal(tr.box[k], f.box) and tr.flagged[k] == f.flagged:
            same_box += 1
            assert np.array_equal(got_sink.frames[i], ref_sink.frames[i]), f"frame {i}: annotated pixels differ"
        elif f.embedded:
            assert np.abs(tr.box[k] - f.box).max() <= 1
    assert same_box >= 0.8 * sum(f.embedded for f in ref.frames)
    processed = {f.frame_index for f in ref.frames}
    for i in range(140):
        if i not in processed:
            assert np.array_equal(got_sink.frames[i], frames[i]), f"frame {i} must pass through untouched"


def test_device_overlay_stream_equals_the_host_overlay(analyzer):
    """SURVEY.md 8f (device overlay): with device_overlay=True the boxes and captions are drawn by trl_overlay on the
    resident frames; the stream handed to the writer is the same, pixel for pixel, as with OpenCV on the host copies
    (240x320 frames: the 'AI Detected - Frame n' caption does not fit, so the pending-caption hand-back is exercised too;
    480x640: it fits and is drawn on the device)."""
    for (h, w, seed) in ((240, 320, 31), (480, 640, 33)):
        clip = SyntheticClip(h, w, 30, 140, n_faces=(1, 1), face_h=(0.3 * h, 0.45 * h), jitter=1.6, seed=seed)
        frames = [f for f in clip]
        host_sink, dev_sink = _FrameSink(), _FrameSink()
        tr_h = M.analyze_stream(iter([f.copy() for f in frames]), 30, w, h, writer=host_sink, analyzer=analyzer, chunk=8,
                                device_overlay=False)
        tr_d = M.analyze_stream(iter([f.copy() for f in frames]), 30, w, h, writer=dev_sink, analyzer=analyzer, chunk=8,
                                device_overlay=True)
        assert tr_h.flagged == tr_d.flagged and tr_h.score == tr_d.score
        assert sum(tr_h.flagged) > 0 and sum(s is not None and not f for s, f in zip(tr_h.sim, tr_h.flagged)) > 0
        assert len(host_sink.frames) == len(dev_sink.frames) == 140
        for i in range(140):
            assert np.array_equal(host_sink.frames[i], dev_sink.frames[i]), f"{h}x{w} frame {i}: overlays differ"

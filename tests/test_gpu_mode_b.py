"""GPU: mode B (SURVEY.md 8f row 4, Appendix A "Mode-B extras") -- upstream extract_face (INTER_AREA, margin),
fixed_image_standardization and keep_all embedding -- against oracle/mode_b.py, through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H
from oracle import mode_b as OB
from truely_b200 import model as M
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu
vp = M._vp


@pytest.fixture(scope="module")
def analyzer_b():
    an = M.Analyzer(device=0, mode="b")
    yield an
    an.close()


def _extract(an, frames, boxes, S, margin):
    """trl_extract_face on frames [B,H,W,3] with one box per frame."""
    B, Hh, Ww, _ = frames.shape
    d = torch.from_numpy(frames).cuda()
    bx = torch.zeros((B, 5), dtype=torch.float32, device="cuda")
    bx[:, :4] = torch.from_numpy(np.asarray(boxes, np.float32)).cuda()
    nf = torch.ones(B, dtype=torch.int32, device="cuda")
    box_int = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    valid = torch.zeros(B, dtype=torch.uint8, device="cuda")
    crops = torch.zeros((B, S, S, 3), dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(an.stream):
        an._check(an.lib.trl_extract_face(an.ctx, vp(d), B, Hh, Ww, vp(bx), 5, vp(nf), S, margin, vp(box_int), vp(valid), vp(crops),
                                          an._sptr()))
    an.stream.synchronize()
    return box_int.cpu().numpy(), valid.cpu().numpy(), crops.cpu().numpy()


@pytest.mark.parametrize("S,margin", [(160, 0), (160, 32), (80, 0)])
def test_extract_face_is_bit_exact_with_opencv_inter_area(analyzer, S, margin):
    rng = np.random.default_rng(7)
    Hh, Ww = 720, 1280
    frame = rng.integers(0, 256, (Hh, Ww, 3), dtype=np.uint8)
    boxes = [
        [100.0, 50.0, 420.0, 370.0],       # 320 x 320: exactly 2 x 2 blocks at S = 160 (rounding special case)
        [10.0, 20.0, 490.0, 340.0],        # 480 x 320: 3 x 2 integer ratios
        [0.0, 0.0, 640.0, 640.0],          # 4 x 4
        [200.3, 100.7, 461.2, 361.9],      # general area path
        [333.9, 12.1, 555.5, 700.2],       # tall
        [600.5, 300.5, 660.5, 371.5],      # smaller than S on both axes: enlarging path
        [700.0, 100.0, 790.0, 500.0],      # mixed: one axis enlarged, one reduced
        [-30.5, -12.0, 250.0, 300.0],      # clipped at the top-left corner
        [1100.0, 500.0, 1400.0, 900.0],    # clipped at the bottom-right corner
        [640.0, 360.0, 801.0, 521.0],      # 161 x 161: ratio just above one
        [900.2, 10.0, 900.9, 400.0],       # empty after truncation
        [50.0, 600.0, 213.0, 719.9],
    ]
    rng2 = np.random.default_rng(8)
    for _ in range(20):
        x1, y1 = rng2.uniform(-20, Ww - 40), rng2.uniform(-20, Hh - 40)
        boxes.append([x1, y1, x1 + rng2.uniform(15, 500), y1 + rng2.uniform(15, 500)])
    frames = np.broadcast_to(frame, (len(boxes), Hh, Ww, 3)).copy()
    box_int, valid, crops = _extract(analyzer, frames, boxes, S, margin)
    n_paths = {"empty": 0, "ok": 0}
    for k, b in enumerate(boxes):
        face, bb = OB.extract_face(frame, np.asarray(b, np.float32), S, margin)
        assert list(box_int[k]) == bb, f"box {k}: {list(box_int[k])} vs {bb}"
        if face is None:
            assert valid[k] == 0 and not crops[k].any()
            n_paths["empty"] += 1
        else:
            assert valid[k] == 1
            assert np.array_equal(crops[k], face), f"box {k} {b}: max diff {np.abs(crops[k].astype(int) - face.astype(int)).max()}"
            n_paths["ok"] += 1
    assert n_paths["empty"] >= 1 and n_paths["ok"] >= 25


def test_mode_b_embedding_matches_oracle(analyzer):
    """fixed_image_standardization folded into the stem: cosine >= 0.999 vs the fp32 oracle on standardised 160x160 crops."""
    clip = SyntheticClip(360, 640, 30, 40, n_faces=(1, 1), face_h=(120.0, 200.0), jitter=1.0, seed=41)
    fn = H.oracle_facenet()
    crops = []
    for i in (0, 5, 11, 23):
        f = clip.frame(i)
        box = clip.faces(i)[0].box()
        face, _ = OB.extract_face(f, np.asarray(box, np.float32), 160, 0)
        crops.append(face)
    crops = np.stack(crops)
    d = torch.from_numpy(crops).cuda()
    emb = torch.empty((len(crops), 512), dtype=torch.float32, device="cuda")
    with torch.cuda.stream(analyzer.stream):
        analyzer._check(analyzer.lib.trl_facenet_norm(analyzer.ctx, vp(d), len(crops), 160, 1, vp(emb), analyzer._sptr()))
    analyzer.stream.synchronize()
    got = emb.cpu().numpy()
    for k, face in enumerate(crops):
        with torch.no_grad():
            ref = fn(OB.fixed_image_standardization(face).unsqueeze(0)).numpy().ravel()
        assert H.cosine(got[k], ref) >= 0.999, f"crop {k}"
        assert abs(np.linalg.norm(got[k]) - 1.0) < 1e-4


def test_mode_b_whole_loop_matches_oracle(analyzer_b):
    """Analyzer(mode='b'): detect -> extract_face(160, margin 0) -> standardise -> FaceNet -> consistency, against the same
    loop on the oracle (the reference's run loop with the crop of server/model.py:55-58 replaced by upstream's)."""
    clip = SyntheticClip(360, 640, 30, 120, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=1.2, seed=19)
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()])
    res = analyzer_b.process_frames(frames, detail=False)
    mt, fn = H.oracle_mtcnn(), H.oracle_facenet()
    prev, n_cmp = None, 0
    for k, f in enumerate(frames):
        boxes, _ = mt.detect(f)
        has = boxes is not None and len(boxes) > 0
        face, bb = (OB.extract_face(f, boxes[0], 160, 0) if has else (None, None))
        assert bool(res.valid[k]) == (face is not None)
        if face is None:
            continue
        assert np.abs(res.box[k] - np.asarray(bb)).max() <= 1
        with torch.no_grad():
            e = fn(OB.fixed_image_standardization(face).unsqueeze(0)).numpy().ravel()
        assert H.cosine(res.emb[k], e) >= 0.999, f"frame {k}"
        if prev is not None:
            sim = float(np.dot(e, prev) / (np.linalg.norm(e) * np.linalg.norm(prev)))
            assert abs(float(res.sim[k]) - sim) < 1e-3, f"frame {k}: sim {res.sim[k]} vs {sim}"
            n_cmp += 1
        prev = e
    assert n_cmp >= 25


def test_keep_all_embeds_every_face(analyzer_b):
    """keep_all on 1080p frames with 4-8 faces (BASELINE.json configs[3]): every box the cascade returns is cropped and
    embedded; face order = the detector's (largest first); prefix offsets partition the batch."""
    clip = SyntheticClip(1080, 1920, 60, 64, n_faces=(4, 8), face_h=(60.0, 300.0), seed=9)
    frames = np.stack([clip.frame(i) for i in (0, 24)])
    got = analyzer_b.embed_all_faces(frames)
    mt, fn = H.oracle_mtcnn(), H.oracle_facenet()
    total = 0
    for k, f in enumerate(frames):
        boxes, _ = mt.detect(f)
        n_ref = 0 if boxes is None else len(boxes)
        assert len(got[k]) == n_ref and n_ref >= 4
        ref = OB.embed_faces(f, boxes, fn, 160, 0)
        for (gb, ge), (rb, re_) in zip(got[k], ref):
            assert np.abs(np.asarray(gb) - np.asarray(rb)).max() <= 1
            if re_ is None:
                assert ge is None
            elif np.array_equal(np.asarray(gb), np.asarray(rb)):
                assert H.cosine(ge, re_) >= 0.999
            total += 1
    assert total >= 8
    from truely_b200 import _lib as L
    with pytest.raises(L.TrlError) as e:                                 # more faces than the caller made room for
        analyzer_b.embed_all_faces(frames, max_faces=3)
    assert e.value.code == L.TRL_E_CAPACITY

"""CPU: the oracle against its closures, the committed golden vectors and numpy restatements of the exact arithmetic."""
import os

import cv2
import numpy as np
import pytest
import torch

import helpers as H
from oracle import mtcnn as OM
from oracle.inception_resnet_v1 import InceptionResnetV1
from oracle.reference_run import final_score, frame_stride, reference_run_frames
from truely_b200 import weights as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_counts_close_the_architecture():
    # the only cross-checks available without facenet_pytorch
    c = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert c(OM.PNet()) == 6632
    assert c(OM.RNet()) == 100178
    assert c(OM.ONet()) == 389040
    assert c(InceptionResnetV1()) == 23482624


def test_upstream_state_dict_names_load_strict():
    state, _ = W.load_mtcnn_state()
    for net, cls in (("pnet", OM.PNet), ("rnet", OM.RNet), ("onet", OM.ONet)):
        sd = {k[len(net) + 1:]: torch.from_numpy(v) for k, v in state.items() if k.startswith(net + ".")}
        cls().load_state_dict(sd, strict=True)
        assert [k for k, _ in W.MTCNN_SHAPES[net]] == list(sd.keys())


def test_pyramid_scales_match_survey_counts():
    # 9 / 11 / 12 scales and total pyramid pixels at 360p / 720p / 1080p
    for (h, w), n, px in (((360, 640), 9, 167785), ((720, 1280), 11, 669638), ((1080, 1920), 12, 1504462)):
        s = OM.pyramid_scales(h, w)
        assert len(s) == n
        assert sum(int(h * k + 1) * int(w * k + 1) for k in s) == px


def test_area_resample_is_two_divisions():
    """imresample == (integer window sum / kh) / kw in fp32: the form the CUDA kernels implement bit-exactly."""
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (1, 97, 131, 3), dtype=np.uint8)
    t = torch.as_tensor(img).permute(0, 3, 1, 2).type(torch.float32)
    for oh, ow in ((59, 79), (13, 18), (24, 24), (120, 140)):
        ref = OM.imresample(t, (oh, ow))[0].numpy()
        f = img[0].astype(np.int64)
        out = np.zeros((3, oh, ow), np.float32)
        for i in range(oh):
            y0, y1 = (i * 97) // oh, -((-(i + 1) * 97) // oh)
            for j in range(ow):
                x0, x1 = (j * 131) // ow, -((-(j + 1) * 131) // ow)
                s = f[y0:y1, x0:x1].sum(axis=(0, 1)).astype(np.float32)
                out[:, i, j] = (s / np.float32(y1 - y0)) / np.float32(x1 - x0)
        assert np.array_equal(out, ref)


def emu_cv2_resize_linear_u8(src, dw, dh):
    """numpy restatement of cv2.resize(INTER_LINEAR) on uint8 (what crop_align_kernel implements)."""
    sh, sw = src.shape[:2]

    def coefs(ssize, dsize, clamp):
        scale = 1.0 / (dsize / ssize)
        idx, a0, a1 = np.zeros(dsize, np.int64), np.zeros(dsize, np.int64), np.zeros(dsize, np.int64)
        for d in range(dsize):
            fx = np.float32((d + 0.5) * scale - 0.5)
            sx = int(np.floor(fx))
            fx = np.float32(fx - sx)
            if clamp:
                if sx < 0:
                    fx, sx = np.float32(0), 0
                if sx >= ssize - 1:
                    fx, sx = np.float32(0), ssize - 1
            idx[d] = sx
            a1[d] = int(np.rint(np.float32(fx * np.float32(2048))))
            a0[d] = int(np.rint(np.float32((np.float32(1) - fx) * np.float32(2048))))
        return idx, a0, a1

    xi, xa0, xa1 = coefs(sw, dw, True)
    yi, ya0, ya1 = coefs(sh, dh, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    hp = s[:, xi, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    out = (((ya0[:, None, None] * (hp[y0] >> 4)) >> 16) + ((ya1[:, None, None] * (hp[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def test_cv2_resize_restatement_is_bit_exact():
    rng = np.random.default_rng(1)
    for t in range(120):
        h, w = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        if t % 5 == 0:
            h = int(rng.integers(1, 20))
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (80, 80)), emu_cv2_resize_linear_u8(src, 80, 80)), (h, w)


def test_generate_bounding_box_division_is_fp32():
    """(2*c + 1) / scale is evaluated in float32 with scale rounded to float32 (pnet.cu relies on it)."""
    for scale in OM.pyramid_scales(720, 1280):
        c = torch.arange(0, 700, dtype=torch.float32)
        ref = ((2 * c + 1) / scale).floor().numpy()
        mine = np.floor((2 * c.numpy() + np.float32(1)) / np.float32(scale))
        assert np.array_equal(ref, mine)
        ref2 = ((2 * c + 12 - 1 + 1) / scale).floor().numpy()
        mine2 = np.floor((2 * c.numpy() + np.float32(12)) / np.float32(scale))
        assert np.array_equal(ref2, mine2)


def test_score_state_machine_cases():
    # server/model.py:83-95 literal cases
    assert final_score(0, 0, 0, 30, 4) == 0
    assert final_score(0, 0, 100, 30, 4) == 0
    assert frame_stride(30) == 4 and frame_stride(60) == 8 and frame_stride(24) == 3 and frame_stride(29) == 4 and frame_stride(5) == 1
    # 960 frames, 240 processed, 100 flagged, final run 20: pct 41.67, conf min(41.67*1.333,100)=55.6, weight .5 (960 > 900)
    assert final_score(100, 20, 960, 30, 4) == int(min(100 / 240 * 100 + min(100 / 240 * 100 * (20 / 15), 100) * 0.5, 100))
    assert final_score(240, 240, 960, 30, 4) == 100
    assert final_score(10, 0, 200, 30, 4) == int(10 / 50 * 100)


def _golden_clip():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_golden_detection():
    mg = _golden_clip()
    g = np.load(os.path.join(GOLD, "mtcnn_detect.npz"))
    mt = H.oracle_mtcnn()
    boxes, probs = mt.detect(mg.golden_clip().frame(0))
    assert len(boxes) == len(g["a_boxes"])
    assert np.allclose(boxes, g["a_boxes"], atol=2e-2)      # fp32 conv summation order may differ across hosts
    assert np.allclose(probs, g["a_probs"], atol=1e-4)
    assert np.allclose(OM.pyramid_scales(360, 640), g["a_scales"], rtol=0, atol=0)


def test_oracle_reproduces_golden_run_prefix():
    """First 40 frames of the golden clip: boxes, embeddings and similarities of the oracle loop."""
    mg = _golden_clip()
    g = np.load(os.path.join(GOLD, "reference_run.npz"))
    clip = mg.golden_clip()
    frames = (clip.frame(i) for i in range(40))
    tr = reference_run_frames(frames, clip.fps, clip.width, clip.height, H.oracle_mtcnn(), H.oracle_facenet())
    assert len(tr.frames) == 10
    for k, f in enumerate(tr.frames):
        assert f.n_faces == g["n_faces"][k]
        assert f.embedded == bool(g["embedded"][k])
        if f.embedded:
            assert np.abs(f.box - g["box"][k]).max() <= 1
            assert H.cosine(f.emb, g["emb"][k]) > 0.99999
            if f.sim is not None:
                assert abs(f.sim - g["sim"][k]) < 2e-5
    # whole-clip integers recorded in the fixture
    assert int(g["score"]) == final_score(int(g["flagged_count"]), int(g["final_run"]), int(g["frame_count"]), 30, 4)


def test_bf16_storage_keeps_embeddings_within_tolerance():
    """predicts on CPU that the GPU path's bf16 activations stay inside cos >= 0.999."""
    mg = _golden_clip()
    clip = mg.golden_clip()
    fn = H.oracle_facenet()
    xs = []
    for i in (0, 4, 8):
        f = clip.frame(i)
        x1, y1, x2, y2 = [int(v) for v in clip.faces(i)[0].box()]
        xs.append(torch.from_numpy(cv2.resize(f[y1:y2, x1:x2], (80, 80))).permute(2, 0, 1).float().div(255))
    x = torch.stack(xs)
    with torch.no_grad():
        e32 = fn(x).numpy()
        with H.bf16_storage_sim():
            e16 = fn(x).numpy()
    for a, b in zip(e32, e16):
        assert H.cosine(a, b) > 0.9995


def test_bundled_clip_golden_pins_the_oracle():
    """BASELINE.json configs[0]: the reference's bundled clip (fixture copy, tests/golden/make_bundled_golden.py).  The
    oracle, re-run here on the first 20 processed frames of the same OpenCV decode, reproduces the committed trace."""
    import cv2
    from oracle.reference_run import reference_run
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(gold_dir, "bundled_clip_run.npz"))
    assert int(g["frame_count"]) == 960 and len(g["frame_index"]) == 240 and int(g["stride"]) == 4
    assert (int(g["width"]), int(g["height"]), int(g["fps"])) == (640, 360, 30)
    tr = reference_run(os.path.join(gold_dir, "bundled_veo3_360p.mp4"), None, H.oracle_mtcnn(), H.oracle_facenet(), max_frames=80)
    assert len(tr.frames) == 20
    for k, f in enumerate(tr.frames):
        assert f.frame_index == int(g["frame_index"][k]) and f.n_faces == int(g["n_faces"][k])
        assert f.embedded == bool(g["embedded"][k])
        if f.embedded:
            assert np.array_equal(f.box, g["box"][k])
            assert np.allclose(f.emb, g["emb"][k], atol=2e-5)
            if f.sim is not None:
                assert abs(f.sim - float(g["sim"][k])) < 1e-5


def test_inter_area_restatement_is_bit_exact_with_opencv():
    """oracle.mode_b.resize_area_u8 (what crop_area_kernel mirrors) against the installed cv2.resize(INTER_AREA): the
    integer-ratio fast path (incl. the 2x2 rounding special case), the general float area path, and the bilinear path with
    area coefficients for enlarged / mixed crops."""
    from oracle.mode_b import resize_area_u8
    rng = np.random.default_rng(0)
    cases = [(320, 320), (480, 320), (160, 160), (161, 200), (200, 260), (233, 177), (640, 640), (80, 95), (100, 100),
             (159, 161), (161, 159), (400, 123), (50, 300), (800, 480), (163, 163), (319, 321), (21, 23)]
    cases += [(int(rng.integers(20, 500)), int(rng.integers(20, 500))) for _ in range(12)]
    for h, w in cases:
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for S in (160, 80):
            assert np.array_equal(resize_area_u8(src, S, S), cv2.resize(src, (S, S), interpolation=cv2.INTER_AREA)), (h, w, S)


def test_mode_b_extract_face_semantics():
    """upstream extract_face: margin-adjusted truncated box, INTER_AREA resize, standardisation (x - 127.5) / 128."""
    from oracle.mode_b import extract_box, extract_face, fixed_image_standardization
    assert extract_box([10.7, 20.2, 110.9, 150.5], 160, 0, 640, 360) == [10, 20, 110, 150]
    assert extract_box([-5.0, -3.0, 700.0, 400.0], 160, 0, 640, 360) == [0, 0, 640, 360]
    b = extract_box([100.0, 100.0, 200.0, 260.0], 160, 32, 640, 360)          # margin 32: 25 px wider, 40 px taller
    assert b == [87, 80, 212, 280]
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
    face, bb = extract_face(img, [100.3, 50.2, 300.9, 310.1], 160, 0)
    assert face.shape == (160, 160, 3) and bb == [100, 50, 300, 310]
    t = fixed_image_standardization(face)
    assert t.shape == (3, 160, 160) and float(t.min()) >= -127.5 / 128 and float(t.max()) <= 127.5 / 128
    assert extract_face(img, [700.0, 10.0, 800.0, 90.0], 160, 0)[0] is None       # box outside the frame: empty crop

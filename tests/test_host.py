"""CPU: host logic of the drop-in and the C-ABI surface (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import truely_b200  # noqa: F401
from truely_b200 import _lib as L
from truely_b200 import model as M
from truely_b200 import weights as W
from oracle import reference_run as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "truely_b200.h")).read()
    declared = set(re.findall(r"\b(trl_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))


def test_weight_blob_sizes_agree_with_the_library():
    lib = L.load()
    assert lib.trl_facenet_blob_len() == W.facenet_blob_size()
    state, src = W.load_mtcnn_state()
    assert W.pack_mtcnn(state, "pnet").size == 6632
    assert W.pack_mtcnn(state, "rnet").size == 100178
    assert W.pack_mtcnn(state, "onet").size == 389040


def test_facenet_pack_matches_blob_len_and_folding():
    sd, _ = W.load_facenet_state()
    blob = W.pack_facenet(sd)
    assert blob.size == W.facenet_blob_size() and blob.dtype == np.float32
    convs, resids, head = W.fold_facenet(sd)
    assert len(convs) == 111 and len(resids) == 21          # 111 BasicConv2d + 21 up-projections = 132 Conv2d layers
    # folded conv == conv followed by eval-mode BN, on random data
    name, w, b = convs[1]
    x = torch.randn(2, 32, 9, 9)
    y = torch.nn.functional.conv2d(x, torch.from_numpy(w).permute(0, 3, 1, 2), torch.from_numpy(b))
    bn = torch.nn.BatchNorm2d(32, eps=1e-3).eval()
    bn.weight.data = torch.from_numpy(sd[f"{name}.bn.weight"]); bn.bias.data = torch.from_numpy(sd[f"{name}.bn.bias"])
    bn.running_mean = torch.from_numpy(sd[f"{name}.bn.running_mean"]); bn.running_var = torch.from_numpy(sd[f"{name}.bn.running_var"])
    ref = bn(torch.nn.functional.conv2d(x, torch.from_numpy(sd[f"{name}.conv.weight"])))
    assert torch.allclose(y, ref, atol=1e-4, rtol=1e-4)


def test_synthetic_weights_are_machine_independent():
    a = W.synth_facenet_state()
    b = W.synth_facenet_state()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    # numpy PCG64 stream: first draw of the first conv is fixed for all time
    first = np.random.default_rng(20180402).standard_normal((32, 3, 3, 3), dtype=np.float32)[0, 0, 0, 0]
    assert np.isclose(a["conv2d_1a.conv.weight"][0, 0, 0, 0], first * np.sqrt(2.0 / 27), rtol=1e-6)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback_create_fails_loudly():
    lib = L.load()
    w = L.Weights()
    ctx = C.c_void_p()
    rc = lib.trl_create(0, C.byref(w), None, C.byref(ctx))
    assert rc == L.TRL_E_CUDA and not ctx.value
    assert b"no CPU fallback" in lib.trl_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.Analyzer(device=0)


def test_run_guards_return_zero_like_the_reference(tmp_path, capsys):
    # server/model.py:20-22 -- missing / empty input -> print + 0 (no GPU needed: guards come first)
    assert M.run(str(tmp_path / "missing.mp4"), str(tmp_path / "o.mp4")) == 0
    assert "doesn't exist or is empty" in capsys.readouterr().out
    empty = tmp_path / "empty.mp4"
    empty.write_bytes(b"")
    assert M.run(str(empty), str(tmp_path / "o.mp4")) == 0
    junk = tmp_path / "junk.mp4"
    junk.write_bytes(b"not a video" * 100)
    assert M.run(str(junk), str(tmp_path / "o.mp4")) == 0
    out = capsys.readouterr().out
    assert "couldn't open video file" in out or "Invalid video properties" in out


def test_run_length_and_score_match_the_reference_logic():
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        fps = int(rng.choice([5, 24, 25, 30, 60]))
        stride = M.frame_stride(fps)
        assert stride == R.frame_stride(fps)
        below = rng.random(n) < rng.uniform(0.2, 0.98)
        rl = M.RunLength()
        run = flagged = 0
        for b in below:                                  # server/model.py:62-70 restated inline
            run = run + 1 if b else 0
            if run > 15:
                flagged += 1
            assert rl.step(bool(b)) == (run > 15)
        frame_count = n * stride - int(rng.integers(0, stride))
        assert rl.deepfake_count == run and rl.deep_fake_frame_count == flagged
        assert M.final_score(flagged, run, frame_count, fps, stride) == R.final_score(flagged, run, frame_count, fps, stride)


def test_chunk_schedule_covers_every_frame_once():
    from truely_b200.model import chunk_schedule
    for n in (1, 7, 89, 90, 91, 250, 450, 1000):
        for chunk in (1, 8, 90):
            for ramp in (False, True):
                r = chunk_schedule(n, chunk, ramp=ramp)
                assert r[0][0] == 0 and r[-1][1] == n
                assert all(a < b and b - a <= chunk for a, b in r)
                assert all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
    r = chunk_schedule(450, 90, ramp=True)
    sz = [b - a for a, b in r]
    assert sz[0] == 22 and 90 // 10 <= sz[-1] < 45                      # short first copy, last cascade on a small chunk
    assert all(sz[i] >= sz[i + 1] for i in range(sz.index(90), len(sz) - 1))      # the tail only shrinks
    k = sz.index(max(sz))
    assert all(sz[i] >= sz[i + 1] for i in range(k, len(sz) - 1))       # after the full-size chunks the tail only shrinks


def test_synthetic_weights_are_opt_in(monkeypatch, tmp_path):
    """run() must never score a video with stand-in networks silently.  Without upstream files and
    without TRUELY_ALLOW_SYNTHETIC=1 the loaders raise; an installed facenet_pytorch wheel's data/ dir is searched."""
    monkeypatch.setenv("TRUELY_WEIGHTS_DIR", str(tmp_path))          # empty
    monkeypatch.setenv("TORCH_HOME", str(tmp_path / "torch_home"))   # empty
    monkeypatch.delenv("TRUELY_ALLOW_SYNTHETIC", raising=False)
    with pytest.raises(W.MissingWeightsError):
        W.load_mtcnn_state()
    with pytest.raises(W.MissingWeightsError):
        W.load_facenet_state()
    monkeypatch.setenv("TRUELY_ALLOW_SYNTHETIC", "1")
    assert W.load_mtcnn_state()[1] == "synthetic" and W.load_facenet_state()[1] == "synthetic"
    # upstream-format files are picked up: two of three MTCNN nets present is still "missing" (no half-upstream cascade)
    state, _ = W.load_mtcnn_state()
    for net in ("pnet", "rnet"):
        torch.save({k[len(net) + 1:]: torch.from_numpy(v) for k, v in state.items() if k.startswith(net + ".")},
                   str(tmp_path / f"{net}.pt"))
    monkeypatch.delenv("TRUELY_ALLOW_SYNTHETIC", raising=False)
    with pytest.raises(W.MissingWeightsError) as e:
        W.load_mtcnn_state()
    assert "onet.pt" in str(e.value)
    torch.save({k[5:]: torch.from_numpy(v) for k, v in state.items() if k.startswith("onet.")}, str(tmp_path / "onet.pt"))
    got, src = W.load_mtcnn_state()
    assert src == "upstream" and all(np.array_equal(got[k], state[k]) for k in state)
    # an installed facenet_pytorch wheel keeps {p,r,o}net.pt in <package>/data: that directory is on the search path
    pkg = tmp_path / "site" / "facenet_pytorch"
    (pkg / "data").mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    monkeypatch.syspath_prepend(str(tmp_path / "site"))
    import importlib
    importlib.invalidate_caches()
    assert str(pkg / "data") in W._weights_dirs()


def test_annotation_pixels_equal_the_reference_drawing():
    """a12 (server/model.py:66-74): model.annotate_frame draws exactly what the reference loop draws -- checked by running
    the oracle loop with stub models that dictate box and similarity, annotate=True, against annotate_frame on copies."""
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, size=(120, 160, 3), dtype=np.uint8) for _ in range(40)]
    boxes = [np.array([20 + k % 7, 30 + k % 5, 90 + k % 11, 100 + k % 3], np.float32) for k in range(40)]

    class StubMtcnn:
        k = 0

        def detect(self, frame):
            b = boxes[self.k][None]
            self.k += 1
            return b, np.array([0.99], np.float32)

    class StubFacenet:
        """alternating embeddings: every comparison is far below 0.99 -> the run passes 15 and frames get flagged"""
        k = 0

        def __call__(self, x):
            e = torch.zeros(1, 512)
            e[0, self.k % 2] = 1.0
            self.k += 1
            return e

    class Sink:
        def __init__(self):
            self.frames = []

        def write(self, f):
            self.frames.append(f.copy())

    sink = Sink()
    ref = R.reference_run_frames(iter([f.copy() for f in frames]), 7, 160, 120, StubMtcnn(), StubFacenet(), writer=sink, annotate=True)
    assert sum(f.flagged for f in ref.frames) > 0 and sum((not f.flagged) and f.sim is not None for f in ref.frames) > 0
    for k, ft in enumerate(ref.frames):
        mine = frames[k].copy()
        if ft.sim is not None:
            M.annotate_frame(mine, ft.box, ft.flagged, ft.frame_index)
        assert np.array_equal(mine, sink.frames[k]), f"frame {k}: annotated pixels differ"


def test_score_clips_equals_per_clip_scores():
    rng = np.random.default_rng(11)
    lens = [40, 1, 75, 0, 33]
    clips = [(m, 8 * m) for m in lens]
    n = sum(lens)
    valid = (rng.random(n) > 0.1).astype(np.uint8)
    below = (rng.random(n) > 0.2).astype(np.uint8)
    has = valid.copy()
    a = 0
    for m in lens:                       # a clip's first face-bearing frame has nothing to compare with
        idx = np.flatnonzero(valid[a:a + m])
        if len(idx):
            has[a + idx[0]] = 0
        a += m
    scores, flagged = M.score_clips(valid, has, below, clips, 60, 8)
    a = 0
    for i, m in enumerate(lens):
        sc, fl, _ = M.score_from_flags(valid[a:a + m], has[a:a + m], below[a:a + m], 8 * m, 60, 8)
        assert scores[i] == sc and flagged[a:a + m] == fl
        a += m
    assert M.clip_start_mask(clips).tolist() == [1 if i in (0, 40, 41, 116) else 0 for i in range(n)]
    with pytest.raises(ValueError):
        M.score_clips(valid[:-1], has[:-1], below[:-1], clips, 60, 8)


def test_shard_record_layout_agrees_with_the_library():
    from truely_b200 import dist as D
    lib = L.load()
    for n_max in (0, 1, 15, 16, 17, 451, 2401):
        assert lib.trl_shard_record_bytes(n_max) == D.record_bytes(n_max)
        assert D.record_bytes(n_max) % 16 == 0


def test_flag_comparison_helper_only_frees_in_band_decisions():
    import helpers as Hh
    rng = np.random.default_rng(0)
    n = 200
    valid = [True] * n
    ref_sim = [None] + [float(rng.choice([0.95, 0.995, 0.9895])) for _ in range(n - 1)]

    def machine(sims):
        rl, fl = M.RunLength(), [False]
        for s in sims[1:]:
            fl.append(rl.step(s < 0.99))
        return fl, M.final_score(rl.deep_fake_frame_count, rl.deepfake_count, n * 4, 30, 4)

    fl, sc = machine(ref_sim)
    assert Hh.assert_flags_match_outside_band(valid, ref_sim, fl, sc, n * 4, 4, 30, valid, ref_sim, fl, sc) > 0
    moved = [None] + [(s + 0.001 if abs(s - 0.99) < 1e-3 else s) for s in ref_sim[1:]]      # in-band frames flip side
    fl2, sc2 = machine(moved)
    assert fl2 != fl
    Hh.assert_flags_match_outside_band(valid, moved, fl2, sc2, n * 4, 4, 30, valid, ref_sim, fl, sc)
    with pytest.raises(AssertionError):                                                      # flags not following the sims
        Hh.assert_flags_match_outside_band(valid, moved, fl, sc2, n * 4, 4, 30, valid, ref_sim, fl, sc)
    far = list(ref_sim)
    far[5] = 0.5 if ref_sim[5] > 0.99 else 0.999                                            # an out-of-band flip
    fl3, sc3 = machine(far)
    if abs(ref_sim[5] - 0.99) >= 1e-3:
        with pytest.raises(AssertionError):
            Hh.assert_flags_match_outside_band(valid, far, fl3, sc3, n * 4, 4, 30, valid, ref_sim, fl, sc)


def test_vectorised_score_equals_the_frame_by_frame_machine():
    """model.score_from_flags (numpy) against RunLength.step fed frame by frame and the oracle's final_score, random flags."""
    rng = np.random.default_rng(21)
    for trial in range(200):
        n = int(rng.integers(0, 400))
        valid = (rng.random(n) > rng.uniform(0, 0.5)).astype(np.uint8)
        has = valid & (rng.random(n) > 0.1).astype(np.uint8)
        below = (rng.random(n) > rng.uniform(0.02, 0.6)).astype(np.uint8)
        stride = int(rng.integers(1, 9))
        frame_count = n * stride - int(rng.integers(0, stride)) if n else 0
        fps = int(rng.choice([7, 24, 30, 60]))
        rl, flagged = M.RunLength(), []
        for v, h, b in zip(valid, has, below):
            flagged.append(rl.step(bool(b)) if (v and h) else False)
        want = R.final_score(rl.deep_fake_frame_count, rl.deepfake_count, max(frame_count, 0), fps, stride)
        score, fl, got = M.score_from_flags(valid, has, below, max(frame_count, 0), fps, stride)
        assert fl == flagged and score == want
        assert (got.deepfake_count, got.deep_fake_frame_count) == (rl.deepfake_count, rl.deep_fake_frame_count)


def test_overlay_stamps_reproduce_opencv():
    """The stamps behind trl_overlay (overlay.py::build_stamps, made with OpenCV's own rasteriser) and the rectangle rule,
    applied in numpy exactly as overlay.cu applies them, against cv2.rectangle + cv2.putText(LINE_AA) of
    server/model.py:67-74 on random frames, boxes and frame numbers (captions that would be clipped are completed on the
    host, as the product does)."""
    import cv2
    from truely_b200 import overlay as O
    st = O.stamps()
    assert len(st.boxes) == 12 and st.digit_advance > 0 and st.idx.max() <= st.lut.shape[0]
    rng = np.random.default_rng(0)
    n_device_text = 0
    for trial in range(240):
        h, w = [(360, 640), (720, 1280), (96, 176), (233, 417)][trial % 4]
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        x1 = int(rng.integers(0, w - 2)); x2 = int(rng.integers(x1 + 1, w + 1))
        y1 = int(rng.integers(0, h - 2)); y2 = int(rng.integers(y1 + 1, h + 1))
        state = int(rng.integers(1, 3))
        fi = int(rng.choice([0, 7, 10, 48, 123, 999, 1234, 56789, 100000]))
        ref = frame.copy()
        if state == O.STATE_AI:
            cv2.rectangle(ref, (x1, y1), (x2, y2), (0, 0, 255), 2)
            cv2.putText(ref, f"AI Detected - Frame {fi}", (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 0, 255), 2, cv2.LINE_AA)
        else:
            cv2.rectangle(ref, (x1, y1), (x2, y2), (0, 255, 0), 2)
            cv2.putText(ref, "Real Frame", (x1, y1 - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 255, 0), 2, cv2.LINE_AA)
        got = frame.copy()
        if O.apply_numpy(got, (x1, y1, x2, y2), state, fi, st):
            O.draw_text_host(got, (x1, y1, x2, y2), state, fi)
        else:
            n_device_text += 1
        assert np.array_equal(got, ref), f"trial {trial}: {h}x{w} state {state} box {(x1, y1, x2, y2)} frame {fi}"
    assert n_device_text > 60

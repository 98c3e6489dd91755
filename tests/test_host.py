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
    assert sz[0] == 22 and 90 // 4 <= sz[-1] < 45                       # short first copy, last cascade on a quarter of a chunk
    k = sz.index(max(sz))
    assert all(sz[i] >= sz[i + 1] for i in range(k, len(sz) - 1))       # after the full-size chunks the tail only shrinks

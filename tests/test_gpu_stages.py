"""GPU: stage-isolated parity of every kernel against the CPU oracle on the same seeded inputs, through the C ABI.

Bars (BASELINE.json north_star): bit-exact for byte / integer / index work (pyramid, crop-resample, crop-align, NMS);
P/R/O-Net outputs within 2e-5 absolute (fp32, different summation order than oneDNN); embedding cosine >= 0.999.
"""
import ctypes as C
import os

import cv2
import numpy as np
import pytest
import torch

import helpers as H
from oracle import mtcnn as OM
from oracle.reference_run import to_tensor_u8
from truely_b200 import _lib as L
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ok(an, rc):
    assert rc == 0, an.lib.trl_last_error(an.ctx).decode()


def frames_tensor(frames):
    return torch.from_numpy(np.ascontiguousarray(np.stack(frames))).cuda()


def geometry(an, H_, W_):
    sc = (C.c_double * L.MAX_SCALES)()
    arrs = [(C.c_int * L.MAX_SCALES)() for _ in range(4)]
    n = an.lib.trl_pyramid_geometry(an.ctx, H_, W_, sc, *arrs)
    assert n > 0
    return [sc[i] for i in range(n)], [[a[i] for i in range(n)] for a in arrs]


@pytest.fixture(scope="module")
def clip_frames():
    clip = SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=0.5, seed=3)
    multi = SyntheticClip(540, 960, 60, 64, n_faces=(3, 3), face_h=(60.0, 160.0), seed=5)
    return [clip.frame(i) for i in (0, 8, 20)], [multi.frame(i) for i in (3, 11)]


# ----------------------------------------------------------------------------- K1
@pytest.mark.parametrize("shape", [(360, 640), (233, 417), (720, 1280)])
def test_pyramid_bit_exact(analyzer, shape):
    an = analyzer
    rng = np.random.default_rng(7)
    h, w = shape
    B = 2
    frames = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    scales, (hs, ws, oh, ow) = geometry(an, h, w)
    assert scales == OM.pyramid_scales(h, w)                     # doubles, exactly
    total = sum(3 * a * b for a, b in zip(hs, ws))
    d_f = torch.from_numpy(frames).cuda()
    d_out = torch.empty(B * total, dtype=torch.float32, device="cuda")
    ok(an, an.lib.trl_pyramid(an.ctx, vp(d_f), B, h, w, vp(d_out), None))
    torch.cuda.synchronize()
    out = d_out.cpu()
    t = torch.from_numpy(frames).permute(0, 3, 1, 2).type(torch.float32)
    off = 0
    for k, s in enumerate(scales):
        ref = (OM.imresample(t, (int(h * s + 1), int(w * s + 1))) - 127.5) * 0.0078125
        assert ref.shape[2:] == (hs[k], ws[k])
        got = out[off:off + B * 3 * hs[k] * ws[k]].view(B, 3, hs[k], ws[k])
        assert torch.equal(got, ref.contiguous()), f"level {k}"
        off += B * 3 * hs[k] * ws[k]


# ----------------------------------------------------------------------------- K2
def test_pnet_maps_match_oracle(analyzer, clip_frames):
    an = analyzer
    pnet = H.oracle_mtcnn().pnet
    frames = clip_frames[0]
    t = torch.from_numpy(np.stack(frames)).permute(0, 3, 1, 2).type(torch.float32)
    for s in OM.pyramid_scales(360, 640)[::2]:
        im = ((OM.imresample(t, (int(360 * s + 1), int(640 * s + 1))) - 127.5) * 0.0078125).contiguous()
        with torch.no_grad():
            reg, prob = pnet(im)
        B, _, hs, ws = im.shape
        oh, ow = prob.shape[2:]
        d_in = im.cuda()
        d_prob = torch.empty((B, oh, ow), dtype=torch.float32, device="cuda")
        d_reg = torch.empty((B, 4, oh, ow), dtype=torch.float32, device="cuda")
        ok(an, an.lib.trl_pnet(an.ctx, vp(d_in), B, hs, ws, vp(d_prob), vp(d_reg), None))
        torch.cuda.synchronize()
        assert (d_prob.cpu() - prob[:, 1]).abs().max().item() < 2e-5, f"scale {s}"
        assert (d_reg.cpu() - reg).abs().max().item() < 2e-5, f"scale {s}"


# tile geometry of the persistent P-Net kernel: outputs smaller than one 16x32 tile, exact tile multiples, one cell more
# than a multiple (a second, nearly empty tile row / column), the 12x12 minimum, many frames with few tiles each
@pytest.mark.parametrize("B,hs,ws", [(1, 12, 12), (3, 13, 29), (2, 42, 74), (2, 44, 76), (1, 47, 83), (5, 121, 67), (37, 20, 33)])
def test_pnet_maps_tile_edges(analyzer, B, hs, ws):
    an = analyzer
    pnet = H.oracle_mtcnn().pnet
    g = torch.Generator().manual_seed(1000 * hs + ws)
    im = (torch.rand((B, 3, hs, ws), generator=g) * 2 - 1).contiguous()
    with torch.no_grad():
        reg, prob = pnet(im)
    oh, ow = prob.shape[2:]
    assert (oh, ow) == ((hs - 2 + 1) // 2 - 4, (ws - 2 + 1) // 2 - 4)
    d_in = im.cuda()
    d_prob = torch.full((B, oh, ow), -1.0, dtype=torch.float32, device="cuda")
    d_reg = torch.full((B, 4, oh, ow), -9.0, dtype=torch.float32, device="cuda")
    ok(an, an.lib.trl_pnet(an.ctx, vp(d_in), B, hs, ws, vp(d_prob), vp(d_reg), None))
    torch.cuda.synchronize()
    assert (d_prob.cpu() - prob[:, 1]).abs().max().item() < 2e-5
    assert (d_reg.cpu() - reg).abs().max().item() < 2e-5


def test_pnet_is_deterministic_across_launches(analyzer):
    """the persistent kernel's tile -> CTA assignment must not leak into the maps (no stale shared memory / TMEM)"""
    an = analyzer
    g = torch.Generator().manual_seed(5)
    im = (torch.rand((4, 3, 90, 150), generator=g) * 2 - 1).cuda().contiguous()
    oh, ow = (90 - 2 + 1) // 2 - 4, (150 - 2 + 1) // 2 - 4
    outs = []
    for _ in range(3):
        d_prob = torch.empty((4, oh, ow), dtype=torch.float32, device="cuda")
        d_reg = torch.empty((4, 4, oh, ow), dtype=torch.float32, device="cuda")
        ok(an, an.lib.trl_pnet(an.ctx, vp(im), 4, 90, 150, vp(d_prob), vp(d_reg), None))
        torch.cuda.synchronize()
        outs.append((d_prob.clone(), d_reg.clone()))
    for pr, rg in outs[1:]:
        assert torch.equal(pr, outs[0][0]) and torch.equal(rg, outs[0][1])


# ----------------------------------------------------------------------------- K4 / K5
def _random_boxes(rng, n, spread, ties=False):
    c = rng.uniform(0, spread, (n, 2)).astype(np.float32)
    wh = rng.uniform(8, 80, (n, 2)).astype(np.float32)
    boxes = np.concatenate([c, c + wh], 1).astype(np.float32)
    scores = rng.uniform(0.6, 1.0, n).astype(np.float32)
    if ties:
        scores = np.round(scores * 20).astype(np.float32) / 20
    return boxes, scores


@pytest.mark.parametrize("n,ties", [(0, False), (1, False), (31, False), (33, True), (257, False), (700, True), (2048, False)])
def test_nms_matches_torchvision_and_numpy(analyzer, n, ties):
    from torchvision.ops import nms as tv_nms
    an = analyzer
    rng = np.random.default_rng(100 + n)
    boxes, scores = _random_boxes(rng, n, spread=60.0 * max(1.0, np.sqrt(max(n, 1) / 30.0)), ties=ties)
    d_b, d_s = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    d_keep = torch.zeros(max(n, 1), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    for thr in (0.5, 0.7):
        ok(an, an.lib.trl_nms(an.ctx, vp(d_b), vp(d_s), n, thr, 0, vp(d_keep), vp(d_n), None))
        torch.cuda.synchronize()
        got = d_keep[: int(d_n.item())].cpu().numpy()
        ref = tv_nms(torch.from_numpy(boxes).view(-1, 4), torch.from_numpy(scores), thr).numpy()
        assert np.array_equal(got, ref), f"mode 0 thr {thr}"
    if not ties and n > 0:       # numpy argsort (quicksort) is not stable, so ties are excluded here
        ok(an, an.lib.trl_nms(an.ctx, vp(d_b), vp(d_s), n, 0.7, 1, vp(d_keep), vp(d_n), None))
        torch.cuda.synchronize()
        got = d_keep[: int(d_n.item())].cpu().numpy()
        ref = OM.nms_numpy(boxes, scores, 0.7, "Min")
        assert np.array_equal(got, ref.astype(np.int64))


# ----------------------------------------------------------------------------- K7
def test_crop_resample_bit_exact(analyzer, clip_frames):
    an = analyzer
    frames = clip_frames[1]
    B, h, w = len(frames), 540, 960
    rng = np.random.default_rng(9)
    n = 40
    pad = np.zeros((n, 4), np.int32)
    img = rng.integers(0, B, n).astype(np.int32)
    for k in range(n):
        side = int(rng.integers(3, 400))
        x = int(rng.integers(1, w - 2)); y = int(rng.integers(1, h - 2))
        pad[k] = (y, min(h, y + side), x, min(w, x + int(side * rng.uniform(0.5, 1.5))))
    pad[0] = (1, h, 1, w)            # whole frame
    pad[1] = (5, 5, 7, 7)            # 1x1 crop
    t = torch.from_numpy(np.stack(frames)).permute(0, 3, 1, 2).type(torch.float32)
    d_f = frames_tensor(frames)
    d_pad, d_img = torch.from_numpy(pad).cuda(), torch.from_numpy(img).cuda()
    for size in (24, 48):
        d_out = torch.empty((n, 3, size, size), dtype=torch.float32, device="cuda")
        ok(an, an.lib.trl_crop_resample(an.ctx, vp(d_f), B, h, w, vp(d_pad), vp(d_img), n, size, vp(d_out), None))
        torch.cuda.synchronize()
        ref, valid = OM._crop_resample(t, img, pad[:, 0], pad[:, 1], pad[:, 2], pad[:, 3], size)
        assert all(valid)
        assert torch.equal(d_out.cpu(), ref.contiguous())


# ----------------------------------------------------------------------------- K8 / K9
def test_rnet_onet_match_oracle(analyzer):
    an = analyzer
    mt = H.oracle_mtcnn()
    rng = np.random.default_rng(4)
    for net, size, fn in ((mt.rnet, 24, an.lib.trl_rnet), (mt.onet, 48, an.lib.trl_onet)):
        n = 37
        x = torch.from_numpy(((rng.integers(0, 256, (n, 3, size, size)).astype(np.float32) - 127.5) * 0.0078125).astype(np.float32))
        with torch.no_grad():
            out = net(x)
        reg, prob = out[0], out[-1][:, 1]
        d_x = x.cuda()
        d_prob = torch.empty(n, dtype=torch.float32, device="cuda")
        d_reg = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        ok(an, fn(an.ctx, vp(d_x), n, vp(d_prob), vp(d_reg), None))
        torch.cuda.synchronize()
        assert (d_prob.cpu() - prob).abs().max().item() < 2e-5
        assert (d_reg.cpu() - reg).abs().max().item() < 5e-5


# ----------------------------------------------------------------------------- K10
def test_crop_align_bit_exact_with_cv2(analyzer, clip_frames):
    an = analyzer
    frames = clip_frames[1]
    B0, h, w = len(frames), 540, 960
    rng = np.random.default_rng(12)
    n = 24
    fr = [frames[i % B0] for i in range(n)]
    boxes = np.zeros((n, 5), np.float32)
    nf = np.ones(n, np.int32)
    for k in range(n):
        x1, y1 = rng.uniform(-30, w - 20), rng.uniform(-30, h - 20)
        boxes[k, :4] = (x1, y1, x1 + rng.uniform(2, 420), y1 + rng.uniform(2, 420))
    boxes[0, :4] = (10.7, 20.2, 90.9, 100.1)        # exactly 80x80 after truncation: identity resize
    boxes[1, :4] = (50.5, 60.5, 50.9, 200.0)        # empty after truncation -> invalid
    nf[2] = 0                                       # no face
    d_f = frames_tensor(fr)
    d_boxes, d_nf = torch.from_numpy(boxes).cuda(), torch.from_numpy(nf).cuda()
    d_bi = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_crops = torch.zeros((n, 80, 80, 3), dtype=torch.uint8, device="cuda")
    ok(an, an.lib.trl_crop_align(an.ctx, vp(d_f), n, h, w, vp(d_boxes), 5, vp(d_nf), vp(d_bi), vp(d_valid), vp(d_crops), None))
    torch.cuda.synchronize()
    bi, valid, crops = d_bi.cpu().numpy(), d_valid.cpu().numpy(), d_crops.cpu().numpy()
    from oracle.reference_run import clamp_box
    for k in range(n):
        if nf[k] == 0:
            assert valid[k] == 0
            continue
        b = clamp_box(boxes[k, :4], w, h)
        assert np.array_equal(bi[k], b)
        good = b[2] > b[0] and b[3] > b[1]
        assert bool(valid[k]) == bool(good)
        if good:
            ref = cv2.resize(fr[k][b[1]:b[3], b[0]:b[2]], (80, 80))
            assert np.array_equal(crops[k], ref), f"crop {k} box {b}"


# ----------------------------------------------------------------------------- K11
def _face_crops(n, seed):
    xs = []
    for i in range(n):
        clip = SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(80.0, 200.0), jitter=0.5, seed=seed + i)
        f = clip.frame(i)
        x1, y1, x2, y2 = [int(max(0, v)) for v in clip.faces(i)[0].box()]
        xs.append(cv2.resize(f[y1:y2, x1:x2], (80, 80)))
    return np.stack(xs)


def _run_facenet(an, crops):
    n, S = crops.shape[0], crops.shape[1]
    d_c = torch.from_numpy(crops).cuda()
    d_e = torch.empty((n, 512), dtype=torch.float32, device="cuda")
    ok(an, an.lib.trl_facenet(an.ctx, vp(d_c), n, S, vp(d_e), None))
    torch.cuda.synchronize()
    return d_e.cpu().numpy()


def _oracle_emb(crops):
    fn = H.oracle_facenet()
    x = torch.stack([to_tensor_u8(c) for c in crops])
    with torch.no_grad():
        return fn(x).numpy()


def test_facenet_simt_embedding_cosine(analyzer_simt):
    crops = _face_crops(9, 40)
    got, ref = _run_facenet(analyzer_simt, crops), _oracle_emb(crops)
    cos = [H.cosine(a, b) for a, b in zip(got, ref)]
    assert min(cos) >= 0.999, cos
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


def test_facenet_umma_layers_match_simt(analyzer, analyzer_simt):
    """tcgen05 implicit GEMM vs the direct-conv kernels, layer by layer on identical inputs (both bf16 storage,
    fp32 accumulation): differences are summation-order only."""
    crops = _face_crops(5, 70)
    _run_facenet(analyzer_simt, crops)
    _run_facenet(analyzer, crops)
    n = crops.shape[0]
    lib = analyzer.lib
    nl = lib.trl_debug_facenet_num_layers(analyzer.ctx)
    assert nl == lib.trl_debug_facenet_num_layers(analyzer_simt.ctx) and nl > 100
    worst = []
    for idx in range(nl):
        name = C.create_string_buffer(64)
        dims = (C.c_int * 4)()
        assert lib.trl_debug_facenet_layer(analyzer.ctx, idx, name, 64, dims) == 0
        hh, ww, cc = dims[0], dims[1], dims[2]
        a = np.zeros((n, hh, ww, cc), np.uint16)
        b = np.zeros((n, hh, ww, cc), np.uint16)
        assert lib.trl_debug_facenet_output(analyzer.ctx, idx, n, a.ctypes.data_as(C.c_void_p)) == 0
        assert lib.trl_debug_facenet_output(analyzer_simt.ctx, idx, n, b.ctypes.data_as(C.c_void_p)) == 0
        fa = torch.from_numpy(a.view(np.int16)).view(torch.bfloat16).float()
        fb = torch.from_numpy(b.view(np.int16)).view(torch.bfloat16).float()
        # (buffers are re-used by later blocks, so for repeated blocks this compares the last writer's output)
        err = (fa - fb).norm().item()
        scale = fb.norm().item() + 1e-6
        worst.append((err / scale, name.value.decode(), err, scale))
    bad = [w for w in worst if w[0] > 0.02]
    assert not bad, f"layers deviating: {bad[:8]}"


def test_facenet_umma_embedding_cosine(analyzer):
    crops = _face_crops(9, 40)
    got, ref = _run_facenet(analyzer, crops), _oracle_emb(crops)
    cos = [H.cosine(a, b) for a, b in zip(got, ref)]
    assert min(cos) >= 0.999, cos
    # a batch that is not a multiple of any tile size, and batch 1
    for n in (1, 3):
        g = _run_facenet(analyzer, crops[:n])
        assert min(H.cosine(a, b) for a, b in zip(g, ref[:n])) >= 0.999


def test_facenet_160_mode_b(analyzer):
    """North-star variant: 160x160 crops.  Same weights, oracle at 160."""
    an = analyzer
    crops80 = _face_crops(4, 90)
    crops = np.stack([cv2.resize(c, (160, 160)) for c in crops80])
    got, ref = _run_facenet(an, crops), _oracle_emb(crops)
    assert min(H.cosine(a, b) for a, b in zip(got, ref)) >= 0.999
    _run_facenet(an, crops80)        # back to the reference size (re-plans)


# ----------------------------------------------------------------------------- K12
def test_consistency_matches_numpy(analyzer):
    an = analyzer
    rng = np.random.default_rng(2)
    B = 50
    emb = rng.standard_normal((B, 512)).astype(np.float32)
    emb[1:] = 0.97 * emb[:-1] + 0.03 * emb[1:] * rng.uniform(0.2, 6, (B - 1, 1)).astype(np.float32)
    valid = (rng.random(B) > 0.25).astype(np.uint8)
    valid[:3] = (0, 0, 1)
    for halo in (None, rng.standard_normal(512).astype(np.float32)):
        d_e, d_v = torch.from_numpy(emb).cuda(), torch.from_numpy(valid).cuda()
        d_h = torch.from_numpy(halo).cuda() if halo is not None else None
        d_sim = torch.zeros(B, device="cuda"); d_b = torch.zeros(B, dtype=torch.uint8, device="cuda")
        d_hs = torch.zeros(B, dtype=torch.uint8, device="cuda")
        d_le = torch.zeros(512, device="cuda"); d_lv = torch.zeros(1, dtype=torch.uint8, device="cuda")
        ok(an, an.lib.trl_consistency(an.ctx, vp(d_e), vp(d_v), B, vp(d_h), None, 0.99, vp(d_sim), vp(d_b), vp(d_hs), vp(d_le),
                                      vp(d_lv), None))
        torch.cuda.synchronize()
        sim, below, has = d_sim.cpu().numpy(), d_b.cpu().numpy(), d_hs.cpu().numpy()
        prev = halo
        for i in range(B):
            if not valid[i]:
                assert has[i] == 0
                continue
            if prev is None:
                assert has[i] == 0
            else:
                ref = np.dot(emb[i], prev) / (np.linalg.norm(emb[i]) * np.linalg.norm(prev))   # server/model.py:61
                assert has[i] == 1 and abs(sim[i] - ref) < 2e-6
                if abs(ref - 0.99) > 1e-5:
                    assert below[i] == (ref < 0.99)
            prev = emb[i]
        assert d_lv.item() == 1 and np.array_equal(d_le.cpu().numpy(), prev)


def test_facenet_valid_only_equals_full_batch_on_the_valid_rows(analyzer):
    """trl_facenet_valid packs the face-bearing crops on the device (batch size read from device memory by every kernel of
    the pass) -- same bits as embedding everything, zeros for the frames without a face, any pattern of holes."""
    import ctypes as C
    from truely_b200.model import _vp
    an = analyzer
    rng = np.random.default_rng(3)
    for n, S, pattern in ((37, 80, "random"), (64, 80, "none"), (9, 160, "all"), (130, 80, "first_last"), (5, 80, "empty")):
        crops = torch.from_numpy(rng.integers(0, 256, (n, S, S, 3), dtype=np.uint8)).cuda()
        if pattern == "random":
            v = (rng.random(n) > 0.4).astype(np.uint8)
        elif pattern == "none":
            v = np.ones(n, np.uint8)
        elif pattern == "all":
            v = np.ones(n, np.uint8); v[3] = 0
        elif pattern == "first_last":
            v = np.zeros(n, np.uint8); v[0] = v[-1] = 1
        else:
            v = np.zeros(n, np.uint8)
        valid = torch.from_numpy(v).cuda()
        full = torch.empty((n, 512), dtype=torch.float32, device="cuda")
        part = torch.full((n, 512), 7.0, dtype=torch.float32, device="cuda")
        with torch.cuda.stream(an.stream):
            an._check(an.lib.trl_facenet(an.ctx, _vp(crops), n, S, _vp(full), an._sptr()))
            an._check(an.lib.trl_facenet_valid(an.ctx, _vp(crops), _vp(valid), n, S, _vp(part), an._sptr()))
        an.stream.synchronize()
        full, part = full.cpu().numpy(), part.cpu().numpy()
        for i in range(n):
            if v[i]:
                assert np.array_equal(part[i], full[i]), f"{pattern}: row {i}"
            else:
                assert not part[i].any(), f"{pattern}: row {i} must be zero"


def test_overlay_kernel_equals_opencv(analyzer):
    """trl_overlay (server/model.py:67-74): rectangle + anti-aliased caption, bit exact with cv2.rectangle /
    cv2.putText on random frames, boxes (clamped ones and boxes at the frame edge included) and frame numbers; captions
    that would leave the frame come back as pending and are completed with cv2.putText on the host."""
    from truely_b200 import model as M
    from truely_b200 import overlay as O
    an = analyzer
    rng = np.random.default_rng(5)
    for (h, w) in ((360, 640), (720, 1280), (96, 176), (233, 417)):
        B = 24
        frames = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
        box = np.zeros((B, 4), np.int32)
        state = rng.integers(0, 3, B).astype(np.uint8)
        fidx = rng.choice([0, 7, 10, 48, 123, 999, 1234, 56789, 100000], B).astype(np.int32)
        for b in range(B):
            x1 = int(rng.integers(0, w - 2)); x2 = int(rng.integers(x1 + 1, w + 1))
            y1 = int(rng.integers(0, h - 2)); y2 = int(rng.integers(y1 + 1, h + 1))
            box[b] = (x1, y1, x2, y2)
        box[0] = (0, 0, w, h); box[1] = (w - 1, h - 1, w, h); box[2] = (0, 40, 50, 90)
        d_frames = torch.from_numpy(frames).cuda()
        d_box = torch.from_numpy(box).cuda()
        pend = an.overlay_device(d_frames, d_box, state, fidx)
        an.stream.synchronize()
        got = d_frames.cpu().numpy()
        pend = pend.cpu().numpy()
        n_dev = 0
        for b in range(B):
            ref = frames[b].copy()
            if state[b] != 0:
                M.annotate_frame(ref, box[b], state[b] == 2, int(fidx[b]))
            out = got[b].copy()
            if pend[b]:
                O.draw_text_host(out, box[b], int(state[b]), int(fidx[b]))
            elif state[b] != 0:
                n_dev += 1
            assert np.array_equal(out, ref), f"{h}x{w} frame {b} state {state[b]} box {box[b]}"
        if w >= 640:
            assert n_dev > 0


def test_facenet_upstream_weights_parity_when_present(analyzer):
    """ADVICE (round 1): the bf16 trunk is demonstrated on the seeded stand-in network only.  When the upstream vggface2
    checkpoint is installed (TRUELY_WEIGHTS_DIR / TORCH_HOME / the facenet_pytorch package data) this holds the tcgen05
    path to what a 1e-3 band around the 0.99 similarity threshold needs -- 1 - cos <= 1e-4 against the fp32 oracle and
    consecutive-frame similarities within 1e-3 -- and otherwise reports that bar as unverified (skip)."""
    if analyzer.facenet_source == "synthetic":
        pytest.skip("upstream vggface2 weights absent: bf16-trunk parity with real weights unverified (stand-in network only)")
    crops = _face_crops(12, 40)
    got, ref = _run_facenet(analyzer, crops), _oracle_emb(crops)
    cos = [H.cosine(a, b) for a, b in zip(got, ref)]
    assert 1.0 - min(cos) <= 1e-4, cos
    sim_g = [float(np.dot(got[i], got[i + 1])) for i in range(len(got) - 1)]
    sim_r = [float(np.dot(ref[i], ref[i + 1]) / (np.linalg.norm(ref[i]) * np.linalg.norm(ref[i + 1]))) for i in range(len(ref) - 1)]
    assert max(abs(a - b) for a, b in zip(sim_g, sim_r)) < 1e-3


def test_overlay_empty_batch_and_untouched_frames(analyzer):
    """trl_overlay: B = 0 is a no-op, frames with state 0 come back bit for bit, a missing stamp table is an error."""
    an = analyzer
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (3, 120, 200, 3), dtype=np.uint8)
    d_frames = torch.from_numpy(frames).cuda()
    d_box = torch.tensor([[10, 10, 60, 70]] * 3, dtype=torch.int32, device="cuda")
    pend = an.overlay_device(d_frames, d_box, [0, 0, 0], [1, 2, 3])
    an.stream.synchronize()
    assert np.array_equal(d_frames.cpu().numpy(), frames) and not pend.cpu().numpy().any()
    st = torch.zeros(1, dtype=torch.uint8, device="cuda")
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert an.lib.trl_overlay(an.ctx, vp(d_frames), 0, 120, 200, vp(d_box), vp(st), vp(idx), None, None) == 0
    assert an.lib.trl_overlay(an.ctx, None, 1, 120, 200, vp(d_box), vp(st), vp(idx), None, None) == L.TRL_E_INVALID

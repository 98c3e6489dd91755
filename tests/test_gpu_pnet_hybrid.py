"""GPU: the hybrid P-Net (single-pass tensor-core screen + exact fp32 re-evaluation of the screened cells), through the C ABI.

* the fp16 hi / lo pair pyramid restores the fp32 pyramid (bit-exact test: test_gpu_stages.py) to 2^-22 relative;
* the screen's logit map is close enough to the fp32 oracle that the screening margin (0.05 in probability) is never
  touched: |sigmoid(d) - prob| stays below a quarter of it;
* the cascade with the hybrid P-Net (pnet_precision 2 and 3) produces the candidates of the 3-term kernel
  (pnet_precision 0): same counts at every stage, same boxes.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H
from oracle import mtcnn as OM
from truely_b200 import _lib as L
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu


def vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ok(an, rc):
    assert rc == 0, an.lib.trl_last_error(an.ctx).decode()


@pytest.fixture(scope="module")
def analyzers():
    from truely_b200.model import Analyzer
    return {p: Analyzer(device=0, pnet_precision=p) for p in (0, 2, 3)}


def pair_layout(an, h, w):
    per = C.c_longlong()
    off = (C.c_longlong * L.MAX_SCALES)()
    pitch = (C.c_int * L.MAX_SCALES)()
    n = an.lib.trl_pyramid_pairs_size(an.ctx, h, w, C.byref(per), off, pitch)
    assert n > 0
    return per.value, [off[i] for i in range(n)], [pitch[i] for i in range(n)]


def run_pairs(an, frames):
    B, h, w, _ = frames.shape
    per, off, pitch = pair_layout(an, h, w)
    d_f = torch.from_numpy(frames).cuda()
    d_hi = torch.zeros((B * per, 8), dtype=torch.float16, device="cuda")
    d_lo = torch.zeros((B * per, 8), dtype=torch.float16, device="cuda")
    ok(an, an.lib.trl_pyramid_pairs(an.ctx, vp(d_f), B, h, w, vp(d_hi), vp(d_lo), None))
    torch.cuda.synchronize()
    return d_hi, d_lo, off, pitch


@pytest.mark.parametrize("shape", [(360, 640), (233, 417), (720, 1280), (1080, 1920), (96, 176), (2160, 3840), (1081, 1923)])
def test_pair_pyramid_restores_the_fp32_pyramid(analyzer, shape):
    an = analyzer
    h, w = shape
    B = 2 if h * w <= 1920 * 1080 else 1
    frames = np.random.default_rng(11).integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    d_hi, d_lo, off, pitch = run_pairs(an, frames)
    hi, lo = d_hi.cpu().float(), d_lo.cpu().float()
    t = torch.from_numpy(frames).permute(0, 3, 1, 2).type(torch.float32)
    for k, s in enumerate(OM.pyramid_scales(h, w)):
        hs, ws = int(h * s + 1), int(w * s + 1)
        ref = (OM.imresample(t, (hs, ws)) - 127.5) * 0.0078125                     # [B,3,hs,ws]
        n = B * hs * pitch[k]
        lvl_hi = hi[off[k] * B: off[k] * B + n].view(B, hs, 2 * pitch[k], 4)       # pixel-major: 4 halves per pixel
        lvl_lo = lo[off[k] * B: off[k] * B + n].view(B, hs, 2 * pitch[k], 4)
        got = (lvl_hi + lvl_lo)[:, :, :ws, :3].permute(0, 3, 1, 2)
        err = (got - ref).abs()
        assert (err <= ref.abs() * 2.0 ** -22 + 1e-9).all(), f"level {k}: max err {err.max().item()}"
        assert (lvl_hi[:, :, :ws, 3] == 0).all() and (lvl_lo[:, :, :ws, 3] == 0).all()
        # hi is the round-to-nearest fp16 of the value (the screen's operand)
        assert torch.equal(lvl_hi[:, :, :ws, :3].permute(0, 3, 1, 2), ref.half().float()), f"level {k}"


@pytest.mark.parametrize("shape,faces", [((360, 640), 1), ((233, 417), 2), ((540, 960), 3)])
def test_screen_logits_track_the_fp32_maps(analyzer, shape, faces):
    an = analyzer
    h, w = shape
    clip = SyntheticClip(h, w, 30, 64, n_faces=(faces, faces), face_h=(0.15 * h, 0.4 * h), jitter=0.5, seed=21)
    frames = np.stack([clip.frame(i) for i in (0, 12)])
    B = frames.shape[0]
    d_hi, _, off, pitch = run_pairs(an, frames)
    scales = OM.pyramid_scales(h, w)
    pnet = H.oracle_mtcnn().pnet
    t = torch.from_numpy(frames).permute(0, 3, 1, 2).type(torch.float32)
    refs, total = [], 0
    for s in scales:
        im = ((OM.imresample(t, (int(h * s + 1), int(w * s + 1))) - 127.5) * 0.0078125).contiguous()
        with torch.no_grad():
            _, prob = pnet(im)
        refs.append(prob[:, 1])
        total += prob[:, 1].numel()
    d_logit = torch.full((total,), float("nan"), dtype=torch.float32, device="cuda")
    ok(an, an.lib.trl_pnet_screen_maps(an.ctx, vp(d_hi), B, h, w, vp(d_logit), None))
    torch.cuda.synchronize()
    logit = d_logit.cpu()
    o, worst = 0, 0.0
    for k, ref in enumerate(refs):
        got = torch.sigmoid(logit[o:o + ref.numel()].view_as(ref))
        assert torch.isfinite(got).all(), f"level {k}: unwritten cells"
        worst = max(worst, (got - ref).abs().max().item())
        o += ref.numel()
    assert worst < 0.0125, f"screen probability error {worst} is not small against the 0.05 margin"


def detect(an, frames):
    B, h, w, _ = frames.shape
    d_f = torch.from_numpy(frames).cuda()
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_boxes = torch.zeros((B, an.box_cap, 5), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    ok(an, an.lib.trl_detect(an.ctx, vp(d_f), B, h, w, vp(d_n), vp(d_boxes), vp(d_counts), None))
    torch.cuda.synchronize()
    an.check_capacity()
    return d_n.cpu().numpy(), d_boxes.cpu().numpy(), d_counts.cpu().numpy()


@pytest.mark.parametrize("shape,faces,nf", [((360, 640), 1, 6), ((233, 417), 2, 3), ((720, 1280), 1, 4), ((1080, 1920), 6, 2)])
def test_hybrid_cascade_equals_the_three_term_cascade(analyzers, shape, faces, nf):
    h, w = shape
    clip = SyntheticClip(h, w, 30, 200, n_faces=(faces, faces), face_h=(0.12 * h, 0.35 * h), jitter=0.5, seed=31)
    frames = np.stack([clip.frame(7 * i) for i in range(nf)])
    n0, b0, c0 = detect(analyzers[0], frames)
    assert n0.sum() > 0
    for prec in (2, 3):
        n1, b1, c1 = detect(analyzers[prec], frames)
        assert np.array_equal(c0, c1), f"precision {prec}: stage counts differ\n{c0}\n{c1}"
        assert np.array_equal(n0, n1)
        for i in range(nf):
            assert np.allclose(b0[i, :n0[i]], b1[i, :n1[i]], rtol=0, atol=2e-3), f"precision {prec} frame {i}"


def test_hybrid_cascade_on_noise_and_low_thresholds(analyzers):
    """textured input (many near-threshold cells) exercises the screen list and the refine kernel much harder"""
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (2, 45, 80, 3), dtype=np.uint8)
    frames = np.ascontiguousarray(np.repeat(np.repeat(base, 8, axis=1), 8, axis=2))      # 360 x 640 blocks
    frames = np.clip(frames.astype(np.int16) + rng.integers(-20, 21, frames.shape), 0, 255).astype(np.uint8)
    n0, b0, c0 = detect(analyzers[0], frames)
    for prec in (2, 3):
        n1, b1, c1 = detect(analyzers[prec], frames)
        assert np.array_equal(c0, c1), f"precision {prec}: stage counts differ\n{c0}\n{c1}"
        assert np.array_equal(n0, n1)

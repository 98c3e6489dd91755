"""Host side of the device overlay (csrc/overlay.cu, ``trl_overlay``): the stamps of the reference's two captions.

The reference annotates a compared frame with ``cv2.rectangle(..., 2)`` and an anti-aliased ``cv2.putText``
(server/model.py:67-74).  Anti-aliased strokes blend into the frame, so what a caption does to a pixel is a function of
the pixel's previous value alone; ``build_stamps`` asks OpenCV's own rasteriser for these functions (the caption on the
256 constant backgrounds) and stores them as de-duplicated 256-entry tables plus one index map per caption -- the kernel
then reproduces ``cv2.putText`` bit for bit wherever the caption lies inside the frame.  ``apply_numpy`` is the same
procedure in numpy (CPU tests pin the tables against cv2 on random frames without a GPU).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import cv2
import numpy as np

from . import _lib as L

FONT = cv2.FONT_HERSHEY_SIMPLEX
REAL_TEXT, REAL_SCALE = "Real Frame", 0.5                    # server/model.py:73
AI_PREFIX, AI_SCALE, AI_ORG = "AI Detected - Frame ", 1.0, (10, 30)      # server/model.py:68
THICKNESS = 2
STATE_NONE, STATE_REAL, STATE_AI = 0, 1, 2
_PAD = 8


@dataclass
class Stamps:
    lut: np.ndarray            # uint8 [n_lut, 2, 256]: towards 0 / towards 255
    boxes: list                # 12 x (ox, oy, w, h, idx_off): "Real Frame", the prefix, the digits 0-9 in the first slot
    idx: np.ndarray            # uint16, concatenated index maps (0 = untouched, k = table k - 1)
    digit_advance: int


def _tables(text: str, scale: float):
    """(lut_pix uint8 [h, w, 2, 256], (org_x, org_y)) of ``text`` on a canvas that holds it with a margin."""
    (tw, th), base = cv2.getTextSize(text, FONT, scale, THICKNESS)
    w, h = tw + 2 * _PAD + 4, th + base + 2 * _PAD + 4
    org = (_PAD, _PAD + th + 2)
    out = np.empty((h, w, 2, 256), np.uint8)
    for b in range(256):
        img = np.full((h, w, 3), b, np.uint8)
        cv2.putText(img, text, org, FONT, scale, (0, 255, 0), THICKNESS, cv2.LINE_AA)
        out[:, :, 0, b] = img[:, :, 0]          # channel blended towards 0
        out[:, :, 1, b] = img[:, :, 1]          # channel blended towards 255
    return out, org


def build_stamps() -> Stamps:
    ident = np.arange(256, dtype=np.uint8)
    rows, maps = [], []          # per stamp: (tables of its touched pixels [n, 512]), (ox, oy, w, h, mask)

    def add(lut_pix, org, base=None):
        touched = (lut_pix != ident).any(axis=(2, 3))
        if base is not None:                                   # only what the extra glyph adds to the prefix
            touched &= (lut_pix != base).any(axis=(2, 3))
        ys, xs = np.nonzero(touched)
        y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
        sub = touched[y0:y1, x0:x1]
        rows.append(lut_pix[y0:y1, x0:x1][sub].reshape(-1, 512))
        maps.append((int(x0 - org[0]), int(y0 - org[1]), int(x1 - x0), int(y1 - y0), sub))

    real, org = _tables(REAL_TEXT, REAL_SCALE)
    add(real, org)
    # all captions of the second kind on one canvas size: prefix + one digit
    prefix_only = None
    for d in [None] + list(range(10)):
        text = AI_PREFIX + ("0" if d is None else str(d))
        pix, org = _tables(text, AI_SCALE)
        if d is None:
            # the prefix alone, on the canvas of prefix + digit: render it there
            (tw, th), base = cv2.getTextSize(text, FONT, AI_SCALE, THICKNESS)
            h, w = pix.shape[:2]
            prefix_only = np.empty_like(pix)
            for b in range(256):
                img = np.full((h, w, 3), b, np.uint8)
                cv2.putText(img, AI_PREFIX, org, FONT, AI_SCALE, (0, 255, 0), THICKNESS, cv2.LINE_AA)
                prefix_only[:, :, 0, b] = img[:, :, 0]
                prefix_only[:, :, 1, b] = img[:, :, 1]
            add(prefix_only, org)
        else:
            add(pix, org, base=prefix_only)
    adv = (cv2.getTextSize(AI_PREFIX + "00", FONT, AI_SCALE, THICKNESS)[0][0]
           - cv2.getTextSize(AI_PREFIX + "0", FONT, AI_SCALE, THICKNESS)[0][0])
    allrows = np.concatenate(rows)
    uniq, inv = np.unique(allrows, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    assert len(uniq) < 65535
    boxes, idx, pos, off = [], [], 0, 0
    for ox, oy, w, h, sub in maps:
        m = np.zeros((h, w), np.uint16)
        n = int(sub.sum())
        m[sub] = inv[pos:pos + n] + 1
        pos += n
        boxes.append((ox, oy, w, h, off))
        idx.append(m.reshape(-1))
        off += w * h
    return Stamps(lut=np.ascontiguousarray(uniq.reshape(-1, 2, 256)), boxes=boxes, idx=np.concatenate(idx), digit_advance=int(adv))


_STAMPS = None


def stamps() -> Stamps:
    global _STAMPS
    if _STAMPS is None:
        _STAMPS = build_stamps()
    return _STAMPS


def register(lib, ctx, st: Stamps | None = None):
    """Upload the stamps into a context (``trl_overlay_set_stamps``)."""
    st = st or stamps()
    arr = (L.Stamp * 12)(*[L.Stamp(*b) for b in st.boxes])
    rc = lib.trl_overlay_set_stamps(ctx, st.lut.ctypes.data_as(C.c_void_p), int(st.lut.shape[0]), arr, 12,
                                    st.idx.ctypes.data_as(C.c_void_p), int(st.idx.size), st.digit_advance)
    if rc != 0:
        raise L.TrlError(rc, lib.trl_last_error(ctx).decode())


# ----------------------------------------------------------------------------- numpy restatement of the kernel (tests)

def rectangle_mask(h: int, w: int, x1: int, y1: int, x2: int, y2: int) -> np.ndarray:
    """Pixels cv2.rectangle(img, (x1, y1), (x2, y2), colour, 2) paints: within one pixel of the outline, minus the four
    outer corner pixels."""
    yy, xx = np.mgrid[0:h, 0:w]
    outer = (xx >= x1 - 1) & (xx <= x2 + 1) & (yy >= y1 - 1) & (yy <= y2 + 1)
    inner = (xx >= x1 + 2) & (xx <= x2 - 2) & (yy >= y1 + 2) & (yy <= y2 - 2)
    corner = ((xx == x1 - 1) | (xx == x2 + 1)) & ((yy == y1 - 1) | (yy == y2 + 1))
    return outer & ~inner & ~corner


def _inside(box, ox, oy, w, h):
    sx, sy, sw, sh, _ = box
    return ox + sx - 2 >= 0 and oy + sy - 2 >= 0 and ox + sx + sw + 2 <= w and oy + sy + sh + 2 <= h


def _apply(frame, st, k, ox, oy, targets):
    sx, sy, sw, sh, off = st.boxes[k]
    m = st.idx[off:off + sw * sh].reshape(sh, sw)
    ys, xs = np.nonzero(m)
    for c in range(3):
        v = frame[oy + sy + ys, ox + sx + xs, c]
        frame[oy + sy + ys, ox + sx + xs, c] = st.lut[m[ys, xs] - 1, targets[c], v]


def apply_numpy(frame: np.ndarray, box, state: int, frame_index: int, st: Stamps | None = None) -> bool:
    """What overlay_kernel does to one frame, in place.  Returns True when the caption was left to the host (clipped)."""
    st = st or stamps()
    if state == STATE_NONE:
        return False
    h, w = frame.shape[:2]
    x1, y1, x2, y2 = (int(v) for v in box)
    colour = (0, 255, 0) if state == STATE_REAL else (0, 0, 255)
    frame[rectangle_mask(h, w, x1, y1, x2, y2)] = colour
    targets = tuple(1 if c else 0 for c in colour)
    if state == STATE_REAL:
        if not _inside(st.boxes[0], x1, y1 - 10, w, h):
            return True
        _apply(frame, st, 0, x1, y1 - 10, targets)
        return False
    digits = [int(ch) for ch in str(max(0, int(frame_index)))]
    ox, oy = AI_ORG
    ok = _inside(st.boxes[1], ox, oy, w, h) and all(
        _inside(st.boxes[2 + d], ox + k * st.digit_advance, oy, w, h) for k, d in enumerate(digits))
    if not ok:
        return True
    _apply(frame, st, 1, ox, oy, targets)
    for k, d in enumerate(digits):
        _apply(frame, st, 2 + d, ox + k * st.digit_advance, oy, targets)
    return False


def draw_text_host(frame, box, state: int, frame_index: int):
    """The caption alone, with OpenCV (frames whose caption the device left pending)."""
    if state == STATE_AI:
        cv2.putText(frame, f"{AI_PREFIX}{frame_index}", AI_ORG, FONT, AI_SCALE, (0, 0, 255), THICKNESS, cv2.LINE_AA)
    elif state == STATE_REAL:
        cv2.putText(frame, REAL_TEXT, (int(box[0]), int(box[1]) - 10), FONT, REAL_SCALE, (0, 255, 0), THICKNESS, cv2.LINE_AA)

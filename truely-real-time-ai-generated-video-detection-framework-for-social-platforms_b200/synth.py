"""Synthetic clips for the BASELINE.json configs.

Host-side, seeded, numpy + OpenCV drawing only.  A clip is a low-frequency noise
background (bilinear-upsampled U[0,255] grid) with one or more elliptical "faces"
(skin-toned ellipse, hair cap, dark eye and mouth blobs) moving on smooth Lissajous
paths, plus per-frame additive N(0,2) sensor noise.  ``jitter > 0`` perturbs the
facial layout independently per frame, which is what drives consecutive-frame
embedding similarity below the reference's 0.99 threshold (server/model.py:16,62)
and exercises the run-length counter; ``jitter = 0`` gives a temporally stable face.

Frames are BGR uint8 [H, W, 3], the layout ``cv2.VideoCapture.read`` hands the
reference loop (server/model.py:43).
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np

__all__ = ["FaceSpec", "render_face", "make_background", "SyntheticClip", "CONFIGS"]


@dataclass
class FaceSpec:
    cx: float
    cy: float
    h: float                     # face height in px; width = aspect * h
    aspect: float = 0.76
    skin: tuple = (120.0, 150.0, 200.0)   # BGR
    eye_dx: float = 0.19         # fractions of width / height
    eye_dy: float = -0.10
    eye_w: float = 0.10
    eye_h: float = 0.045
    mouth_dy: float = 0.26
    mouth_w: float = 0.20
    mouth_h: float = 0.05
    hair: float = 0.30           # fraction of the height covered by the hair cap
    tilt: float = 0.0            # degrees

    @property
    def w(self) -> float:
        return self.aspect * self.h

    def box(self):
        """Ground-truth box (x1, y1, x2, y2) of the ellipse."""
        return (self.cx - self.w / 2, self.cy - self.h / 2, self.cx + self.w / 2, self.cy + self.h / 2)


def _ipt(x, y, shift=4):
    s = 1 << shift
    return int(round(x * s)), int(round(y * s))


def render_face(img: np.ndarray, f: FaceSpec) -> None:
    """Draw one face into ``img`` (BGR uint8, in place) with sub-pixel anti-aliased ellipses."""
    sh = 4
    s = 1 << sh
    w, h = f.w, f.h
    ax = (int(round(w / 2 * s)), int(round(h / 2 * s)))
    c = _ipt(f.cx, f.cy, sh)
    skin = tuple(float(v) for v in f.skin)
    cv2.ellipse(img, c, ax, f.tilt, 0, 360, skin, -1, cv2.LINE_AA, sh)
    # hair cap: upper arc of the same ellipse, dark
    hair_col = (30.0, 35.0, 45.0)
    span = float(np.degrees(np.arccos(max(-1.0, min(1.0, 1.0 - 2.0 * f.hair)))))
    cv2.ellipse(img, c, ax, f.tilt, 270 - span, 270 + span, hair_col, -1, cv2.LINE_AA, sh)
    ca, sa = np.cos(np.radians(f.tilt)), np.sin(np.radians(f.tilt))

    def at(dx, dy):
        return f.cx + dx * ca - dy * sa, f.cy + dx * sa + dy * ca

    eye_ax = (max(1, int(round(f.eye_w * w * s))), max(1, int(round(f.eye_h * h * s))))
    for sgn in (-1.0, 1.0):
        ex, ey = at(sgn * f.eye_dx * w, f.eye_dy * h)
        cv2.ellipse(img, _ipt(ex, ey, sh), eye_ax, f.tilt, 0, 360, (25.0, 25.0, 30.0), -1, cv2.LINE_AA, sh)
    mx, my = at(0.0, f.mouth_dy * h)
    m_ax = (max(1, int(round(f.mouth_w * w * s))), max(1, int(round(f.mouth_h * h * s))))
    cv2.ellipse(img, _ipt(mx, my, sh), m_ax, f.tilt, 0, 360, (50.0, 40.0, 120.0), -1, cv2.LINE_AA, sh)
    nx, ny = at(0.0, 0.07 * h)
    n_ax = (max(1, int(round(0.035 * w * s))), max(1, int(round(0.07 * h * s))))
    cv2.ellipse(img, _ipt(nx, ny, sh), n_ax, f.tilt, 0, 360,
                (skin[0] * 0.8, skin[1] * 0.8, skin[2] * 0.85), -1, cv2.LINE_AA, sh)


def make_background(rng: np.random.Generator, h: int, w: int, cell: int = 8) -> np.ndarray:
    """Low-frequency noise: U[0,255] on a (h/cell, w/cell) grid, bilinear-upsampled."""
    gh, gw = max(2, h // cell + 1), max(2, w // cell + 1)
    grid = rng.integers(0, 256, (gh, gw, 3), dtype=np.uint8)
    return cv2.resize(grid, (w, h), interpolation=cv2.INTER_LINEAR)


class SyntheticClip:
    """Deterministic synthetic clip; ``frame(i)`` renders frame ``i`` on demand."""

    def __init__(self, height: int, width: int, fps: int, n_frames: int, n_faces=(1, 1),
                 face_h=(160.0, 260.0), jitter: float = 0.0, seed: int = 0, noise_sigma: float = 2.0):
        self.height, self.width, self.fps, self.n_frames = height, width, fps, n_frames
        self.jitter = jitter
        self.seed = seed
        rng = np.random.default_rng(seed)
        self.bg = make_background(rng, height, width)
        k = int(rng.integers(n_faces[0], n_faces[1] + 1))
        self.tracks = []
        for j in range(k):
            fh = float(rng.uniform(*face_h))
            # Lissajous path confined so the face stays inside the frame
            ax = max(1.0, (width - 0.8 * fh) / 2 - 8)
            ay = max(1.0, (height - fh) / 2 - 8)
            self.tracks.append(dict(
                h=fh,
                cx0=width / 2.0, cy0=height / 2.0, ax=ax * rng.uniform(0.3, 0.95), ay=ay * rng.uniform(0.3, 0.95),
                fx=rng.uniform(0.05, 0.2), fy=rng.uniform(0.05, 0.2), px=rng.uniform(0, 2 * np.pi),
                py=rng.uniform(0, 2 * np.pi),
                skin=(120 + rng.uniform(-15, 15), 150 + rng.uniform(-15, 15), 200 + rng.uniform(-15, 15)),
                eye_dx=0.19 + rng.uniform(-0.02, 0.02), mouth_dy=0.26 + rng.uniform(-0.02, 0.02),
                aspect=0.76 + rng.uniform(-0.04, 0.04),
            ))
        # a small bank of sensor-noise frames, re-used with per-frame rolls
        self._noise = np.clip(np.rint(rng.normal(0.0, noise_sigma, (4, height, width, 3))), -127, 127).astype(np.int8) \
            if noise_sigma > 0 else None

    def faces(self, i: int):
        t = i / float(self.fps)
        out = []
        jr = np.random.default_rng((self.seed + 1) * 1_000_003 + i)
        for tr in self.tracks:
            f = FaceSpec(
                cx=tr["cx0"] + tr["ax"] * np.sin(2 * np.pi * tr["fx"] * t + tr["px"]),
                cy=tr["cy0"] + tr["ay"] * np.sin(2 * np.pi * tr["fy"] * t + tr["py"]),
                h=tr["h"] * (1.0 + 0.05 * np.sin(2 * np.pi * 0.1 * t)),
                aspect=tr["aspect"], skin=tr["skin"], eye_dx=tr["eye_dx"], mouth_dy=tr["mouth_dy"])
            if self.jitter > 0:
                j = self.jitter
                f.eye_dx += jr.uniform(-0.05, 0.05) * j
                f.eye_dy += jr.uniform(-0.05, 0.05) * j
                f.mouth_dy += jr.uniform(-0.06, 0.06) * j
                f.mouth_w *= 1.0 + jr.uniform(-0.5, 0.5) * j
                f.hair = float(np.clip(f.hair + jr.uniform(-0.12, 0.12) * j, 0.05, 0.6))
                f.skin = tuple(float(np.clip(c + jr.uniform(-40, 40) * j, 0, 255)) for c in f.skin)
                f.tilt = float(jr.uniform(-12, 12) * j)
            out.append(f)
        return out

    def frame(self, i: int) -> np.ndarray:
        img = self.bg.copy()
        for f in self.faces(i):
            render_face(img, f)
        if self._noise is not None:
            n = self._noise[i % len(self._noise)]
            n = np.roll(n, (i * 37) % self.height, axis=0)
            img = np.clip(img.astype(np.int16) + n, 0, 255).astype(np.uint8)
        return img

    def __iter__(self):
        for i in range(self.n_frames):
            yield self.frame(i)

    def processed_indices(self):
        stride = max(1, int(self.fps / 7))          # server/model.py:40
        return list(range(0, self.n_frames, stride))


# BASELINE.json "configs" -> generator arguments
CONFIGS = {
    "720p30_single": dict(height=720, width=1280, fps=30, n_frames=1800, n_faces=(1, 1), face_h=(160.0, 260.0)),
    "1080p60_multi": dict(height=1080, width=1920, fps=60, n_frames=600, n_faces=(4, 8), face_h=(60.0, 300.0)),
    "360p30_single": dict(height=360, width=640, fps=30, n_frames=960, n_faces=(1, 1), face_h=(80.0, 130.0)),
}

"""Oracle, mode B (TEST INFRASTRUCTURE ONLY): the upstream-canonical face crop of facenet_pytorch 2.6.0 that the north
star words ("crop/resize/prewhiten to 160x160"), as opposed to what the reference actually does (cv2.resize to 80x80 +
to_tensor, server/model.py:55-58; SURVEY.md section 0 and Appendix A "Mode-B extras").

Restated from upstream ``models/utils/detect_face.py`` (not on disk, parity unpinned like the rest of the oracle):

* ``extract_face(img, box, image_size=160, margin=0)``: margin-adjusted integer box, ``crop_resize`` =
  ``cv2.resize(img[y1:y2, x1:x2], (image_size, image_size), interpolation=cv2.INTER_AREA)`` for ndarray input,
  ``F.to_tensor(np.float32(face))`` (float input: no /255);
* ``fixed_image_standardization(x) = (x - 127.5) / 128.0`` (what ``MTCNN.forward`` applies with ``post_process=True``);
* ``keep_all``: every detected box is cropped and embedded, not only the largest.

``resize_area_u8`` restates OpenCV's INTER_AREA for uint8 (the three code paths of cv::resize: integer-ratio fast path,
general float area path, and the bilinear fixed-point path with "area" coefficients that INTER_AREA degrades to whenever
an axis is enlarged).  It is pinned bit for bit against the installed ``cv2.resize`` in tests/test_oracle.py and is what
``crop_area_kernel`` (csrc/preproc.cu) mirrors.
"""
from __future__ import annotations

import math

import cv2
import numpy as np
import torch

COEF_SCALE = 1 << 11          # INTER_RESIZE_COEF_SCALE


def extract_box(box, image_size: int, margin: int, width: int, height: int):
    """Integer crop box of upstream extract_face: the arithmetic runs on the numpy float32 scalars of ``box``."""
    box = np.asarray(box, np.float32)
    m = [margin * (box[2] - box[0]) / (image_size - margin), margin * (box[3] - box[1]) / (image_size - margin)]
    return [int(max(box[0] - m[0] / 2, 0)), int(max(box[1] - m[1] / 2, 0)),
            int(min(box[2] + m[0] / 2, width)), int(min(box[3] + m[1] / 2, height))]


def extract_face(img: np.ndarray, box, image_size: int = 160, margin: int = 0):
    """-> (uint8 [S,S,3] resized crop or None when the box is empty, integer box)."""
    h, w = img.shape[:2]
    b = extract_box(box, image_size, margin, w, h)
    face = img[b[1]:b[3], b[0]:b[2]]
    if face.size == 0:
        return None, b
    return cv2.resize(face, (image_size, image_size), interpolation=cv2.INTER_AREA).copy(), b


def fixed_image_standardization(face_u8: np.ndarray) -> torch.Tensor:
    """F.to_tensor(np.float32(face)) then (x - 127.5) / 128.0  -> float32 [3,S,S]."""
    t = torch.from_numpy(np.float32(face_u8)).permute(2, 0, 1).contiguous()
    return (t - 127.5) / 128.0


def embed_faces(frame: np.ndarray, boxes, facenet, image_size: int = 160, margin: int = 0):
    """keep_all: embeddings of every box of one frame -> list of (int box, float32[512] or None)."""
    out = []
    for box in boxes:
        face, b = extract_face(frame, box, image_size, margin)
        if face is None:
            out.append((b, None))
            continue
        with torch.no_grad():
            e = facenet(fixed_image_standardization(face).unsqueeze(0)).detach().numpy().flatten()
        out.append((b, e))
    return out


# ----------------------------------------------------------------------------- OpenCV INTER_AREA, restated


def _area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab: [(dst index, src index, float32 weight)] in accumulation order."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _sat_u8(v):
    """saturate_cast<uchar>(float): round half to even, clamp."""
    return np.clip(np.rint(v.astype(np.float32)).astype(np.int64), 0, 255).astype(np.uint8)


def _linear_area_coeffs(ssize: int, dsize: int, scale: float, inv_scale: float):
    idx = np.zeros(dsize, np.int64)
    c = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        s = math.floor(d * scale)
        f = np.float32((d + 1) - (s + 1) * inv_scale)
        f = np.float32(0.0) if f <= 0 else np.float32(f - math.floor(f))
        if s < 0:
            f, s = np.float32(0.0), 0
        if s >= ssize - 1:
            f, s = np.float32(0.0), ssize - 1
        idx[d] = s
        c[d] = (int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(COEF_SCALE)))),
                int(np.rint(np.float32(f * np.float32(COEF_SCALE)))))
    return idx, c


def resize_area_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 [h,w,c], bit exact."""
    sh, sw, cn = src.shape
    inv_x, inv_y = dw / sw, dh / sh
    scale_x, scale_y = 1.0 / inv_x, 1.0 / inv_y
    if scale_x >= 1 and scale_y >= 1:
        isx, isy = int(np.rint(scale_x)), int(np.rint(scale_y))
        eps = np.finfo(np.float64).eps
        if abs(scale_x - isx) < eps and abs(scale_y - isy) < eps:            # integer ratios: resizeAreaFast_
            s = src.astype(np.int64)
            acc = np.zeros((dh, dw, cn), np.int64)
            for ky in range(isy):
                for kx in range(isx):
                    acc += s[ky:ky + dh * isy:isy, kx:kx + dw * isx:isx][:dh, :dw]
            if isx == 2 and isy == 2:
                return ((acc + 2) >> 2).astype(np.uint8)
            return _sat_u8(acc.astype(np.float32) * np.float32(1.0 / (isx * isy)))
        xtab, ytab = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)    # resizeArea_<uchar, float>
        out = np.zeros((dh, dw, cn), np.uint8)
        sf = src.astype(np.float32)
        summ = np.zeros((dw, cn), np.float32)
        prev = ytab[0][0]
        for dy, sy, beta in ytab:
            buf = np.zeros((dw, cn), np.float32)
            row = sf[sy]
            for dx, sx, alpha in xtab:
                buf[dx] = buf[dx] + row[sx] * alpha          # separate fp32 multiply and add, in table order
            if dy != prev:
                out[prev] = _sat_u8(summ)
                summ = beta * buf
                prev = dy
            else:
                summ = summ + beta * buf
        out[prev] = _sat_u8(summ)
        return out
    # an axis is enlarged: bilinear fixed-point resize with area-mode coefficients
    xi, xc = _linear_area_coeffs(sw, dw, scale_x, inv_x)
    yi, yc = _linear_area_coeffs(sh, dh, scale_y, inv_y)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    out = np.zeros((dh, dw, cn), np.uint8)
    for dy in range(dh):
        sy0 = int(yi[dy])
        sy1 = min(sy0 + 1, sh - 1)
        r0 = s[sy0][xi] * xc[:, 0:1] + s[sy0][x1] * xc[:, 1:2]
        r1 = s[sy1][xi] * xc[:, 0:1] + s[sy1][x1] * xc[:, 1:2]
        b0, b1 = int(yc[dy][0]), int(yc[dy][1])
        out[dy] = np.clip((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)
    return out

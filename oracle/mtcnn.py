"""Oracle: MTCNN face detector, CPU fp32 restatement (TEST INFRASTRUCTURE ONLY).

Follows the algorithm the reference invokes at ``server/model.py:18`` (``MTCNN()``)
and ``server/model.py:47`` (``mtcnn.detect(frame)``), i.e. facenet_pytorch==2.6.0
``models/mtcnn.py`` (PNet/RNet/ONet/MTCNN.detect) and
``models/utils/detect_face.py`` (detect_face, generateBoundingBox, bbreg, rerec,
pad, imresample, batched_nms_numpy, fixed_batch_process) as restated in
SURVEY.md Appendix A.  The package is not on disk; see oracle/__init__.py
("parity unpinned").  Attribute names match upstream so upstream ``.pt``
state-dicts load with ``strict=True``.

The cascade is organised as three stage functions that also record a trace of
every intermediate the CUDA path is compared against stage by stage.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn.functional import interpolate
from torchvision.ops import batched_nms

__all__ = ["PNet", "RNet", "ONet", "MTCNN", "detect_face", "pyramid_scales"]


class PNet(nn.Module):
    """12x12 fully-convolutional proposal net (upstream models/mtcnn.py PNet)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 10, kernel_size=3)
        self.prelu1 = nn.PReLU(10)
        self.pool1 = nn.MaxPool2d(2, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(10, 16, kernel_size=3)
        self.prelu2 = nn.PReLU(16)
        self.conv3 = nn.Conv2d(16, 32, kernel_size=3)
        self.prelu3 = nn.PReLU(32)
        self.conv4_1 = nn.Conv2d(32, 2, kernel_size=1)
        self.softmax4_1 = nn.Softmax(dim=1)
        self.conv4_2 = nn.Conv2d(32, 4, kernel_size=1)

    def forward(self, x):
        x = self.pool1(self.prelu1(self.conv1(x)))
        x = self.prelu2(self.conv2(x))
        x = self.prelu3(self.conv3(x))
        a = self.softmax4_1(self.conv4_1(x))
        b = self.conv4_2(x)
        return b, a


class RNet(nn.Module):
    """24x24 refinement net (upstream models/mtcnn.py RNet)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 28, kernel_size=3)
        self.prelu1 = nn.PReLU(28)
        self.pool1 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(28, 48, kernel_size=3)
        self.prelu2 = nn.PReLU(48)
        self.pool2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv3 = nn.Conv2d(48, 64, kernel_size=2)
        self.prelu3 = nn.PReLU(64)
        self.dense4 = nn.Linear(576, 128)
        self.prelu4 = nn.PReLU(128)
        self.dense5_1 = nn.Linear(128, 2)
        self.softmax5_1 = nn.Softmax(dim=1)
        self.dense5_2 = nn.Linear(128, 4)

    def forward(self, x):
        x = self.pool1(self.prelu1(self.conv1(x)))
        x = self.pool2(self.prelu2(self.conv2(x)))
        x = self.prelu3(self.conv3(x))
        # flatten order is (W, H, C): upstream permutes before the view
        x = x.permute(0, 3, 2, 1).contiguous()
        x = self.prelu4(self.dense4(x.view(x.shape[0], -1)))
        a = self.softmax5_1(self.dense5_1(x))
        b = self.dense5_2(x)
        return b, a


class ONet(nn.Module):
    """48x48 output net (upstream models/mtcnn.py ONet)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, kernel_size=3)
        self.prelu1 = nn.PReLU(32)
        self.pool1 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3)
        self.prelu2 = nn.PReLU(64)
        self.pool2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=3)
        self.prelu3 = nn.PReLU(64)
        self.pool3 = nn.MaxPool2d(2, 2, ceil_mode=True)
        self.conv4 = nn.Conv2d(64, 128, kernel_size=2)
        self.prelu4 = nn.PReLU(128)
        self.dense5 = nn.Linear(1152, 256)
        self.prelu5 = nn.PReLU(256)
        self.dense6_1 = nn.Linear(256, 2)
        self.softmax6_1 = nn.Softmax(dim=1)
        self.dense6_2 = nn.Linear(256, 4)
        self.dense6_3 = nn.Linear(256, 10)

    def forward(self, x):
        x = self.pool1(self.prelu1(self.conv1(x)))
        x = self.pool2(self.prelu2(self.conv2(x)))
        x = self.pool3(self.prelu3(self.conv3(x)))
        x = self.prelu4(self.conv4(x))
        x = x.permute(0, 3, 2, 1).contiguous()
        x = self.prelu5(self.dense5(x.view(x.shape[0], -1)))
        a = self.softmax6_1(self.dense6_1(x))
        b = self.dense6_2(x)
        c = self.dense6_3(x)
        return b, c, a


# --------------------------------------------------------------------------- helpers


def pyramid_scales(h: int, w: int, minsize: int = 20, factor: float = 0.709):
    """Scale list, in Python doubles exactly as upstream detect_face builds it."""
    m = 12.0 / minsize
    minl = min(h, w) * m
    scale_i = m
    scales = []
    while minl >= 12:
        scales.append(scale_i)
        scale_i = scale_i * factor
        minl = minl * factor
    return scales


def imresample(img, sz):
    """upstream detect_face.imresample: area interpolation (adaptive average)."""
    return interpolate(img, size=sz, mode="area")


def generate_bounding_box(reg, probs, scale, thresh):
    """upstream generateBoundingBox: P-Net cells with prob >= thresh -> 9-col rows."""
    stride = 2
    cellsize = 12
    reg = reg.permute(1, 0, 2, 3)
    mask = probs >= thresh
    mask_inds = mask.nonzero()
    image_inds = mask_inds[:, 0]
    score = probs[mask]
    reg = reg[:, mask].permute(1, 0)
    bb = mask_inds[:, 1:].type(reg.dtype).flip(1)
    q1 = ((stride * bb + 1) / scale).floor()
    q2 = ((stride * bb + cellsize - 1 + 1) / scale).floor()
    return torch.cat([q1, q2, score.unsqueeze(1), reg], dim=1), image_inds


def bbreg(boundingbox, reg):
    """upstream bbreg: regression with the +1 width/height convention."""
    if reg.shape[1] == 1:
        reg = torch.reshape(reg, (reg.shape[2], reg.shape[3]))
    w = boundingbox[:, 2] - boundingbox[:, 0] + 1
    h = boundingbox[:, 3] - boundingbox[:, 1] + 1
    b1 = boundingbox[:, 0] + reg[:, 0] * w
    b2 = boundingbox[:, 1] + reg[:, 1] * h
    b3 = boundingbox[:, 2] + reg[:, 2] * w
    b4 = boundingbox[:, 3] + reg[:, 3] * h
    boundingbox[:, :4] = torch.stack([b1, b2, b3, b4]).permute(1, 0)
    return boundingbox


def rerec(bbox):
    """upstream rerec: make boxes square around their centre (in place)."""
    h = bbox[:, 3] - bbox[:, 1]
    w = bbox[:, 2] - bbox[:, 0]
    side = torch.max(w, h)
    bbox[:, 0] = bbox[:, 0] + w * 0.5 - side * 0.5
    bbox[:, 1] = bbox[:, 1] + h * 0.5 - side * 0.5
    bbox[:, 2:4] = bbox[:, :2] + side.repeat(2, 1).permute(1, 0)
    return bbox


def pad(boxes, w, h):
    """upstream pad: trunc to int, clamp to the 1-based image rectangle."""
    boxes = boxes.trunc().int().cpu().numpy()
    x = boxes[:, 0]
    y = boxes[:, 1]
    ex = boxes[:, 2]
    ey = boxes[:, 3]
    x[x < 1] = 1
    y[y < 1] = 1
    ex[ex > w] = w
    ey[ey > h] = h
    return y, ey, x, ex


def nms_numpy(boxes, scores, threshold, method):
    """upstream nms_numpy (+1 areas, ascending argsort, keep o <= threshold)."""
    if boxes.size == 0:
        return np.empty((0, 3))
    x1 = boxes[:, 0].copy()
    y1 = boxes[:, 1].copy()
    x2 = boxes[:, 2].copy()
    y2 = boxes[:, 3].copy()
    s = scores
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.argsort(s)
    pick = np.zeros_like(s, dtype=np.int16)
    counter = 0
    while order.size > 0:
        i = order[-1]
        pick[counter] = i
        counter += 1
        idx = order[0:-1]
        xx1 = np.maximum(x1[i], x1[idx]).copy()
        yy1 = np.maximum(y1[i], y1[idx]).copy()
        xx2 = np.minimum(x2[i], x2[idx]).copy()
        yy2 = np.minimum(y2[i], y2[idx]).copy()
        w = np.maximum(0.0, xx2 - xx1 + 1).copy()
        h = np.maximum(0.0, yy2 - yy1 + 1).copy()
        inter = w * h
        if method == "Min":
            o = inter / np.minimum(area[i], area[idx])
        else:
            o = inter / (area[i] + area[idx] - inter)
        order = order[np.where(o <= threshold)]
    return pick[:counter].copy()


def batched_nms_numpy(boxes, scores, idxs, threshold, method):
    """upstream batched_nms_numpy: coordinate-offset trick then nms_numpy."""
    device = boxes.device
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=device)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    boxes_for_nms = boxes + offsets[:, None]
    boxes_for_nms = boxes_for_nms.cpu().numpy()
    scores = scores.cpu().numpy()
    keep = nms_numpy(boxes_for_nms, scores, threshold, method)
    return torch.as_tensor(keep, dtype=torch.long, device=device)


def fixed_batch_process(im_data, model, batch_size=512):
    """upstream fixed_batch_process: run in chunks of 512 and concatenate."""
    out = []
    for i in range(0, len(im_data), batch_size):
        out.append(model(im_data[i:(i + batch_size)]))
    return tuple(torch.cat(v, dim=0) for v in zip(*out))


def _crop_resample(imgs, image_inds, y, ey, x, ex, size):
    """Stage-2/3 input: clipped (not zero padded) crop -> area resample -> normalise."""
    im_data = []
    valid = []
    for k in range(len(y)):
        ok = bool(ey[k] > (y[k] - 1) and ex[k] > (x[k] - 1))
        valid.append(ok)
        if ok:
            img_k = imgs[image_inds[k], :, (y[k] - 1):ey[k], (x[k] - 1):ex[k]].unsqueeze(0)
            im_data.append(imresample(img_k, (size, size)))
    if not im_data:
        return torch.zeros(0, 3, size, size), valid
    im_data = torch.cat(im_data, dim=0)
    return (im_data - 127.5) * 0.0078125, valid


# --------------------------------------------------------------------------- cascade


def detect_face(imgs, minsize, pnet, rnet, onet, threshold, factor, device=None, trace=None):
    """upstream detect_face restated; returns (batch_boxes, batch_points).

    ``trace`` (optional dict) receives every intermediate tensor, keyed by stage.
    """
    if isinstance(imgs, np.ndarray):
        imgs = torch.as_tensor(imgs.copy(), device=device)
    imgs = torch.as_tensor(imgs, device=device)
    if imgs.dim() == 3:
        imgs = imgs.unsqueeze(0)
    model_dtype = next(pnet.parameters()).dtype
    imgs = imgs.permute(0, 3, 1, 2).type(model_dtype)
    batch_size = len(imgs)
    h, w = imgs.shape[2:4]
    scales = pyramid_scales(h, w, minsize, factor)
    tr = trace if trace is not None else {}
    tr["scales"] = scales

    # ---- stage 1: P-Net over the pyramid, per-scale NMS 0.5
    boxes, image_inds, scale_picks = [], [], []
    tr["pyramid"], tr["pnet_prob"], tr["pnet_reg"], tr["s1_per_scale"] = [], [], [], []
    offset = 0
    for scale in scales:
        im_data = imresample(imgs, (int(h * scale + 1), int(w * scale + 1)))
        im_data = (im_data - 127.5) * 0.0078125
        reg, probs = pnet(im_data)
        if trace is not None:
            tr["pyramid"].append(im_data.contiguous().clone())
            tr["pnet_prob"].append(probs[:, 1].contiguous().clone())
            tr["pnet_reg"].append(reg.contiguous().clone())
        boxes_scale, image_inds_scale = generate_bounding_box(reg, probs[:, 1], scale, threshold[0])
        boxes.append(boxes_scale)
        image_inds.append(image_inds_scale)
        pick = batched_nms(boxes_scale[:, :4], boxes_scale[:, 4], image_inds_scale, 0.5)
        if trace is not None:
            tr["s1_per_scale"].append((boxes_scale.clone(), image_inds_scale.clone(), pick.clone()))
        scale_picks.append(pick + offset)
        offset += boxes_scale.shape[0]

    boxes = torch.cat(boxes, dim=0)
    image_inds = torch.cat(image_inds, dim=0)
    scale_picks = torch.cat(scale_picks, dim=0)
    boxes, image_inds = boxes[scale_picks], image_inds[scale_picks]

    # ---- cross-scale NMS 0.7, stage-1 regression (no +1), square, pad
    pick = batched_nms(boxes[:, :4], boxes[:, 4], image_inds, 0.7)
    boxes, image_inds = boxes[pick], image_inds[pick]
    regw = boxes[:, 2] - boxes[:, 0]
    regh = boxes[:, 3] - boxes[:, 1]
    qq1 = boxes[:, 0] + boxes[:, 5] * regw
    qq2 = boxes[:, 1] + boxes[:, 6] * regh
    qq3 = boxes[:, 2] + boxes[:, 7] * regw
    qq4 = boxes[:, 3] + boxes[:, 8] * regh
    boxes = torch.stack([qq1, qq2, qq3, qq4, boxes[:, 4]]).permute(1, 0)
    boxes = rerec(boxes)
    y, ey, x, ex = pad(boxes, w, h)
    if trace is not None:
        tr["s1_boxes"] = boxes.clone()
        tr["s1_inds"] = image_inds.clone()
        tr["s1_pad"] = np.stack([y, ey, x, ex], axis=1).copy() if len(y) else np.zeros((0, 4), np.int32)

    # ---- stage 2: R-Net on 24x24 crops
    if len(boxes) > 0:
        im_data, valid = _crop_resample(imgs, image_inds, y, ey, x, ex, 24)
        if not all(valid):
            # upstream would raise on the shape mismatch; degenerate boxes are dropped here
            keep = torch.as_tensor(valid)
            boxes, image_inds = boxes[keep], image_inds[keep]
        if trace is not None:
            tr["rnet_in"] = im_data.clone()
        out = fixed_batch_process(im_data, rnet) if len(im_data) else (torch.zeros(0, 4), torch.zeros(0, 2))
        out0 = out[0].permute(1, 0)
        out1 = out[1].permute(1, 0)
        score = out1[1, :]
        if trace is not None:
            tr["rnet_score"] = score.clone()
            tr["rnet_reg"] = out[0].clone()
        ipass = score > threshold[1]
        boxes = torch.cat((boxes[ipass, :4], score[ipass].unsqueeze(1)), dim=1)
        image_inds = image_inds[ipass]
        mv = out0[:, ipass].permute(1, 0)
        pick = batched_nms(boxes[:, :4], boxes[:, 4], image_inds, 0.7)
        boxes, image_inds, mv = boxes[pick], image_inds[pick], mv[pick]
        boxes = bbreg(boxes, mv)
        boxes = rerec(boxes)
    if trace is not None:
        tr["s2_boxes"] = boxes.clone()
        tr["s2_inds"] = image_inds.clone()

    # ---- stage 3: O-Net on 48x48 crops, "Min" NMS
    points = torch.zeros(0, 5, 2, device=device)
    if len(boxes) > 0:
        y, ey, x, ex = pad(boxes, w, h)
        im_data, valid = _crop_resample(imgs, image_inds, y, ey, x, ex, 48)
        if not all(valid):
            keep = torch.as_tensor(valid)
            boxes, image_inds = boxes[keep], image_inds[keep]
        if trace is not None:
            tr["s2_pad"] = np.stack([y, ey, x, ex], axis=1).copy()
            tr["onet_in"] = im_data.clone()
        out = fixed_batch_process(im_data, onet) if len(im_data) else (
            torch.zeros(0, 4), torch.zeros(0, 10), torch.zeros(0, 2))
        out0 = out[0].permute(1, 0)
        out1 = out[1].permute(1, 0)
        out2 = out[2].permute(1, 0)
        score = out2[1, :]
        if trace is not None:
            tr["onet_score"] = score.clone()
            tr["onet_reg"] = out[0].clone()
        points = out1
        ipass = score > threshold[2]
        points = points[:, ipass]
        boxes = torch.cat((boxes[ipass, :4], score[ipass].unsqueeze(1)), dim=1)
        image_inds = image_inds[ipass]
        mv = out0[:, ipass].permute(1, 0)
        w_i = boxes[:, 2] - boxes[:, 0] + 1
        h_i = boxes[:, 3] - boxes[:, 1] + 1
        points_x = w_i.repeat(5, 1) * points[:5, :] + boxes[:, 0].repeat(5, 1) - 1
        points_y = h_i.repeat(5, 1) * points[5:10, :] + boxes[:, 1].repeat(5, 1) - 1
        points = torch.stack((points_x, points_y)).permute(2, 1, 0)
        boxes = bbreg(boxes, mv)
        pick = batched_nms_numpy(boxes[:, :4], boxes[:, 4], image_inds, 0.7, "Min")
        boxes, image_inds, points = boxes[pick], image_inds[pick], points[pick]

    boxes = boxes.cpu().numpy()
    points = points.cpu().numpy()
    image_inds = image_inds.cpu().numpy()
    batch_boxes, batch_points = [], []
    for b_i in range(batch_size):
        sel = np.where(image_inds == b_i)
        batch_boxes.append(boxes[sel].copy())
        batch_points.append(points[sel].copy())
    if trace is not None:
        tr["final"] = [b.copy() for b in batch_boxes]
    return batch_boxes, batch_points


class MTCNN(nn.Module):
    """upstream MTCNN with the defaults the reference uses (``MTCNN()``, server/model.py:18).

    Only ``detect`` is restated: the reference never calls ``forward``/``extract``.
    """

    def __init__(self, image_size=160, margin=0, min_face_size=20, thresholds=(0.6, 0.7, 0.7),
                 factor=0.709, post_process=True, select_largest=True, keep_all=False, device=None):
        super().__init__()
        self.image_size = image_size
        self.margin = margin
        self.min_face_size = min_face_size
        self.thresholds = list(thresholds)
        self.factor = factor
        self.post_process = post_process
        self.select_largest = select_largest
        self.keep_all = keep_all
        self.pnet = PNet()
        self.rnet = RNet()
        self.onet = ONet()
        self.device = torch.device("cpu")
        self.eval()

    def detect(self, img, trace=None):
        """Returns (boxes [N,4] float32 sorted largest-area first | None, probs)."""
        with torch.no_grad():
            batch_boxes, _ = detect_face(img, self.min_face_size, self.pnet, self.rnet, self.onet,
                                         self.thresholds, self.factor, self.device, trace=trace)
        box = batch_boxes[0]
        if len(box) == 0:
            return None, [None]
        if self.select_largest:
            order = np.argsort((box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1]))[::-1]
            box = box[order]
        return box[:, :4], box[:, 4]

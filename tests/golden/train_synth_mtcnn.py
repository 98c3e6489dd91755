"""Fixture generator: train the oracle's P/R/O-Net on the synthetic-face generator.

The pretrained ``{p,r,o}net.pt`` ship inside the facenet_pytorch wheel, which is absent
offline, and random MTCNN weights give meaningless candidate
densities.  This script produces *trained-like* stand-in weights: the
three nets (oracle/mtcnn.py modules, upstream layer names) are fitted on crops of the
synthetic faces of ``synth.py`` with the usual MTCNN targets (face/non-face
cross-entropy on IoU>=0.65 / <0.3 crops, box-offset regression on IoU>=0.4 crops).
Output: ``<package>/data/synth_mtcnn.npz`` (float32 state dict, keys ``pnet.conv1.weight`` ...).
Both the oracle and the CUDA path load that same file, so parity does not depend on how
good the detector is; the fit only makes the candidate counts per stage realistic.

Run:  python tests/golden/train_synth_mtcnn.py   (CPU, ~10 min on 8 cores, seeded)
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
import time

import cv2
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import truely_b200  # noqa: E402,F401
from truely_b200.synth import FaceSpec, render_face  # noqa: E402
from oracle.mtcnn import ONet, PNet, RNet  # noqa: E402

OUT = os.path.join(ROOT, truely_b200.__name__ if False else
                   "truely-real-time-ai-generated-video-detection-framework-for-social-platforms_b200",
                   "data", "synth_mtcnn.npz")


def iou(a, b):
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    x2, y2 = min(a[2], b[2]), min(a[3], b[3])
    inter = max(0.0, x2 - x1) * max(0.0, y2 - y1)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def random_face(rng, cx, cy, h):
    return FaceSpec(
        cx=cx, cy=cy, h=h, aspect=0.76 + rng.uniform(-0.05, 0.05),
        skin=(120 + rng.uniform(-25, 25), 150 + rng.uniform(-25, 25), 200 + rng.uniform(-25, 25)),
        eye_dx=0.19 + rng.uniform(-0.04, 0.04), eye_dy=-0.10 + rng.uniform(-0.04, 0.04),
        mouth_dy=0.26 + rng.uniform(-0.05, 0.05), mouth_w=0.20 * (1 + rng.uniform(-0.4, 0.4)),
        hair=float(np.clip(0.30 + rng.uniform(-0.12, 0.12), 0.05, 0.6)), tilt=float(rng.uniform(-12, 12)))


def background(rng, R):
    """Noise background as a crop of the frame background seen at a random pyramid scale."""
    cell = float(np.exp(rng.uniform(np.log(0.6), np.log(24.0))))      # canvas px per noise cell
    g = max(2, int(np.ceil(R / cell)) + 2)
    grid = rng.integers(0, 256, (g, g, 3), dtype=np.uint8)
    big = cv2.resize(grid, (int(np.ceil(g * cell)) + 1,) * 2, interpolation=cv2.INTER_LINEAR)
    o = int(rng.integers(0, max(1, big.shape[0] - R)))
    p = int(rng.integers(0, max(1, big.shape[1] - R)))
    out = big[o:o + R, p:p + R]
    if out.shape[0] < R or out.shape[1] < R:
        out = cv2.resize(out, (R, R))
    return np.ascontiguousarray(out)


def make_sample(rng, size, kind):
    """-> (uint8 [size,size,3], label in {1,0,-1}, reg[4]).  kind: 'pos' | 'neg' | 'any'."""
    R = size * 4
    crop = (0.0, 0.0, float(R), float(R))
    for _ in range(50):
        img = background(rng, R)
        if kind == "neg" and rng.uniform() < 0.35:
            label, reg = 0, np.zeros(4, np.float32)          # pure background
            break
        if kind == "pos":
            fh = R * rng.uniform(0.85, 1.25)
            cx, cy = R / 2 + rng.uniform(-0.12, 0.12) * R, R / 2 + rng.uniform(-0.12, 0.12) * R
        else:
            fh = R * float(np.exp(rng.uniform(np.log(0.15), np.log(6.0))))
            cx, cy = R / 2 + rng.uniform(-0.8, 0.8) * max(R, fh * 0.6), R / 2 + rng.uniform(-0.8, 0.8) * max(R, fh * 0.6)
        f = random_face(rng, cx, cy, fh)
        render_face(img, f)
        gt = f.box()
        v = iou(crop, gt)
        if v >= 0.65:
            label = 1
        elif v >= 0.4:
            label = -1
        elif v < 0.3:
            label = 0
        else:
            continue
        if kind == "pos" and label == 0:
            continue
        if kind == "neg" and label != 0:
            continue
        reg = np.array([gt[0] / R, gt[1] / R, (gt[2] - R) / R, (gt[3] - R) / R], np.float32)
        break
    else:
        label, reg = 0, np.zeros(4, np.float32)
    img = np.clip(img.astype(np.int16) + np.rint(rng.normal(0, 2.0, img.shape)), 0, 255).astype(np.uint8)
    small = cv2.resize(img, (size, size), interpolation=cv2.INTER_AREA)
    return small, label, reg


def make_dataset(seed, size, n):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, size, size, 3), np.uint8)
    L = np.zeros(n, np.int64)
    G = np.zeros((n, 4), np.float32)
    for i in range(n):
        kind = ("pos", "neg", "any", "pos", "neg")[i % 5]
        X[i], L[i], G[i] = make_sample(rng, size, kind)
    return X, L, G


def net_outputs(net, x):
    out = net(x)
    reg, prob = out[0], out[-1]
    if prob.dim() == 4:
        prob, reg = prob[:, :, 0, 0], reg[:, :, 0, 0]
    return reg, prob


def train(net, X, L, G, steps, batch, lr, seed, name, smooth):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    xs = ((torch.from_numpy(X).permute(0, 3, 1, 2).float() - 127.5) * 0.0078125).contiguous()
    ls, gs = torch.from_numpy(L), torch.from_numpy(G)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps)
    net.train()
    t0 = time.time()
    for it in range(steps):
        idx = torch.randint(0, len(xs), (batch,), generator=g)
        x, l, gt = xs[idx], ls[idx], gs[idx]
        if torch.rand((), generator=g) < 0.5:                 # horizontal flip augmentation
            x = x.flip(3)
            gt = torch.stack([-gt[:, 2], gt[:, 1], -gt[:, 0], gt[:, 3]], 1)
        reg, prob = net_outputs(net, x)
        cls_mask = l >= 0
        p1 = prob[:, 1].clamp(1e-7, 1 - 1e-7)
        # label smoothing keeps the logits moderate: no saturated (tied) probabilities, and the hard
        # negatives that cross the 0.6 / 0.7 thresholds give candidate densities like a real detector's
        tgt = l[cls_mask].float() * (1.0 - 2.0 * smooth) + smooth
        cls_loss = F.binary_cross_entropy(p1[cls_mask], tgt)
        reg_mask = l != 0
        reg_loss = F.mse_loss(reg[reg_mask], gt[reg_mask]) if reg_mask.any() else reg.sum() * 0
        loss = cls_loss + 0.5 * reg_loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if it % 500 == 0 or it == steps - 1:
            with torch.no_grad():
                acc = ((p1 > 0.5) == (l == 1))[cls_mask].float().mean().item()
            print(f"[{name}] it {it:5d} loss {loss.item():.4f} cls {cls_loss.item():.4f} reg {reg_loss.item():.5f} "
                  f"acc {acc:.3f}  {time.time() - t0:.0f}s", flush=True)
    net.eval()
    return net


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = [("pnet", PNet, 12, 90000, 6000, 512, 2e-3, 0.20),
           ("rnet", RNet, 24, 60000, 3000, 256, 1e-3, 0.10),
           ("onet", ONet, 48, 30000, 2000, 128, 1e-3, 0.03)]
    state = {}
    for k, (name, cls, size, n, steps, batch, lr, smooth) in enumerate(cfg):
        t0 = time.time()
        X, L, G = make_dataset(1234 + k, size, n)
        print(f"[{name}] dataset {X.shape} pos {(L == 1).sum()} part {(L == -1).sum()} neg {(L == 0).sum()} "
              f"in {time.time() - t0:.0f}s", flush=True)
        net = train(cls(), X, L, G, steps, batch, lr, 99 + k, name, smooth)
        for key, val in net.state_dict().items():
            state[f"{name}.{key}"] = val.detach().cpu().numpy().astype(np.float32)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez(OUT, **state)
    print("wrote", OUT, sum(v.size for v in state.values()), "params")


if __name__ == "__main__":
    main()

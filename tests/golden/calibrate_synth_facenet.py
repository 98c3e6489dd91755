"""Fixture generator: BatchNorm statistics for the seeded stand-in InceptionResnetV1.

``weights.synth_facenet_state`` draws He-normal conv kernels from a fixed numpy seed;
this script measures, once, the per-channel BatchNorm statistics of that network on a
seeded batch of synthetic face crops (prepared exactly like the reference prepares its
FaceNet input: crop -> cv2.resize 80x80 -> /255, BGR; server/model.py:55-58) and writes
them to ``<package>/data/synth_facenet_bn.npz``.  The measurement runs the oracle module in
float64 so the stored float32 statistics do not depend on the host's conv kernels.

SURVEY.md section 7 H1: default init collapses (every embedding identical) and naive
calibration is chaotic.  ``BETA`` shifts every BN output up so most units stay in the linear
region of the ReLU, which keeps the network smooth in its input, and ``LAST_ALPHA`` conditions
the final BatchNorm1d (see below).  Measured with these defaults on synthetic clips: consecutive-frame
cosine 0.994 (stable face) / 0.986 (jitter=1.0), bf16-storage vs fp32 1-cos <= 4e-5.

Run:  python tests/golden/calibrate_synth_facenet.py
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import truely_b200  # noqa: E402,F401
from truely_b200 import weights as W  # noqa: E402
from truely_b200.synth import SyntheticClip  # noqa: E402
from oracle.inception_resnet_v1 import InceptionResnetV1  # noqa: E402

BETA = float(os.environ.get("CAL_BETA", "2.0"))
GAMMA = float(os.environ.get("CAL_GAMMA", "1.0"))
# last_bn.running_mean = LAST_ALPHA * measured mean.  alpha = 1 would centre the embedding (population cosine ~ 0);
# a negative alpha leaves a common component so that similarities of consecutive synthetic frames straddle the
# reference's 0.99 threshold (stable clip ~0.994, jitter-1.0 clip ~0.986) and bf16 storage costs ~4e-5 in cosine.
LAST_ALPHA = float(os.environ.get("CAL_LAST_ALPHA", "-2.0"))
OUT = os.path.join(W.DATA_DIR, "synth_facenet_bn.npz")


def face_crops(n, seed, jitter=0.0, size=80):
    """n crops [n,3,size,size] float32 in [0,1], BGR, from n different synthetic clips/frames."""
    out = []
    for i in range(n):
        clip = SyntheticClip(360, 640, 30, 64, n_faces=(1, 1), face_h=(80.0, 200.0), jitter=jitter, seed=seed + i)
        j = (i * 7) % 64
        frame = clip.frame(j)
        x1, y1, x2, y2 = clip.faces(j)[0].box()
        rng = np.random.default_rng(seed * 977 + i)
        dx, dy = rng.uniform(-0.06, 0.06, 2) * (x2 - x1)
        x1, x2 = int(max(0, x1 + dx)), int(min(640, x2 + dx))
        y1, y2 = int(max(0, y1 + dy)), int(min(360, y2 + dy))
        face = cv2.resize(frame[y1:y2, x1:x2], (size, size))
        out.append(torch.from_numpy(face).permute(2, 0, 1).float().div(255))
    return torch.stack(out)


def build_model(sd):
    m = InceptionResnetV1()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    return m.eval()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = W.synth_facenet_state()                      # uncalibrated if the npz is absent
    for k in list(sd):
        if k.endswith("bn.weight") or k == "last_bn.weight":
            sd[k] = np.full_like(sd[k], GAMMA)
        if k.endswith("bn.bias"):
            sd[k] = np.full_like(sd[k], BETA)
        if k == "last_bn.bias":
            sd[k] = np.zeros_like(sd[k])
        if k.endswith("running_mean"):
            sd[k] = np.zeros_like(sd[k])
        if k.endswith("running_var"):
            sd[k] = np.ones_like(sd[k])
    m = build_model(sd).double()
    bns = [mod for mod in m.modules() if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d))]
    for b in bns:
        b.momentum = None           # cumulative average -> running stats = batch stats of the one batch
        b.reset_running_stats()
        b.train()
    x = face_crops(96, seed=7000).double()
    with torch.no_grad():
        m(x)
    m.eval()
    out = {}
    for name, mod in m.named_modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            out[f"{name}.weight"] = mod.weight.detach().float().numpy()
            out[f"{name}.bias"] = mod.bias.detach().float().numpy()
            mean = mod.running_mean.detach().float().numpy()
            out[f"{name}.running_mean"] = mean * np.float32(LAST_ALPHA) if name == "last_bn" else mean
            out[f"{name}.running_var"] = mod.running_var.detach().float().numpy()
    np.savez(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()

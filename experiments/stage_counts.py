"""Candidate counts per cascade stage on the bench workloads (what R-Net / O-Net / NMS actually process).
Usage (GPU box): python experiments/stage_counts.py [workload]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import truely_b200  # noqa: E402,F401
from truely_b200 import model  # noqa: E402
from truely_b200.synth import CONFIGS, SyntheticClip  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "720p30_single"
cfg = dict(CONFIGS[name])
clip = SyntheticClip(**cfg, jitter=1.2, seed=0)
stride = max(1, int(clip.fps / 7))
frames = np.stack([clip.frame(i) for i in range(0, min(cfg["n_frames"], 90 * stride), stride)])
an = model.Analyzer(device=0)
res = an.process_frames(frames, detail=True)
c = res.counts.astype(np.float64)
names = ["after stage-1 NMS (per frame)", "R-Net inputs", "O-Net inputs", "faces"]
print(f"{name}: {len(frames)} frames")
for k, n in enumerate(names):
    print(f"  {n:32s} mean {c[:, k].mean():8.2f}  max {c[:, k].max():6.0f}")

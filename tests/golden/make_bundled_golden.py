"""Fixture generator for BASELINE.json configs[0]: the reference's bundled test clip through the CPU oracle.

The reference's only fixture is `test/Google's new AI video tool Veo 3 is WILD! - Impekable (360p, h264).mp4`
(640x360, 30 fps, 960 frames -> stride 4 -> 240 processed frames; SURVEY.md 8d "Config 1").  /root/reference does not exist
on the GPU box, so this script (run once in the build container, where it does)
  1. copies the clip, byte for byte, to tests/golden/bundled_veo3_360p.mp4 (a data fixture, not source), and
  2. runs oracle.reference_run over it (same OpenCV decode as server/model.py:23,43) with the committed stand-in weights
     and writes the per-frame trace + score to tests/golden/bundled_clip_run.npz -- the golden the GPU path is compared
     with on the same decoded frames (tests/test_gpu_e2e.py::test_bundled_clip_matches_oracle_golden), and that the
     oracle itself is pinned against (tests/test_oracle.py::test_bundled_clip_golden_pins_the_oracle).
No expected output for this clip exists in the reference (README gives none): the golden is the oracle's, parity unpinned.

Run:  python tests/golden/make_bundled_golden.py
"""
import glob
import os
import shutil
import sys
import time

os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle.reference_run import reference_run  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(OUT, "bundled_veo3_360p.mp4")


def main():
    src = glob.glob("/root/reference/test/*.mp4")
    if src:
        shutil.copyfile(src[0], FIXTURE)
    assert os.path.exists(FIXTURE), "bundled clip not found"
    torch.set_num_threads(os.cpu_count() or 1)
    mt, fn = H.oracle_mtcnn(), H.oracle_facenet()
    t0 = time.perf_counter()
    tr = reference_run(FIXTURE, None, mt, fn)
    dt = time.perf_counter() - t0
    fr = tr.frames
    np.savez_compressed(
        os.path.join(OUT, "bundled_clip_run.npz"),
        fps=np.array(tr.fps), width=np.array(tr.width), height=np.array(tr.height), stride=np.array(tr.stride),
        frame_index=np.array([f.frame_index for f in fr]), n_faces=np.array([f.n_faces for f in fr]),
        box=np.array([f.box if f.box is not None else [0, 0, 0, 0] for f in fr]),
        box_f=np.array([f.box_f if f.box_f is not None else [0, 0, 0, 0] for f in fr], np.float32),
        embedded=np.array([f.embedded for f in fr]), sim=np.array([np.nan if f.sim is None else f.sim for f in fr], np.float32),
        run=np.array([f.run for f in fr]), flagged=np.array([f.flagged for f in fr]),
        emb=np.array([f.emb if f.emb is not None else np.zeros(512, np.float32) for f in fr], np.float32),
        score=np.array(tr.score), flagged_count=np.array(tr.flagged_count), final_run=np.array(tr.final_run),
        frame_count=np.array(tr.frame_count))
    sims = [f.sim for f in fr if f.sim is not None]
    print(f"bundled clip: {tr.frame_count} frames, {len(fr)} processed, {sum(f.embedded for f in fr)} with a face, "
          f"faces/frame max {max(f.n_faces for f in fr)}, score {tr.score}, flagged {tr.flagged_count}, "
          f"sims in band {sum(abs(s - 0.99) < 1e-3 for s in sims)}/{len(sims)}, oracle {dt:.1f} s on {os.cpu_count()} cores")


if __name__ == "__main__":
    main()

"""Fixture generator: golden vectors of the CPU oracle on seeded synthetic inputs (committed as tests/golden/*.npz).

The reference ships no golden vectors and facenet_pytorch cannot be imported offline, so these
are outputs of the *oracle* (oracle/), generated in this container with the committed stand-in weights; they pin the
oracle against drift and give the GPU tests fixed targets that do not require re-running the oracle.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from truely_b200.synth import SyntheticClip  # noqa: E402
from oracle.reference_run import reference_run_frames  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def golden_clip(seed=21, jitter=1.6, n_frames=200):
    return SyntheticClip(360, 640, 30, n_frames, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=jitter, seed=seed)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    mt, fn = H.oracle_mtcnn(), H.oracle_facenet()
    # 1. MTCNN on two frames (single face 360p, multi face 540p)
    det = {}
    for tag, clip, idx in (("a", golden_clip(), 0), ("b", SyntheticClip(540, 960, 60, 16, n_faces=(3, 3), face_h=(60.0, 160.0), seed=5), 3)):
        tr = {}
        boxes, probs = mt.detect(clip.frame(idx), trace=tr)
        det[f"{tag}_boxes"] = boxes.astype(np.float32)
        det[f"{tag}_probs"] = np.asarray(probs, np.float32)
        det[f"{tag}_counts"] = np.array([sum(len(x[0]) for x in tr["s1_per_scale"]), len(tr["s1_boxes"]), len(tr["s2_boxes"]), len(boxes)])
        det[f"{tag}_scales"] = np.asarray(tr["scales"], np.float64)
    np.savez(os.path.join(OUT, "mtcnn_detect.npz"), **det)
    # 2. whole reference loop on a short jittery clip (similarities on both sides of 0.99)
    clip = golden_clip()
    tr = reference_run_frames(iter(clip), clip.fps, clip.width, clip.height, mt, fn)
    fr = tr.frames
    np.savez(os.path.join(OUT, "reference_run.npz"),
             frame_index=np.array([f.frame_index for f in fr]), n_faces=np.array([f.n_faces for f in fr]),
             box=np.array([f.box if f.box is not None else [0, 0, 0, 0] for f in fr]),
             box_f=np.array([f.box_f if f.box_f is not None else [0, 0, 0, 0] for f in fr], np.float32),
             embedded=np.array([f.embedded for f in fr]), sim=np.array([np.nan if f.sim is None else f.sim for f in fr], np.float32),
             run=np.array([f.run for f in fr]), flagged=np.array([f.flagged for f in fr]),
             emb=np.array([f.emb if f.emb is not None else np.zeros(512, np.float32) for f in fr], np.float32),
             score=np.array(tr.score), flagged_count=np.array(tr.flagged_count), final_run=np.array(tr.final_run),
             frame_count=np.array(tr.frame_count))
    print("golden written; score", tr.score, "flagged", tr.flagged_count, "sims", np.round([f.sim for f in fr if f.sim is not None], 4))


if __name__ == "__main__":
    main()

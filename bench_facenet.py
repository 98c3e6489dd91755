#!/usr/bin/env python
"""BASELINE.json configs[2]: FaceNet InceptionResnetV1-only embedding sweep (tensor-core roofline).

Random uint8 crops resident in HBM -> trl_facenet -> embeddings; CUDA-event timed.  Algorithmic work =
2 x 1,417.7 MMAC = 2.835 GFLOP per 160x160 crop, 2 x 233.3 MMAC = 0.467 GFLOP per 80x80 crop.
Prints one JSON object per (S, B) and a summary; `--out` also writes them to a file (profiles/).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="160,80")
    ap.add_argument("--batches", default="64,128,256,512,1024,2048,4096")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import truely_b200  # noqa: F401
    from truely_b200.model import Analyzer, _vp
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
        {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    an = Analyzer(device=0)
    rows = []
    for S in [int(v) for v in args.sizes.split(",")]:
        flops = 2 * (1417.7e6 if S == 160 else 233.3e6)
        for B in [int(v) for v in args.batches.split(",")]:
            g = torch.Generator(device="cuda").manual_seed(0)
            crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda", generator=g)
            emb = torch.empty((B, 512), dtype=torch.float32, device="cuda")
            with torch.cuda.stream(an.stream):
                for _ in range(2):
                    an._check(an.lib.trl_facenet(an.ctx, _vp(crops), B, S, _vp(emb), an._sptr()))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(an.stream)
                for _ in range(args.iters):
                    an._check(an.lib.trl_facenet(an.ctx, _vp(crops), B, S, _vp(emb), an._sptr()))
                e1.record(an.stream)
            an.stream.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            tf = flops * B / (ms * 1e-3) / 1e12
            row = {"crop": S, "batch": B, "ms": ms, "crops_per_s": B / (ms * 1e-3), "tflops_bf16": tf,
                   "frac_of_measured_burst": tf / peaks["bf16_tflops"], "frac_of_measured_sustained": tf / peaks["bf16_tflops_sustained"],
                   "norm_ok": bool(torch.allclose(emb.norm(dim=1), torch.ones(B, device="cuda"), atol=1e-4))}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del crops, emb
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

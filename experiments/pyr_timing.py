"""Times trl_pyramid_pairs alone (CUDA events) and, with TRL_PYR_GENERIC=1, the general kernel on the same frames.
usage: python experiments/pyr_timing.py [H W B]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
import truely_b200  # noqa: F401,E402
from truely_b200 import _lib as L  # noqa: E402
from truely_b200.model import Analyzer  # noqa: E402

H, W, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (720, 1280, 225)


def vp(t):
    return C.c_void_p(t.data_ptr())


def run(tag):
    an = Analyzer(device=0)
    per = C.c_longlong()
    off = (C.c_longlong * L.MAX_SCALES)()
    pitch = (C.c_int * L.MAX_SCALES)()
    n = an.lib.trl_pyramid_pairs_size(an.ctx, H, W, C.byref(per), off, pitch)
    frames = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (B, H, W, 3), dtype=np.uint8)).cuda()
    hi = torch.zeros((B * per.value, 8), dtype=torch.float16, device="cuda")
    lo = torch.zeros_like(hi)
    for _ in range(3):
        assert an.lib.trl_pyramid_pairs(an.ctx, vp(frames), B, H, W, vp(hi), vp(lo), None) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        an.lib.trl_pyramid_pairs(an.ctx, vp(frames), B, H, W, vp(hi), vp(lo), None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag}: {H}x{W} B={B}: {ms:.3f} ms per launch, {n} levels, hi checksum {int(hi.view(torch.int16).to(torch.int64).sum())}, "
          f"lo checksum {int(lo.view(torch.int16).to(torch.int64).sum())}")
    return hi, lo


if __name__ == "__main__":
    os.environ["TRL_PYR_GENERIC"] = "0"
    h1, l1 = run("fast   ")
    os.environ["TRL_PYR_GENERIC"] = "1"
    h0, l0 = run("generic")
    print("hi equal:", bool(torch.equal(h1, h0)), " lo equal:", bool(torch.equal(l1, l0)))

"""Timeline of one host-buffer (e2e) step: when each chunk's H2D copy and cascade start / end on the device.
Usage (GPU box): python experiments/e2e_timeline.py [chunk]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import truely_b200  # noqa: E402,F401
from truely_b200 import model as M  # noqa: E402
import bench  # noqa: E402

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 90
an = M.Analyzer(device=0)
clip, stride, pinned, n_total, n_local = bench.make_frames("720p30_single", 0, 1, torch)
H, W = clip.height, clip.width
stage = torch.empty((3, chunk, H, W, 3), dtype=torch.uint8, device="cuda:0")
for _ in range(2):
    an.analyze_resident(pinned, chunk=chunk, h2d=True, dev_frames=stage)
torch.cuda.synchronize()

# re-implementation of the loop of analyze_resident with events around every piece
t = torch
out = an._res_out
cs = an._copy_stream
chunks = M.chunk_schedule(n_local, chunk, ramp=True)
ev = lambda: t.cuda.Event(enable_timing=True)
t0 = ev()
c_s, c_e, k_s, k_e = [ev() for _ in chunks], [ev() for _ in chunks], [ev() for _ in chunks], [ev() for _ in chunks]
f_s, f_e = ev(), ev()
t0.record(an.stream)
cs.wait_stream(an.stream)
for k, (a, b) in enumerate(chunks):
    with t.cuda.stream(cs):
        if k >= 3:
            cs.wait_event(k_e[k - 3])
        c_s[k].record(cs)
        stage[k % 3, : b - a].copy_(pinned[a:b], non_blocking=True)
        c_e[k].record(cs)
    with t.cuda.stream(an.stream):
        an.stream.wait_event(c_e[k])
        k_s[k].record(an.stream)
        an._check(an.lib.trl_detect_align(an.ctx, M._vp(stage[k % 3, : b - a]), b - a, H, W, M._vp(out["box"][a:b]),
                                          M._vp(out["valid"][a:b]), M._vp(out["nfaces"][a:b]), M._vp(out["crops"][a:b]), an._sptr()))
        k_e[k].record(an.stream)
with t.cuda.stream(an.stream):
    f_s.record(an.stream)
    an._check(an.lib.trl_facenet(an.ctx, M._vp(out["crops"]), n_local, an.crop_size, M._vp(out["emb"]), an._sptr()))
    f_e.record(an.stream)
torch.cuda.synchronize()
print("chunk  frames   copy start..end (ms)     cascade start..end (ms)")
for k, (a, b) in enumerate(chunks):
    print(f"{k:3d}   {b - a:4d}    {t0.elapsed_time(c_s[k]):7.2f} .. {t0.elapsed_time(c_e[k]):7.2f}      "
          f"{t0.elapsed_time(k_s[k]):7.2f} .. {t0.elapsed_time(k_e[k]):7.2f}")
print(f"facenet {t0.elapsed_time(f_s):7.2f} .. {t0.elapsed_time(f_e):7.2f}")

"""Per-layer device time of the FaceNet forward pass against each layer's tensor-pipe and HBM floors.
Usage (GPU box): python experiments/facenet_layers.py [batch] [crop]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import truely_b200  # noqa: E402,F401
from truely_b200.model import Analyzer, _vp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 160
an = Analyzer(device=0)
crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda")
emb = torch.empty((B, 512), dtype=torch.float32, device="cuda")
MAXS = 256
info = (C.c_int * (MAXS * 8))()
ms = (C.c_float * MAXS)()
fn = an.lib.trl_debug_facenet_step_times
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p]
with torch.cuda.stream(an.stream):
    ns = fn(an.ctx, _vp(crops), B, S, _vp(emb), MAXS, info, ms, an._sptr())
assert ns > 0, ns
PEAK = 1361.6e12 / 2          # MAC/s, measured sustained bf16
HBM = 6550.7e9
tot = tot_ideal = 0.0
agg = {}
for i in range(ns):
    kind, cin, cout, kh, kw, h, bn, bk = info[i * 8:i * 8 + 8]
    t = ms[i] * 1e-3
    if kind == 1:
        key = ("pool", cin, cout, kh, kw, h, bn)
        macs, byts = 0, 0
    else:
        key = ("conv", cin, cout, kh, kw, h, bn)
        macs = B * h * h * cout * cin * kh * kw
        byts = 2 * B * h * h * (cout + cin)          # write out + read in once (bf16), weights negligible
    a = agg.setdefault(key, [0, 0.0, 0, 0])
    a[0] += 1; a[1] += t; a[2] += macs; a[3] += byts
    tot += t
print(f"batch {B} crop {S}: {ns} steps, {tot * 1e3:.2f} ms")
print(f"{'layer':34s} {'n':>3s} {'ms':>8s} {'share':>6s} {'tensor':>7s} {'hbm':>6s}")
for key, (cnt, t, macs, byts) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    kindn, cin, cout, kh, kw, h, bn = key
    print(f"{kindn} {cin:4d}->{cout:4d} {kh}x{kw} @{h:3d} bn{bn:3d}   {cnt:3d} {t * 1e3:8.3f} {100 * t / tot:5.1f}% "
          f"{100 * macs / PEAK / t if t else 0:6.1f}% {100 * byts / HBM / t if t else 0:5.1f}%")

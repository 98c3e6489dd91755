import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
import truely_b200
from truely_b200 import model as M
an = M.Analyzer(device=0)
path = "tests/golden/bundled_veo3_360p.mp4"
M.run_trace(path, None, analyzer=an)
t0 = time.perf_counter(); tr = M.run_trace(path, None, analyzer=an); dt = time.perf_counter() - t0
print("total", dt, tr.timings)
pr = cProfile.Profile(); pr.enable(); tr = M.run_trace(path, None, analyzer=an); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:5000])
# decode alone
import cv2
t0 = time.perf_counter(); cap = cv2.VideoCapture(path); n = 0
while True:
    ok, f = cap.read()
    if not ok: break
    n += 1
print("decode only", n, time.perf_counter() - t0)

# with the annotated output video (what server.py calls): decode + GPU + annotate + encode
import tempfile
outp = os.path.join(tempfile.mkdtemp(), "out.mp4")
M.run_trace(path, outp, analyzer=an)
t0 = time.perf_counter(); tr = M.run_trace(path, outp, analyzer=an); dt = time.perf_counter() - t0
print("with writer: total", dt, tr.timings, "output bytes", os.path.getsize(outp))
pr = cProfile.Profile(); pr.enable(); tr = M.run_trace(path, outp, analyzer=an); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(8); print(s.getvalue()[:2500])

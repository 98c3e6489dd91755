"""Concurrent requests through the persistent service (service.py) against sequential run() calls on the bundled clip."""
import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
import truely_b200  # noqa
from truely_b200 import model as M, service as S
path = "tests/golden/bundled_veo3_360p.mp4"
an = M.Analyzer(device=0)
with contextlib.redirect_stdout(io.StringIO()):
    M.run_trace(path, None, analyzer=an)
    t0 = time.perf_counter()
    for _ in range(4):
        tr = M.run_trace(path, None, analyzer=an)
    t_seq = time.perf_counter() - t0
    svc = S.AnalysisService(an)
    S.run_trace_with_service(svc, path, None)
    t0 = time.perf_counter()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(lambda _: S.run_trace_with_service(svc, path, None).score, range(4)))
    t_svc = time.perf_counter() - t0
    t0 = time.perf_counter()
    r1 = S.run_trace_with_service(svc, path, None)
    t_one = time.perf_counter() - t0
    stats = dict(svc.stats)
    svc.close()
print(f"4 x run() sequential: {t_seq:.3f} s; 4 concurrent requests through the service: {t_svc:.3f} s; one request through the service: {t_one:.3f} s")
print("scores", tr.score, res, r1.score, "service stats", stats)

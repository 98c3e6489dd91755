"""CPU: the batching service's host logic (truely_b200/service.py) with the GPU pass replaced by a stub -- chunks of
concurrent requests are merged, every request sees its own results in its own order, a request's halo chains through its
chunks only, oversize / odd-shaped jobs wait without reordering a request."""
import threading

import numpy as np

import truely_b200  # noqa: F401
from truely_b200 import model as M
from truely_b200 import service as SV


class _StubService(SV.AnalysisService):
    """_run_batch without CUDA: a frame's 'embedding' is its first pixel value; sim = previous value of the SAME request."""

    def __init__(self, **kw):
        self.batches = []
        self.an = None
        self.max_batch_frames = kw.get("max_batch_frames")
        self.linger_s = kw.get("linger_s", 0.0)
        from collections import deque
        self._dq, self._cond, self._next_id, self._lock = deque(), threading.Condition(), 0, threading.Lock()
        self.stats = dict(batches=0, chunks=0, frames=0, max_chunks_in_batch=0)
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def _budget(self, h, w):
        return self.max_batch_frames or 64

    def _run_batch(self, batch):
        self.batches.append([(j.req.id, j.n) for j in batch])
        out = []
        for j in batch:
            vals = np.asarray(j.frames)[:, 0, 0, 0].astype(np.float32)
            sim = np.full(j.n, np.nan, np.float32)
            has = np.zeros(j.n, np.uint8)
            prev = j.req.halo
            for i in range(j.n):
                if prev is not None:
                    sim[i], has[i] = prev, 1
                prev = float(vals[i])
            j.req.halo = prev
            out.append(SV.ChunkResult(nfaces=np.ones(j.n, np.int32), box=np.zeros((j.n, 4), np.int32), valid=np.ones(j.n, np.uint8),
                                      emb=np.repeat(vals[:, None], 512, 1), sim=sim, below=(sim < 0.99).astype(np.uint8), has_sim=has))
        return out


def _clip(tag, n, h=4, w=6):
    return [np.full((h, w, 3), (tag * 50 + i) % 251, np.uint8) for i in range(n)]


def test_concurrent_requests_are_batched_and_keep_their_own_chains():
    svc = _StubService(linger_s=0.02)
    traces = {}

    def worker(tag):
        traces[tag] = SV.analyze_stream_service(svc, iter(_clip(tag, 40)), 7, 6, 4, chunk=5, keep_emb=True)   # fps 7 -> stride 1

    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=60)
    svc.close()
    assert svc.stats["max_chunks_in_batch"] >= 2, "concurrent chunks were never merged"
    assert svc.stats["frames"] == 4 * 40
    for tag, tr in traces.items():
        vals = [float((tag * 50 + i) % 251) for i in range(40)]
        assert [e[0] for e in tr.emb] == vals                                 # own frames, own order
        assert tr.sim[0] is None and tr.sim[1:] == vals[:-1]                  # halo chains through this request's chunks only
        assert tr.frame_count == 40 and tr.frame_index == list(range(40))


def test_budget_and_mixed_shapes_never_reorder_a_request():
    svc = _StubService(max_batch_frames=8)
    r1, r2, r3 = svc.open_request(4, 6), svc.open_request(4, 6), svc.open_request(8, 8)
    with svc._cond:                                                           # enqueue atomically: one dispatcher pass sees all
        pass
    f = []
    big = np.zeros((6, 4, 6, 3), np.uint8)
    small = np.zeros((2, 4, 6, 3), np.uint8)
    other = np.zeros((3, 8, 8, 3), np.uint8)
    svc.linger_s = 0.05
    f.append(r1.submit(big)); f.append(r2.submit(big)); f.append(r2.submit(small)); f.append(r3.submit(other)); f.append(r1.submit(small))
    for x in f:
        x.result(timeout=30)
    svc.close()
    order = [c for b in svc.batches for c in b]
    # per request, chunks ran in submission order
    for rid in (r1.id, r2.id, r3.id):
        mine = [n for (i, n) in order if i == rid]
        assert mine == {r1.id: [6, 2], r2.id: [6, 2], r3.id: [3]}[rid]
    for b in svc.batches:
        assert sum(n for _, n in b) <= 8 or len(b) == 1                       # budget respected
    assert all(len({i for i, _ in b} & {r3.id}) == 0 or len(b) == 1 for b in svc.batches)    # the 8x8 job ran alone
    assert M.frame_stride(7) == 1

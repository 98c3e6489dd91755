"""GPU: the multi-request batching service -- videos analysed concurrently through one shared context,
their chunks merged into common cascade / FaceNet batches, give bit-identical traces to analysing each one alone."""
import os
import threading

import cv2
import numpy as np
import pytest

from truely_b200 import model as M
from truely_b200 import service as SV
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu


def _same(a: M.Trace, b: M.Trace):
    assert a.frame_count == b.frame_count and a.frame_index == b.frame_index
    assert a.valid == b.valid and a.nfaces == b.nfaces
    assert all(np.array_equal(x, y) for x, y in zip(a.box, b.box))
    assert all(np.array_equal(x, y) for x, y in zip(a.emb, b.emb))
    assert a.sim == b.sim and a.flagged == b.flagged and a.score == b.score


def test_concurrent_streams_equal_solo_runs(analyzer):
    clips = [SyntheticClip(240, 320, 30, 160, n_faces=(1, 1), face_h=(70.0, 110.0), jitter=1.6, seed=31),
             SyntheticClip(240, 320, 30, 120, n_faces=(1, 1), face_h=(70.0, 110.0), jitter=0.3, seed=32),
             SyntheticClip(240, 320, 30, 100, n_faces=(0, 0), seed=33),                       # a video without any face
             SyntheticClip(360, 640, 30, 80, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=1.0, seed=34)]   # another frame size
    frames = [[f for f in c] for c in clips]
    solo = [M.analyze_stream(iter(fr), 30, c.width, c.height, analyzer=analyzer, chunk=6, keep_emb=True) for fr, c in zip(frames, clips)]
    svc = SV.AnalysisService(analyzer, linger_s=0.01)
    got = [None] * len(clips)

    def worker(k):
        got[k] = SV.analyze_stream_service(svc, iter(frames[k]), 30, clips[k].width, clips[k].height, chunk=6, keep_emb=True)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(len(clips))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    svc.close()
    assert svc.stats["max_chunks_in_batch"] >= 2, "chunks of concurrent requests were never merged into one GPU pass"
    for a, b in zip(got, solo):
        _same(a, b)
    assert sum(solo[0].flagged) > 0 and not any(solo[2].valid)


def test_run_many_on_files_equals_run(analyzer, tmp_path):
    paths = []
    for k, seed in enumerate((41, 42)):
        clip = SyntheticClip(240, 320, 30, 90, n_faces=(1, 1), face_h=(70.0, 110.0), jitter=1.5, seed=seed)
        src = str(tmp_path / f"in{k}.mp4")
        wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 30, (320, 240))
        for f in clip:
            wr.write(f)
        wr.release()
        paths.append(src)
    M._ANALYZER = analyzer
    ref = [M.run(p, str(tmp_path / f"ref{k}.mp4")) for k, p in enumerate(paths)]
    svc = SV.AnalysisService(analyzer)
    got = svc.run_many([(p, str(tmp_path / f"out{k}.mp4")) for k, p in enumerate(paths)])
    svc.close()
    assert got == ref
    for k in range(2):
        out = str(tmp_path / f"out{k}.mp4")
        assert os.path.exists(out) and os.path.getsize(out) > 0
        cap = cv2.VideoCapture(out)
        assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 90
        cap.release()

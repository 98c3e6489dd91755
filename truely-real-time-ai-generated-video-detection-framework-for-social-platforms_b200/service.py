"""Persistent service context with multi-request batching.

The reference rebuilds both models inside every ``run`` call (server/model.py:18-19) and its FastAPI handlers call ``run``
synchronously from ``async def`` bodies (server/server.py:585, 611, 813, 856), so concurrent ``/analyze-*`` requests
serialise on the event loop.  This module keeps ONE GPU context per process and lets any number of caller threads use it
at the same time: every caller decodes its own video and hands chunks of processed frames to the service; a dispatcher
thread takes whatever chunks are waiting -- from whichever requests, as long as they share a frame size -- and runs them
as ONE cascade call and ONE FaceNet batch (the per-frame work, server/model.py:47-59, carries no state between frames).
What does carry state, the previous face embedding and the run-length counter (server/model.py:37-39, 60-70), stays per
request: the consistency kernel runs on each request's slice with that request's own halo, in submission order, and the
counter lives in the caller's thread.  Results are bit-identical to analysing each video alone
(tests/test_gpu_service.py); the only thing that changes is how many frames share a kernel launch.

``run_with_service(service, video_path_one, video_path_two)`` is ``run`` on top of the service;
``AnalysisService.run_many([...])`` analyses several videos concurrently from a thread pool.
"""
from __future__ import annotations

import os
import threading
import time
from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass

import cv2
import numpy as np

from . import _lib as L
from . import model as M

_vp = M._vp


@dataclass
class ChunkResult:
    """Per processed frame outputs of one submitted chunk (numpy, host)."""
    nfaces: np.ndarray
    box: np.ndarray
    valid: np.ndarray
    emb: np.ndarray
    sim: np.ndarray
    below: np.ndarray
    has_sim: np.ndarray


class _Job:
    __slots__ = ("req", "frames", "n", "future")


class Request:
    """One video being analysed through the service (one per run() call)."""

    def __init__(self, service, rid, height, width):
        self.service, self.id, self.height, self.width = service, rid, height, width
        self.halo = None                # device tensors (emb [512], valid [1]) of this request's last face-bearing frame
        self.closed = False

    def submit(self, frames) -> Future:
        """frames: uint8 [n,H,W,3] (numpy or a pinned torch tensor).  Chunks of one request are analysed in submission
        order.  Returns a Future of ChunkResult."""
        return self.service._submit(self, frames)

    def close(self):
        self.closed = True
        self.halo = None


class AnalysisService:
    """One Analyzer shared by concurrent callers; chunks waiting at the same time are batched into one GPU pass."""

    def __init__(self, analyzer: M.Analyzer | None = None, max_batch_frames: int | None = None, linger_s: float = 0.0):
        self.an = analyzer or M.get_analyzer()
        self.max_batch_frames = max_batch_frames
        self.linger_s = linger_s        # how long the dispatcher waits for more chunks once it has one (0: take what is there)
        self._dq = deque()              # waiting jobs in submission order (None = shut down), guarded by _cond
        self._cond = threading.Condition()
        self._next_id = 0
        self._lock = threading.Lock()
        self.stats = dict(batches=0, chunks=0, frames=0, max_chunks_in_batch=0)
        self._thread = threading.Thread(target=self._loop, name="truely-b200-service", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ caller side
    def open_request(self, height: int, width: int) -> Request:
        with self._lock:
            self._next_id += 1
            return Request(self, self._next_id, height, width)

    def _submit(self, req: Request, frames) -> Future:
        if req.closed:
            raise RuntimeError("request is closed")
        job = _Job()
        job.req, job.frames, job.n, job.future = req, frames, int(frames.shape[0]), Future()
        if job.n == 0:
            job.future.set_result(None)
            return job.future
        if tuple(frames.shape[1:]) != (req.height, req.width, 3):
            raise ValueError(f"chunk shape {tuple(frames.shape)} does not match the request's {req.height}x{req.width} frames")
        with self._cond:
            self._dq.append(job)
            self._cond.notify()
        return job.future

    def close(self):
        with self._cond:
            self._dq.append(None)
            self._cond.notify()
        self._thread.join(timeout=30)

    # ------------------------------------------------------------------ dispatcher
    def _budget(self, height, width):
        if self.max_batch_frames:
            return self.max_batch_frames
        return max(8, int(256e6 // max(1, height * width * 3)))          # ~256 MB of frames per GPU pass

    def _take_batch(self):
        """Blocks until a job waits, then takes, in submission order, every waiting job that has the first one's frame size
        and fits the frame budget.  Jobs that do not fit stay where they are (a request's chunks never overtake each
        other: once one of its chunks is passed over, its later ones are too)."""
        with self._cond:
            while not self._dq:
                self._cond.wait()
            if self._dq[0] is None:
                return None
        if self.linger_s > 0:
            time.sleep(self.linger_s)
        with self._cond:
            first = self._dq[0]
            shape = (first.req.height, first.req.width)
            budget = self._budget(*shape)
            batch, keep, frames, passed = [], deque(), 0, set()
            for j in self._dq:
                if j is None:
                    keep.append(j)
                    continue
                fits = (j.req.height, j.req.width) == shape and (not batch or frames + j.n <= budget) and j.req.id not in passed
                if fits:
                    batch.append(j)
                    frames += j.n
                else:
                    passed.add(j.req.id)
                    keep.append(j)
            self._dq = keep
        return batch

    def _loop(self):
        while True:
            batch = self._take_batch()
            if batch is None:
                return
            try:
                results = self._run_batch(batch)
                for job, res in zip(batch, results):
                    job.future.set_result(res)
            except BaseException as e:          # noqa: BLE001 -- every waiting caller must see the failure
                for job in batch:
                    if not job.future.done():
                        job.future.set_exception(e)
            self.stats["batches"] += 1
            self.stats["chunks"] += len(batch)
            self.stats["frames"] += sum(j.n for j in batch)
            self.stats["max_chunks_in_batch"] = max(self.stats["max_chunks_in_batch"], len(batch))

    # ------------------------------------------------------------------ one merged GPU pass (CUDA; overridable for CPU tests)
    def _run_batch(self, batch):
        an, t = self.an, self.an.torch
        H, W = batch[0].req.height, batch[0].req.width
        n_tot = sum(j.n for j in batch)
        S = an.crop_size
        dev = f"cuda:{an.device}"
        buf = getattr(self, "_buf", None)
        if buf is None or buf["frames"].shape[0] < n_tot or tuple(buf["frames"].shape[1:3]) != (H, W):
            cap = max(n_tot, self._budget(H, W))
            buf = an.alloc_outputs(cap)
            buf["frames"] = t.empty((cap, H, W, 3), dtype=t.uint8, device=dev)
            buf["crops"] = t.empty((cap, S, S, 3), dtype=t.uint8, device=dev)
            self._buf = buf
        with t.cuda.stream(an.stream):
            a = 0
            for j in batch:
                src = j.frames if t.is_tensor(j.frames) else t.from_numpy(np.ascontiguousarray(j.frames, dtype=np.uint8))
                buf["frames"][a:a + j.n].copy_(src, non_blocking=True)
                a += j.n
            # the stateless per-frame work of every waiting request as one batch
            an._check(an.lib.trl_detect_align(an.ctx, _vp(buf["frames"]), n_tot, H, W, _vp(buf["box"]), _vp(buf["valid"]),
                                              _vp(buf["nfaces"]), _vp(buf["crops"]), an._sptr()))
            an._check(an.lib.trl_facenet_valid(an.ctx, _vp(buf["crops"]), _vp(buf["valid"]), n_tot, S, _vp(buf["emb"]), an._sptr()))
            # the stateful part per request, on its own slice with its own halo
            a = 0
            outs = []
            for j in batch:
                sl = slice(a, a + j.n)
                he, hv = j.req.halo if j.req.halo is not None else (None, None)
                last_emb = t.empty(L.EMB_DIM, dtype=t.float32, device=dev)
                last_valid = t.empty(1, dtype=t.uint8, device=dev)
                an._check(an.lib.trl_consistency(
                    an.ctx, _vp(buf["emb"][sl]), _vp(buf["valid"][sl]), j.n, _vp(he), _vp(hv), M.THRESHOLD_FACE_SIMILARITY,
                    _vp(buf["sim"][sl]), _vp(buf["below"][sl]), _vp(buf["has_sim"][sl]), _vp(last_emb), _vp(last_valid), an._sptr()))
                j.req.halo = (last_emb, last_valid)
                outs.append({k: buf[k][sl].to("cpu", non_blocking=True) for k in ("nfaces", "box", "valid", "emb", "sim", "below", "has_sim")})
                a += j.n
        an.stream.synchronize()
        an.check_capacity()
        return [ChunkResult(**{k: v.numpy() for k, v in o.items()}) for o in outs]

    # ------------------------------------------------------------------ convenience
    def run_many(self, jobs, max_workers: int | None = None):
        """jobs: [(video_path_one, video_path_two), ...] analysed concurrently -> list of int scores, in order."""
        with ThreadPoolExecutor(max_workers=max_workers or max(1, len(jobs))) as ex:
            futs = [ex.submit(run_with_service, self, a, b) for a, b in jobs]
            return [f.result() for f in futs]


def analyze_stream_service(service: AnalysisService, frame_iter, fps: int, width: int, height: int, writer=None,
                           chunk: int | None = None, keep_emb: bool = False) -> M.Trace:
    """model.analyze_stream on top of the service: same trace, same annotated output (server/model.py:42-77)."""
    stride = M.frame_stride(fps)
    if chunk is None:
        chunk = max(4, min(64, int(96e6 // max(1, width * height * 3))))
    tr = M.Trace(stride=stride, frame_index=[], valid=[], box=[], sim=[], flagged=[], nfaces=[], emb=[] if keep_emb else None,
                 timings=dict(decode_s=0.0, submit_s=0.0, finish_s=0.0, capacity_retries=0))
    rl = M.RunLength()
    req = service.open_request(height, width)
    pending = deque()
    frame_count = 0

    def finish(item):
        fut, frames, proc_pos, proc_idx = item
        res = fut.result() if fut is not None else None
        k = 0
        for pos, frame in enumerate(frames):
            if k < len(proc_pos) and proc_pos[k] == pos:
                fidx = proc_idx[k]
                valid = bool(res.valid[k])
                flagged = False
                if valid and res.has_sim[k]:
                    flagged = rl.step(bool(res.below[k]))
                    if writer is not None:
                        M.annotate_frame(frame, res.box[k], flagged, fidx)
                tr.frame_index.append(fidx)
                tr.valid.append(valid)
                tr.box.append(res.box[k].copy())
                tr.sim.append(float(res.sim[k]) if res.has_sim[k] else None)
                tr.flagged.append(flagged)
                tr.nfaces.append(int(res.nfaces[k]))
                if keep_emb:
                    tr.emb.append(res.emb[k].copy())
                k += 1
            if writer is not None:
                writer.write(frame)

    cur_frames, cur_proc, proc_pos, proc_idx = [], [], [], []
    for frame in frame_iter:
        if frame_count % stride == 0:
            proc_pos.append(len(cur_frames))
            proc_idx.append(frame_count)
            cur_proc.append(frame)
        cur_frames.append(frame if writer is not None else None)
        frame_count += 1
        if len(cur_proc) == chunk:
            pending.append((req.submit(np.stack(cur_proc)), cur_frames, proc_pos, proc_idx))
            cur_frames, cur_proc, proc_pos, proc_idx = [], [], [], []
            while len(pending) > 1:
                finish(pending.popleft())
    if cur_frames:
        fut = req.submit(np.stack(cur_proc)) if cur_proc else None
        pending.append((fut, cur_frames, proc_pos, proc_idx))
    while pending:
        finish(pending.popleft())
    req.close()
    tr.frame_count = frame_count
    tr.flagged_count = rl.deep_fake_frame_count
    tr.final_run = rl.deepfake_count
    tr.score = M.final_score(rl.deep_fake_frame_count, rl.deepfake_count, frame_count, fps, stride)
    return tr


def run_trace_with_service(service: AnalysisService, video_path_one: str, video_path_two: str | None,
                           keep_emb: bool = False) -> M.Trace:
    """model.run_trace with the GPU work going through the shared service (same guards, prints and side effects)."""
    start_time = time.time()
    if not os.path.exists(video_path_one) or os.path.getsize(video_path_one) == 0:
        print(f"Error: Input video file {video_path_one} doesn't exist or is empty")
        return M.Trace()
    cap = cv2.VideoCapture(video_path_one)
    if not cap.isOpened():
        print(f"Error: OpenCV couldn't open video file {video_path_one}")
        return M.Trace()
    fps = int(cap.get(cv2.CAP_PROP_FPS))
    width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    if width <= 0 or height <= 0 or fps <= 0:
        print(f"Error: Invalid video properties: width={width}, height={height}, fps={fps}")
        cap.release()
        return M.Trace()
    out = None
    try:
        out = M._open_writer(video_path_two, fps, width, height) if video_path_two is not None else None
        tr = analyze_stream_service(service, M._video_frames(cap), fps, width, height, writer=out, keep_emb=keep_emb)
        print(f"Total Execution Time: {time.time() - start_time} seconds")
    finally:
        cap.release()
        if out is not None:
            out.release()
    if tr.frame_count == 0:
        print("Error: No frames were processed")
        tr.score = 0
    return tr


def run_with_service(service: AnalysisService, video_path_one: str, video_path_two: str) -> int:
    """Drop-in for reference server/model.py::run that is safe to call from many threads at once."""
    return run_trace_with_service(service, video_path_one, video_path_two).score

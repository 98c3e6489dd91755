"""Host side of the drop-in: ``run(video_path_one, video_path_two) -> int`` (reference server/model.py:11-95).

Same signature, guards, printed messages, side effects (annotated copy of every frame) and integer
score as the reference.  What differs is where the per-frame work happens: processed frames
(server/model.py:46) are batched, copied to the GPU from pinned memory and pushed through
``libtruely_b200.so`` (MTCNN cascade -> crop-align -> FaceNet -> consecutive-embedding cosine), while
this module keeps what the reference also does on the host: OpenCV decode / annotate / encode and the
run-length state machine + score (server/model.py:62-70, 83-95), which is exact integer logic.

PyTorch is used for device buffers, pinned staging buffers and the CUDA stream only.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from collections import deque
from dataclasses import dataclass

import cv2
import numpy as np

from . import _lib as L
from . import weights as W

THRESHOLD_FACE_SIMILARITY = 0.99      # server/model.py:16
THRESHOLD_FRAMES_FOR_DEEPFAKE = 15    # server/model.py:17
CROP_SIZE = 80                        # server/model.py:41


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@dataclass
class BatchResult:
    """Per processed frame outputs of one batch (numpy, host)."""
    nfaces: np.ndarray      # int32 [B]
    box: np.ndarray         # int32 [B,4]  truncated + clamped box of the largest face (server/model.py:49-53)
    valid: np.ndarray       # uint8 [B]    1 = a face was embedded
    emb: np.ndarray         # float32 [B,512]
    sim: np.ndarray         # float32 [B]  NaN where there was nothing to compare with
    below: np.ndarray       # uint8 [B]    sim < threshold
    has_sim: np.ndarray     # uint8 [B]
    box_f: np.ndarray | None = None    # float32 [B,4] untruncated (detail mode only)
    boxes: np.ndarray | None = None    # float32 [B,cap,5]            (detail mode only)
    counts: np.ndarray | None = None   # int32 [B,4] candidates per stage (detail mode only)


class Analyzer:
    """Process-global GPU context: weights resident on the device, reusable across ``run`` calls
    (the reference rebuilds both models on every call, server/model.py:18-19; behaviourally invisible)."""

    def __init__(self, device: int = 0, facenet_impl: int | None = None, crop_size: int | None = None,
                 cand_cap_scale: int | None = None, cand_cap_frame: int | None = None, box_cap_frame: int | None = None,
                 pnet_precision: int | None = None, mode: str = "reference", margin: int = 0):
        """``mode``: "reference" = the crop path of server/model.py:49-58 (80x80, INTER_LINEAR, /255: what ``run`` uses);
        "b" = upstream facenet_pytorch's own crop as the north star words it (extract_face: INTER_AREA to 160x160 with
        ``margin``, fixed_image_standardization), SURVEY.md section 0 / Appendix A "Mode-B extras"."""
        if mode not in ("reference", "b"):
            raise ValueError("mode must be 'reference' or 'b'")
        if crop_size is None:
            crop_size = CROP_SIZE if mode == "reference" else 160
        import torch
        self.torch = torch
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("truely_b200: no CUDA device visible; the hot path has no CPU fallback")
        self.device = device
        torch.cuda.set_device(device)
        cfg = L.Config()
        self.lib.trl_default_config(C.byref(cfg))
        cfg.crop_size = crop_size
        if facenet_impl is None:
            facenet_impl = int(os.environ.get("TRUELY_FACENET_IMPL", "0"))
        cfg.facenet_impl = facenet_impl
        if pnet_precision is None:
            pnet_precision = int(os.environ.get("TRUELY_PNET_PRECISION", "3"))
        cfg.pnet_precision = pnet_precision
        cfg.mode = 0 if mode == "reference" else 1
        cfg.margin = margin
        self.mode, self.margin = mode, margin
        if cand_cap_scale:
            cfg.cand_cap_scale = cand_cap_scale
        if cand_cap_frame:
            cfg.cand_cap_frame = cand_cap_frame
        if box_cap_frame:
            cfg.box_cap_frame = box_cap_frame
        self.cfg = cfg
        mt, self.mtcnn_source = W.load_mtcnn_state()
        fn, self.facenet_source = W.load_facenet_state()
        blobs = [W.pack_mtcnn(mt, "pnet"), W.pack_mtcnn(mt, "rnet"), W.pack_mtcnn(mt, "onet"), W.pack_facenet(fn)]
        w = L.Weights()
        fp = C.POINTER(C.c_float)
        w.h_pnet, w.pnet_len = blobs[0].ctypes.data_as(fp), blobs[0].size
        w.h_rnet, w.rnet_len = blobs[1].ctypes.data_as(fp), blobs[1].size
        w.h_onet, w.onet_len = blobs[2].ctypes.data_as(fp), blobs[2].size
        w.h_facenet, w.facenet_len = blobs[3].ctypes.data_as(fp), blobs[3].size
        ctx = C.c_void_p()
        rc = self.lib.trl_create(device, C.byref(w), C.byref(cfg), C.byref(ctx))
        if rc != L.TRL_OK:
            raise L.TrlError(rc, self.lib.trl_last_error(None).decode())
        self.ctx = ctx
        self.pnet_precision = int(self.lib.trl_pnet_precision(ctx))       # the mode in effect (the library may fall back to 0)
        self.stream = torch.cuda.Stream(device=device)
        self.box_cap = cfg.box_cap_frame
        self.crop_size = crop_size

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.trl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != L.TRL_OK:
            raise L.TrlError(rc, self.lib.trl_last_error(self.ctx).decode())

    def check_capacity(self):
        """Call after a stream synchronise: raises TRL_E_CAPACITY if a candidate list overflowed."""
        detail = (C.c_int * 4)()
        self._check(self.lib.trl_check_capacity(self.ctx, detail))

    # capacities of the overflow-retry path: the global-memory NMS limit of the library (include/truely_b200.h)
    BIG_CAPS = (16384, 16384, 2048)

    def set_capacity(self, cand_cap_scale: int, cand_cap_frame: int, box_cap_frame: int):
        """trl_set_capacity: slow (synchronises, drops the workspace)."""
        self._check(self.lib.trl_set_capacity(self.ctx, cand_cap_scale, cand_cap_frame, box_cap_frame))
        self.cfg.cand_cap_scale, self.cfg.cand_cap_frame, self.cfg.box_cap_frame = cand_cap_scale, cand_cap_frame, box_cap_frame
        self.box_cap = box_cap_frame

    def detect_align_uncapped(self, d_frames, box, valid, nfaces, crops):
        """Overflow retry (upstream detect_face has no candidate cap; the fast path has): the cascade + crop-align of
        ``d_frames`` again, one frame per call, with the capacities raised to the library's maximum -- groups above 2048
        candidates then run through the global-memory NMS.  Synchronous and slow by design; the fast-path capacities
        are restored afterwards.  Raises TRL_E_CAPACITY only if even 16384 candidates per group are not enough."""
        saved = (self.cfg.cand_cap_scale, self.cfg.cand_cap_frame, self.cfg.box_cap_frame)
        B, H, Wd, _ = d_frames.shape
        self.stream.synchronize()
        try:
            self.check_capacity()          # drop a flag left by work that was still in flight (the caller redoes all of it)
        except L.TrlError as e:
            if e.code != L.TRL_E_CAPACITY:
                raise
        self.set_capacity(*self.BIG_CAPS)
        try:
            with self.torch.cuda.stream(self.stream):
                for i in range(B):
                    self._check(self.lib.trl_detect_align(
                        self.ctx, _vp(d_frames[i:i + 1]), 1, H, Wd, _vp(box[i:i + 1]), _vp(valid[i:i + 1]),
                        _vp(nfaces[i:i + 1]), _vp(crops[i:i + 1]), self._sptr()))
            self.stream.synchronize()
            self.check_capacity()
        finally:
            self.set_capacity(*saved)

    def overlay_device(self, d_frames, d_box, state, frame_index):
        """Annotate resident frames in place (``trl_overlay``; server/model.py:67-74 on the device, bit exact with
        OpenCV).  ``d_frames`` uint8 [n,H,W,3] and ``d_box`` int32 [n,4] live on the device; ``state`` (0 untouched /
        1 "Real Frame" / 2 "AI Detected") and ``frame_index`` are host sequences.  Enqueued on ``self.stream``; returns
        the device uint8 [n] tensor of frames whose caption would leave the frame and is therefore left to the host
        (``overlay.draw_text_host``)."""
        from . import overlay as O
        t = self.torch
        if not getattr(self, "_overlay_ready", False):
            O.register(self.lib, self.ctx)
            self._overlay_ready = True
        n, H, W = int(d_frames.shape[0]), int(d_frames.shape[1]), int(d_frames.shape[2])
        dev = f"cuda:{self.device}"
        h_state = t.as_tensor(np.asarray(state, dtype=np.uint8)).pin_memory()
        h_index = t.as_tensor(np.asarray(frame_index, dtype=np.int32)).pin_memory()
        with t.cuda.stream(self.stream):
            d_state = h_state.to(dev, non_blocking=True)
            d_index = h_index.to(dev, non_blocking=True)
            pending = t.zeros(n, dtype=t.uint8, device=dev)
            self._check(self.lib.trl_overlay(self.ctx, _vp(d_frames), n, H, W, _vp(d_box), _vp(d_state), _vp(d_index),
                                             _vp(pending), self._sptr()))
        self._overlay_keep = (h_state, h_index, d_state, d_index)      # alive until the stream has consumed them
        return pending

    def launch_count(self) -> int:
        return int(self.lib.trl_launch_count(self.ctx))

    def _sptr(self):
        return C.c_void_p(self.stream.cuda_stream)

    def set_profiling(self, on: bool):
        self._check(self.lib.trl_set_profiling(self.ctx, 1 if on else 0))

    def read_stage_times(self):
        """-> ({stage name: summed ms}, n_calls) since the last read; synchronises the stream first."""
        self.stream.synchronize()
        ms = (C.c_float * L.NUM_STAGES)()
        calls = self.lib.trl_read_stage_times(self.ctx, ms)
        names = []
        for i in range(L.NUM_STAGES):
            b = C.create_string_buffer(32)
            self.lib.trl_stage_name(i, b, 32)
            names.append(b.value.decode())
        return {n: float(ms[i]) for i, n in enumerate(names)}, calls

    def alloc_outputs(self, B):
        t, dev = self.torch, f"cuda:{self.device}"
        return dict(
            nfaces=t.empty(B, dtype=t.int32, device=dev), box=t.empty((B, 4), dtype=t.int32, device=dev),
            valid=t.empty(B, dtype=t.uint8, device=dev), emb=t.empty((B, L.EMB_DIM), dtype=t.float32, device=dev),
            sim=t.empty(B, dtype=t.float32, device=dev), below=t.empty(B, dtype=t.uint8, device=dev),
            has_sim=t.empty(B, dtype=t.uint8, device=dev),
            last_emb=t.zeros(L.EMB_DIM, dtype=t.float32, device=dev), last_valid=t.zeros(1, dtype=t.uint8, device=dev))

    # ------------------------------------------------------------------ device-resident API
    def process_device(self, d_frames, out, halo=None, thr=THRESHOLD_FACE_SIMILARITY):
        """One fused batch on device tensors (uint8 [B,H,W,3]); asynchronous on ``self.stream``.
        ``halo`` = (emb float32[512], valid uint8[1]) device tensors of the preceding range, or None."""
        B, H, Wd, _ = d_frames.shape
        he, hv = (halo if halo is not None else (None, None))
        self._check(self.lib.trl_process(
            self.ctx, _vp(d_frames), B, H, Wd, _vp(he), _vp(hv), thr, _vp(out["box"]), _vp(out["valid"]),
            _vp(out["emb"]), _vp(out["sim"]), _vp(out["below"]), _vp(out["has_sim"]), _vp(out["nfaces"]),
            _vp(out["last_emb"]), _vp(out["last_valid"]), self._sptr()))

    # ------------------------------------------------------------------ host-buffer API
    def process_frames(self, frames: np.ndarray, halo=None, detail: bool = True,
                       thr: float = THRESHOLD_FACE_SIMILARITY) -> BatchResult:
        """frames: uint8 [B,H,W,3] BGR on the host.  Synchronous convenience wrapper (tests, smoke)."""
        t = self.torch
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        B, H, Wd, _ = frames.shape
        dev = f"cuda:{self.device}"
        with t.cuda.stream(self.stream):
            d_frames = t.from_numpy(frames).pin_memory().to(dev, non_blocking=True)
            out = self.alloc_outputs(B)
            he = hv = None
            if halo is not None:
                he = t.from_numpy(np.ascontiguousarray(halo, np.float32)).to(dev)
                hv = t.ones(1, dtype=t.uint8, device=dev)
            boxes = counts = None
            if detail:
                boxes = t.zeros((B, self.box_cap, 5), dtype=t.float32, device=dev)
                counts = t.zeros((B, 4), dtype=t.int32, device=dev)
                crops = t.empty((B, self.crop_size, self.crop_size, 3), dtype=t.uint8, device=dev)
                self._check(self.lib.trl_detect(self.ctx, _vp(d_frames), B, H, Wd, _vp(out["nfaces"]), _vp(boxes),
                                                _vp(counts), self._sptr()))
                self._check(self.lib.trl_crop_align(self.ctx, _vp(d_frames), B, H, Wd, _vp(boxes), self.box_cap * 5,
                                                    _vp(out["nfaces"]), _vp(out["box"]), _vp(out["valid"]), _vp(crops),
                                                    self._sptr()))
                self._check(self.lib.trl_facenet_valid(self.ctx, _vp(crops), _vp(out["valid"]), B, self.crop_size, _vp(out["emb"]),
                                                       self._sptr()))
                self._check(self.lib.trl_consistency(self.ctx, _vp(out["emb"]), _vp(out["valid"]), B, _vp(he), _vp(hv), thr,
                                                     _vp(out["sim"]), _vp(out["below"]), _vp(out["has_sim"]),
                                                     _vp(out["last_emb"]), _vp(out["last_valid"]), self._sptr()))
            else:
                self.process_device(d_frames, out, (he, hv) if halo is not None else None, thr)
        self.stream.synchronize()
        self.check_capacity()
        res = BatchResult(nfaces=out["nfaces"].cpu().numpy(), box=out["box"].cpu().numpy(), valid=out["valid"].cpu().numpy(),
                          emb=out["emb"].cpu().numpy(), sim=out["sim"].cpu().numpy(), below=out["below"].cpu().numpy(),
                          has_sim=out["has_sim"].cpu().numpy())
        if detail:
            res.boxes = boxes.cpu().numpy()
            res.box_f = res.boxes[:, 0, :4].copy()
            res.counts = counts.cpu().numpy()
        return res


    def embed_all_faces(self, frames: np.ndarray, max_faces: int | None = None):
        """keep_all (mode B, BASELINE.json configs[3]: 4-8 faces per frame): MTCNN.detect on every frame, then EVERY
        detected box is cropped (extract_face: INTER_AREA to crop_size, margin) and embedded with
        fixed_image_standardization, batched over all faces of the batch.  Synchronous convenience wrapper.
        -> list over frames of [(int box [4], float32 embedding [512] or None for an empty crop), ...], largest box first."""
        t = self.torch
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        B, H, Wd, _ = frames.shape
        dev = f"cuda:{self.device}"
        S = self.crop_size
        cap = self.box_cap
        if max_faces is None:
            max_faces = B * min(cap, 16)
        with t.cuda.stream(self.stream):
            d_frames = t.from_numpy(frames).pin_memory().to(dev, non_blocking=True)
            nfaces = t.empty(B, dtype=t.int32, device=dev)
            boxes = t.zeros((B, cap, 5), dtype=t.float32, device=dev)
            self._check(self.lib.trl_detect(self.ctx, _vp(d_frames), B, H, Wd, _vp(nfaces), _vp(boxes), None, self._sptr()))
            face_off = t.empty(B + 1, dtype=t.int32, device=dev)
            face_frame = t.full((max_faces,), -1, dtype=t.int32, device=dev)
            box_int = t.zeros((max_faces, 4), dtype=t.int32, device=dev)
            valid = t.zeros(max_faces, dtype=t.uint8, device=dev)
            crops = t.zeros((max_faces, S, S, 3), dtype=t.uint8, device=dev)
            emb = t.zeros((max_faces, L.EMB_DIM), dtype=t.float32, device=dev)
            self._check(self.lib.trl_extract_faces_all(
                self.ctx, _vp(d_frames), B, H, Wd, _vp(boxes), cap * 5, cap, _vp(nfaces), S, self.margin, max_faces,
                _vp(face_off), _vp(face_frame), _vp(box_int), _vp(valid), _vp(crops), self._sptr()))
            self._check(self.lib.trl_facenet_norm(self.ctx, _vp(crops), max_faces, S, 1, _vp(emb), self._sptr()))
        self.stream.synchronize()
        self.check_capacity()
        off = face_off.cpu().numpy()
        bi, va, em = box_int.cpu().numpy(), valid.cpu().numpy(), emb.cpu().numpy()
        out = []
        for b in range(B):
            out.append([(bi[i].copy(), em[i].copy() if va[i] else None) for i in range(off[b], off[b + 1])])
        return out

    # ------------------------------------------------------------------ clip-level API on staged frames
    def analyze_resident(self, frames, chunk: int = 90, host_out=None, halo=None, h2d: bool = False, dev_frames=None,
                         thr: float = THRESHOLD_FACE_SIMILARITY, pipeline: bool = True, clip_start=None):
        """All processed frames of a clip (or of this rank's range) in one go.

        ``frames``: uint8 [N,H,W,3] tensor, either on the device (``h2d=False``) or in pinned host memory
        (``h2d=True``: chunks are copied to the device inside this call on a copy stream, multi-buffered in
        ``dev_frames`` [nbuf,chunk,H,W,3], so the copies of the next chunks overlap the cascade on chunk k).
        The MTCNN cascade + crop-align run chunk by chunk (bounded workspace; with ``pipeline`` the latency-bound tail
        of chunk k runs under the pyramid of chunk k+1, trl_detect_align_async); all N crops are then embedded by ONE
        FaceNet call (large-M GEMMs) and compared by one consistency call.  Nothing synchronises with the host.
        ``clip_start`` (uint8 [N] device tensor, optional): 1 on the first processed frame of every clip when the range
        holds several clips back to back (BASELINE.json configs[4]); the embedding chain is cut there, like the locals of
        a fresh run() call (server/model.py:37-39).
        Returns the dict of device outputs [N, ...]; if ``host_out`` (pinned tensors keyed like the outputs) is given,
        those per-frame results are copied back asynchronously as well.
        """
        t = self.torch
        N, H, Wd, _ = frames.shape
        S = self.crop_size
        dev = f"cuda:{self.device}"
        out = getattr(self, "_res_out", None)
        if out is None or out["valid"].shape[0] < N:
            out = self.alloc_outputs(N)
            out["crops"] = t.empty((N, S, S, 3), dtype=t.uint8, device=dev)
            self._res_out = out
        if h2d and getattr(self, "_copy_stream", None) is None:
            self._copy_stream = t.cuda.Stream(device=self.device)
        he, hv = (halo if halo is not None else (None, None))
        chunks = chunk_schedule(N, chunk, ramp=h2d)
        nbuf = dev_frames.shape[0] if h2d else 0
        if h2d and pipeline and nbuf < 2:
            # a pipelined tail still reads staging buffer k while chunk k+1 is copied in: one buffer cannot be recycled safely
            raise ValueError("analyze_resident(h2d=True, pipeline=True) needs at least two staging buffers (dev_frames.shape[0] >= 2)")
        if h2d:
            cs = self._copy_stream
            copied = [t.cuda.Event() for _ in chunks]
            consumed = [t.cuda.Event() for _ in chunks]
            cs.wait_stream(self.stream)
        for k, (a, b) in enumerate(chunks):
            if h2d:
                with t.cuda.stream(cs):
                    if k >= nbuf:
                        cs.wait_event(consumed[k - nbuf])       # staging buffer k % nbuf is free again
                    dev_frames[k % nbuf, : b - a].copy_(frames[a:b], non_blocking=True)
                    copied[k].record(cs)
                d = dev_frames[k % nbuf, : b - a]
            else:
                d = frames[a:b]
            with t.cuda.stream(self.stream):
                if h2d:
                    self.stream.wait_event(copied[k])
                fn = self.lib.trl_detect_align_async if pipeline else self.lib.trl_detect_align
                self._check(fn(
                    self.ctx, _vp(d), b - a, H, Wd, _vp(out["box"][a:b]), _vp(out["valid"][a:b]), _vp(out["nfaces"][a:b]),
                    _vp(out["crops"][a:b]), self._sptr()))
                if h2d:
                    # pipelined: the tail of chunk k still reads its frames until call k+1 has ordered it before its own
                    # P-Net, so staging buffer k is released one call later (the last ones after the FaceNet call's join)
                    if not pipeline:
                        consumed[k].record(self.stream)
                    elif k >= 1:
                        consumed[k - 1].record(self.stream)
        with t.cuda.stream(self.stream):
            # one FaceNet batch for the whole range: per-chunk calls (tried, to hide FaceNet under the next copy) make the
            # step launch bound on the host (~110 launches per call) and were 8 ms slower end to end
            self._check(self.lib.trl_facenet_valid(self.ctx, _vp(out["crops"]), _vp(out["valid"]), N, S, _vp(out["emb"]),
                                                   self._sptr()))
            self._check(self.lib.trl_consistency_clips(
                self.ctx, _vp(out["emb"]), _vp(out["valid"]), N, _vp(clip_start), _vp(he), _vp(hv), thr, _vp(out["sim"]),
                _vp(out["below"]), _vp(out["has_sim"]), _vp(out["last_emb"]), _vp(out["last_valid"]), self._sptr()))
            if host_out is not None:
                for key, h in host_out.items():
                    h[:N].copy_(out[key][:N], non_blocking=True)
        return out


# tail rule of chunk_schedule (next = SCHED_A * size + SCHED_B frames); environment overrides are for tuning runs only
SCHED_A = float(os.environ.get("TRL_SCHED_A", 0.35))
SCHED_B = float(os.environ.get("TRL_SCHED_B", 6))
SCHED_STEPS = int(os.environ.get("TRL_SCHED_STEPS", 6))
SCHED_MIN_DIV = int(os.environ.get("TRL_SCHED_MIN_DIV", 10))


def chunk_schedule(n: int, chunk: int, ramp: bool = False):
    """[(start, end)] ranges of at most ``chunk`` frames.  With ``ramp`` (host frames: the H2D copy of chunk k+1 overlaps
    the cascade on chunk k) the chunks grow at the start and shrink towards the end: the copy of the first chunk and the
    cascade of the last one are the parts of the pipeline that nothing overlaps.  With the round-2 kernels a cascade costs
    about 0.4 ms + 0.016 ms/frame and a copy 0.05 ms/frame (720p, B200), so a chunk keeps up with the copy of its successor
    if the successor is at least ~0.32 of its size + 8 frames: the tail shrinks by next = 0.35 size + 6 (at most six steps,
    not below a tenth of a chunk; 90 -> 37, 18, 12, 10, 9 frames).  Measured on the bench clip (experiments/variants/sweep_sched.sh):
    24.50 ms per 450 frames against 24.66 ms with the round-1 rule (0.6 size + 8, not below a quarter), H2D alone 22.42 ms."""
    if not ramp or n <= chunk:
        return [(a, min(n, a + chunk)) for a in range(0, n, chunk)]
    head = [max(1, chunk // 4), max(1, chunk // 2)]
    tail, t = [], chunk
    while True:
        t = int(SCHED_A * t + SCHED_B)
        if t >= chunk or t < chunk // SCHED_MIN_DIV or (tail and t >= tail[-1]) or len(tail) == SCHED_STEPS:
            break
        tail.append(t)
    if sum(head) + sum(tail) + chunk > n:
        tail = tail[-1:] if tail and sum(head) + tail[-1] < n else []
    rest = n - sum(head) - sum(tail)
    full, rem = divmod(rest, chunk)
    if rem and rem + head[1] <= chunk:
        head[1] += rem                   # a small remainder rides with the second chunk
        rem = 0
    sizes = head + ([rem] if rem else []) + [chunk] * full + tail
    out, a = [], 0
    for sz in sizes:
        out.append((a, a + sz))
        a += sz
    return out


def score_from_flags(valid, has_sim, below, frame_count: int, fps: int, stride: int):
    """K13 on the host: run-length machine over the per-frame flags (server/model.py:62-70) + score (83-95).

    Vectorised (the flags of a whole clip -- or of all ranks' shards -- arrive at once): among the compared frames
    (valid and has_sim) the counter after frame i is the length of the run of `below` ending at i, a frame is flagged when
    that length exceeds 15.  Same integers as feeding RunLength.step frame by frame (tests/test_host.py)."""
    v = np.asarray(valid).astype(bool) & np.asarray(has_sim).astype(bool)
    idx = np.flatnonzero(v)
    flagged = np.zeros(len(v), bool)
    rl = RunLength()
    if len(idx):
        b = np.asarray(below)[idx].astype(np.int64)
        c = np.cumsum(b)
        run = c - np.maximum.accumulate(np.where(b == 0, c, 0))     # consecutive `below` frames ending here
        f = run > THRESHOLD_FRAMES_FOR_DEEPFAKE
        flagged[idx] = f
        rl.deepfake_count = int(run[-1])
        rl.deep_fake_frame_count = int(f.sum())
    return final_score(rl.deep_fake_frame_count, rl.deepfake_count, frame_count, fps, stride), flagged.tolist(), rl


def score_clips(valid, has_sim, below, clips, fps: int, stride: int):
    """Many clips back to back (BASELINE.json configs[4]): ``clips`` = [(n_processed, frame_count), ...] in batch order.
    Each clip gets the score a separate run() call would return: the run-length counters start from zero at every clip
    (server/model.py:37-39).  -> (scores [n_clips], flagged list for the whole batch)."""
    scores, flagged, a = [], [], 0
    for n_proc, frame_count in clips:
        sc, fl, _ = score_from_flags(valid[a:a + n_proc], has_sim[a:a + n_proc], below[a:a + n_proc], frame_count, fps, stride)
        scores.append(sc)
        flagged.extend(fl)
        a += n_proc
    if a != len(valid):
        raise ValueError(f"clips cover {a} processed frames, the flag arrays hold {len(valid)}")
    return scores, flagged


def clip_start_mask(clips):
    """uint8 numpy mask over the processed frames of the batch: 1 on the first processed frame of every clip."""
    n = sum(c[0] for c in clips)
    m = np.zeros(n, np.uint8)
    a = 0
    for n_proc, _ in clips:
        if n_proc > 0:
            m[a] = 1
        a += n_proc
    return m


_ANALYZER = None


def get_analyzer() -> Analyzer:
    global _ANALYZER
    if _ANALYZER is None:
        _ANALYZER = Analyzer(device=int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("TRUELY_USE_LOCAL_RANK") else 0)
    return _ANALYZER


# ---------------------------------------------------------------------- K13: run-length state machine + score (host)

class RunLength:
    """server/model.py:37-39, 62-70: deepfake_count / deep_fake_frame_count, exact integer logic."""

    def __init__(self):
        self.deepfake_count = 0
        self.deep_fake_frame_count = 0

    def step(self, below: bool) -> bool:
        """Feed one compared frame; returns True if the frame is flagged (counted)."""
        if below:
            self.deepfake_count += 1
        else:
            self.deepfake_count = 0
        if self.deepfake_count > THRESHOLD_FRAMES_FOR_DEEPFAKE:
            self.deep_fake_frame_count += 1
            return True
        return False


def final_score(deep_fake_frame_count: int, deepfake_count: int, frame_count: int, fps: int, stride: int) -> int:
    """server/model.py:83-95."""
    if frame_count == 0:
        return 0
    total_processed_frames = (frame_count + stride - 1) // stride      # == sum(1 for i in range(frame_count) if i % stride == 0)
    if total_processed_frames == 0:
        return 0
    deepfake_percentage = (deep_fake_frame_count / total_processed_frames) * 100
    confidence_factor = min(deepfake_percentage * (deepfake_count / THRESHOLD_FRAMES_FOR_DEEPFAKE), 100)
    if frame_count > fps * 30:
        weighted_score = min(deepfake_percentage + confidence_factor * 0.5, 100)
    else:
        weighted_score = min(deepfake_percentage + confidence_factor * 0.3, 100)
    return max(0, min(100, int(weighted_score)))


def frame_stride(fps: int) -> int:
    return max(1, int(fps / 7))   # server/model.py:40


# ---------------------------------------------------------------------- streaming driver

@dataclass
class Trace:
    score: int = 0
    frame_count: int = 0
    stride: int = 1
    flagged_count: int = 0
    final_run: int = 0
    frame_index: list = None
    valid: list = None
    box: list = None
    sim: list = None
    flagged: list = None
    nfaces: list = None
    emb: list = None
    timings: dict = None


def staging_empty(torch, shape, write_combined: bool = True):
    """uint8 host staging tensor for frames on their way to the GPU (trl_host_alloc): page-locked and, by default,
    write-combined -- the host only writes decoded frames into it (server/model.py:43 yields them one by one) and the
    copy engine reads it without snooping the CPU caches.  Do not read it back on the CPU (uncached reads are slow).
    TRL_STAGING=pinned selects plain page-locked memory."""
    import weakref
    if os.environ.get("TRL_STAGING", "").lower() == "pinned":
        write_combined = False
    lib = L.load()
    n_bytes = 1
    for s in shape:
        n_bytes *= int(s)
    if n_bytes == 0:
        return torch.empty(shape, dtype=torch.uint8, pin_memory=True)
    ptr = C.c_void_p()
    rc = lib.trl_host_alloc(n_bytes, 1 if write_combined else 0, C.byref(ptr))
    if rc != L.TRL_OK:
        raise L.TrlError(rc, f"trl_host_alloc({n_bytes} bytes) failed")
    buf = (C.c_uint8 * n_bytes).from_address(ptr.value)
    ten = torch.frombuffer(buf, dtype=torch.uint8).view(*shape)
    weakref.finalize(buf, lib.trl_host_free, ptr.value)       # buf lives as long as any tensor viewing it
    return ten


_PER_FRAME = ("nfaces", "box", "valid", "emb", "sim", "below", "has_sim", "frames", "crops")


def annotate_frame(frame, box, flagged: bool, frame_index: int):
    """server/model.py:66-74, in place: a compared frame gets the red box + "AI Detected - Frame n" when it is counted as
    suspicious, the green box + "Real Frame" otherwise (same colours, thickness, fonts, scales and anchor points)."""
    x1, y1, x2, y2 = int(box[0]), int(box[1]), int(box[2]), int(box[3])
    if flagged:
        cv2.rectangle(frame, (x1, y1), (x2, y2), (0, 0, 255), 2)
        cv2.putText(frame, f"AI Detected - Frame {frame_index}", (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 0, 255), 2, cv2.LINE_AA)
    else:
        cv2.rectangle(frame, (x1, y1), (x2, y2), (0, 255, 0), 2)
        cv2.putText(frame, "Real Frame", (x1, y1 - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 255, 0), 2, cv2.LINE_AA)


class _Chunk:
    __slots__ = ("frames", "proc_pos", "proc_idx", "pinned", "pinned_np", "out", "host", "event", "n", "halo")


def analyze_stream(frame_iter, fps: int, width: int, height: int, writer=None, analyzer: Analyzer | None = None,
                   chunk: int | None = None, keep_emb: bool = False, device_overlay: bool | None = None) -> Trace:
    """The hot loop of server/model.py:42-77 as a two-deep software pipeline.

    While the GPU works on chunk k (H2D copy + trl_process, all asynchronous), the host decodes chunk k+1;
    when chunk k's small result arrays are back, its frames are annotated and written in order.
    ``frame_iter`` yields BGR uint8 frames; ``writer`` (optional) is a cv2.VideoWriter.
    ``device_overlay`` (default: environment TRUELY_DEVICE_OVERLAY=1): boxes and captions are drawn by ``trl_overlay`` on the
    chunk's resident frames, which are then read back for the encoder, instead of by OpenCV on the host copies -- the
    same pixels either way (tests/test_gpu_e2e.py); the host path stays the default because the decoded frames already
    are on the host and this box has no NVENC to hand the resident frames to.
    """
    if device_overlay is None:
        device_overlay = os.environ.get("TRUELY_DEVICE_OVERLAY", "0") == "1"
    device_overlay = bool(device_overlay) and writer is not None
    an = analyzer or get_analyzer()
    t = an.torch
    stride = frame_stride(fps)
    if chunk is None:
        chunk = max(4, min(64, int(96e6 // max(1, width * height * 3))))   # ~96 MB of pinned frames per chunk
    tr = Trace(stride=stride, frame_index=[], valid=[], box=[], sim=[], flagged=[], nfaces=[], emb=[] if keep_emb else None,
               timings=dict(decode_s=0.0, submit_s=0.0, finish_s=0.0, capacity_retries=0))
    rl = RunLength()
    dev = f"cuda:{an.device}"
    pending = deque()
    free_bufs = []
    halo = None
    frame_count = 0

    # chunk buffers (page-locked staging, device frames, result arrays) outlive the call: allocating and, above all, freeing
    # page-locked memory costs ~0.25 s per buffer, more than a short clip takes to analyse (experiments/run_api_profile.py)
    buf_key = (chunk, height, width, bool(keep_emb))
    cache = getattr(an, "_stream_bufs", None)
    if cache is None or cache[0] != buf_key:
        cache = (buf_key, [])
        an._stream_bufs = cache
    free_bufs.extend(cache[1])
    del cache[1][:]

    def new_chunk():
        c = _Chunk()
        c.frames, c.proc_pos, c.proc_idx, c.n, c.halo = [], [], [], 0, None
        if free_bufs:
            c.pinned, c.out, c.host = free_bufs.pop()
            c.pinned_np = c.pinned.numpy()
        else:
            c.pinned = staging_empty(t, (chunk, height, width, 3))
            c.pinned_np = c.pinned.numpy()
            c.out = an.alloc_outputs(chunk)
            c.out["frames"] = t.empty((chunk, height, width, 3), dtype=t.uint8, device=dev)
            # this chunk's own copy of the incoming halo (the producing chunk's buffers are recycled two chunks later)
            c.out["halo_emb"] = t.zeros(L.EMB_DIM, dtype=t.float32, device=dev)
            c.out["halo_valid"] = t.zeros(1, dtype=t.uint8, device=dev)
            c.host = {k: t.empty(c.out[k].shape, dtype=c.out[k].dtype, pin_memory=True)
                      for k in ("nfaces", "box", "valid", "sim", "below", "has_sim")}
            if keep_emb:
                c.host["emb"] = t.empty(c.out["emb"].shape, dtype=t.float32, pin_memory=True)
        return c

    def submit(c):
        nonlocal halo
        t0 = time.perf_counter()
        n = c.n
        if n > 0:
            with t.cuda.stream(an.stream):
                d_frames = c.out["frames"][:n]
                d_frames.copy_(c.pinned[:n], non_blocking=True)
                if halo is not None:
                    c.out["halo_emb"].copy_(halo[0], non_blocking=True)
                    c.out["halo_valid"].copy_(halo[1], non_blocking=True)
                else:
                    c.out["halo_valid"].zero_()
                c.halo = (c.out["halo_emb"], c.out["halo_valid"])
                view = {k: (v[:n] if k in _PER_FRAME else v) for k, v in c.out.items()}
                an.process_device(d_frames, view, c.halo)
                halo = (c.out["last_emb"], c.out["last_valid"])
                for k, h in c.host.items():
                    h[:n].copy_(c.out[k][:n], non_blocking=True)
                c.event = t.cuda.Event()
                c.event.record(an.stream)
        else:
            c.event = None
        pending.append(c)
        tr.timings["submit_s"] += time.perf_counter() - t0

    def redo_uncapped(q, halo_in):
        """capacity overflow: chunk q again through the uncapped slow path, chained on ``halo_in``"""
        n = q.n
        if n == 0:
            return halo_in
        S = an.crop_size
        if "crops" not in q.out:
            q.out["crops"] = t.empty((chunk, S, S, 3), dtype=t.uint8, device=dev)
        o = q.out
        an.detect_align_uncapped(o["frames"][:n], o["box"][:n], o["valid"][:n], o["nfaces"][:n], o["crops"][:n])
        he, hv = halo_in if halo_in is not None else (None, None)
        with t.cuda.stream(an.stream):
            an._check(an.lib.trl_facenet_valid(an.ctx, _vp(o["crops"]), _vp(o["valid"]), n, S, _vp(o["emb"]), an._sptr()))
            an._check(an.lib.trl_consistency(
                an.ctx, _vp(o["emb"]), _vp(o["valid"]), n, _vp(he), _vp(hv), THRESHOLD_FACE_SIMILARITY, _vp(o["sim"]),
                _vp(o["below"]), _vp(o["has_sim"]), _vp(o["last_emb"]), _vp(o["last_valid"]), an._sptr()))
            for k, h in q.host.items():
                h[:n].copy_(o[k][:n], non_blocking=True)
        an.stream.synchronize()
        q.event = None
        return (o["last_emb"], o["last_valid"])

    def finish(c):
        t0 = time.perf_counter()
        if c.event is not None:
            c.event.synchronize()
            try:
                an.check_capacity()
            except L.TrlError as e:
                if e.code != L.TRL_E_CAPACITY:
                    raise
                # A candidate list overflowed somewhere in the chunks submitted so far (the flag is per context, and a
                # truncated list may also have produced the halo of the next chunk): redo this chunk and every chunk
                # already in flight behind it through the uncapped path, in order, re-chaining the halo.
                tr.timings["capacity_retries"] += 1
                h = c.halo if c.n > 0 else None
                for q in [c] + list(pending):
                    h = redo_uncapped(q, h) if q.n > 0 else h
        host = {k: v[:c.n].numpy() for k, v in c.host.items()}
        drawn = None
        if device_overlay and c.n > 0:
            # the run-length state of every processed frame first (host, exact integers), then one overlay launch on the
            # chunk's resident frames and their copy back for the encoder
            rl2 = RunLength()
            rl2.deepfake_count, rl2.deep_fake_frame_count = rl.deepfake_count, rl.deep_fake_frame_count
            states = np.zeros(c.n, np.uint8)
            for j in range(c.n):
                if host["valid"][j] and host["has_sim"][j]:
                    states[j] = 2 if rl2.step(bool(host["below"][j])) else 1
            if "host_frames_out" not in c.out:       # page-locked and cacheable (the staging buffer is write-combined)
                c.out["host_frames_out"] = t.empty((chunk, height, width, 3), dtype=t.uint8, pin_memory=True)
            pend = an.overlay_device(c.out["frames"][:c.n], c.out["box"][:c.n], states, c.proc_idx[:c.n])
            with t.cuda.stream(an.stream):
                c.out["host_frames_out"][:c.n].copy_(c.out["frames"][:c.n], non_blocking=True)
                pend_h = pend.cpu()
            an.stream.synchronize()
            drawn = (c.out["host_frames_out"][:c.n].numpy(), states, pend_h.numpy())
        k = 0
        for pos, frame in enumerate(c.frames):
            if k < c.n and c.proc_pos[k] == pos:
                fidx = c.proc_idx[k]
                valid = bool(host["valid"][k])
                box = host["box"][k]
                flagged = False
                if valid and host["has_sim"][k]:
                    flagged = rl.step(bool(host["below"][k]))
                    if drawn is not None:
                        frame = drawn[0][k]
                        if drawn[2][k]:
                            from . import overlay as O
                            O.draw_text_host(frame, box, int(drawn[1][k]), fidx)
                    elif writer is not None:
                        annotate_frame(frame, box, flagged, fidx)
                tr.frame_index.append(fidx)
                tr.valid.append(valid)
                tr.box.append(box.copy())
                tr.sim.append(float(host["sim"][k]) if host["has_sim"][k] else None)
                tr.flagged.append(flagged)
                tr.nfaces.append(int(host["nfaces"][k]))
                if keep_emb:
                    tr.emb.append(host["emb"][k].copy())
                k += 1
            if writer is not None:
                writer.write(frame)                               # server/model.py:77: every frame, in order
        c.frames = []
        free_bufs.append((c.pinned, c.out, c.host))
        tr.timings["finish_s"] += time.perf_counter() - t0

    cur = new_chunk()
    t_dec = time.perf_counter()
    for frame in frame_iter:
        tr.timings["decode_s"] += time.perf_counter() - t_dec
        if frame_count % stride == 0:                             # server/model.py:46
            # plain memcpy into the staging buffer: a torch copy_ goes through the intra-op thread pool, whose spinning
            # workers take the cores the video decoder's threads need (cap.read was 5x slower next to it)
            np.copyto(cur.pinned_np[cur.n], frame)
            cur.proc_pos.append(len(cur.frames))
            cur.proc_idx.append(frame_count)
            cur.n += 1
        if writer is not None:
            cur.frames.append(frame)
        elif frame_count % stride == 0:
            cur.frames.append(None)
        frame_count += 1
        if cur.n == chunk:
            submit(cur)
            while len(pending) > 1:
                finish(pending.popleft())
            cur = new_chunk()
        t_dec = time.perf_counter()
    if cur.n > 0 or cur.frames:
        submit(cur)
    while pending:
        finish(pending.popleft())
    an._stream_bufs[1].extend(free_bufs)                         # keep the buffers for the next call on this shape
    tr.frame_count = frame_count
    tr.flagged_count = rl.deep_fake_frame_count
    tr.final_run = rl.deepfake_count
    tr.score = final_score(rl.deep_fake_frame_count, rl.deepfake_count, frame_count, fps, stride)
    return tr


def _video_frames(cap):
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret:
            break
        yield frame


def _open_writer(path, fps, width, height):
    """Reference: fourcc 'H264' (server/model.py:35-36).  OpenCV wheels without an H.264 encoder cannot open that
    writer; server.py rejects a missing/empty output file, so fall back to 'mp4v' and say so."""
    out = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"H264"), fps, (width, height))
    if not out.isOpened():
        out.release()
        print("Note: this OpenCV build has no H264 encoder; writing the annotated video with fourcc 'mp4v' instead")
        out = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (width, height))
    return out


def run_trace(video_path_one: str, video_path_two: str | None, analyzer: Analyzer | None = None,
              keep_emb: bool = False) -> Trace:
    """``run`` plus the per-frame trace (superset used by the tests).  ``video_path_two=None`` skips the writer."""
    start_time = time.time()
    if not os.path.exists(video_path_one) or os.path.getsize(video_path_one) == 0:
        print(f"Error: Input video file {video_path_one} doesn't exist or is empty")
        return Trace()
    cap = cv2.VideoCapture(video_path_one)
    if not cap.isOpened():
        print(f"Error: OpenCV couldn't open video file {video_path_one}")
        return Trace()
    fps = int(cap.get(cv2.CAP_PROP_FPS))
    width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    if width <= 0 or height <= 0 or fps <= 0:
        print(f"Error: Invalid video properties: width={width}, height={height}, fps={fps}")
        cap.release()
        return Trace()
    out = None
    try:
        out = _open_writer(video_path_two, fps, width, height) if video_path_two is not None else None
        tr = analyze_stream(_video_frames(cap), fps, width, height, writer=out, analyzer=analyzer, keep_emb=keep_emb)
        execution_time = time.time() - start_time
        print(f"Total Execution Time: {execution_time} seconds")
    finally:
        # the reference releases both on its only exit path (server/model.py:81-82); an exception from the GPU path
        # (mapped to HTTP 500 by server.py) must not leak the capture or leave the writer's file handle open
        cap.release()
        if out is not None:
            out.release()
    if tr.frame_count == 0:
        print("Error: No frames were processed")
        tr.score = 0
    return tr


def run(video_path_one: str, video_path_two: str) -> int:
    """Drop-in for reference server/model.py::run (same signature, return value and side effects)."""
    return run_trace(video_path_one, video_path_two).score

"""Oracle: the reference's ``run`` loop on CPU, with a per-frame trace (TEST INFRASTRUCTURE ONLY).

Restates the logic of reference ``server/model.py:11-95``: frame sampling
(``:40,:46``), ``mtcnn.detect`` (``:47``), box truncate+clamp (``:49-53``),
crop + ``cv2.resize`` to 80x80 + ``to_tensor`` (``:55-58``), FaceNet forward
(``:59``), cosine similarity vs the previous face-bearing frame (``:60-61``),
the run-length counter (``:62-70``) and the score (``:83-95``).  It differs from
the reference only in (a) taking already-constructed models (the reference
rebuilds them per call, ``:18-19``), (b) accepting an in-memory frame iterator,
(c) returning the trace, (d) the output video being optional.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import cv2
import numpy as np
import torch

THRESHOLD_FACE_SIMILARITY = 0.99      # server/model.py:16
THRESHOLD_FRAMES_FOR_DEEPFAKE = 15    # server/model.py:17
RESIZE_DIMENSIONS = (80, 80)          # server/model.py:41


@dataclass
class FrameTrace:
    frame_index: int
    n_faces: int = 0
    box_f: np.ndarray | None = None        # float32[4], largest face as detect() returned it
    box: np.ndarray | None = None          # int[4] after astype(int) + clamp
    embedded: bool = False
    emb: np.ndarray | None = None          # float32[512]
    sim: float | None = None
    run: int = 0                           # deepfake_count after this frame
    flagged: bool = False                  # counted in deep_fake_frame_count


@dataclass
class RunTrace:
    score: int = 0
    fps: int = 0
    width: int = 0
    height: int = 0
    stride: int = 1
    frame_count: int = 0
    flagged_count: int = 0
    final_run: int = 0
    frames: list = field(default_factory=list)


def frame_stride(fps: int) -> int:
    """server/model.py:40"""
    return max(1, int(fps / 7))


def to_tensor_u8(face: np.ndarray) -> torch.Tensor:
    """torchvision.transforms.functional.to_tensor for uint8 HWC (server/model.py:58)."""
    return torch.from_numpy(np.ascontiguousarray(face)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


def clamp_box(box_f: np.ndarray, width: int, height: int) -> np.ndarray:
    """server/model.py:49-53"""
    box = box_f.astype(int)
    box[0] = max(0, box[0])
    box[1] = max(0, box[1])
    box[2] = min(width, box[2])
    box[3] = min(height, box[3])
    return box


def final_score(deep_fake_frame_count: int, deepfake_count: int, frame_count: int, fps: int, stride: int) -> int:
    """server/model.py:83-95"""
    if frame_count == 0:
        return 0
    total_processed_frames = sum(1 for i in range(frame_count) if i % stride == 0)
    if total_processed_frames == 0:
        return 0
    deepfake_percentage = (deep_fake_frame_count / total_processed_frames) * 100
    confidence_factor = min(deepfake_percentage * (deepfake_count / THRESHOLD_FRAMES_FOR_DEEPFAKE), 100)
    if frame_count > fps * 30:
        weighted_score = min(deepfake_percentage + confidence_factor * 0.5, 100)
    else:
        weighted_score = min(deepfake_percentage + confidence_factor * 0.3, 100)
    return max(0, min(100, int(weighted_score)))


def consistency_step(emb, prev, run):
    """server/model.py:60-65 -> (sim, run)."""
    sim = np.dot(emb, prev) / (np.linalg.norm(emb) * np.linalg.norm(prev))
    run = run + 1 if sim < THRESHOLD_FACE_SIMILARITY else 0
    return sim, run


def embed_frame(frame, box, facenet):
    """server/model.py:55-59 -> float32[512] or None."""
    face = frame[box[1]:box[3], box[0]:box[2]]
    if face.size == 0:
        return None
    face = cv2.resize(face, RESIZE_DIMENSIONS)
    face_tensor = to_tensor_u8(face).unsqueeze(0)
    with torch.no_grad():
        return facenet(face_tensor).detach().numpy().flatten()


def reference_run_frames(frames, fps: int, width: int, height: int, mtcnn, facenet,
                         writer=None, annotate: bool = False) -> RunTrace:
    """The hot loop of server/model.py:42-77 over an iterable of BGR uint8 frames."""
    tr = RunTrace(fps=fps, width=width, height=height, stride=frame_stride(fps))
    deepfake_count = 0
    deep_fake_frame_count = 0
    prev = None
    frame_count = 0
    for frame in frames:
        if frame_count % tr.stride == 0:
            ft = FrameTrace(frame_index=frame_count)
            boxes, _ = mtcnn.detect(frame)
            if boxes is not None and len(boxes) > 0:
                ft.n_faces = len(boxes)
                ft.box_f = boxes[0].copy()
                box = clamp_box(boxes[0], width, height)
                ft.box = box.copy()
                if box[2] > box[0] and box[3] > box[1]:
                    emb = embed_frame(frame, box, facenet)
                    if emb is not None:
                        ft.embedded = True
                        ft.emb = emb
                        if prev is not None:
                            sim, deepfake_count = consistency_step(emb, prev, deepfake_count)
                            ft.sim = float(sim)
                            if deepfake_count > THRESHOLD_FRAMES_FOR_DEEPFAKE:
                                deep_fake_frame_count += 1
                                ft.flagged = True
                                if annotate:
                                    cv2.rectangle(frame, (box[0], box[1]), (box[2], box[3]), (0, 0, 255), 2)
                                    cv2.putText(frame, f"AI Detected - Frame {frame_count}", (10, 30),
                                                cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 0, 255), 2, cv2.LINE_AA)
                            elif annotate:
                                cv2.rectangle(frame, (box[0], box[1]), (box[2], box[3]), (0, 255, 0), 2)
                                cv2.putText(frame, "Real Frame", (box[0], box[1] - 10),
                                            cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 255, 0), 2, cv2.LINE_AA)
                        prev = emb
            ft.run = deepfake_count
            tr.frames.append(ft)
        frame_count += 1
        if writer is not None:
            writer.write(frame)
    tr.frame_count = frame_count
    tr.flagged_count = deep_fake_frame_count
    tr.final_run = deepfake_count
    tr.score = final_score(deep_fake_frame_count, deepfake_count, frame_count, fps, tr.stride)
    return tr


def _video_frames(cap):
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret:
            break
        yield frame


def reference_run(video_path_one: str, video_path_two: str | None, mtcnn, facenet,
                  fourcc: str = "H264", max_frames: int | None = None) -> RunTrace:
    """server/model.py:11-95 with guards; ``video_path_two=None`` skips the writer."""
    if not os.path.exists(video_path_one) or os.path.getsize(video_path_one) == 0:
        print(f"Error: Input video file {video_path_one} doesn't exist or is empty")
        return RunTrace()
    cap = cv2.VideoCapture(video_path_one)
    if not cap.isOpened():
        print(f"Error: OpenCV couldn't open video file {video_path_one}")
        return RunTrace()
    fps = int(cap.get(cv2.CAP_PROP_FPS))
    width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    if width <= 0 or height <= 0 or fps <= 0:
        print(f"Error: Invalid video properties: width={width}, height={height}, fps={fps}")
        cap.release()
        return RunTrace()
    writer = None
    if video_path_two is not None:
        writer = cv2.VideoWriter(video_path_two, cv2.VideoWriter_fourcc(*fourcc), fps, (width, height))
    frames = _video_frames(cap)
    if max_frames is not None:
        import itertools
        frames = itertools.islice(frames, max_frames)
    tr = reference_run_frames(frames, fps, width, height, mtcnn, facenet, writer=writer,
                              annotate=writer is not None)
    cap.release()
    if writer is not None:
        writer.release()
    if tr.frame_count == 0:
        print("Error: No frames were processed")
    return tr

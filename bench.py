#!/usr/bin/env python
"""Benchmark of the Truely visual-analysis hot path on B200 (contract: see the task's bench.py section).

A "step" is one pass of the hot path (MTCNN cascade -> crop-align -> FaceNet -> consistency -> run-length score) over
every processed frame of the workload clip: BASELINE.json configs[1], a synthetic 720p 30 fps 60 s clip with one
face = 1800 frames, stride 4 -> 450 processed frames per GPU (weak scaling: with N GPUs the clip is N times longer
and sharded by contiguous frame ranges with the embedding halo, dist.py).

  value : processed frames/s, frames resident in HBM when the timed region starts (device timed, max over ranks)
  e2e   : the same through the host-buffer API: pinned host frames -> H2D inside the timed region -> results D2H
  roofline / stages : per-stage device time (CUDA events on the launching stream) and the dominant kernel's roof
  cpu_baseline : the CPU oracle (port of the reference path) on a bounded sample of the same frames
  --impl reference : the reference's CPU path (oracle port; facenet_pytorch is not installable offline) timed alone
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# offline there are no upstream weight files: the bench runs on the seeded stand-ins and says so in config.weights
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")

# The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on file
# descriptor 1 when the first communicator is created), so everything that is not the result line goes to stderr.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

import numpy as np  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WORKLOADS = {
    "720p30_single": dict(kind="synth", cfg="720p30_single", baseline_config=1,
                          desc="synthetic 720p 30 fps 60 s clip, single face (BASELINE.json configs[1])"),
    "1080p60_multi": dict(kind="synth", cfg="1080p60_multi", baseline_config=3,
                          desc="synthetic 1080p 60 fps clip, 4-8 faces (BASELINE.json configs[3])"),
    "360p30_single": dict(kind="synth", cfg="360p30_single", baseline_config=None,
                          desc="synthetic 360p 30 fps clip, single face (shape of the bundled test clip)"),
    "bundled_360p": dict(kind="file", path=os.path.join(GOLDEN, "bundled_veo3_360p.mp4"), baseline_config=0,
                         desc="the reference's bundled test clip 'Veo 3' 360p h264, 960 frames -> 240 processed "
                              "(BASELINE.json configs[0]; fixture copy tests/golden/bundled_veo3_360p.mp4)"),
    "clips1080p": dict(kind="clips", cfg="1080p60_multi", baseline_config=4, processed_per_clip=75, distinct=4,
                       desc="batch of synthetic 1080p 60 fps 10 s clips (600 frames -> 75 processed each, 4-8 faces), laid end to "
                            "end and frame-sharded with clip-boundary reset (BASELINE.json configs[4]: 256 clips at 8 GPUs)"),
}
METRIC = "frames/sec MTCNN+FaceNet consistency"
UNIT = "processed frames/s"
REF_SAMPLE_NOTE = "reference arm: a bounded sample of the first processed frames of this workload per step, all host threads"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def recorded_traffic(kernel, workload, frames_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest profiles/r*_traffic.json
    (written by profiles/summarize_ncu.py traffic from an `ncu --set full` capture).  None when no capture of this
    kernel at this launch shape is on record -- never a stale literal."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            recs = json.load(open(path))
        except (OSError, ValueError):
            continue
        for r in recs:
            if r.get("kernel") == kernel and r.get("workload") == workload and abs(r.get("frames_per_launch", -1) - frames_per_launch) < 0.5:
                return dict(bytes=float(r["dram_bytes_per_launch"]), source=f"{os.path.relpath(path, ROOT)} <- {r.get('source', '?')}")
    return None


class ClockSampler:
    """SM clock + throttle reasons during the timed region: NVML every 5 ms (a query takes well under a millisecond),
    nvidia-smi every 200 ms if NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.stop, self.th = index, threading.Event(), None
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        try:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.handle))
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop.wait(0.005)

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")] if out else []
                if r and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace(".", "").isdigit():
                    self.mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                    reasons=sorted(self.reasons), samples=len(self.sm), source=self.source)


def pnet_work(geom):
    """Algorithmic FLOPs and compulsory bytes of the P-Net pyramid per frame (2 x MACs of conv1..conv4), from the
    library's own pyramid geometry (trl_pyramid_geometry: one (hs, ws) per scale)."""
    macs = byts = cells = tmacs = 0
    for hs, ws in geom:
        c1h, c1w = hs - 2, ws - 2
        ph, pw = (c1h + 1) // 2, (c1w + 1) // 2
        oh, ow = ph - 4, pw - 4
        macs += c1h * c1w * 10 * 27 + (ph - 2) * (pw - 2) * 16 * 90 + oh * ow * 32 * 144 + oh * ow * 32 * 6
        tmacs += (ph - 2) * (pw - 2) * 16 * 90 + oh * ow * 32 * 144       # conv2 + conv3: the tensor-pipe layers
        byts += 3 * hs * ws * 4
        cells += oh * ow
    return 2 * macs, byts, cells, 2 * tmacs


def pyramid_geometry(an, H, W):
    import ctypes as C
    from truely_b200 import _lib as L
    hs, ws = (C.c_int * L.MAX_SCALES)(), (C.c_int * L.MAX_SCALES)()
    n = an.lib.trl_pyramid_geometry(an.ctx, H, W, None, hs, ws, None, None)
    if n < 0:
        raise RuntimeError(f"trl_pyramid_geometry({H}x{W}) -> {n}")
    return [(hs[k], ws[k]) for k in range(n)]


class Workload:
    """This rank's share of the workload: processed frames in page-locked staging memory + what the scoring needs."""

    def __init__(self, name, args, rank, world, torch):
        from truely_b200.synth import CONFIGS, SyntheticClip
        w = WORKLOADS[name]
        self.name, self.desc, self.kind = name, w["desc"], w["kind"]
        self.clips = None                  # [(n_processed, frame_count)] of the WHOLE batch (clip workloads)
        self.clip_start_local = None       # numpy uint8 [n_local]
        if w["kind"] == "synth":
            cfg = dict(CONFIGS[w["cfg"]])
            cfg["n_frames"] *= world                                   # weak scaling: one N-times longer clip
            clip = SyntheticClip(**cfg, jitter=1.2, seed=0)
            self.H, self.W, self.fps = clip.height, clip.width, clip.fps
            self.stride = max(1, int(clip.fps / 7))
            idx = list(range(0, cfg["n_frames"], self.stride))
            per = len(idx) // world
            mine = idx[rank * per:(rank + 1) * per]
            self.n_local, self.n_total, self.frame_count = len(mine), per * world, cfg["n_frames"]
            self._render = [(clip, i) for i in mine]
            self._sample = [(clip, i) for i in idx[:per]]              # rank 0's range: what the CPU arm samples
        elif w["kind"] == "file":
            import cv2
            cap = cv2.VideoCapture(w["path"])
            if not cap.isOpened():
                raise FileNotFoundError(w["path"])
            self.fps = int(cap.get(cv2.CAP_PROP_FPS))
            self.W, self.H = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
            self.stride = max(1, int(self.fps / 7))
            frames, k = [], 0
            t0 = time.perf_counter()
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                if k % self.stride == 0:
                    frames.append(f)
                k += 1
            cap.release()
            self.decode_s = time.perf_counter() - t0
            # every rank analyses one copy of the clip; with N ranks the batch is N clips (clip-boundary reset)
            self.n_local, self.n_total, self.frame_count = len(frames), len(frames) * world, k
            self._frames = frames
            if world > 1:
                self.clips = [(len(frames), k)] * world
        else:
            cfg = dict(CONFIGS[w["cfg"]])
            ppc, distinct, cpg = w["processed_per_clip"], w["distinct"], args.clips_per_gpu
            self.fps, self.H, self.W = cfg["fps"], cfg["height"], cfg["width"]
            self.stride = max(1, int(self.fps / 7))
            cfg["n_frames"] = ppc * self.stride                         # 600 frames = 10 s at 60 fps
            self._distinct = [SyntheticClip(**cfg, jitter=1.2, seed=100 + d) for d in range(distinct)]
            self.clips = [(ppc, cfg["n_frames"])] * (cpg * world)
            self.n_local, self.n_total, self.frame_count = cpg * ppc, cpg * ppc * world, cfg["n_frames"]
            self._clip_ids = [(rank * cpg + j) % distinct for j in range(cpg)]
            self._ppc = ppc
            # the host-buffer (e2e) path runs on the first clips of every rank only: page-locking all 32 x 466 MB per rank
            # (120 GB on an 8-GPU box) is refused by the boxes of this pool
            self.e2e_clips = min(cpg, args.e2e_clips_per_gpu)
        self.n_e2e, self.clips_e2e = self.n_local, self.clips
        if self.clips is not None:
            from truely_b200.model import clip_start_mask
            m = clip_start_mask(self.clips)
            a = rank * self.n_local
            self.clip_start_local = m[a:a + self.n_local].copy()
            if self.kind == "clips":
                self.n_e2e = self.e2e_clips * self._ppc
                self.clips_e2e = [self.clips[0]] * (self.e2e_clips * world)

    def fill(self, pinned, torch):
        if self.kind == "synth":
            for k, (clip, i) in enumerate(self._render):
                pinned[k].copy_(torch.from_numpy(clip.frame(i)))
        elif self.kind == "file":
            for k, f in enumerate(self._frames):
                pinned[k].copy_(torch.from_numpy(f))
        else:
            # host staging holds the first e2e_clips clips of this rank; resident() tiles the rest on the device
            self._cache = {}                                            # distinct clip -> its processed frames, rendered once
            for j, d in enumerate(self._clip_ids[:self.e2e_clips]):
                for k, f in enumerate(self._clip_frames(d, torch)):
                    pinned[j * self._ppc + k].copy_(f)

    def _clip_frames(self, d, torch):
        if d not in self._cache:
            clip = self._distinct[d]
            self._cache[d] = [torch.from_numpy(clip.frame(i)) for i in clip.processed_indices()]
        return self._cache[d]

    def resident(self, pinned, dev, torch):
        """All n_local processed frames of this rank in HBM (before any timed region)."""
        if self.kind != "clips":
            return pinned.to(dev)
        d = torch.empty((self.n_local, self.H, self.W, 3), dtype=torch.uint8, device=dev)
        on_dev = {}
        for j, c in enumerate(self._clip_ids):
            if c not in on_dev:
                on_dev[c] = torch.stack(self._clip_frames(c, torch)).to(dev)
            d[j * self._ppc:(j + 1) * self._ppc].copy_(on_dev[c])
        self._cache = None
        return d

    def sample_frames(self, n):
        """The first n processed frames of rank 0's range as numpy arrays (regenerated: write-combined staging memory is
        slow to read back) -- what the CPU arm and the parity check run the oracle on."""
        if self.kind == "synth":
            return [clip.frame(i) for clip, i in self._sample[:n]]
        if self.kind == "file":
            return [f.copy() for f in self._frames[:n]]
        clip = self._distinct[0]                                        # rank 0's first clip is distinct clip 0
        return [clip.frame(i) for i in clip.processed_indices()[:n]]

    def config(self, crop, weights):
        """Describes the WORKLOAD only, identically in both arms (`--impl ours` / `--impl reference`)."""
        return {"workload": self.desc, "baseline_config": WORKLOADS[self.name]["baseline_config"], "frame": [self.H, self.W],
                "fps": self.fps, "stride": self.stride, "processed_frames_per_gpu": self.n_local,
                "clips_per_gpu": (len(self.clips) // max(1, self.n_total // self.n_local)) if self.clips else None,
                "e2e_processed_frames_per_gpu": self.n_e2e,
                "crop": crop, "weights": weights, "reference_arm_sample": REF_SAMPLE_NOTE}


def staging_tensor(torch, shape, staging, rank):
    if staging == "wc":
        from truely_b200.model import staging_empty
        return staging_empty(torch, shape, write_combined=True)
    if staging.startswith("interleave"):
        from truely_b200.dist import interleaved_staging
        pinned, desc = interleaved_staging(torch, shape, "all" if staging.endswith("all") else "socket")
        print(f"[bench] rank {rank}: staging {desc}", file=sys.stderr)
        return pinned
    return torch.empty(shape, dtype=torch.uint8, pin_memory=True)


def oracle_models():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    return helpers.oracle_mtcnn(), helpers.oracle_facenet(), helpers


def pick_cpu_threads(torch, probe):
    """The CPU arm uses "all the host threads it can use" -- which on a box of SMT siblings / a busy multi-rank launch is not
    always os.cpu_count(): 32 torch threads ran the oracle slower than 16 on the round-1 boxes.  Time `probe()` once with
    every candidate and keep the fastest; the chosen count is what `cpu_baseline.cores` reports."""
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for n in sorted({cores, max(1, cores // 2)}, reverse=True):
        torch.set_num_threads(n)
        probe()                                   # warm the thread pool at this size
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def weight_sources():
    from truely_b200 import weights as W
    return {"mtcnn": W.load_mtcnn_state()[1], "facenet": W.load_facenet_state()[1]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores.  facenet_pytorch (the reference's arithmetic) cannot
    be installed offline, so this is the oracle port of it (oracle/, `kind: port`), with every host thread, on a bounded
    sample of the same workload per step."""
    if rank != 0:
        return
    import torch
    from oracle.reference_run import reference_run_frames
    import truely_b200  # noqa: F401
    wl = Workload(args.workload, args, 0, world, torch)
    per_step = min(args.cpu_frames_per_step, wl.n_local)
    mt, fn, _ = oracle_models()
    frames = wl.sample_frames(per_step)

    def step():
        # fps=7 -> stride 1: every frame handed in is a processed frame (the sample already is every stride-th frame)
        return reference_run_frames(iter([f.copy() for f in frames]), 7, wl.W, wl.H, mt, fn)

    cores = pick_cpu_threads(torch, lambda: reference_run_frames(iter([f.copy() for f in frames[:3]]), 7, wl.W, wl.H, mt, fn))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f"first {per_step} processed frames of the workload per step, {args.steps} steps; oracle port of the reference path "
              f"(MTCNN + crop + InceptionResnetV1 + consistency, torch CPU fp32, {cores} of {os.cpu_count()} host threads: the faster of all / half)")
    _emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "bundled clip" if wl.kind == "file" else "synthetic",
        "config": wl.config(80, weight_sources()),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def parity_check(wl, mt, fn, helpers, ref_trace, out, n, thr=0.99):
    """North-star tolerances on the bench's own frames: the oracle trace of the cpu_baseline sample against the GPU
    outputs of the same processed frames (the first n of rank 0's range, taken from the last timed `value` step)."""
    from oracle.reference_run import clamp_box
    g = {k: out[k][:n].cpu().numpy() for k in ("nfaces", "box", "valid", "emb", "sim", "has_sim", "below")}
    res = dict(frames=n, face_count_equal=0, frames_with_face=0, min_box_iou=None, max_box_px_diff=0, min_emb_cosine=None,
               max_sim_abs_diff=0.0, in_band_frames=0, flag_mismatches_outside_band=0)
    ious, coss = [], []
    for k, f in enumerate(ref_trace.frames[:n]):
        res["face_count_equal"] += int(int(g["nfaces"][k]) == f.n_faces)
        if f.embedded != bool(g["valid"][k]):
            res["flag_mismatches_outside_band"] += 1
            continue
        if not f.embedded:
            continue
        res["frames_with_face"] += 1
        ious.append(helpers.box_iou(g["box"][k].astype(np.float64), f.box.astype(np.float64)))
        res["max_box_px_diff"] = max(res["max_box_px_diff"], int(np.abs(g["box"][k] - f.box).max()))
        coss.append(helpers.cosine(g["emb"][k], f.emb))
        if f.sim is not None:
            if not g["has_sim"][k]:
                res["flag_mismatches_outside_band"] += 1
                continue
            res["max_sim_abs_diff"] = max(res["max_sim_abs_diff"], abs(float(g["sim"][k]) - f.sim))
            if abs(f.sim - thr) < 1e-3:
                res["in_band_frames"] += 1
            elif bool(g["below"][k]) != (f.sim < thr):
                res["flag_mismatches_outside_band"] += 1
    res["min_box_iou"] = min(ious) if ious else None
    res["min_emb_cosine"] = min(coss) if coss else None
    res["pass"] = bool(res["face_count_equal"] == n and res["flag_mismatches_outside_band"] == 0 and
                       (not ious or min(ious) >= 0.95) and (not coss or min(coss) >= 0.999) and res["max_sim_abs_diff"] < 1e-3)
    res["tolerances"] = "same face count; box IoU >= 0.95; embedding cosine >= 0.999; |sim diff| < 1e-3; identical sim<0.99 decisions outside the 1e-3 band"
    return res


def facenet_sweep(an, torch, pk, batches=(64, 128, 256, 512, 1024, 2048, 4096), sizes=(160, 80), iters=5):
    """BASELINE.json configs[2]: InceptionResnetV1-only sweep over crop batches resident in HBM
    (random uint8 crops, seed 0) at 160x160 and at the reference's 80x80, F.to_tensor normalisation (x / 255, norm 0); the
    largest batch also with fixed_image_standardization ((x - 127.5) / 128, norm 1: folded into the stem, same work)."""
    from truely_b200.model import _vp
    rows = []
    for S in sizes:
        flops = 2 * (1417.7e6 if S == 160 else 233.3e6)
        for B in batches:
            g = torch.Generator(device="cuda").manual_seed(0)
            crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda", generator=g)
            emb = torch.empty((B, 512), dtype=torch.float32, device="cuda")
            for norm in ((0, 1) if B == batches[-1] else (0,)):
                with torch.cuda.stream(an.stream):
                    for _ in range(3):
                        an._check(an.lib.trl_facenet_norm(an.ctx, _vp(crops), B, S, norm, _vp(emb), an._sptr()))
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(an.stream)
                    for _ in range(iters):
                        an._check(an.lib.trl_facenet_norm(an.ctx, _vp(crops), B, S, norm, _vp(emb), an._sptr()))
                    e1.record(an.stream)
                an.stream.synchronize()
                ms = e0.elapsed_time(e1) / iters
                tf = flops * B / (ms * 1e-3) / 1e12
                rows.append({"crop": S, "batch": B, "norm": "x/255" if norm == 0 else "(x-127.5)/128", "ms": ms, "tflops_bf16": tf,
                             "frac_of_sustained": tf / pk["bf16_sustained"], "frac_of_burst": tf / pk["bf16_burst"]})
            del crops, emb
    return rows


def candidate_counts(an, wl, n_local, np):
    """Mean / max number of candidates per frame after each cascade stage on the first frames of the workload (SURVEY.md 8d:
    they define the NMS / R-Net / O-Net work; with the synthetic MTCNN weights they are a property of those weights)."""
    smp = np.stack(wl.sample_frames(min(32, n_local)))
    c = an.process_frames(smp, detail=True).counts.astype(np.float64)
    names = ("after_per_scale_nms", "rnet_inputs", "onet_inputs", "faces")
    return {"frames": int(smp.shape[0]), **{n: {"mean": round(float(c[:, k].mean()), 2), "max": int(c[:, k].max())} for k, n in enumerate(names)}}


def run_api_timing(path, an):
    """The public entry point on a video file, decode included: model.run_trace(path, None) = run() without the output
    video (the decode-inclusive number; decode is OpenCV on the host, server/model.py:23,43)."""
    import contextlib
    import io
    from truely_b200 import model as M
    with contextlib.redirect_stdout(io.StringIO()):
        M.run_trace(path, None, analyzer=an)                      # warm-up (file cache, workspace for this shape)
        t0 = time.perf_counter()
        tr = M.run_trace(path, None, analyzer=an)
        dt = time.perf_counter() - t0
    n = len(tr.frame_index)
    return {"video": os.path.basename(path), "processed_frames": n, "frames": tr.frame_count, "seconds": dt,
            "processed_frames_per_s": n / dt, "host_decode_s": tr.timings.get("decode_s"), "score": tr.score}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="720p30_single", choices=list(WORKLOADS))
    ap.add_argument("--chunk", type=int, default=0, help="frames per host->device copy / cascade call of the e2e path (0 = by frame size)")
    ap.add_argument("--resident-chunk", type=int, default=0,
                    help="frames per cascade call when the frames are already in HBM (`value`); 0 = by frame size")
    ap.add_argument("--clips-per-gpu", type=int, default=32, help="clips1080p: clips per GPU (32 x 8 GPUs = the 256 of configs[4])")
    ap.add_argument("--e2e-clips-per-gpu", type=int, default=8,
                    help="clips1080p: clips per GPU whose frames are page-locked on the host for the e2e measurement")
    ap.add_argument("--cpu-frames", type=int, default=48, help="processed frames in the cpu_baseline / parity_check sample")
    ap.add_argument("--cpu-frames-per-step", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-facenet-sweep", action="store_true")
    ap.add_argument("--staging", default="wc", choices=["wc", "pinned", "interleave-socket", "interleave-all"],
                    help="host staging memory of the e2e path: write-combined page-locked (default) or plain page-locked")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import truely_b200  # noqa: F401
    from truely_b200 import model as M
    from truely_b200.dist import ShardedAnalyzer

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    from truely_b200.dist import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    an = M.Analyzer(device=local_rank)
    wl = Workload(args.workload, args, rank, world, torch)
    H, W, stride, n_local = wl.H, wl.W, wl.stride, wl.n_local
    # frames per cascade call: sized so a call's frames stay near 250 MB (e2e) / 620 MB (resident), as tuned at 720p
    if not args.chunk:
        args.chunk = max(8, min(90, int(round(90 * (720 * 1280) / (H * W)))))
    if not args.resident_chunk:
        args.resident_chunk = max(8, min(225, int(round(225 * (720 * 1280) / (H * W)))))
    dev = f"cuda:{local_rank}"
    n_e2e = wl.n_e2e
    pinned = staging_tensor(torch, (n_e2e, H, W, 3), args.staging, rank)
    wl.fill(pinned, torch)
    d_frames = wl.resident(pinned, dev, torch)                  # resident in HBM before the timed region
    stage_buf = torch.empty((3, args.chunk, H, W, 3), dtype=torch.uint8, device=dev)     # triple-buffered H2D staging
    host_out = {k: torch.empty(n_local, dtype=torch.uint8, pin_memory=True) for k in ("valid", "has_sim", "below")}
    clip_start = torch.from_numpy(wl.clip_start_local).to(dev) if wl.clip_start_local is not None else None
    sharded = ShardedAnalyzer(an, None) if world > 1 else None
    last = {}

    def step(h2d):
        src = pinned if h2d else d_frames
        chunk = args.chunk if h2d else args.resident_chunk
        n = n_e2e if h2d else n_local
        clips = wl.clips_e2e if h2d else wl.clips
        cs = clip_start[:n] if clip_start is not None else None
        if sharded is not None:
            score, flagged, out = sharded.analyze(src, n, wl.frame_count, wl.fps, stride, chunk=chunk, h2d=h2d,
                                                  dev_frames=stage_buf, clip_start=cs, clips=clips)
        else:
            out = an.analyze_resident(src, chunk=chunk, host_out=host_out, h2d=h2d, dev_frames=stage_buf, clip_start=cs)
            an.stream.synchronize()
            v, s, b = host_out["valid"].numpy()[:n], host_out["has_sim"].numpy()[:n], host_out["below"].numpy()[:n]
            if clips is not None:
                score, flagged = M.score_clips(v, s, b, clips, wl.fps, stride)
            else:
                score, flagged, _ = M.score_from_flags(v, s, b, wl.frame_count, wl.fps, stride)
        last["score"], last["flagged"], last["out"] = score, int(sum(flagged)), out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(h2d, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(an.stream)
        for _ in range(k):
            step(h2d)
        e1.record(an.stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step(False)
    an.check_capacity()
    launches0 = an.launch_count()
    with ClockSampler(local_rank) as cs:
        ms_total = timed(False, args.steps)
    launches = an.launch_count() - launches0
    clocks = cs.summary()
    value = args.steps * n_local * world / (ms_total / 1e3)
    resident = {"score": last["score"], "flagged": last["flagged"]}
    out_resident = {k: last["out"][k][:n_local].clone() for k in ("nfaces", "box", "valid", "emb", "sim", "has_sim", "below")}
    # per-stage device times: a second pass of the same steps with CUDA events around every stage.  The events need a
    # serial schedule, so this pass runs the cascade un-pipelined (trl_detect_align_async degrades to trl_detect_align
    # while profiling is on); the stage times therefore add up to a little more than ms_per_step.
    an.set_profiling(True)
    an.read_stage_times()
    barrier()
    for _ in range(args.steps):
        step(False)
    barrier()
    stage_ms, calls = an.read_stage_times()
    an.set_profiling(False)

    # ---- e2e: host (pinned) frames in, flags out, through the same public API
    for _ in range(min(args.warmup, 2)):
        step(True)
    ms_e2e = timed(True, args.steps)
    e2e_value = args.steps * n_e2e * world / (ms_e2e / 1e3)
    h2d_bytes = n_e2e * H * W * 3
    d2h_bytes = 3 * n_e2e
    # the two paths (frames resident / frames from the host) must agree on the result, and so must all ranks
    if n_e2e == n_local:
        result_consistent = (resident["score"] == last["score"] and resident["flagged"] == last["flagged"])
    else:       # the e2e pass ran on the first clips of every rank only: their scores must equal the resident pass's
        ne = wl.e2e_clips
        per = len(wl.clips) // world
        want = [sc for r in range(world) for sc in resident["score"][r * per:r * per + ne]]
        result_consistent = (last["score"] == want)
    if world > 1:
        scores = [None] * world
        dist.all_gather_object(scores, (last["score"], last["flagged"]))
        result_consistent = result_consistent and all(sc == scores[0] for sc in scores)
    # the PCIe floor of the e2e number: the same pinned -> device copies with no compute behind them
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        for k, (a, b) in enumerate(M.chunk_schedule(n_e2e, args.chunk, ramp=True)):
            stage_buf[k % 3, : b - a].copy_(pinned[a:b], non_blocking=True)
    c1.record()
    barrier()
    ms_h2d_only = c0.elapsed_time(c1) / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert result_consistent, f"scores differ between paths / ranks: resident {resident}, e2e {last.get('score')}"

    # ---- per-stage numbers and the dominant kernel's roofline
    pk = peaks()
    per_step = {k: v / args.steps for k, v in stage_ms.items()}
    geom = pyramid_geometry(an, H, W)
    flops_pnet, bytes_pnet, _, tflops_pnet = pnet_work(geom)
    hybrid = an.pnet_precision == 3
    pyr_px = bytes_pnet // 12
    pyr_write_bytes = pyr_px * (16 if hybrid else 12)      # hybrid: fp16 hi + lo pair images (8 + 8 B per pixel), else fp32 planar
    if hybrid:
        bytes_pnet = pyr_px * 8                            # the screen reads the hi image once
    S = an.crop_size
    flops_facenet = 2 * (233.3e6 if S == 80 else 1417.7e6)
    stages = {}
    for name, ms in per_step.items():
        d = {"ms_per_step": ms}
        if ms > 0:
            if name == "pyramid":
                d["algo_bytes_per_frame"] = 3 * H * W + pyr_write_bytes
                d["achieved_gbs"] = d["algo_bytes_per_frame"] * n_local / (ms * 1e-3) / 1e9
                d["frac_of_hbm_peak"] = d["achieved_gbs"] / pk["hbm_gbs"]
            if name == "pnet":
                d["algo_flops_per_frame"] = flops_pnet
                d["achieved_tflops_fp32"] = flops_pnet * n_local / (ms * 1e-3) / 1e12
                d["algo_bytes_per_frame"] = bytes_pnet
                d["achieved_gbs"] = bytes_pnet * n_local / (ms * 1e-3) / 1e9
            if name == "facenet":
                d["algo_flops_per_crop"] = flops_facenet
                d["achieved_tflops_bf16"] = flops_facenet * n_local / (ms * 1e-3) / 1e12
                d["frac_of_sustained_bf16"] = d["achieved_tflops_bf16"] / pk["bf16_sustained"]
        stages[name] = d
    dom = max(per_step, key=per_step.get)
    n_chunks = len(M.chunk_schedule(n_local, args.resident_chunk))
    dom_ms_launch = per_step[dom] / max(1, n_chunks)                  # average duration of one launch of the stage (one chunk)
    frames_per_launch = n_local / n_chunks
    if dom == "facenet":
        ach = flops_facenet * n_local / (per_step[dom] * 1e-3) / 1e12
        roof = {"kernel": "facenet (conv_umma_kernel x103 + stem/pool/head)", "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": None}
    elif dom == "pnet":
        ach = flops_pnet * n_local / (per_step[dom] * 1e-3) / 1e12
        if hybrid:
            # pnet2_kernel: conv1 / conv2 / conv3 as tcgen05 implicit GEMMs in one fp16 pass (fp32 accumulate) that only
            # screens; refine_kernel re-evaluates the near-threshold cells exactly in fp32.  HBM traffic: the fp16 hi pair
            # image (8 B per pyramid pixel) read once.
            roof = {"kernel": "pnet2_kernel (+ refine_kernel)", "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": None,
                    "note": "algorithmic fp32 FLOPs of P-Net (all four conv layers) over the measured bf16 tensor peak; the stage "
                            "time covers the single-pass fp16 tcgen05 screen and the exact fp32 re-evaluation of the screened "
                            "cells (candidates and scores are those of an fp32 P-Net).  The screen's MMAs are N <= 64 wide and "
                            "bound by their shared-memory operand fetch (~40-48 cycles each whatever N is), not by the MMA "
                            "rate; HBM side: %.0f GB/s = %.3f of the HBM peak"
                            % (bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9,
                               bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9 / pk["hbm_gbs"])}
        else:
            # conv2 + conv3 (85 % of the FLOPs) run on the tensor pipe, conv1 on the FMA pipe; HBM traffic is the pyramid read
            # once (ncu: traffic == algorithmic bytes), far from the HBM roof
            roof = {"kernel": "pnet_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_sustained"], "traffic": None,
                    "note": "algorithmic fp32 FLOPs over the measured bf16 tensor peak. fp32-level parity (P-Net maps within 2e-5 of "
                            "the fp32 oracle) is kept with a 3-term fp16 operand split, so the tensor pipe executes 3x the "
                            "algorithmic conv2/conv3 FLOPs: %.1f TFLOP/s executed on the tensor pipe; HBM side: %.0f GB/s = %.3f of "
                            "the HBM peak" % (3 * tflops_pnet * n_local / (per_step[dom] * 1e-3) / 1e12,
                                              bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9,
                                              bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9 / pk["hbm_gbs"])}
        roof["algorithmic_bytes_per_launch"] = bytes_pnet * frames_per_launch
    elif dom == "pyramid":
        # K1 materialises the pyramid (then count it): reads the frame once, writes every level -- fp16 hi + lo
        # pair images (16 B per pyramid pixel) for the hybrid P-Net, fp32 planar (12 B) otherwise
        ab = 3 * H * W + pyr_write_bytes
        ach = ab * n_local / (per_step[dom] * 1e-3) / 1e9
        roof = {"kernel": "pyramid_sep_kernel", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None,
                "note": "algorithmic bytes = frame read once (3 H W) + pyramid written once (%d B per pyramid pixel, %d pixels per "
                        "frame); the kernel is bound by the SM's load/store data pipe (shared-memory wavefronts of the horizontal pass) and "
                        "integer issue on its exact window sums (ncu: LSU data-pipe wavefronts 73-82 %% of peak, issue slots 66-76 %% "
                        "busy, DRAM traffic == algorithmic bytes), not by HBM" % (16 if hybrid else 12, pyr_px)}
        roof["algorithmic_bytes_per_launch"] = ab * frames_per_launch
    elif dom in ("rnet", "onet"):
        # workloads with many candidates per frame (real video): an fp32 FMA kernel, one candidate per CTA.  Roof = the FMA
        # pipe (SMs x 128 lanes x 2 flop x clock; outside the hbm / tensor pair the contract names for the headline workload)
        smp = np.stack(wl.sample_frames(min(32, n_local)))
        cnt = an.process_frames(smp, detail=True).counts.astype(np.float64)
        per_frame = float(cnt[:, 1].mean() if dom == "rnet" else cnt[:, 2].mean())
        flops_c = 3.06e6 if dom == "rnet" else 25.8e6                       # SURVEY.md 8d, per candidate
        prop = torch.cuda.get_device_properties(torch.cuda.current_device())
        peak = prop.multi_processor_count * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
        ach = flops_c * per_frame * n_local / (per_step[dom] * 1e-3) / 1e12
        roof = {"kernel": dom + "_kernel", "bound": "fma", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": None, "note": "%.1f %s inputs per frame on the first %d frames of the workload; fp32 FMA peak from the SM "
                                         "count and clock" % (per_frame, "R-Net" if dom == "rnet" else "O-Net", smp.shape[0])}
    else:
        ach = (3 * H * W) * n_local / (per_step[dom] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None}
    kname = {"pnet": "pnet2_kernel" if hybrid else "pnet_kernel", "pyramid": "pyramid_sep_kernel", "facenet": "conv_umma_kernel"}.get(dom, dom)
    rec = recorded_traffic(kname, args.workload, frames_per_launch)
    if rec is not None:
        roof["traffic"] = rec["bytes"]
        roof["traffic_unit"] = "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
        roof["traffic_source"] = rec["source"]
    roof["peak_source"] = pk["source"] if roof["bound"] != "fma" else "SM count x 128 fp32 lanes x 2 x max SM clock"
    roof["launch_ms"] = dom_ms_launch
    roof["frames_per_launch"] = frames_per_launch

    cpu_baseline = parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.reference_run import reference_run_frames
        mt, fn, helpers = oracle_models()
        nfr = min(args.cpu_frames, n_local)
        frames = wl.sample_frames(nfr)
        cores = pick_cpu_threads(torch, lambda: reference_run_frames(iter([f.copy() for f in frames[:3]]), 7, W, H, mt, fn))
        t0 = time.perf_counter()
        ref_trace = reference_run_frames(iter(frames), 7, W, H, mt, fn)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": nfr / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {nfr} processed frames of the workload, oracle MTCNN+FaceNet+consistency (torch CPU fp32, {cores} of {os.cpu_count()} host threads: the faster of all / half)"}
        # the same trace is the checker of the GPU numbers above (never the thing measured)
        parity = parity_check(wl, mt, fn, helpers, ref_trace, out_resident, nfr)
    sweep = None
    if world == 1 and not args.no_facenet_sweep:
        sweep = facenet_sweep(an, torch, pk)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tensor-core FaceNet (fp32 accumulate) + MTCNN in fp32 (P-Net: fp16 tensor-core screen, fp32 re-evaluation of every near-threshold cell) + u8/int pre-processing",
        "data": "bundled clip" if wl.kind == "file" else "synthetic",
        "config": wl.config(S, {"mtcnn": an.mtcnn_source, "facenet": an.facenet_source}),
        "run_config": {"chunk": args.resident_chunk, "e2e_chunk": args.chunk,
                       "cascade": "tail of chunk k (NMS, crops, R-Net, O-Net, crop-align) on a second stream under the pyramid of chunk k+1",
                       "cache": "inputs larger than L2 (%.2f GB of frames per step per GPU)" % (n_local * H * W * 3 / 1e9),
                       "sharding": ("contiguous frame ranges; one all-gather of shard records (flags + boundary embeddings), "
                                    "cross-shard comparisons resolved on device") if world > 1 else "single GPU",
                       "clips_in_batch": len(wl.clips) if wl.clips else 1,
                       "host_cpus_bound_per_rank": numa,
                       "host_staging": {"wc": "page-locked write-combined (trl_host_alloc)", "pinned": "page-locked"}.get(args.staging, args.staging)},
        "video_frames_per_s": value * stride,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps, "h2d_only_ms_per_step": ms_h2d_only,
                "h2d_only_gbs": h2d_bytes / (ms_h2d_only * 1e-3) / 1e9},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "stages": stages,
        "cpu_baseline": cpu_baseline,
        "parity_check": parity,
        "facenet_sweep": sweep,
        "result": {"score": last.get("score") if wl.clips is None else None,
                   "clip_scores": (last.get("score")[:8] if wl.clips is not None else None),
                   "flagged_frames": last.get("flagged"), "paths_and_ranks_agree": bool(result_consistent)},
    }
    if wl.kind == "file":
        line["host_decode_s"] = wl.decode_s
    if world == 1:
        line["candidates_per_frame"] = candidate_counts(an, wl, n_local, np)
        if wl.kind == "file":
            line["run_api"] = run_api_timing(WORKLOADS[args.workload]["path"], an)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

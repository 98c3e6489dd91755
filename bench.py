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

WORKLOADS = {
    "720p30_single": dict(cfg="720p30_single", desc="synthetic 720p 30 fps 60 s clip, single face (BASELINE.json configs[1])"),
    "1080p60_multi": dict(cfg="1080p60_multi", desc="synthetic 1080p 60 fps clip, 4-8 faces (BASELINE.json configs[3])"),
    "360p30_single": dict(cfg="360p30_single", desc="synthetic 360p 30 fps clip, single face (shape of the bundled test clip)"),
}
METRIC = "frames/sec MTCNN+FaceNet consistency"
UNIT = "processed frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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


def pnet_work(H, W):
    """Algorithmic FLOPs and compulsory bytes of the P-Net pyramid per frame (2 x MACs of conv1..conv4)."""
    from oracle.mtcnn import pyramid_scales  # geometry only
    macs = byts = cells = tmacs = 0
    for s in pyramid_scales(H, W):
        hs, ws = int(H * s + 1), int(W * s + 1)
        c1h, c1w = hs - 2, ws - 2
        ph, pw = (c1h + 1) // 2, (c1w + 1) // 2
        oh, ow = ph - 4, pw - 4
        macs += c1h * c1w * 10 * 27 + (ph - 2) * (pw - 2) * 16 * 90 + oh * ow * 32 * 144 + oh * ow * 32 * 6
        tmacs += (ph - 2) * (pw - 2) * 16 * 90 + oh * ow * 32 * 144       # conv2 + conv3: the tensor-pipe layers
        byts += 3 * hs * ws * 4
        cells += oh * ow
    return 2 * macs, byts, cells, 2 * tmacs


def make_frames(cfg_name, rank, world, torch, staging="wc"):
    """This rank's processed frames of the (world x longer) clip, in page-locked host staging memory
    (staging="wc": write-combined, the product's default staging buffer, model.staging_empty; "pinned": plain)."""
    from truely_b200.synth import CONFIGS, SyntheticClip
    cfg = dict(CONFIGS[cfg_name])
    per_rank_frames = cfg["n_frames"]
    cfg["n_frames"] = per_rank_frames * world
    clip = SyntheticClip(**cfg, jitter=1.2, seed=0)
    stride = max(1, int(clip.fps / 7))
    idx = list(range(0, cfg["n_frames"], stride))
    per = len(idx) // world
    mine = idx[rank * per:(rank + 1) * per]
    shape = (len(mine), clip.height, clip.width, 3)
    if staging == "wc":
        from truely_b200.model import staging_empty
        pinned = staging_empty(torch, shape, write_combined=True)
    elif staging.startswith("interleave"):
        from truely_b200.dist import interleaved_staging
        pinned, desc = interleaved_staging(torch, shape, "all" if staging.endswith("all") else "socket")
        print(f"[bench] rank {rank}: staging {desc}", file=sys.stderr)
    else:
        pinned = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
    for k, i in enumerate(mine):
        pinned[k].copy_(torch.from_numpy(clip.frame(i)))
    return clip, stride, pinned, len(idx), per


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle.reference_run import reference_run_frames
    from truely_b200.synth import CONFIGS, SyntheticClip
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(CONFIGS[WORKLOADS[args.workload]["cfg"]])
    clip = SyntheticClip(**cfg, jitter=1.2, seed=0)
    stride = max(1, int(clip.fps / 7))
    per_step = args.cpu_frames_per_step
    mt, fn = helpers.oracle_mtcnn(), helpers.oracle_facenet()
    frames = [clip.frame(i * stride) for i in range(per_step)]

    def step():
        # fps=7 -> stride 1: every frame handed in is a processed frame (the sample already is every stride-th frame)
        return reference_run_frames(iter([f.copy() for f in frames]), 7, clip.width, clip.height, mt, fn)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = f"{per_step} processed frames of the workload clip per step (oracle MTCNN+FaceNet+consistency, torch CPU fp32)"
    _emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], "frame": [clip.height, clip.width], "stride": stride},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="720p30_single", choices=list(WORKLOADS))
    ap.add_argument("--chunk", type=int, default=90, help="frames per host->device copy / cascade call of the e2e path")
    ap.add_argument("--resident-chunk", type=int, default=225,
                    help="frames per cascade call when the frames are already in HBM (`value`); bounded by the workspace only")
    ap.add_argument("--cpu-frames", type=int, default=48, help="processed frames in the cpu_baseline sample")
    ap.add_argument("--cpu-frames-per-step", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--staging", default="wc", choices=["wc", "pinned", "interleave-socket", "interleave-all"],
                    help="host staging memory of the e2e path: write-combined page-locked (default) or plain page-locked")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
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
    clip, stride, pinned, n_proc_total, n_local = make_frames(WORKLOADS[args.workload]["cfg"], rank, world, torch, args.staging)
    H, W = clip.height, clip.width
    frame_count = clip.n_frames
    dev = f"cuda:{local_rank}"
    d_frames = pinned.to(dev)                                   # resident in HBM before the timed region
    stage_buf = torch.empty((3, args.chunk, H, W, 3), dtype=torch.uint8, device=dev)     # triple-buffered H2D staging
    host_out = {k: torch.empty(n_local, dtype=torch.uint8, pin_memory=True) for k in ("valid", "has_sim", "below")}
    sharded = ShardedAnalyzer(an, None) if world > 1 else None
    last = {}

    def step(h2d):
        src = pinned if h2d else d_frames
        if sharded is not None:
            score, flagged, _ = sharded.analyze(src, n_local + 1, frame_count, clip.fps, stride,
                                                chunk=args.chunk if h2d else args.resident_chunk, h2d=h2d,
                                                dev_frames=stage_buf)
        else:
            an.analyze_resident(src, chunk=args.chunk if h2d else args.resident_chunk, host_out=host_out, h2d=h2d,
                                dev_frames=stage_buf)
            an.stream.synchronize()
            score, flagged, _ = M.score_from_flags(host_out["valid"].numpy(), host_out["has_sim"].numpy(),
                                                   host_out["below"].numpy(), frame_count, clip.fps, stride)
        last["score"], last["flagged"] = score, int(sum(flagged))

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
    e2e_value = args.steps * n_local * world / (ms_e2e / 1e3)
    h2d_bytes = n_local * H * W * 3
    d2h_bytes = 3 * n_local
    # the PCIe floor of the e2e number: the same pinned -> device copies with no compute behind them
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        for k, (a, b) in enumerate(M.chunk_schedule(n_local, args.chunk, ramp=True)):
            stage_buf[k % 3, : b - a].copy_(pinned[a:b], non_blocking=True)
    c1.record()
    barrier()
    ms_h2d_only = c0.elapsed_time(c1) / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-stage numbers and the dominant kernel's roofline
    pk = peaks()
    per_step = {k: v / args.steps for k, v in stage_ms.items()}
    flops_pnet, bytes_pnet, _, tflops_pnet = pnet_work(H, W)
    S = an.crop_size
    flops_facenet = 2 * (233.3e6 if S == 80 else 1417.7e6)
    stages = {}
    for name, ms in per_step.items():
        d = {"ms_per_step": ms}
        if ms > 0:
            if name == "pyramid":
                pyr_px = bytes_pnet // 12
                d["algo_bytes_per_frame"] = 3 * H * W + pyr_px * 12
                d["achieved_gbs"] = d["algo_bytes_per_frame"] * n_local / (ms * 1e-3) / 1e9
            if name == "pnet":
                d["algo_flops_per_frame"] = flops_pnet
                d["achieved_tflops_fp32"] = flops_pnet * n_local / (ms * 1e-3) / 1e12
                d["algo_bytes_per_frame"] = bytes_pnet
                d["achieved_gbs"] = bytes_pnet * n_local / (ms * 1e-3) / 1e9
            if name == "facenet":
                d["algo_flops_per_crop"] = flops_facenet
                d["achieved_tflops_bf16"] = flops_facenet * n_local / (ms * 1e-3) / 1e12
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
        # conv2 + conv3 (85 % of the FLOPs) run on the tensor pipe (conv2 mma.sync, conv3 tcgen05), conv1 on the FMA pipe;
        # HBM traffic is the pyramid read once (ncu: traffic == algorithmic bytes), far from the HBM roof
        ach = flops_pnet * n_local / (per_step[dom] * 1e-3) / 1e12
        roof = {"kernel": "pnet_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_sustained"], "traffic": None,
                "note": "algorithmic fp32 FLOPs over the measured bf16 tensor peak. fp32-level parity (P-Net maps within 2e-5 of "
                        "the fp32 oracle) is kept with a 3-term fp16 operand split, so the tensor pipe executes 3x the "
                        "algorithmic conv2/conv3 FLOPs: %.1f TFLOP/s executed on the tensor pipe; HBM side: %.0f GB/s = %.3f of "
                        "the HBM peak" % (3 * tflops_pnet * n_local / (per_step[dom] * 1e-3) / 1e12,
                                          bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9,
                                          bytes_pnet * n_local / (per_step[dom] * 1e-3) / 1e9 / pk["hbm_gbs"])}
    else:
        ach = (3 * H * W) * n_local / (per_step[dom] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None}
    if dom == "pnet" and args.workload == "720p30_single" and frames_per_launch == 225:
        # dram__bytes_read.sum + dram__bytes_write.sum of one pnet_kernel launch (225 frames), ncu --set full capture
        # summarised in profiles/r01e_pnet_full.md (1.8176 GB + 4.5 MB): the fp32 pyramid of the chunk read exactly once
        roof["traffic"] = 1822.1e6
        roof["traffic_unit"] = "bytes/launch (ncu, profiles/r01e_pnet_full.md)"
        roof["algorithmic_bytes_per_launch"] = bytes_pnet * frames_per_launch
    roof["peak_source"] = pk["source"]
    roof["launch_ms"] = dom_ms_launch
    roof["frames_per_launch"] = frames_per_launch

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers
        from oracle.reference_run import reference_run_frames
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        mt, fn = helpers.oracle_mtcnn(), helpers.oracle_facenet()
        nfr = min(args.cpu_frames, n_local)
        frames = [clip.frame(i * stride) for i in range(nfr)]       # regenerated: the staging buffer is write-combined (slow to read back)
        reference_run_frames(iter(frames[:2]), 7, W, H, mt, fn)            # warm-up
        t0 = time.perf_counter()
        reference_run_frames(iter(frames), 7, W, H, mt, fn)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": nfr / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {nfr} processed frames of the workload clip, oracle MTCNN+FaceNet+consistency (torch CPU fp32)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tensor-core FaceNet (fp32 accumulate) + fp32 MTCNN + u8/int pre-processing", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], "frame": [H, W], "fps": clip.fps, "stride": stride,
                   "processed_frames_per_gpu": n_local, "chunk": args.resident_chunk, "e2e_chunk": args.chunk,
                   "cascade": "tail of chunk k (NMS, crops, R-Net, O-Net, crop-align) on a second stream under the pyramid of chunk k+1",
                   "crop": S,
                   "weights": {"mtcnn": an.mtcnn_source, "facenet": an.facenet_source},
                   "cache": "inputs larger than L2 (%.2f GB of frames per step per GPU)" % (n_local * H * W * 3 / 1e9),
                   "sharding": "contiguous frame ranges + embedding halo all-gather" if world > 1 else "single GPU",
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
        "result": {"score": last.get("score"), "flagged_frames": last.get("flagged")},
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

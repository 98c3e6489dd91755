"""Frame-range sharding over the GPUs of a box.

Per-frame work (MTCNN, crop, FaceNet) is independent, so rank r simply takes a contiguous range of the processed
frames -- of one long clip, or of a batch of clips laid end to end (BASELINE.json configs[4]).  Two pieces of state
cross shard boundaries (reference server/model.py:37-39), and both stop at clip boundaries:

* ``previous_face_encoding``: the first face-bearing frame of a shard is compared with the last face-bearing frame
  *before* the shard -- usually the previous rank's last frame, but if that shard's tail (or all of it) has no face,
  one further back (frames without a face are skipped, server/model.py:56-75) -- unless a clip starts in between.
* ``deepfake_count``: a run-length over a whole clip.  Per-frame flags (3 bytes per processed frame) are gathered and
  the scan + score (exact integer logic) run on the host, identically on every rank, per clip.

One collective per analysis: every rank condenses its range into a fixed-size record (``trl_shard_pack``: flags, the
embedding of the first frame still waiting for a predecessor, the outgoing halo), the records are all-gathered
(``torch.distributed``: NCCL over NVLink on the box, gloo in the CPU tests; a few KB per rank, latency bound, nothing to
fuse with), ``trl_shard_resolve`` finishes the cross-shard comparisons on the gathered buffer on the device, and ONE
device->host copy brings the flags of the whole batch back.  No host synchronisation before that copy.
"""
from __future__ import annotations

import os

import numpy as np

EMB_DIM = 512
_HDR_BYTES = 8 * 4


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced split of n processed frames."""
    return (n * rank) // world, (n * (rank + 1)) // world


def record_pad(n_max: int) -> int:
    return (n_max + 15) & ~15


def record_bytes(n_max: int) -> int:
    """Size of one rank's record (mirrors trl_shard_record_bytes, include/truely_b200.h; checked in tests/test_host.py)."""
    return _HDR_BYTES + 2 * EMB_DIM * 4 + 3 * record_pad(n_max)


def unpack_flags(all_records: np.ndarray, n_max: int):
    """``all_records``: uint8 [world, record_bytes(n_max)] (host).  -> (valid, has_sim, below) of the whole batch in frame
    order, using every rank's true length."""
    world = all_records.shape[0]
    npad = record_pad(n_max)
    off = _HDR_BYTES + 2 * EMB_DIM * 4
    v, s, b = [], [], []
    for r in range(world):
        n = int(all_records[r, :4].view(np.int32)[0])
        fl = all_records[r, off:off + 3 * npad].reshape(3, npad)
        v.append(fl[0, :n]); s.append(fl[1, :n]); b.append(fl[2, :n])
    return np.concatenate(v), np.concatenate(s), np.concatenate(b)


class ShardedAnalyzer:
    """One process per GPU; each rank analyses its contiguous range of the batch's processed frames."""

    def __init__(self, analyzer, group=None):
        import torch.distributed as dist
        self.an = analyzer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._bufs = {}

    # ---- the three device steps around the collective (CUDA: C ABI; the CPU tests substitute numpy restatements)
    def _pack(self, out, n_local, n_max, clip_start):
        from . import model as M
        an, t = self.an, self.an.torch
        key = ("rec", n_max)
        if key not in self._bufs:
            rb = record_bytes(n_max)
            assert rb == an.lib.trl_shard_record_bytes(n_max)
            dev = f"cuda:{an.device}"
            self._bufs[key] = (t.empty(rb, dtype=t.uint8, device=dev), t.empty((self.world, rb), dtype=t.uint8, device=dev),
                               t.empty((self.world, rb), dtype=t.uint8, pin_memory=True))
        rec, _, _ = self._bufs[key]
        an._check(an.lib.trl_shard_pack(an.ctx, M._vp(out["emb"]), M._vp(out["valid"]), M._vp(out["has_sim"]), M._vp(out["below"]),
                                        M._vp(clip_start), n_local, n_max, M._vp(rec), an._sptr()))
        return rec

    def _gather(self, rec, n_max):
        import torch.distributed as dist
        _, allr, _ = self._bufs[("rec", n_max)]
        dist.all_gather_into_tensor(allr, rec, group=self.group)
        return allr

    def _resolve(self, allr, n_max, thr, out):
        from . import model as M
        an = self.an
        an._check(an.lib.trl_shard_resolve(an.ctx, M._vp(allr), self.world, self.rank, n_max, thr, M._vp(out["sim"]),
                                           M._vp(out["below"]), M._vp(out["has_sim"]), an._sptr()))

    def _stream_ctx(self):
        return self.an.torch.cuda.stream(self.an.stream)

    def _to_host(self, allr, n_max):
        _, _, host = self._bufs[("rec", n_max)]
        host.copy_(allr, non_blocking=True)
        self.an.stream.synchronize()          # the only host synchronisation of the analysis
        return host.numpy()

    def analyze(self, local_frames, n_max: int, frame_count: int, fps: int, stride: int, chunk: int = 90,
                h2d: bool = False, dev_frames=None, thr: float = 0.99, clip_start=None, clips=None):
        """``local_frames``: this rank's processed frames (device tensor, or pinned host tensor with h2d=True).
        One clip: returns (score, flagged list for the whole clip, device outputs of the local range).
        A batch of clips (``clips`` = [(n_processed, frame_count), ...] for the WHOLE batch, ``clip_start`` = this rank's
        slice of model.clip_start_mask(clips) on the device): returns (list of per-clip scores, flagged, outputs)."""
        from . import model as M
        an = self.an
        n_local = local_frames.shape[0]
        if n_local > n_max:
            raise ValueError(f"local range of {n_local} frames exceeds n_max={n_max}")
        out = an.analyze_resident(local_frames, chunk=chunk, halo=None, h2d=h2d, dev_frames=dev_frames, thr=thr,
                                  clip_start=clip_start)
        with self._stream_ctx():
            rec = self._pack(out, n_local, n_max, clip_start)
            allr = self._gather(rec, n_max)
            self._resolve(allr, n_max, thr, out)
            host = self._to_host(allr, n_max)
        v, s, b = unpack_flags(host, n_max)
        if clips is not None:
            scores, flagged = M.score_clips(v, s, b, clips, fps, stride)
            return scores, flagged, out
        score, flagged, _ = M.score_from_flags(v, s, b, frame_count, fps, stride)
        return score, flagged, out


def bind_to_gpu_numa_node(index: int):
    """Multi-rank runs: pin this process (and therefore its pinned host buffers, first touch) to the CPUs NVML reports as
    local to GPU `index`, so that the ranks' H2D streams do not all cross the same socket link.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def numa_topology():
    """{node: set(cpus)} and {node: socket} from sysfs (empty dicts when the kernel exposes no NUMA information)."""
    import glob
    import re
    nodes, socket_of = {}, {}
    for path in glob.glob("/sys/devices/system/node/node[0-9]*"):
        n = int(re.search(r"node(\d+)$", path).group(1))
        cpus = set()
        try:
            for part in open(os.path.join(path, "cpulist")).read().strip().split(","):
                if not part:
                    continue
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        except OSError:
            continue
        if not cpus:
            continue
        nodes[n] = cpus
        try:
            socket_of[n] = int(open(f"/sys/devices/system/cpu/cpu{min(cpus)}/topology/physical_package_id").read())
        except OSError:
            socket_of[n] = 0
    return nodes, socket_of


def interleaved_staging(torch, shape, scope: str = "socket"):
    """Experimental host staging buffer whose pages are interleaved over several NUMA nodes (scope "socket": the nodes
    of the socket this process runs on; "all": every node) and then page-locked with cudaHostRegister.  With sub-NUMA
    clustering a rank bound to its GPU's local CPUs otherwise puts its whole staging buffer behind the two or three
    memory channels of one sub-node.  Returns (tensor, description)."""
    import ctypes as C
    import mmap
    nodes, socket_of = numa_topology()
    here = None
    cpus_now = os.sched_getaffinity(0)
    for n, cpus in nodes.items():
        if cpus & cpus_now:
            here = n
            break
    if scope == "socket" and here is not None:
        use = sorted(n for n in nodes if socket_of[n] == socket_of[here])
    else:
        use = sorted(nodes)
    n_bytes = 1
    for s in shape:
        n_bytes *= int(s)
    libc = C.CDLL(None, use_errno=True)
    set_ok = False
    if len(use) > 1:
        mask = 0
        for n in use:
            mask |= 1 << n
        maxnode = max(use) + 2
        arr = (C.c_ulong * ((maxnode + 63) // 64))(*[(mask >> (64 * i)) & (2**64 - 1) for i in range((maxnode + 63) // 64)])
        set_ok = libc.syscall(238, 3, arr, C.c_ulong(maxnode)) == 0          # set_mempolicy(MPOL_INTERLEAVE)
    mm = mmap.mmap(-1, n_bytes)                                               # anonymous, first touch decides the node
    ten = torch.frombuffer(mm, dtype=torch.uint8)
    ten.zero_()                                                               # touch every page under the policy
    if set_ok:
        libc.syscall(238, 0, None, C.c_ulong(0))                              # back to MPOL_DEFAULT
    addr = C.addressof(C.c_char.from_buffer(mm))
    err = torch.cuda.cudart().cudaHostRegister(addr, n_bytes, 0)
    if int(err) != 0:
        raise RuntimeError(f"cudaHostRegister failed: {err}")
    return ten.view(*shape), f"interleaved over NUMA nodes {use} (policy set: {set_ok}), cudaHostRegister"

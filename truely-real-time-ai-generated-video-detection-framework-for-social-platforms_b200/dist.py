"""Frame-range sharding of one clip over the GPUs of a box.

Per-frame work (MTCNN, crop, FaceNet) is independent, so rank r simply takes a contiguous range of the
processed frames.  Two pieces of state cross shard boundaries (reference server/model.py:37-39):

* ``previous_face_encoding``: the first face-bearing frame of a shard is compared with the last face-bearing
  frame *before* the shard -- usually the previous rank's last frame, but if that shard's tail (or all of it)
  has no face, one further back (frames without a face are skipped, server/model.py:56-75).  Every rank
  publishes ``(has_any, last_valid_embedding)``; one all-gather of 513 floats per rank; each rank picks the
  nearest preceding rank that has one.  This is the "one-frame embedding halo", exact across faceless gaps.
* ``deepfake_count``: a run-length over the whole clip.  Per-frame flags (3 bytes per processed frame) are
  all-gathered and the scan + score (exact integer logic) run on the host, identically on every rank.

The collectives go through ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests); they move
kilobytes, so they are latency bound and there is nothing to fuse them with.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced split of n processed frames."""
    return (n * rank) // world, (n * (rank + 1)) // world


def exchange_halo(last_emb, last_valid, group=None):
    """All-gather (valid, embedding) and return (halo_emb [512] or None, gathered [world, 513]).

    ``last_emb``: float32 [512] tensor, ``last_valid``: uint8/bool [1] tensor, both on the device the backend
    needs (CUDA for NCCL, CPU for gloo).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.cat([last_valid.to(torch.float32).reshape(1), last_emb.to(torch.float32).reshape(-1)])
    allv = torch.empty((world, mine.numel()), dtype=torch.float32, device=mine.device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(allv, mine, group=group)
    else:
        dist.all_gather(list(allv.unbind(0)), mine, group=group)
    halo = None
    flags = allv[:, 0].cpu().numpy()                # one tiny D2H on CUDA: choosing the source rank is host logic
    for r in range(rank - 1, -1, -1):
        if flags[r] != 0:
            halo = allv[r, 1:].contiguous()
            break
    return halo, allv


def gather_flags(valid, has_sim, below, n_local: int, n_max: int, group=None):
    """All-gather the per-frame flags.  Each rank contributes uint8 [3, n_max] (padded); returns three numpy
    arrays for the whole clip in frame order, using the true local lengths."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = valid.device
    buf = torch.zeros((3, n_max + 4), dtype=torch.uint8, device=dev)
    buf[0, :n_local], buf[1, :n_local], buf[2, :n_local] = valid[:n_local], has_sim[:n_local], below[:n_local]
    # local length rides along in the last 4 bytes (little endian)
    buf[0, n_max:] = torch.tensor(list(int(n_local).to_bytes(4, "little")), dtype=torch.uint8, device=dev)
    allb = torch.empty((world,) + tuple(buf.shape), dtype=torch.uint8, device=dev)
    if buf.is_cuda:
        dist.all_gather_into_tensor(allb, buf, group=group)
    else:
        dist.all_gather(list(allb.unbind(0)), buf, group=group)
    h = allb.cpu().numpy()
    v, s, b = [], [], []
    for r in range(world):
        n = int.from_bytes(bytes(h[r, 0, n_max:].tolist()), "little")
        v.append(h[r, 0, :n]); s.append(h[r, 1, :n]); b.append(h[r, 2, :n])
    return np.concatenate(v), np.concatenate(s), np.concatenate(b)


class ShardedAnalyzer:
    """One process per GPU; each rank analyses its contiguous range of the clip's processed frames."""

    def __init__(self, analyzer, group=None):
        import torch.distributed as dist
        self.an = analyzer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def analyze(self, local_frames, n_max: int, frame_count: int, fps: int, stride: int, chunk: int = 90,
                h2d: bool = False, dev_frames=None, thr: float = 0.99):
        """``local_frames``: this rank's processed frames (device tensor, or pinned host tensor with h2d=True).
        Returns (score, flagged list for the whole clip, device outputs of the local range)."""
        import ctypes as C
        from . import model as M
        an, t = self.an, self.an.torch
        n_local = local_frames.shape[0]
        out = an.analyze_resident(local_frames, chunk=chunk, halo=None, h2d=h2d, dev_frames=dev_frames)
        with t.cuda.stream(an.stream):
            halo, _ = exchange_halo(out["last_emb"], out["last_valid"], self.group)
            if halo is not None:
                # re-evaluate K12 on the local range with the halo: only the first face-bearing frame changes
                an._check(an.lib.trl_consistency(
                    an.ctx, M._vp(out["emb"]), M._vp(out["valid"]), n_local, M._vp(halo), C.c_void_p(0), thr,
                    M._vp(out["sim"]), M._vp(out["below"]), M._vp(out["has_sim"]), C.c_void_p(0), C.c_void_p(0), an._sptr()))
            v, s, b = gather_flags(out["valid"], out["has_sim"], out["below"], n_local, n_max, self.group)
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

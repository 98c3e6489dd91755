"""CPU, world_size 2 and 3 over gloo: the frame-range sharding logic (one all-gather of shard records, halo selection
across faceless shards and clip boundaries, flag unpacking, run-length + per-clip score) gives exactly the single-process
answer.  The three device steps around the collective (K12, trl_shard_pack, trl_shard_resolve) are CUDA-only in the
product; here they are restated in numpy (mirroring csrc/consistency.cu) so the host logic can be exercised without a
GPU.  The CUDA kernels themselves are checked against the same single-process answer in tests/test_gpu_dist.py."""
import contextlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import truely_b200  # noqa: F401
from truely_b200 import dist as D
from truely_b200 import model as M


def np_consistency(emb, valid, halo, thr=0.99, clip_start=None):
    """numpy restatement of consistency_kernel + last_valid_kernel (csrc/consistency.cu)."""
    n = len(valid)
    sim = np.full(n, np.nan, np.float32)
    below = np.zeros(n, np.uint8)
    has = np.zeros(n, np.uint8)
    prev = halo
    for i in range(n):
        if clip_start is not None and clip_start[i]:
            prev = None
        if not valid[i]:
            continue
        if prev is not None:
            s = float(np.dot(emb[i], prev) / (np.linalg.norm(emb[i]) * np.linalg.norm(prev)))
            sim[i], has[i], below[i] = s, 1, 1 if s < thr else 0
        prev = emb[i]
    return sim, below, has, prev


def np_pack(emb, valid, has, below, clip_start, n_local, n_max):
    """numpy restatement of shard_pack_kernel."""
    rec = np.zeros(D.record_bytes(n_max), np.uint8)
    first = -1
    for i in range(n_local):
        if clip_start is not None and clip_start[i]:
            break
        if valid[i]:
            first = i
            break
    j, blocked = n_local - 1, 0
    while j >= 0 and not valid[j]:
        if clip_start is not None and clip_start[j]:
            blocked = 1
            break
        j -= 1
    if blocked:
        j = -1
    rec[:32].view(np.int32)[:4] = (n_local, first, 1 if j >= 0 else 0, blocked)
    fe = rec[32:32 + 2048].view(np.float32)
    le = rec[32 + 2048:32 + 4096].view(np.float32)
    if first >= 0:
        fe[:] = emb[first]
    if j >= 0:
        le[:] = emb[j]
    npad = D.record_pad(n_max)
    fl = rec[32 + 4096:].reshape(3, npad)
    fl[0, :n_local], fl[1, :n_local], fl[2, :n_local] = valid[:n_local], has[:n_local], below[:n_local]
    return rec


def np_resolve(allr, world, rank, n_max, thr, sim, below, has):
    """numpy restatement of shard_resolve_kernel (in place on allr and the local outputs)."""
    npad = D.record_pad(n_max)
    for q in range(1, world):
        hdr = allr[q, :32].view(np.int32)
        first = int(hdr[1])
        if first < 0:
            continue
        src = -1
        for r in range(q - 1, -1, -1):
            h = allr[r, :32].view(np.int32)
            if h[2]:
                src = r
                break
            if h[3]:
                break
        if src < 0:
            continue
        cur = allr[q, 32:32 + 2048].view(np.float32)
        prev = allr[src, 32 + 2048:32 + 4096].view(np.float32)
        s = float(np.dot(cur, prev) / (np.linalg.norm(cur) * np.linalg.norm(prev)))
        fl = allr[q, 32 + 4096:].reshape(3, npad)
        fl[1, first], fl[2, first] = 1, 1 if s < thr else 0
        if q == rank:
            sim[first], has[first], below[first] = s, 1, fl[2, first]


class _FakeAnalyzer:
    """Stands in for model.Analyzer on the CPU: 'analyses' precomputed embeddings with the numpy K12."""
    torch = torch
    stream = None

    def __init__(self, emb, valid):
        self.emb, self.valid = emb, valid

    def analyze_resident(self, local_frames, chunk=90, halo=None, h2d=False, dev_frames=None, thr=0.99, clip_start=None):
        a = int(local_frames[0]) if len(local_frames) else 0         # "frames" are their global indices here
        n = len(local_frames)
        e, v = self.emb[a:a + n], self.valid[a:a + n]
        sim, below, has, _ = np_consistency(e, v, None, thr, clip_start)
        return dict(emb=e, valid=v, sim=sim, below=below, has_sim=has)


class CpuSharded(D.ShardedAnalyzer):
    def _stream_ctx(self):
        return contextlib.nullcontext()

    def _pack(self, out, n_local, n_max, clip_start):
        return torch.from_numpy(np_pack(out["emb"], out["valid"], out["has_sim"], out["below"], clip_start, n_local, n_max))

    def _gather(self, rec, n_max):
        allr = torch.empty((self.world, rec.numel()), dtype=torch.uint8)
        dist.all_gather(list(allr.unbind(0)), rec, group=self.group)
        return allr

    def _resolve(self, allr, n_max, thr, out):
        np_resolve(allr.numpy(), self.world, self.rank, n_max, thr, out["sim"], out["below"], out["has_sim"])

    def _to_host(self, allr, n_max):
        return allr.numpy()


def _make_case(seed, n):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, 512)).astype(np.float32)
    for i in range(1, n):
        a = 0.05 if rng.random() < 0.04 else rng.uniform(0.2, 0.5)      # cos = 1/sqrt(1+a^2): mostly < 0.99, rarely above
        emb[i] = emb[i - 1] / np.linalg.norm(emb[i - 1]) + a * emb[i] / np.linalg.norm(emb[i])
    valid = (rng.random(n) > 0.3).astype(np.uint8)
    return emb, valid


def _worker(rank, world, port, seed, n, holes, clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emb, valid = _make_case(seed, n)
    for a, b in holes:
        valid[a:b] = 0
    a, b = D.shard_range(n, rank, world)
    sh = CpuSharded(_FakeAnalyzer(emb, valid))
    n_max = (n + world - 1) // world + 1
    cs = M.clip_start_mask(clips)[a:b] if clips else None
    res, flagged, out = sh.analyze(np.arange(a, b), n_max, n * 4, 30, 4, clip_start=cs, clips=clips)
    q.put((rank, res, [bool(f) for f in flagged], out["sim"].tolist(), out["has_sim"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(world, seed, n, holes, clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, n, holes, clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


@pytest.mark.parametrize("world,holes", [(2, []), (2, [(40, 75)]), (3, [(30, 70)]), (3, [(0, 45)])])
def test_sharded_equals_single_process(world, holes):
    seed, n = 5, 100
    emb, valid = _make_case(seed, n)
    for a, b in holes:
        valid[a:b] = 0
    sim, below, has, _ = np_consistency(emb, valid, None)
    ref_score, ref_flagged, rl = M.score_from_flags(valid, has, below, n * 4, 30, 4)
    assert sum(ref_flagged) > 0, "case must flag something to be meaningful"
    res = _run_world(world, seed, n, holes, None)
    for rank, score, flagged, lsim, lhas in res:
        assert score == ref_score and flagged == [bool(f) for f in ref_flagged]
        a, b = D.shard_range(n, rank, world)
        assert lhas == has[a:b].tolist()                                   # the local outputs were patched too
        assert np.array_equal(np.nan_to_num(np.array(lsim, np.float32), nan=-2), np.nan_to_num(sim[a:b], nan=-2))


@pytest.mark.parametrize("world,holes,lengths", [
    (2, [], [50, 50]),                   # clip boundary exactly on the shard boundary
    (2, [(45, 60)], [30, 40, 30]),       # boundary inside a shard, faceless gap across the shard boundary
    (3, [(20, 50)], [33, 1, 40, 26]),    # a one-frame clip, a whole shard boundary inside a faceless stretch
    (3, [], [100]),                      # one clip: same as the plain path
    (3, [(30, 72)], [34, 66]),           # clip starts at 34 = first frame of rank 1's range (33..66): halo must not cross
])
def test_sharded_clips_equal_per_clip_runs(world, holes, lengths):
    """BASELINE.json configs[4]: a batch of clips laid end to end and frame-sharded == run() per clip."""
    seed, n = 9, 100
    assert sum(lengths) == n
    emb, valid = _make_case(seed, n)
    for a, b in holes:
        valid[a:b] = 0
    clips = [(m, m * 4) for m in lengths]
    ref_scores, ref_flagged, a = [], [], 0
    for m in lengths:                                                      # every clip on its own, fresh state
        sim, below, has, _ = np_consistency(emb[a:a + m], valid[a:a + m], None)
        sc, fl, _ = M.score_from_flags(valid[a:a + m], has, below, m * 4, 30, 4)
        ref_scores.append(sc)
        ref_flagged += [bool(f) for f in fl]
        a += m
    # the clip-aware kernel restatement on the whole batch agrees with the per-clip runs
    sim, below, has, _ = np_consistency(emb, valid, None, clip_start=M.clip_start_mask(clips))
    sc2, fl2 = M.score_clips(valid, has, below, clips, 30, 4)
    assert sc2 == ref_scores and [bool(f) for f in fl2] == ref_flagged
    for rank, scores, flagged, _, _ in _run_world(world, seed, n, holes, clips):
        assert scores == ref_scores and flagged == ref_flagged


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 450, 19200):
        for w in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1

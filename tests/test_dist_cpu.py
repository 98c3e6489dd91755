"""CPU, world_size 2 and 3 over gloo: the frame-range sharding logic (halo selection across faceless shards, flag
gather, run-length + score) gives exactly the single-process answer.  The per-frame similarity itself is CUDA-only
in the product; here it is restated in numpy so the host logic can be exercised without a GPU."""
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


def _np_consistency(emb, valid, halo, thr=0.99):
    n = len(valid)
    sim = np.full(n, np.nan, np.float32)
    below = np.zeros(n, np.uint8)
    has = np.zeros(n, np.uint8)
    prev = halo
    for i in range(n):
        if not valid[i]:
            continue
        if prev is not None:
            s = float(np.dot(emb[i], prev) / (np.linalg.norm(emb[i]) * np.linalg.norm(prev)))
            sim[i], has[i], below[i] = s, 1, 1 if s < thr else 0
        prev = emb[i]
    return sim, below, has, prev


def _make_case(seed, n):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, 512)).astype(np.float32)
    for i in range(1, n):
        a = 0.05 if rng.random() < 0.04 else rng.uniform(0.2, 0.5)      # cos = 1/sqrt(1+a^2): mostly < 0.99, rarely above
        emb[i] = emb[i - 1] / np.linalg.norm(emb[i - 1]) + a * emb[i] / np.linalg.norm(emb[i])
    valid = (rng.random(n) > 0.3).astype(np.uint8)
    return emb, valid


def _worker(rank, world, port, seed, n, holes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emb, valid = _make_case(seed, n)
    for a, b in holes:
        valid[a:b] = 0
    a, b = D.shard_range(n, rank, world)
    le, lv = emb[a:b], valid[a:b]
    sim, below, has, last = _np_consistency(le, lv, None)
    last_emb = torch.from_numpy(last if last is not None else np.zeros(512, np.float32))
    last_valid = torch.tensor([1 if last is not None else 0], dtype=torch.uint8)
    halo, _ = D.exchange_halo(last_emb, last_valid)
    if halo is not None:
        sim, below, has, _ = _np_consistency(le, lv, halo.numpy())
    n_max = (n + world - 1) // world + 1
    v, s, bl = D.gather_flags(torch.from_numpy(lv), torch.from_numpy(has), torch.from_numpy(below), b - a, n_max)
    score, flagged, rl = M.score_from_flags(v, s, bl, n * 4, 30, 4)
    q.put((rank, score, [bool(f) for f in flagged], rl.deepfake_count))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,holes", [(2, []), (2, [(40, 75)]), (3, [(30, 70)]), (3, [(0, 45)])])
def test_sharded_equals_single_process(world, holes):
    seed, n = 5, 100
    emb, valid = _make_case(seed, n)
    for a, b in holes:
        valid[a:b] = 0
    sim, below, has, _ = _np_consistency(emb, valid, None)
    ref_score, ref_flagged, rl = M.score_from_flags(valid, has, below, n * 4, 30, 4)
    assert sum(ref_flagged) > 0, "case must flag something to be meaningful"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, n, holes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, score, flagged, final_run in res:
        assert score == ref_score and flagged == [bool(f) for f in ref_flagged] and final_run == rl.deepfake_count


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 450, 19200):
        for w in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1

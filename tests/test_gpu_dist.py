"""GPU: the multi-GPU exchange step (csrc/consistency.cu: clip-aware K12, trl_shard_pack, trl_shard_resolve) against the
single-range answer, bit for bit.  On one GPU the ranks are emulated by running pack per range into one buffer (what the
all-gather would deliver) and resolve once per rank; with two or more GPUs the real thing runs over NCCL."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

import helpers as H  # noqa: F401
from truely_b200 import dist as D
from truely_b200 import model as M
from truely_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu
vp = M._vp


def _consistency(an, emb, valid, clip_start, halo=None, thr=0.99):
    n = emb.shape[0]
    t = torch
    out = dict(sim=t.empty(n, dtype=t.float32, device="cuda"), below=t.empty(n, dtype=t.uint8, device="cuda"),
               has_sim=t.empty(n, dtype=t.uint8, device="cuda"), last_emb=t.zeros(512, device="cuda"),
               last_valid=t.zeros(1, dtype=t.uint8, device="cuda"))
    he, hv = halo if halo is not None else (None, None)
    with t.cuda.stream(an.stream):
        an._check(an.lib.trl_consistency_clips(an.ctx, vp(emb), vp(valid), n, vp(clip_start), vp(he), vp(hv), thr, vp(out["sim"]),
                                               vp(out["below"]), vp(out["has_sim"]), vp(out["last_emb"]), vp(out["last_valid"]),
                                               an._sptr()))
    an.stream.synchronize()
    return out


def _case(seed, n, holes):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, 512)).astype(np.float32)
    for i in range(1, n):
        a = 0.05 if rng.random() < 0.04 else rng.uniform(0.1, 0.5)
        emb[i] = emb[i - 1] / np.linalg.norm(emb[i - 1]) + a * emb[i] / np.linalg.norm(emb[i])
    valid = (rng.random(n) > 0.3).astype(np.uint8)
    for a, b in holes:
        valid[a:b] = 0
    return emb, valid


@pytest.mark.parametrize("world,holes,lengths", [
    (2, [], [100]), (2, [(40, 75)], [100]), (3, [(30, 70)], [100]), (3, [(0, 45)], [100]), (8, [(10, 40)], [100]),
    (2, [], [50, 50]), (2, [(45, 60)], [30, 40, 30]), (3, [(20, 50)], [33, 1, 40, 26]), (3, [(30, 72)], [34, 66]),
    (4, [(60, 100)], [25, 25, 25, 25]), (8, [], [7] * 14 + [2]),
])
def test_shard_pack_resolve_equals_single_range(analyzer, world, holes, lengths):
    an = analyzer
    n = sum(lengths)
    emb_h, valid_h = _case(17 + world, n, holes)
    clips = [(m, 4 * m) for m in lengths]
    cs_h = M.clip_start_mask(clips) if len(lengths) > 1 else None
    emb, valid = torch.from_numpy(emb_h).cuda(), torch.from_numpy(valid_h).cuda()
    cs = torch.from_numpy(cs_h).cuda() if cs_h is not None else None
    ref = _consistency(an, emb, valid, cs)
    # numpy restatement (tests/test_dist_cpu.py) of the same kernel: decisions equal, sims to fp32 rounding
    from test_dist_cpu import np_consistency
    sim_np, below_np, has_np, _ = np_consistency(emb_h, valid_h, None, clip_start=cs_h)
    assert np.array_equal(ref["has_sim"].cpu().numpy(), has_np)
    assert np.allclose(np.nan_to_num(ref["sim"].cpu().numpy(), nan=-2), np.nan_to_num(sim_np, nan=-2), atol=2e-6)
    # per-clip runs (fresh state per clip) give the same bits as the clip-aware kernel on the batch
    a = 0
    for m in lengths:
        one = _consistency(an, emb[a:a + m], valid[a:a + m], None)
        for k in ("sim", "below", "has_sim"):
            x, y = one[k], ref[k][a:a + m]
            if x.is_floating_point():
                x, y = torch.nan_to_num(x, nan=-2.0), torch.nan_to_num(y, nan=-2.0)
            assert torch.equal(x, y), k
        a += m
    # emulated ranks
    n_max = (n + world - 1) // world + 1
    rb = an.lib.trl_shard_record_bytes(n_max)
    assert rb == D.record_bytes(n_max)
    allr = torch.zeros((world, rb), dtype=torch.uint8, device="cuda")
    local = []
    for r in range(world):
        a, b = D.shard_range(n, r, world)
        o = _consistency(an, emb[a:b], valid[a:b], cs[a:b] if cs is not None else None)
        local.append(o)
        with torch.cuda.stream(an.stream):
            an._check(an.lib.trl_shard_pack(an.ctx, vp(emb[a:b]) if b > a else None, vp(valid[a:b]) if b > a else None,
                                            vp(o["has_sim"]) if b > a else None, vp(o["below"]) if b > a else None,
                                            vp(cs[a:b]) if cs is not None and b > a else None, b - a, n_max, vp(allr[r]), an._sptr()))
    an.stream.synchronize()
    gathered = allr.clone()
    for r in range(world):
        mine = gathered.clone()                    # every rank resolves its own copy of the gathered buffer
        o = local[r]
        with torch.cuda.stream(an.stream):
            an._check(an.lib.trl_shard_resolve(an.ctx, vp(mine), world, r, n_max, 0.99, vp(o["sim"]), vp(o["below"]),
                                               vp(o["has_sim"]), an._sptr()))
        an.stream.synchronize()
        v, s, b_ = D.unpack_flags(mine.cpu().numpy(), n_max)
        assert np.array_equal(v, valid_h) and np.array_equal(s, ref["has_sim"].cpu().numpy())
        assert np.array_equal(b_, ref["below"].cpu().numpy())
        a, b = D.shard_range(n, r, world)
        for k in ("sim", "below", "has_sim"):
            x, y = o[k], ref[k][a:b]
            if x.is_floating_point():
                x, y = torch.nan_to_num(x, nan=-2.0), torch.nan_to_num(y, nan=-2.0)
            assert torch.equal(x, y), f"rank {r}: local {k} not patched to the single-range value"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TRUELY_ALLOW_SYNTHETIC="1")
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    an = M.Analyzer(device=rank)
    frames, clips = _nccl_clip()
    n = frames.shape[0]
    a, b = D.shard_range(n, rank, world)
    sh = D.ShardedAnalyzer(an)
    dev = torch.from_numpy(frames[a:b]).to(f"cuda:{rank}")
    n_max = (n + world - 1) // world + 1
    score, flagged, out = sh.analyze(dev, n_max, n * 4, 30, 4, chunk=16)
    cs = torch.from_numpy(M.clip_start_mask(clips)[a:b]).to(f"cuda:{rank}")
    scores, flagged_c, out_c = sh.analyze(dev, n_max, n * 4, 30, 4, chunk=16, clip_start=cs, clips=clips)
    an.stream.synchronize()
    q.put((rank, score, [bool(f) for f in flagged], scores, [bool(f) for f in flagged_c],
           torch.nan_to_num(out_c["sim"][:b - a], nan=-2.0).cpu().numpy(), out_c["emb"][:b - a].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _nccl_clip():
    """48 processed 360p frames with a faceless stretch across the 2-rank shard boundary (frames 20..27 are pure noise)."""
    clip = SyntheticClip(360, 640, 30, 4 * 48, n_faces=(1, 1), face_h=(90.0, 130.0), jitter=1.4, seed=23)
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()])
    rng = np.random.default_rng(1)
    frames[20:28] = rng.integers(0, 256, size=(8, 360, 640, 3), dtype=np.uint8)
    return frames, [(18, 72), (30, 120)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_equals_single_gpu(analyzer):
    """ShardedAnalyzer.analyze over NCCL (2 ranks) == analyze_resident on one GPU, bit for bit: flags, sims, score, one
    clip and a two-clip batch, with a faceless gap across the shard boundary."""
    import torch.multiprocessing as mp
    frames, clips = _nccl_clip()
    n = frames.shape[0]
    d = torch.from_numpy(frames).cuda()
    host = {}
    for name, cs in (("one", None), ("clips", torch.from_numpy(M.clip_start_mask(clips)).cuda())):
        out = analyzer.analyze_resident(d, chunk=16, clip_start=cs)
        analyzer.stream.synchronize()
        host[name] = {k: out[k][:n].cpu().numpy().copy() for k in ("valid", "has_sim", "below", "sim", "emb")}
    assert not host["one"]["valid"][20:28].any() and host["one"]["valid"].sum() >= 36
    score1, flagged1, _ = M.score_from_flags(host["one"]["valid"], host["one"]["has_sim"], host["one"]["below"], n * 4, 30, 4)
    scores_c, flagged_c = M.score_clips(host["clips"]["valid"], host["clips"]["has_sim"], host["clips"]["below"], clips, 30, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, score, flagged, scores, fl_c, sim_local, emb_local in res:
        assert score == score1 and flagged == [bool(f) for f in flagged1]
        assert scores == scores_c and fl_c == [bool(f) for f in flagged_c]
        a, b = D.shard_range(n, rank, 2)
        assert np.array_equal(emb_local, host["clips"]["emb"][a:b])
        assert np.array_equal(sim_local, np.nan_to_num(host["clips"]["sim"][a:b], nan=-2.0))

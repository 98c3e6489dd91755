"""Shared test helpers (CPU): oracle construction and a bf16 storage simulation of FaceNet."""
from __future__ import annotations

import os
import sys
from contextlib import contextmanager

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import truely_b200  # noqa: E402,F401
from truely_b200 import weights as W  # noqa: E402
from oracle.inception_resnet_v1 import BasicConv2d, Block8, Block17, Block35, InceptionResnetV1  # noqa: E402
from oracle.mtcnn import MTCNN  # noqa: E402

_CACHE = {}


def oracle_mtcnn() -> MTCNN:
    """Oracle MTCNN loaded with the same weights the CUDA path loads."""
    if "mtcnn" not in _CACHE:
        state, _ = W.load_mtcnn_state()
        m = MTCNN()
        for net in ("pnet", "rnet", "onet"):
            sd = {k[len(net) + 1:]: torch.from_numpy(v) for k, v in state.items() if k.startswith(net + ".")}
            getattr(m, net).load_state_dict(sd, strict=True)
        _CACHE["mtcnn"] = m.eval()
    return _CACHE["mtcnn"]


def oracle_facenet() -> InceptionResnetV1:
    if "facenet" not in _CACHE:
        sd, _ = W.load_facenet_state()
        m = InceptionResnetV1()
        missing, unexpected = m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()},
                                                strict=False)
        assert not unexpected, unexpected
        assert all(k.endswith("num_batches_tracked") for k in missing), missing
        _CACHE["facenet"] = m.eval()
    return _CACHE["facenet"]


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


@contextmanager
def bf16_storage_sim(trunk_fp32: bool = False):
    """Patch the oracle FaceNet blocks to mimic the CUDA path's storage precision:
    BN folded into bf16 weights, bf16 activations between layers, fp32 accumulation,
    fp32 bias/scale/residual arithmetic.  ``trunk_fp32`` keeps the residual stream in fp32."""
    saved = (BasicConv2d.forward, Block35.forward, Block17.forward, Block8.forward)

    def basic_fwd(self, x):
        bn = self.bn
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = bf16_round(self.conv.weight * s[:, None, None, None])
        b = bn.bias - bn.running_mean * s
        y = torch.nn.functional.conv2d(bf16_round(x), w, b, self.conv.stride, self.conv.padding)
        return bf16_round(torch.relu(y))

    def make_block_fwd(branches):
        def fwd(self, x):
            xin = bf16_round(x)
            outs = [getattr(self, b)(xin) for b in branches]
            cat = torch.cat(outs, 1)
            y = torch.nn.functional.conv2d(bf16_round(cat), bf16_round(self.conv2d.weight), self.conv2d.bias)
            y = y * self.scale + (x if trunk_fp32 else xin)
            if not getattr(self, "noReLU", False):
                y = torch.relu(y)
            return y if trunk_fp32 else bf16_round(y)
        return fwd

    BasicConv2d.forward = basic_fwd
    Block35.forward = make_block_fwd(("branch0", "branch1", "branch2"))
    Block17.forward = make_block_fwd(("branch0", "branch1"))
    Block8.forward = make_block_fwd(("branch0", "branch1"))
    try:
        yield
    finally:
        BasicConv2d.forward, Block35.forward, Block17.forward, Block8.forward = saved


def cosine(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))


def box_iou(a, b):
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    x2, y2 = min(a[2], b[2]), min(a[3], b[3])
    inter = max(0.0, x2 - x1) * max(0.0, y2 - y1)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def assert_flags_match_outside_band(valid, sim, flagged, score, frame_count, stride, fps, ref_valid, ref_sim, ref_flagged,
                                    ref_score, thr: float = 0.99, band: float = 1e-3):
    """BASELINE.json tolerance: 'identical flagged-frame set except frames whose distance lies within 1e-3 of the
    threshold'.  Only the `sim < thr` decision of an in-band frame is free; everything else must be identical.  The
    run-length counter couples frames, so the check replays the reference's machine (server/model.py:62-70) on the
    reference decisions with the in-band ones replaced by the tested path's, and requires the tested flagged list and
    score to equal that replay exactly.  With no in-band frame this is plain equality with the reference.
    Returns the number of in-band frames (never skips an assertion)."""
    from truely_b200.model import RunLength, final_score
    assert [bool(v) for v in valid] == [bool(v) for v in ref_valid], "face-bearing frames differ"
    rl = RunLength()
    replay, n_band = [], 0
    for k in range(len(ref_valid)):
        if not ref_valid[k] or ref_sim[k] is None:
            assert sim[k] is None or (isinstance(sim[k], float) and np.isnan(sim[k])), f"frame {k}: unexpected comparison"
            replay.append(False)
            continue
        assert sim[k] is not None, f"frame {k}: comparison missing"
        below_ref, below_got = ref_sim[k] < thr, sim[k] < thr
        if abs(ref_sim[k] - thr) < band:
            n_band += 1
            decision = below_got
        else:
            assert below_got == below_ref, f"frame {k}: sim {sim[k]} vs reference {ref_sim[k]} on different sides of {thr}"
            decision = below_ref
        replay.append(rl.step(bool(decision)))
    assert [bool(f) for f in flagged] == replay, "flagged-frame set differs outside the tolerance band"
    assert score == final_score(rl.deep_fake_frame_count, rl.deepfake_count, frame_count, fps, stride)
    if n_band == 0:
        assert replay == [bool(f) for f in ref_flagged] and score == ref_score
    return n_band

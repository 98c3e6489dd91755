"""Model weights for the hot path: loading, synthesis, BN folding and packing.

The reference obtains its weights from facenet_pytorch (``MTCNN()`` loads the
``{p,r,o}net.pt`` bundled in the wheel, ``InceptionResnetV1(pretrained="vggface2")``
downloads ``20180402-114759-vggface2.pt``; server/model.py:18-19).  Offline those
files are absent, so this module provides, in order of preference:

1. the upstream state dicts if present (``$TRUELY_WEIGHTS_DIR``, the ``data/`` directory
   of an installed ``facenet_pytorch`` wheel -- where ``MTCNN()`` itself reads
   ``{p,r,o}net.pt`` from -- or ``$TORCH_HOME/checkpoints``, where
   ``InceptionResnetV1(pretrained=...)`` caches its download), key names as in
   SURVEY.md Appendix C;
2. ONLY when ``TRUELY_ALLOW_SYNTHETIC=1`` (set by the tests, ``bench.py`` and
   ``smoke()``; never by ``run``): seeded stand-ins: ``data/synth_mtcnn.npz``
   (P/R/O-Net fitted to the synthetic faces by tests/golden/train_synth_mtcnn.py)
   and a He-normal InceptionResnetV1 (numpy PCG64, machine independent) whose
   BatchNorm running statistics come from ``data/synth_facenet_bn.npz``
   (tests/golden/calibrate_synth_facenet.py).  Without that switch a missing
   upstream file raises ``MissingWeightsError``: a user-facing 0-100 score must never
   come from stand-in networks silently.

Everything here is host-side tensor bookkeeping (fold BatchNorm into the conv
weights in fp32, reorder to the layouts ``include/truely_b200.h`` documents, and
concatenate to flat float32 blobs).  No model is evaluated here.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, "data")
BN_EPS = 1e-3

# ------------------------------------------------------------------ architecture tables

# InceptionResnetV1 BasicConv2d layers in execution order:
# (name, cin, cout, kh, kw, stride, pad_h, pad_w)     [SURVEY.md Appendix B]
def facenet_conv_table():
    t = [
        ("conv2d_1a", 3, 32, 3, 3, 2, 0, 0),
        ("conv2d_2a", 32, 32, 3, 3, 1, 0, 0),
        ("conv2d_2b", 32, 64, 3, 3, 1, 1, 1),
        ("conv2d_3b", 64, 80, 1, 1, 1, 0, 0),
        ("conv2d_4a", 80, 192, 3, 3, 1, 0, 0),
        ("conv2d_4b", 192, 256, 3, 3, 2, 0, 0),
    ]
    for i in range(5):
        p = f"repeat_1.{i}"
        t += [(f"{p}.branch0", 256, 32, 1, 1, 1, 0, 0),
              (f"{p}.branch1.0", 256, 32, 1, 1, 1, 0, 0), (f"{p}.branch1.1", 32, 32, 3, 3, 1, 1, 1),
              (f"{p}.branch2.0", 256, 32, 1, 1, 1, 0, 0), (f"{p}.branch2.1", 32, 32, 3, 3, 1, 1, 1),
              (f"{p}.branch2.2", 32, 32, 3, 3, 1, 1, 1)]
    t += [("mixed_6a.branch0", 256, 384, 3, 3, 2, 0, 0),
          ("mixed_6a.branch1.0", 256, 192, 1, 1, 1, 0, 0), ("mixed_6a.branch1.1", 192, 192, 3, 3, 1, 1, 1),
          ("mixed_6a.branch1.2", 192, 256, 3, 3, 2, 0, 0)]
    for i in range(10):
        p = f"repeat_2.{i}"
        t += [(f"{p}.branch0", 896, 128, 1, 1, 1, 0, 0),
              (f"{p}.branch1.0", 896, 128, 1, 1, 1, 0, 0), (f"{p}.branch1.1", 128, 128, 1, 7, 1, 0, 3),
              (f"{p}.branch1.2", 128, 128, 7, 1, 1, 3, 0)]
    t += [("mixed_7a.branch0.0", 896, 256, 1, 1, 1, 0, 0), ("mixed_7a.branch0.1", 256, 384, 3, 3, 2, 0, 0),
          ("mixed_7a.branch1.0", 896, 256, 1, 1, 1, 0, 0), ("mixed_7a.branch1.1", 256, 256, 3, 3, 2, 0, 0),
          ("mixed_7a.branch2.0", 896, 256, 1, 1, 1, 0, 0), ("mixed_7a.branch2.1", 256, 256, 3, 3, 1, 1, 1),
          ("mixed_7a.branch2.2", 256, 256, 3, 3, 2, 0, 0)]
    for p in [f"repeat_3.{i}" for i in range(5)] + ["block8"]:
        t += [(f"{p}.branch0", 1792, 192, 1, 1, 1, 0, 0),
              (f"{p}.branch1.0", 1792, 192, 1, 1, 1, 0, 0), (f"{p}.branch1.1", 192, 192, 1, 3, 1, 0, 1),
              (f"{p}.branch1.2", 192, 192, 3, 1, 1, 1, 0)]
    return t


# residual up-projection convs (plain Conv2d with bias, no BN): (name, cin, cout, scale)
def facenet_resid_table():
    t = [(f"repeat_1.{i}.conv2d", 96, 256, 0.17) for i in range(5)]
    t += [(f"repeat_2.{i}.conv2d", 256, 896, 0.10) for i in range(10)]
    t += [(f"repeat_3.{i}.conv2d", 384, 1792, 0.20) for i in range(5)]
    t += [("block8.conv2d", 384, 1792, 1.0)]
    return t


MTCNN_SHAPES = {
    "pnet": [("conv1.weight", (10, 3, 3, 3)), ("conv1.bias", (10,)), ("prelu1.weight", (10,)),
             ("conv2.weight", (16, 10, 3, 3)), ("conv2.bias", (16,)), ("prelu2.weight", (16,)),
             ("conv3.weight", (32, 16, 3, 3)), ("conv3.bias", (32,)), ("prelu3.weight", (32,)),
             ("conv4_1.weight", (2, 32, 1, 1)), ("conv4_1.bias", (2,)),
             ("conv4_2.weight", (4, 32, 1, 1)), ("conv4_2.bias", (4,))],
    "rnet": [("conv1.weight", (28, 3, 3, 3)), ("conv1.bias", (28,)), ("prelu1.weight", (28,)),
             ("conv2.weight", (48, 28, 3, 3)), ("conv2.bias", (48,)), ("prelu2.weight", (48,)),
             ("conv3.weight", (64, 48, 2, 2)), ("conv3.bias", (64,)), ("prelu3.weight", (64,)),
             ("dense4.weight", (128, 576)), ("dense4.bias", (128,)), ("prelu4.weight", (128,)),
             ("dense5_1.weight", (2, 128)), ("dense5_1.bias", (2,)),
             ("dense5_2.weight", (4, 128)), ("dense5_2.bias", (4,))],
    "onet": [("conv1.weight", (32, 3, 3, 3)), ("conv1.bias", (32,)), ("prelu1.weight", (32,)),
             ("conv2.weight", (64, 32, 3, 3)), ("conv2.bias", (64,)), ("prelu2.weight", (64,)),
             ("conv3.weight", (64, 64, 3, 3)), ("conv3.bias", (64,)), ("prelu3.weight", (64,)),
             ("conv4.weight", (128, 64, 2, 2)), ("conv4.bias", (128,)), ("prelu4.weight", (128,)),
             ("dense5.weight", (256, 1152)), ("dense5.bias", (256,)), ("prelu5.weight", (256,)),
             ("dense6_1.weight", (2, 256)), ("dense6_1.bias", (2,)),
             ("dense6_2.weight", (4, 256)), ("dense6_2.bias", (4,)),
             ("dense6_3.weight", (10, 256)), ("dense6_3.bias", (10,))],
}

# ------------------------------------------------------------------ loading


class MissingWeightsError(FileNotFoundError):
    """Upstream weight files are absent and the synthetic stand-ins were not explicitly allowed."""


def synthetic_allowed() -> bool:
    return os.environ.get("TRUELY_ALLOW_SYNTHETIC", "").strip().lower() in ("1", "true", "yes")


def _facenet_pytorch_data_dir():
    """``<site-packages>/facenet_pytorch/data`` if that wheel is installed (located without importing it)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("facenet_pytorch")
    except (ImportError, ValueError):
        return None
    if spec is None or not spec.submodule_search_locations:
        return None
    for loc in spec.submodule_search_locations:
        d = os.path.join(loc, "data")
        if os.path.isdir(d):
            return d
    return None


def _weights_dirs():
    dirs = []
    if os.environ.get("TRUELY_WEIGHTS_DIR"):
        dirs.append(os.environ["TRUELY_WEIGHTS_DIR"])
    d = _facenet_pytorch_data_dir()
    if d:
        dirs.append(d)
    torch_home = os.environ.get("TORCH_HOME", os.path.join(os.path.expanduser("~"), ".cache", "torch"))
    dirs.append(os.path.join(torch_home, "checkpoints"))
    return dirs


def _require_synthetic(what, names):
    if not synthetic_allowed():
        raise MissingWeightsError(
            f"{what}: upstream weight file(s) {names} not found in {_weights_dirs()}.  Put them there (or set "
            "TRUELY_WEIGHTS_DIR); the seeded synthetic stand-ins are for tests and benchmarks only and are used "
            "only when TRUELY_ALLOW_SYNTHETIC=1.")


def _find(name):
    for d in _weights_dirs():
        p = os.path.join(d, name)
        if os.path.isfile(p) and os.path.getsize(p) > 0:
            return p
    return None


def _torch_load_np(path):
    import torch
    sd = torch.load(path, map_location="cpu")
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def load_mtcnn_state():
    """-> ({'pnet.conv1.weight': ndarray, ...}, source) for the three nets."""
    paths = {n: _find(f"{n}.pt") for n in ("pnet", "rnet", "onet")}
    if all(paths.values()):
        out = {}
        for n, p in paths.items():
            for k, v in _torch_load_np(p).items():
                out[f"{n}.{k}"] = v.astype(np.float32)
        src = "upstream"
    else:
        # all three or none: a half-upstream cascade would be neither the reference's detector nor the test fixture
        _require_synthetic("MTCNN", [f"{n}.pt" for n, q in paths.items() if q is None])
        p = os.path.join(DATA_DIR, "synth_mtcnn.npz")
        if not os.path.isfile(p):
            raise FileNotFoundError(f"{p} missing: run tests/golden/train_synth_mtcnn.py")
        z = np.load(p)
        out = {k: z[k].astype(np.float32) for k in z.files}
        src = "synthetic"
    for net, shapes in MTCNN_SHAPES.items():
        for key, shp in shapes:
            if tuple(out[f"{net}.{key}"].shape) != shp:
                raise ValueError(f"{net}.{key}: shape {out[f'{net}.{key}'].shape} != {shp}")
    return out, src


def synth_facenet_state(seed: int = 20180402):
    """Seeded stand-in InceptionResnetV1 state dict (numpy, upstream key names).

    Conv kernels are He-normal (std = sqrt(2 / fan_in)); BatchNorm gamma/beta and the
    running statistics are read from ``data/synth_facenet_bn.npz`` (measured once on
    synthetic face crops by the fixture script, so that activations stay O(1) through all
    132 convs and the final embedding is centred like a trained network's).
    """
    rng = np.random.default_rng(seed)
    sd = {}
    bn_path = os.path.join(DATA_DIR, "synth_facenet_bn.npz")
    bn = np.load(bn_path) if os.path.isfile(bn_path) else None
    gain = float(bn["conv_gain"]) if bn is not None and "conv_gain" in bn.files else 1.0

    def bn_params(name, c):
        if bn is not None and f"{name}.running_mean" in bn.files:
            for k in ("weight", "bias", "running_mean", "running_var"):
                sd[f"{name}.{k}"] = bn[f"{name}.{k}"].astype(np.float32)
        else:   # uncalibrated defaults (used only while the fixture itself is being generated)
            sd[f"{name}.weight"] = np.ones(c, np.float32)
            sd[f"{name}.bias"] = np.zeros(c, np.float32)
            sd[f"{name}.running_mean"] = np.zeros(c, np.float32)
            sd[f"{name}.running_var"] = np.ones(c, np.float32)

    for name, cin, cout, kh, kw, _s, _ph, _pw in facenet_conv_table():
        std = gain * np.sqrt(2.0 / (cin * kh * kw))
        sd[f"{name}.conv.weight"] = (rng.standard_normal((cout, cin, kh, kw), dtype=np.float32) * np.float32(std))
        bn_params(f"{name}.bn", cout)
    for name, cin, cout, _scale in facenet_resid_table():
        std = np.sqrt(1.0 / cin)
        sd[f"{name}.weight"] = rng.standard_normal((cout, cin, 1, 1), dtype=np.float32) * np.float32(std)
        sd[f"{name}.bias"] = rng.standard_normal(cout, dtype=np.float32) * np.float32(0.05)
    sd["last_linear.weight"] = rng.standard_normal((512, 1792), dtype=np.float32) * np.float32(np.sqrt(1.0 / 1792))
    bn_params("last_bn", 512)
    return sd


def load_facenet_state():
    p = _find("20180402-114759-vggface2.pt")
    if p is not None:
        sd = {k: v.astype(np.float32) for k, v in _torch_load_np(p).items()
              if not k.startswith("logits") and not k.endswith("num_batches_tracked")}
        return sd, "upstream"
    _require_synthetic("InceptionResnetV1", ["20180402-114759-vggface2.pt"])
    return synth_facenet_state(), "synthetic"


# ------------------------------------------------------------------ folding / packing


def fold_facenet(sd):
    """BatchNorm folded into the convs in fp32.

    Returns (convs, resids, head):
      convs  : list of (name, W[cout,kh,kw,cin] f32, bias[cout] f32) in facenet_conv_table order
      resids : list of (name, W[cout,cin] f32, bias[cout] f32, scale)
      head   : (W[512,1792] f32, bias[512] f32)  = last_linear folded with last_bn
    """
    convs = []
    for name, cin, cout, kh, kw, _s, _ph, _pw in facenet_conv_table():
        w = sd[f"{name}.conv.weight"].astype(np.float32)
        g = sd[f"{name}.bn.weight"].astype(np.float32)
        b = sd[f"{name}.bn.bias"].astype(np.float32)
        m = sd[f"{name}.bn.running_mean"].astype(np.float32)
        v = sd[f"{name}.bn.running_var"].astype(np.float32)
        s = g / np.sqrt(v + np.float32(BN_EPS))
        wf = (w * s[:, None, None, None]).transpose(0, 2, 3, 1)
        convs.append((name, np.ascontiguousarray(wf, np.float32), (b - m * s).astype(np.float32)))
    resids = []
    for name, cin, cout, scale in facenet_resid_table():
        resids.append((name, np.ascontiguousarray(sd[f"{name}.weight"].reshape(cout, cin), np.float32),
                       sd[f"{name}.bias"].astype(np.float32), float(scale)))
    g = sd["last_bn.weight"].astype(np.float32)
    s = g / np.sqrt(sd["last_bn.running_var"].astype(np.float32) + np.float32(BN_EPS))
    hw = sd["last_linear.weight"].astype(np.float32) * s[:, None]
    hb = sd["last_bn.bias"].astype(np.float32) - sd["last_bn.running_mean"].astype(np.float32) * s
    return convs, resids, (np.ascontiguousarray(hw, np.float32), hb.astype(np.float32))


def pack_facenet(sd) -> np.ndarray:
    """Flat float32 blob: for each conv (table order) W then bias; each resid W then bias; head W, bias."""
    convs, resids, head = fold_facenet(sd)
    parts = []
    for _n, w, b in convs:
        parts += [w.ravel(), b.ravel()]
    for _n, w, b, _s in resids:
        parts += [w.ravel(), b.ravel()]
    parts += [head[0].ravel(), head[1].ravel()]
    return np.ascontiguousarray(np.concatenate(parts), np.float32)


def facenet_blob_size() -> int:
    n = 0
    for _name, cin, cout, kh, kw, *_ in facenet_conv_table():
        n += cout * cin * kh * kw + cout
    for _name, cin, cout, _s in facenet_resid_table():
        n += cout * cin + cout
    return n + 512 * 1792 + 512


def pack_mtcnn(state, net: str) -> np.ndarray:
    """Flat float32 blob of one net, tensors in MTCNN_SHAPES order, upstream (PyTorch) layouts."""
    return np.ascontiguousarray(np.concatenate([state[f"{net}.{k}"].astype(np.float32).ravel()
                                                for k, _ in MTCNN_SHAPES[net]]), np.float32)

"""Debug aid: per-level error of the pnet2 screen logits against the fp32 oracle (python experiments/pnet2_debug.py H W faces)."""
import os, sys
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import truely_b200  # noqa
from truely_b200.model import Analyzer
from truely_b200.synth import SyntheticClip
import helpers as H
from oracle import mtcnn as OM
import test_gpu_pnet_hybrid as T

h, w, faces = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
an = Analyzer(device=0)
clip = SyntheticClip(h, w, 30, 64, n_faces=(faces, faces), face_h=(0.15 * h, 0.4 * h), jitter=0.5, seed=21)
frames = np.stack([clip.frame(i) for i in (0, 12)])
B = 2
d_hi, _, off, pitch = T.run_pairs(an, frames)
pnet = H.oracle_mtcnn().pnet
t = torch.from_numpy(frames).permute(0, 3, 1, 2).type(torch.float32)
refs, logits_ref, total = [], [], 0
for s in OM.pyramid_scales(h, w):
    im = ((OM.imresample(t, (int(h * s + 1), int(w * s + 1))) - 127.5) * 0.0078125).contiguous()
    with torch.no_grad():
        x = pnet.prelu1(pnet.conv1(im)); x = pnet.pool1(x); x = pnet.prelu2(pnet.conv2(x)); x = pnet.prelu3(pnet.conv3(x))
        a = pnet.conv4_1(x)
    logits_ref.append(a[:, 1] - a[:, 0])
    total += a[:, 1].numel()
d_logit = torch.full((total,), float("nan"), dtype=torch.float32, device="cuda")
T.ok(an, an.lib.trl_pnet_screen_maps(an.ctx, T.vp(d_hi), B, h, w, T.vp(d_logit), None))
torch.cuda.synchronize()
logit = d_logit.cpu()
o = 0
for k, ref in enumerate(logits_ref):
    got = logit[o:o + ref.numel()].view_as(ref)
    o += ref.numel()
    err = (got - ref).abs()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs()
    idx = np.unravel_index(int(err.argmax()), err.shape)
    print(f"level {k}: map {tuple(ref.shape[1:])} max|dlogit| {err.max().item():.4f} at {idx} (ref {ref[idx].item():.3f} got {got[idx].item():.3f})"
          f" max|dprob| {perr.max().item():.5f}  cells>0.05: {(err > 0.05).sum().item()} max|ref| {ref.abs().max().item():.1f}")
    if (err > 0.05).sum() > 0:
        bad = (err > 0.05).nonzero()
        print("   bad rows", sorted(set(bad[:, 1].tolist()))[:40], "cols", sorted(set(bad[:, 2].tolist()))[:40])

"""Phase timing of pnet_kernel (debug build with -DPNET_TIMING): average clock64 cycles per tile that group A (staging +
conv1) and group B (conv2, conv3 + heads) spend waiting and computing.  Usage (GPU box): python experiments/pnet_timing.py
The timing .so is built next to this script and never replaces the product library."""
import ctypes as C
import glob
import os
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "truely-real-time-ai-generated-video-detection-framework-for-social-platforms_b200")
OUT = os.environ.get("TRL_TIMING_LIB", os.path.join(ROOT, "experiments", "libtruely_b200_timing.so"))


def build():
    srcs = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-DPNET_TIMING", "-shared", "-o", OUT] + srcs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.run(cmd, check=True)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
        sys.exit(0)
    import numpy as np
    import torch
    import truely_b200  # noqa: F401
    from truely_b200 import _lib, model
    from truely_b200.synth import SyntheticClip
    _lib.LIB_PATH = OUT
    lib = _lib.load()
    lib.trl_debug_pnet_timing.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.trl_debug_onet_timing.argtypes = [C.POINTER(C.c_ulonglong)]
    clip = SyntheticClip(720, 1280, 30, 1800, n_faces=(1, 1), seed=0)
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()[:90]])
    an = model.Analyzer(device=0)
    for it in range(3):
        an.process_frames(frames, detail=False)
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 8)()
        lib.trl_debug_pnet_timing(buf)
        n = max(buf[5], 1)
        print(f"iter {it}: {buf[5]} tiles; cycles per tile: group A wait {buf[0] / n:.0f}, conv1 {buf[1] / n:.0f} | "
              f"group B wait {buf[2] / n:.0f}, conv2 {buf[3] / n:.0f}, conv3+heads {buf[4] / n:.0f} (of which waiting for MMA commits {buf[6] / n:.0f})")
        ob = (C.c_ulonglong * 8)()
        lib.trl_debug_onet_timing(ob)
        nc = max(ob[7], 1)
        onames = ["load", "conv1+pool", "conv2+pool", "conv3+pool", "conv4", "dense5", "heads"]
        print(f"        O-Net: {ob[7]} candidates; cycles per candidate: " + ", ".join(f"{onames[i]} {ob[i] / nc:.0f}" for i in range(7)))

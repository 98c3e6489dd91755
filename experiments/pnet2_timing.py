"""Role timing of pnet2_kernel (debug build with -DPNET_TIMING): what the MMA-issue thread waits for and how busy each epilogue
group is, in clock64 cycles per tile.  Usage (GPU box): python experiments/pnet_timing.py --build && python experiments/pnet2_timing.py"""
import ctypes as C
import os
import sys
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.environ.get("TRL_TIMING_LIB", os.path.join(ROOT, "experiments", "libtruely_b200_timing.so"))

if __name__ == "__main__":
    import numpy as np
    import torch
    import truely_b200  # noqa: F401
    from truely_b200 import _lib, model
    from truely_b200.synth import SyntheticClip
    _lib.LIB_PATH = OUT
    lib = _lib.load()
    lib.trl_debug_pnet2_timing.argtypes = [C.POINTER(C.c_ulonglong)]
    clip = SyntheticClip(720, 1280, 30, 1800, n_faces=(1, 1), seed=0)
    nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 225
    frames = np.stack([clip.frame(i) for i in clip.processed_indices()[:nfr]])
    an = model.Analyzer(device=0)
    for it in range(3):
        an.process_frames(frames, detail=False)
        torch.cuda.synchronize()
        b = (C.c_ulonglong * 24)()
        lib.trl_debug_pnet2_timing(b)
        n = max(b[7], 1)
        print(f"iter {it}: {b[7]} tiles; MMA thread {b[0] / n:.0f} cycles per tile, waiting: in_full {b[1] / n:.0f}, p1_ready {b[2] / n:.0f}, "
              f"acc2_empty {b[3] / n:.0f}, acc1_empty {b[4] / n:.0f}, c2_ready {b[5] / n:.0f}, acc3_empty {b[6] / n:.0f} | "
              f"E1 wait {b[8] / n:.0f} busy {b[9] / n:.0f} | E2 wait {b[10] / n:.0f} busy {b[11] / n:.0f} | "
              f"E3 wait {b[12] / n:.0f} busy {b[13] / n:.0f} | TMA wait {b[14] / n:.0f} | issue time conv2 {b[16] / n:.0f} conv1 {b[17] / n:.0f} conv3 {b[18] / n:.0f} | conv3 thread total {b[21] / n:.0f} conv1 thread total {b[22] / n:.0f}")

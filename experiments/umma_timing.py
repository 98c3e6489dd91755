"""Where the persistent FaceNet conv kernel waits (debug build with -DPNET_TIMING): per layer, cycles per k-iteration
the TMA producer thread and the MMA issue thread spend in total and inside their mbarrier waits.
Usage (GPU box): python experiments/umma_timing.py [batch] [crop]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import truely_b200  # noqa: E402,F401
from truely_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.path.join(ROOT, "experiments", "libtruely_b200_timing.so")
from truely_b200.model import Analyzer, _vp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 160
an = Analyzer(device=0)
lib = an.lib
lib.trl_debug_umma_timing.argtypes = [C.POINTER(C.c_ulonglong)]
crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda")
emb = torch.empty((B, 512), dtype=torch.float32, device="cuda")
MAXS = 256
info = (C.c_int * (MAXS * 8))()
ms = (C.c_float * MAXS)()
fn = lib.trl_debug_facenet_step_times
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p]
# the step-times entry point runs the layers one by one; read the counters after each full pass is not enough, so run the
# pass once to warm up, then call layer by layer through the same entry point and difference the counters per layer
with torch.cuda.stream(an.stream):
    ns = fn(an.ctx, _vp(crops), B, S, _vp(emb), MAXS, info, ms, an._sptr())
buf = (C.c_ulonglong * 8)()
lib.trl_debug_umma_timing(buf)       # totals over warm-up + timed pass (2 passes)
n_cta = 148
print(f"batch {B} crop {S}: whole network, per CTA and pass: producer {buf[0] / 2 / n_cta:.0f} cycles of which waiting {buf[1] / 2 / n_cta:.0f}; "
      f"MMA thread {buf[2] / 2 / n_cta:.0f} of which waiting for operands {buf[3] / 2 / n_cta:.0f}, for the accumulator {buf[4] / 2 / n_cta:.0f}")
print(f"k-iterations per CTA and pass {buf[6] / 2 / n_cta:.0f}, tiles {buf[5] / 2 / n_cta:.0f}; "
      f"producer cycles per k-iteration {buf[0] / max(buf[6], 1):.0f} (waiting {buf[1] / max(buf[6], 1):.0f}); "
      f"MMA thread cycles per k-iteration {buf[2] / max(buf[6], 1):.0f} (waiting {buf[3] / max(buf[6], 1):.0f})")

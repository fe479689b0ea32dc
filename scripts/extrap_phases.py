"""Per-phase clock accounting of the extrapolation sweep (debug build)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrmt_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "scripts", "_dbg", "librmt_b200_dbg.so")
import torch
from extrap_probe import case
lib = L.load()
buf = (ctypes.c_ulonglong * 8)()
names = ["phaseA", "wait", "phaseC+idx", "accumulate", "solve", "publish", "-", "targets"]
for label, args in (("N=1025 1 disc R=164 1 layer", (1025, [(0.5, 0.5)], 164, 1, 1)),
                    ("N=4097 64 discs R=164 3 layers", (4097, [((m + .5) / 8, (n + .5) / 8) for n in range(8) for m in range(8)], 164, 3, 1))):
    case(*args); torch.cuda.synchronize()
    lib.rmt_ext_debug_read(buf, 1)
    ms = case(*args); torch.cuda.synchronize()
    lib.rmt_ext_debug_read(buf, 1)
    n = max(buf[7], 1)
    print(label, ": %.3f ms, targets (x2 runs)" % ms, buf[7])
    for k in range(6):
        print("   %-12s %8.0f cycles/target" % (names[k], buf[k] / n))

"""Per-phase clock accounting of the body variant of the extrapolation (debug build, scripts/build_dbg.sh)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import pyrmt_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "scripts", "_dbg", "librmt_b200_dbg.so")
import numpy as np, torch
from pyrmt_b200 import functions as P
from extrap_variants import case
lib = L.load()
buf = (ctypes.c_ulonglong * 16)()
for N in [int(a) for a in sys.argv[1:]] or [4097]:
    X1, X2, phi, dx, dy = case(N, "lattice" if N >= 1025 else "three")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d1, d2, dp = up(X1), up(X2), up(phi)
    P._extrapolate_set_mode("body", 0, 0)
    for _ in range(2):
        P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
    torch.cuda.synchronize()
    lib.rmt_body_debug_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
    e1.record()
    torch.cuda.synchronize()
    lib.rmt_body_debug_read(buf, 1)
    print("N=%d body variant: %.3f ms (ran %s)" % (N, e0.elapsed_time(e1), P._extrapolate_last_mode(N, N)))
    n = max(buf[6], 1)
    for k, nm in enumerate(["chain: wait for record", "chain: header", "chain: pending cells", "chain: accumulate",
                            "chain: solve", "chain: publish"]):
        print("   %-26s %8.0f cycles/target" % (nm, buf[k] / n))
    print("   chain targets", buf[6])
    n = max(buf[11], 1)
    for k, nm in ((8, "prep: wait for queue entry"), (9, "prep: wait for slot"), (10, "prep: phase A")):
        print("   %-26s %8.0f cycles/record" % (nm, buf[k] / n))
    print("   prep records", buf[11])

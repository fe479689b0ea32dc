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
buf = (ctypes.c_ulonglong * 128)()
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
    for l in range(3):
        b = buf[16 * l:16 * l + 16]
        n = max(b[6], 1)
        print(" layer %d: passes %d (paired %d), fallback lookups %d" % (l, b[6], b[7], b[12]))
        print("   chain cycles/pass: wait %.0f  header %.0f  pending %.0f  accumulate %.0f  solve %.0f  publish %.0f"
              % tuple(b[k] / n for k in range(6)))
        n = max(b[11], 1)
        print("   prepare cycles/record: wait entry %.0f  wait slot %.0f  phase A %.0f   (%d records)"
              % (b[8] / n, b[9] / n, b[10] / n, b[11]))
        print("   discovery total cycles: row check + previous-layer wait %.2e  scan + publish %.2e  ring full %.2e"
              % (b[14], b[15], b[13]))
        d = buf[16 * (l + 3):16 * (l + 3) + 16]
        print("   discovery detail (total cycles): metadata %.2e  batch loads %.2e  judge/append %.2e  between %.2e  fence+publish %.2e"
              % (d[0], d[1], d[2], d[3], d[4]))

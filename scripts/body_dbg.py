"""Body variant of the extrapolation on a small case with the watchdog state dumped (debug build)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import pyrmt_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "scripts", "_dbg", "librmt_b200_dbg.so")
import numpy as np, torch
from oracle import rmt_oracle as O
from pyrmt_b200 import functions as P
from extrap_variants import case
lib = L.load()
w = (ctypes.c_int * 192)()
for arg in sys.argv[1:] or ["257"]:
    N = int(arg)
    X1, X2, phi, dx, dy = case(N, "lattice" if N >= 1025 else "three")
    o1, o2 = O.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d1, d2, dp = up(X1), up(X2), up(phi)
    P._extrapolate_set_mode("body", 0, 0)
    reps = int(os.environ.get("REPS", "1"))
    nbad = 0
    for _ in range(reps):
        r1, r2 = P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
        torch.cuda.synchronize()
        nbad += int(bool((r1.cpu().numpy() != o1).any()))
    print("N=%d: %d of %d runs differ from the oracle" % (N, nbad, reps))
    lib.rmt_body_watch_read(w, 1)
    dbg = (ctypes.c_ulonglong * 128)(); lib.rmt_body_debug_read(dbg, 1)
    print("   passes %d paired %d fallback lookups %d" % tuple(sum(dbg[k + 16 * l] for l in range(8)) for k in (6, 7, 12)))
    bad1 = int((r1.cpu().numpy() != o1).sum()); bad2 = int((r2.cpu().numpy() != o2).sum())
    print("N=%d ran %s  mismatches %d %d" % (N, P._extrapolate_last_mode(N, N), bad1, bad2), flush=True)
    for l in range(3):
        for r, nm in enumerate(("chain", "disc", "prep")):
            v = list(w[(l * 3 + r) * 8:(l * 3 + r) * 8 + 8])
            if v[0]:
                print("   watchdog layer %d %s: code %d t/j/qn %d q_count %d consumed %d row_done %d disc_row %d prep_next %d final %d" % (l, nm, *v))
    if bad1:
        jj, ii = np.nonzero(r1.cpu().numpy() != o1)
        print("   first mismatches (j,i):", list(zip(jj[:8].tolist(), ii[:8].tolist())))

"""A few calls of the WENO5 pair advection on the config-4 geometry (timing / ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrmt_b200 import functions as P
from pyrmt_b200.driver import make_case
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
from pyrmt_b200.driver import fsi_step
state, prm = make_case(N, scheme="weno5")
for _ in range(8):
    state, _, _ = fsi_step(state, prm)
a, b, p, X1, X2 = state
print("frac u>=0: %.3f  v>=0: %.3f" % (float((a >= 0).double().mean()), float((b >= 0).double().mean())))
phi = P.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"])
dt = 1e-4
for _ in range(3):
    r = P.advect_reference_map_pair(X1, X2, a, b, prm["X"], prm["Y"], dt, prm["dx"], prm["dy"], phi, "weno5", 0.0, mask_solid=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    r = P.advect_reference_map_pair(X1, X2, a, b, prm["X"], prm["Y"], dt, prm["dx"], prm["dy"], phi, "weno5", 0.0, mask_solid=True)
e1.record(); torch.cuda.synchronize()
print("N=%d weno5 pair: %.3f ms per call" % (N, e0.elapsed_time(e1) / 10))

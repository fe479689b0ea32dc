"""Single-GPU reproduction of what rank r of a P-rank slab run feeds the operators (shapes only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrmt_b200 import functions as F
from pyrmt_b200.driver import make_case
N, P = int(sys.argv[1]), int(sys.argv[2])
state, prm = make_case(N, scheme="weno5")
a, b, p, X1, X2 = state
dx, dy = prm["dx"], prm["dy"]
H = 12
rows = [(N * r) // P for r in range(P + 1)]
for r in range(P):
    r0, r1 = rows[r], rows[r + 1]
    e0, e1 = max(r0 - H, 0), min(r1 + H, N)
    sl = lambda t: t[e0:e1].contiguous()
    phi = F.rebuild_phi_from_reference_map(sl(X1), sl(X2), prm["phi_init"])
    torch.cuda.synchronize(); print("rank", r, "phi ok", tuple(phi.shape), flush=True)
    q1, q2 = F.advect_reference_map_pair(sl(X1), sl(X2), sl(a), sl(b), None, None, 1e-4, dx, dy, phi, "weno5", 0.0, mask_solid=True)
    torch.cuda.synchronize(); print("rank", r, "weno ok", flush=True)
    top = min(512, r0); bot = min(16, N - r1)
    big = lambda t: t[r0 - top:r1 + bot].contiguous()
    phib = F.rebuild_phi_from_reference_map(big(X1), big(X2), prm["phi_init"])
    m = (phib <= 0).double()
    E1, E2 = F.extrapolate_reference_map(big(X1) * m, big(X2) * m, phib, dx, dy, 3, row_offset=r0 - top)
    torch.cuda.synchronize(); print("rank", r, "extrap ok", tuple(phib.shape), F._extrapolate_last_mode(*phib.shape), flush=True)
    st = F.solid_cauchy_stress(sl(X1), sl(X2), dx, dy, 0.1, 0.0, phi)
    torch.cuda.synchronize(); print("rank", r, "stress ok", flush=True)

"""Two calls of the extrapolation (body variant) on the config-4 geometry at 4097^2 (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from pyrmt_b200 import functions as P
from extrap_variants import case
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
X1, X2, phi, dx, dy = case(N, "lattice")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d1, d2, dp = up(X1), up(X2), up(phi)
P._extrapolate_set_mode(sys.argv[2] if len(sys.argv) > 2 else "body", 0, 0)
for _ in range(2):
    P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
torch.cuda.synchronize()
print("ran", P._extrapolate_last_mode(N, N))

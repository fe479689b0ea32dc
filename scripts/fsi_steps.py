"""A few device-resident FSI steps of the config-4 workload (ncu / launch-list target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrmt_b200.driver import make_case, fsi_step
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
state, prm = make_case(N, scheme="weno5")
for _ in range(steps):
    state, dt, _ = fsi_step(state, prm)
torch.cuda.synchronize()
print("ok", N, steps, dt)

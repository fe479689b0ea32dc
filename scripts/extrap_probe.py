"""Micro-benchmark of rmt_extrapolate: how the sweep time depends on the number
and arrangement of discs (chain-latency-bound vs throughput-bound)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrmt_b200 import functions as F
from pyrmt_b200.levelset import DiscSDF

def case(N, centres, Rcells, layers=3, reps=3):
    L = 1.0
    X, Y, dx, dy = F.create_grid(N, N, L, L)
    cx = np.array([c[0] for c in centres]); cy = np.array([c[1] for c in centres])
    R = np.full(len(centres), Rcells * dx)
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    phi = DiscSDF(cx, cy, R)(Xd, Yd)
    X1, X2 = F.mask_solid(Xd, phi), F.mask_solid(Yd, phi)
    F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__":
    g = lambda k: [((m + .5) / k, (n + .5) / k) for n in range(k) for m in range(k)]
    print("N=4097 64 discs R=164 : %.3f ms" % case(4097, g(8), 164))
    print("N=4097  1 disc  R=164 : %.3f ms" % case(4097, [(0.5, 0.5)], 164))
    print("N=4097  1 disc  R=164, 1 layer : %.3f ms" % case(4097, [(0.5, 0.5)], 164, layers=1))
    print("N=4097  8 discs in a row    : %.3f ms" % case(4097, [((m + .5) / 8, 0.5) for m in range(8)], 164))
    print("N=4097  8 discs in a column : %.3f ms" % case(4097, [(0.5, (m + .5) / 8) for m in range(8)], 164))
    print("N=4097  1 disc  R=41  : %.3f ms" % case(4097, [(0.5, 0.5)], 41))
    print("N=4097  no disc       : %.3f ms" % case(4097, [(0.5, 0.5)], 0.0001))
    print("N=1025  1 disc  R=164 : %.3f ms" % case(1025, [(0.5, 0.5)], 164))

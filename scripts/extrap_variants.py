"""Time and cross-check every extrapolation sweep variant on the config-4 geometry (and the small
3-disc test geometry) against the CPU oracle.  Usage: python scripts/extrap_variants.py [N ...]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import rmt_oracle as O
from pyrmt_b200 import functions as P
from pyrmt_b200.driver import disc_lattice

VARIANTS = [("auto", 0), ("body", 0), ("fused8", 0), ("fused16", 0), ("per_layer", 0)]


def case(N, kind):
    X, Y, dx, dy = O.create_grid(N, N, 1.0, 1.0)
    if kind == "lattice":
        cx, cy, R = disc_lattice(8, 1.0, 0.04)
    else:
        cx, cy, R = np.array([0.3, 0.68, 0.5]), np.array([0.3, 0.35, 0.75]), np.array([0.17, 0.12, 0.2])
    phi = O.disc_sdf(X, Y, cx, cy, R)
    m = (phi <= 0).astype(float)
    X1 = (X + 0.03 * np.sin(2.2 * X) * np.cos(1.7 * Y)) * m
    X2 = (Y + 0.02 * np.cos(1.3 * X) * np.sin(2.9 * Y)) * m
    return X1, X2, phi, dx, dy


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128, 257, 513, 1025, 4097]
    out = []
    for N in sizes:
        for kind in (("three",) if N < 1025 else ("lattice",)):
            X1, X2, phi, dx, dy = case(N, kind)
            t0 = time.perf_counter()
            o1, o2 = O.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
            t_or = time.perf_counter() - t0
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            d1, d2, dp = up(X1), up(X2), up(phi)
            for variant, cap in VARIANTS:
                P._extrapolate_set_mode(variant, 0, cap)
                r1, r2 = P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
                torch.cuda.synchronize()
                ran = P._extrapolate_last_mode(N, N)
                ok = bool(np.array_equal(r1.cpu().numpy(), o1) and np.array_equal(r2.cpu().numpy(), o2))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    P.extrapolate_reference_map(d1, d2, dp, dx, dy, 3)
                e1.record()
                torch.cuda.synchronize()
                rec = dict(N=N, kind=kind, variant=variant, ran=ran, bit_exact_vs_oracle=ok,
                           ms=e0.elapsed_time(e1) / 5, oracle_s=t_or)
                print(json.dumps(rec), flush=True)
                out.append(rec)
    P._extrapolate_set_mode("auto", 0, 0)
    return out


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden values of the energy diagnostics (pyRMT/output.py:6-193) and disc_centroid
(benchmarks/common.py:110-115), recorded from the REAL reference (build container only):
    python tests/golden/make_golden_diag.py"""
import os

import numpy as np

from make_golden import HERE, import_reference, smooth


def main():
    F, I, U, C = import_reference()
    import pyRMT.output as OUT
    rng = np.random.default_rng(20240610)
    Nx, Ny = 36, 28
    X, Y, dx, dy = F.create_grid(Nx, Ny, 1.2, 0.9)
    phi = C.initialize_disc(X, Y, 0.5, 0.42, 0.23)
    m = (phi <= 0).astype(float)
    X1, X2 = F.extrapolate_reference_map((X + 0.03 * np.sin(3 * Y)) * m, (0.96 * Y + 0.02 * X) * m, phi, dx, dy, 3)
    a, b = smooth(rng, Y, X, amp=0.4), smooth(rng, Y, X, amp=0.4)
    w_t = 2 * dx
    g = dict(X=X, Y=Y, dx=dx, dy=dy, phi=phi, X1=X1, X2=X2, a=a, b=b, w_t=w_t,
             prm=np.array([1.0, 1.4, 0.02, 0.7, 0.3, 0.05]))        # rho_f rho_s mu_f mu_s kappa eta_s
    g["ke"] = OUT.compute_kinetic_energy(a, b, 1.0, 1.4, phi, w_t, dx, dy)
    g["se"] = OUT.compute_strain_energy(X1, X2, phi, 0.7, dx, dy, kappa=0.3)
    g["se0"] = OUT.compute_strain_energy(X1, X2, phi, 0.7, dx, dy)
    g["diss"] = OUT.compute_viscous_dissipation(a, b, 0.02, phi, w_t, dx, dy, eta_s=0.05)
    g["diss0"] = OUT.compute_viscous_dissipation(a, b, 0.02, phi, w_t, dx, dy)
    g["centroid"] = np.array(C.disc_centroid(phi, X, Y))
    np.savez_compressed(os.path.join(HERE, "diag.npz"), **g)
    print({k: float(g[k]) for k in ("ke", "se", "se0", "diss", "diss0")}, g["centroid"])


if __name__ == "__main__":
    main()

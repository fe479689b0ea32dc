#!/usr/bin/env python
"""Golden vectors for reinitialize_phi_PDE (functions.py:1369-1411), recorded from the REAL reference
(build container only):  python tests/golden/make_golden_reinit.py"""
import os

import numpy as np

from make_golden import HERE, import_reference, smooth


def main():
    F, I, U, C = import_reference()
    rng = np.random.default_rng(20240609)
    Nx, Ny = 36, 28
    X, Y, dx, dy = F.create_grid(Nx, Ny, 1.2, 0.9)
    # a distorted (not signed-distance) level set of a disc, with exact zeros on a few nodes
    phi = (C.initialize_disc(X, Y, 0.55, 0.4, 0.23) * (1.0 + 0.6 * np.sin(5 * X) * np.cos(4 * Y))
           + 0.01 * smooth(rng, Y, X))
    phi[10, 12] = 0.0
    phi[3, 30] = 0.0
    g = dict(dx=dx, dy=dy, phi=phi)
    g["r_none_7"] = F.reinitialize_phi_PDE(phi, dx, dy, 7, None, 0.5)
    g["r_bc_20"] = F.reinitialize_phi_PDE(phi, dx, dy, 20, F.apply_phi_BCs, 0.2)
    g["r_level_set"] = F.reinitialize_level_set(phi, dx, dy, method="pde", num_iters=5, dt_reinit_factor=0.3,
                                                apply_phi_BCs_func=None)
    np.savez_compressed(os.path.join(HERE, "reinit.npz"), **g)
    print("reinit.npz", sum(np.asarray(a).nbytes for a in g.values()) / 1024, "KiB")


if __name__ == "__main__":
    main()

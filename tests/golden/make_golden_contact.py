#!/usr/bin/env python
"""Golden vectors for the two-solid momentum step and the contact force (functions.py:765-895),
recorded from the REAL reference like make_golden.py (build container only):

    python tests/golden/make_golden_contact.py
"""
import os

import numpy as np

from make_golden import HERE, import_reference, smooth


def main():
    F, I, U, C = import_reference()
    rng = np.random.default_rng(20240608)
    Nx, Ny = 36, 28
    X, Y, dx, dy = F.create_grid(Nx, Ny, 1.2, 0.9)
    R = 0.17
    pa = C.initialize_disc(X, Y, 0.43, 0.45, R)
    pb = C.initialize_disc(X, Y, 0.77, 0.47, R)          # gap 0.34 - 2R = 0: touching
    g = dict(X=X, Y=Y, dx=dx, dy=dy, phi_a=pa, phi_b=pb)
    for tag, k_rep, w_c in (("c1", 2.0, 3 * dx), ("c2", 0.7, 5 * dx)):
        fx, fy = F.compute_contact_force(pa, pb, k_rep, w_c, dx, dy)
        g.update({tag + "_k": k_rep, tag + "_w": w_c, tag + "_fx": fx, tag + "_fy": fy})
    ma, mb = (pa <= 0).astype(float), (pb <= 0).astype(float)
    # mildly deformed reference maps so both solid stresses are non-zero
    X1a, X2a = F.extrapolate_reference_map((X + 0.02 * np.sin(3 * Y)) * ma, (Y * 0.97 + 0.01 * X) * ma, pa, dx, dy, 3)
    X1b, X2b = F.extrapolate_reference_map((X * 1.03 - 0.01 * Y) * mb, (Y + 0.015 * np.cos(2 * X)) * mb, pb, dx, dy, 3)
    u = smooth(rng, Y, X, amp=0.3)
    v = smooth(rng, Y, X, amp=0.3)
    p = smooth(rng, Y, X, amp=0.1)
    g.update(X1a=X1a, X2a=X2a, X1b=X1b, X2b=X2b, u=u, v=v, p=p)
    lid = lambda uu, vv: C.no_slip_lid_bc(uu, vv, 1.0)
    ident = lambda uu, vv: (uu.copy(), vv.copy())
    mu_s, kappa, eta_s, dt, rho_s, rho_f, mu_f, w_t = 0.8, 0.3, 0.0, 2e-3, 1.3, 1.0, 0.02, 2 * dx
    g["prm"] = np.array([mu_s, kappa, eta_s, dt, rho_s, rho_f, mu_f, w_t])
    for tag, bc, k_rep, w_c, clamp in (("m_lid", lid, 2.0, 3 * dx, 4.0), ("m_free", ident, 0.0, None, 4.0),
                                      ("m_clamp", lid, 1.5, None, 1.2)):
        un, vn, Jm = F.momentum_step_rk4_2solids(u, v, p, X1a, X2a, X1b, X2b, bc, mu_s, kappa, eta_s, dx, dy, dt,
                                                 rho_s, rho_f, pa, pb, mu_f, w_t, k_rep=k_rep, w_c=w_c,
                                                 detg_clamp=clamp)
        g.update({tag + "_u": un, tag + "_v": vn, tag + "_J": Jm})
    np.savez_compressed(os.path.join(HERE, "contact.npz"), **g)
    print("contact.npz", sum(np.asarray(a).nbytes for a in g.values()) / 1024, "KiB")


if __name__ == "__main__":
    main()

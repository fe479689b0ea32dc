#!/usr/bin/env python
"""Record golden input/output vectors from the REAL reference (pyRMT).

Run in the build container only (needs /root/reference, numba, scipy):

    python tests/golden/make_golden.py

The reference is imported read-only with two empty stub packages (pyamg, h5py)
on sys.path -- neither is reached on the hot path -- and
NUMBA_CACHE_DIR pointed at a writable directory.  Every case stores the inputs
and the reference outputs in tests/golden/<group>.npz; the tests replay the
inputs through the oracle (CPU suite) and through the CUDA path (GPU suite).
Grids are deliberately small and NON-SQUARE so the fixtures stay small and a
transposed index shows up at once.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYRMT_REFERENCE", "/root/reference")


def import_reference():
    stubs = tempfile.mkdtemp(prefix="pyrmt_stubs_")
    for name in ("pyamg", "h5py"):
        os.makedirs(os.path.join(stubs, name))
        open(os.path.join(stubs, name, "__init__.py"), "w").close()
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    sys.path.insert(0, stubs)
    sys.path.insert(0, REF)
    import pyRMT.functions as F
    import pyRMT.interpolators as I
    import pyRMT.utils as U
    sys.path.insert(0, REF)
    import benchmarks.common as C
    return F, I, U, C


def smooth(rng, Y, X, amp=1.0, modes=3, noise=0.0):
    f = np.zeros_like(X)
    for _ in range(modes):
        kx, ky = rng.integers(1, 4, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=2)
        f += rng.uniform(-1, 1) * np.sin(kx * np.pi * X + ph[0]) * np.cos(ky * np.pi * Y + ph[1])
    f *= amp
    if noise:
        f += noise * rng.standard_normal(X.shape)
    return f


def main():
    F, I, U, C = import_reference()
    rng = np.random.default_rng(20240607)
    out = {}

    def save(group, **arrs):
        np.savez_compressed(os.path.join(HERE, group + ".npz"), **arrs)
        out[group] = sum(np.asarray(v).nbytes for v in arrs.values())

    # ------------------------------------------------------------- grid
    Nx, Ny = 36, 28
    X, Y, dx, dy = F.create_grid(Nx, Ny, 1.2, 0.9)
    g = dict(X=X, Y=Y, dx=dx, dy=dy, Nx=Nx, Ny=Ny, Lx=1.2, Ly=0.9)

    # ------------------------------------------------------------- utils
    f = smooth(rng, Y, X, noise=0.05)
    w = smooth(rng, Y, X, noise=0.3)          # sign-changing "velocity"
    w[3, 5] = 0.0
    w[0, 0] = 0.0
    Am = rng.standard_normal((3, 3)); Am = Am @ Am.T + np.eye(3)
    bm = rng.standard_normal(3)
    As = np.ones((3, 3)); bs = np.ones(3)
    save("utils", f=f, w=w, dx=dx, dy=dy,
         gx2=U.grad_central_x_2nd(f, dx), gy2=U.grad_central_y_2nd(f, dy),
         gx4=U.grad_central_x_4th(f, dx), gy4=U.grad_central_y_4th(f, dy),
         up_x=U.diff_upwind_3rd(f, w, dx, 1), up_y=U.diff_upwind_3rd(f, w, dy, 0),
         lap=U.lap_2nd(f, dx, dy),
         A=Am, b=bm, x=U.fast_solve_3x3(Am, bm), As=As, bs=bs, xs=U.fast_solve_3x3(As, bs))

    # ------------------------------------------------------------- interpolators
    xq = rng.uniform(-0.2, 1.4, size=(17, 23))
    yq = rng.uniform(-0.2, 1.1, size=(17, 23))
    xq[0, 0] = np.nan; yq[1, 1] = np.inf; xq[2, 2] = -np.inf
    xq[3, 3] = 1e200; yq[4, 4] = -1e200
    xq[5, 5] = 1.2; yq[5, 5] = 0.9            # exactly the far corner
    xq[6, 6] = 7 * dx; yq[6, 6] = 9 * dy      # exactly on nodes
    save("interp", u=f, xq=xq, yq=yq, dx=dx, dy=dy, Nx=Nx, Ny=Ny,
         bil=I.bilinear_interpolate(f, xq, yq, dx, dy, Nx, Ny),
         bic=I.bicubic_interpolate(f, xq, yq, dx, dy, Nx, Ny),
         cc_in=np.array([0.3, -1.2, 2.5, 0.7, 0.37]),
         cc=np.array([I.cubic_convolution(0.3, -1.2, 2.5, 0.7, 0.37)]))

    # ------------------------------------------------------------- level set + maps
    phi = C.initialize_disc(X, Y, 0.62, 0.44, 0.27)
    solid = (phi <= 0).astype(float)
    # a deformed reference map (smooth, invertible-ish)
    X1 = (X + 0.03 * np.sin(2.2 * X) * np.cos(1.7 * Y)) * solid
    X2 = (Y + 0.02 * np.cos(1.3 * X) * np.sin(2.9 * Y)) * solid

    # ------------------------------------------------------------- extrapolation
    X1e, X2e = F.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
    # second case: body cut by the domain edge + a second body (window clipping, two components)
    phi2 = np.minimum(C.initialize_disc(X, Y, 0.05, 0.40, 0.22),
                      C.initialize_disc(X, Y, 0.90, 0.70, 0.15))
    m2 = (phi2 < 0).astype(float)
    Z1 = (1.3 * X + 0.2 * Y + 0.1 * X * Y) * m2
    Z2 = (-0.4 * X + 0.9 * Y - 0.05 * X * X) * m2
    Z1e, Z2e = F.extrapolate_reference_map(Z1, Z2, phi2, dx, dy, 4)
    # third: no solid at all, and max_layers larger than what the band needs on a tiny grid
    phi3 = np.ones_like(phi)
    W1e, W2e = F.extrapolate_reference_map(X1, X2, phi3, dx, dy, 2)
    save("extrap", dx=dx, dy=dy, phi=phi, X1=X1, X2=X2, X1e=X1e, X2e=X2e, layers=3,
         phi2=phi2, Z1=Z1, Z2=Z2, Z1e=Z1e, Z2e=Z2e, layers2=4,
         phi3=phi3, W1e=W1e, W2e=W2e, layers3=2)

    # ------------------------------------------------------------- advection
    a = smooth(rng, Y, X, amp=0.8)
    b = smooth(rng, Y, X, amp=0.8)
    a[10, 12] = 0.0
    b[10, 12] = 0.0
    dt = 0.2 * dx
    adv = dict(a=a, b=b, X=X, Y=Y, dt=dt, dx=dx, dy=dy, phi=phi, q=X1e, q2=X2e)
    for sch in ("semilagrangian", "semilagrangian_cubic", "central2", "weno5", "conservative"):
        adv["out_" + sch] = F.advect_reference_map(X1e, a, b, X, Y, dt, dx, dy, phi, sch, 0.0)
        adv["out2_" + sch] = F.advect_reference_map(X2e, a, b, X, Y, dt, dx, dy, phi, sch, 1.5 * dx)
    # whole-domain active (phi very negative) so rim fallbacks of WENO5 are exercised
    phin = -np.ones_like(phi)
    adv["full_weno5"] = F.advect_reference_map(f, a, b, X, Y, dt, dx, dy, phin, "weno5", 0.0)
    adv["full_central2"] = F.advect_reference_map(f, a, b, X, Y, dt, dx, dy, phin, "central2", 0.0)
    adv["full_conservative"] = F.advect_reference_map(f, a, b, X, Y, dt, dx, dy, phin, "conservative", 0.0)
    adv["full_q"] = f
    adv["rhs_weno5"] = F._weno5_rhs(f, a, b, dx, dy, phin, 0.0)
    # large dt: backtrace leaves the domain -> clamping
    adv["far_sl"] = F.advect_reference_map(f, a, b, X, Y, 40 * dt, dx, dy, phi, "semilagrangian", 0.0)
    adv["far_dt"] = 40 * dt
    save("advect", **adv)

    # ------------------------------------------------------------- stress / heaviside
    st = dict(X1=X1e, X2=X2e, phi=phi, dx=dx, dy=dy)
    for tag, kw in (("legacy", dict(mu_s=0.7, kappa=0.0)),
                    ("kappa", dict(mu_s=0.7, kappa=2.0)),
                    ("band", dict(mu_s=1.3, kappa=0.5, w_cut=2 * dx, detg_clamp=3.0)),
                    ("iso", dict(mu_s=1.1, kappa=0.4, isochoric=True))):
        r = F.solid_cauchy_stress(X1e, X2e, dx, dy, phi=phi, **kw)
        for nm, arr in zip(("sxx", "sxy", "syy", "J"), r):
            st[f"{tag}_{nm}"] = arr
    # strong compression with clamp (tests/test_stress.py:52-65)
    r = F.solid_cauchy_stress(10.0 * X, Y.copy(), dx, dy, 1.0, 0.0, phi, w_cut=2 * dx, detg_clamp=3.0)
    st["clamp_J"] = r[3]
    w_t = 2.0 * dx
    xs = np.concatenate([np.linspace(-3 * w_t, 3 * w_t, 61), [w_t, -w_t, 0.0]])
    st["hv_x"] = xs
    st["hv_w"] = w_t
    st["hv"] = F.smoothed_heaviside(xs, w_t)
    st["hv_phi"] = F.smoothed_heaviside(phi, w_t)
    save("stress", **st)

    # ------------------------------------------------------------- momentum
    u = smooth(rng, Y, X, amp=0.6)
    v = smooth(rng, Y, X, amp=0.6)
    u[7, 9] = 0.0
    p = smooth(rng, Y, X, amp=0.3)
    lid = lambda uu, vv: C.no_slip_lid_bc(uu, vv, 1.0)
    H = F.smoothed_heaviside(phi, w_t)
    rho_local = (1 - H) * 1.3 + H * 0.9
    exx, exy, eyy, _ = F.solid_cauchy_stress(X1e, X2e, dx, dy, 0.7, 0.0, phi)
    ru, rv = F.velocity_rhs_blended_optimized(
        u, v, p, exx, exy, eyy, dx, dy, phi, 0.02, H,
        U.grad_central_x_2nd(H, dx), U.grad_central_y_2nd(H, dy), rho_local, 0.0, 0.0)
    mom = dict(u=u, v=v, p=p, X1=X1e, X2=X2e, phi=phi, dx=dx, dy=dy, dt=dt, w_t=w_t,
               rhs_u=ru, rhs_v=rv, rhs_mu_f=0.02, rhs_rho_s=1.3, rhs_rho_f=0.9, rhs_mu_s=0.7)
    cases = {
        # name: (bc, mu_s, kappa, eta_s, rho_s, rho_f, mu_f, gamma, stress_band, detg_clamp)
        "lid_eta": ("lid", 0.1, 0.0, 0.01, 1.0, 1.0, 0.01, 0.0, False, 3.0),
        "lid_noeta": ("lid", 0.7, 0.3, 0.0, 1.3, 0.9, 0.02, 0.0, False, 3.0),
        "slip_eta": ("slip", 1.0, 0.0, 0.02, 1.0, 1.0, 1e-3, 0.0, False, 3.0),
        "lid_band": ("lid", 0.4, 0.2, 0.01, 1.1, 1.0, 0.01, 0.0, True, 3.0),
        "lid_gamma": ("lid", 0.1, 0.0, 0.01, 1.0, 1.0, 0.01, 0.05, False, 3.0),
        "wall_fluid": ("wall", 0.0, 0.0, 0.0, 0.0, 1.0, 0.01, 0.0, False, 3.0),
    }
    bcs = {"lid": lid, "slip": C.free_slip_box_bc,
           "wall": lambda uu, vv: (np.where(_rim(uu), 0.0, uu), np.where(_rim(vv), 0.0, vv))}
    for nm, (bcn, mu_s, kap, eta, rs, rf, muf, gam, band, clamp) in cases.items():
        ph = np.ones_like(phi) if nm == "wall_fluid" else phi
        r = F.momentum_step_rk4(u, v, p, X1e, X2e, bcs[bcn], mu_s, kap, eta, dx, dy, dt, rs, rf,
                                ph, muf, w_t, gam, stress_band=band, detg_clamp=clamp)
        mom[nm + "_prm"] = np.array([mu_s, kap, eta, rs, rf, muf, gam, float(band), clamp])
        for k, arr in zip(("un", "vn", "sxx", "sxy", "syy", "J"), r):
            mom[f"{nm}_{k}"] = arr
    mom["curv"] = F.compute_curvature(phi, dx, dy)
    save("momentum", **mom)

    # ------------------------------------------------------------- projection (Neumann)
    eig = F._precompute_poisson_eigenvalues(Nx, Ny, dx, dy)
    rhs = smooth(rng, Y, X, amp=5.0, noise=0.1)
    prj = dict(a=u, b=v, p=p, dx=dx, dy=dy, dt=dt, eig=eig, rhs=rhs, rho=rho_local * 0 + 1.0,
               sol=F._solve_poisson_dct(rhs, eig),
               div=F._compute_divergence(u, v, dx, dy),
               div_rc=F._compute_divergence_rc(u, v, p, dt, 1.0, dx, dy))
    gx, gy = F._compute_pressure_gradient(p, dx, dy)
    prj["gx"], prj["gy"] = gx, gy
    rho_arr = np.ones_like(u) + 1e-13 * rng.standard_normal(u.shape)   # "constant" to 1e-13
    prj["rho_arr"] = rho_arr
    r = F.pressure_projection_amg(u, v, dx, dy, dt, rho_arr, lid, p_prev=p, eigenvalues=eig)
    prj["A_a"], prj["A_b"], prj["A_p"] = r[:3]
    r = F.pressure_projection_amg(u, v, dx, dy, dt, 1.0, C.free_slip_box_bc, p_prev=None, eigenvalues=eig)
    prj["B_a"], prj["B_b"], prj["B_p"] = r[:3]
    r = F.pressure_projection_amg(u, v, dx, dy, dt, 0.8, lid, p_prev=p, eigenvalues=eig)
    prj["C_a"], prj["C_b"], prj["C_p"] = r[:3]
    save("projection", **prj)

    # power-of-two-plus-one grid (FFT path of the CUDA build): 33 x 17
    Xp, Yp, dxp, dyp = F.create_grid(33, 17, 1.0, 0.5)
    eigp = F._precompute_poisson_eigenvalues(33, 17, dxp, dyp)
    rp = smooth(rng, Yp, Xp, amp=3.0, noise=0.2)
    up = smooth(rng, Yp, Xp); vp = smooth(rng, Yp, Xp); pp = smooth(rng, Yp, Xp, amp=0.2)
    r = F.pressure_projection_amg(up, vp, dxp, dyp, 1e-3, 1.0, lid, p_prev=pp, eigenvalues=eigp)
    save("projection_pow2", dx=dxp, dy=dyp, eig=eigp, rhs=rp, sol=F._solve_poisson_dct(rp, eigp),
         a=up, b=vp, p=pp, dt=1e-3, A_a=r[0], A_b=r[1], A_p=r[2])

    # ------------------------------------------------------------- projection (periodic)
    def per_bc(uu, vv):
        uu = uu.copy(); vv = vv.copy()
        uu[:, -1] = uu[:, 0]; vv[:, -1] = vv[:, 0]
        uu[-1, :] = uu[0, :]; vv[-1, :] = vv[0, :]
        return uu, vv
    per = {}
    for tag, (nx, ny) in (("odd", (36, 28)), ("pow2", (33, 17))):
        Xq, Yq, dxq, dyq = F.create_grid(nx, ny, 1.0, 1.0)
        eg, nl = F._precompute_poisson_eigenvalues_periodic(nx, ny, dxq, dyq)
        ua, va = per_bc(smooth(rng, Yq, Xq), smooth(rng, Yq, Xq))
        pa = per_bc(smooth(rng, Yq, Xq, amp=0.2), smooth(rng, Yq, Xq))[0]
        rr = smooth(rng, Yq, Xq, amp=4.0, noise=0.1)
        gxp, gyp = F._compute_pressure_gradient_periodic(pa, dxq, dyq)
        r1 = F.pressure_projection_amg(ua, va, dxq, dyq, 2e-3, np.ones_like(ua), per_bc,
                                       p_prev=pa, eigenvalues=(eg, nl), bc_type='periodic')
        r2 = F.pressure_projection_amg(ua, va, dxq, dyq, 2e-3, 1.0, per_bc,
                                       p_prev=None, eigenvalues=None, bc_type='periodic')
        per.update({f"{tag}_dx": dxq, f"{tag}_dy": dyq, f"{tag}_eig": eg, f"{tag}_null": nl,
                    f"{tag}_a": ua, f"{tag}_b": va, f"{tag}_p": pa, f"{tag}_rhs": rr,
                    f"{tag}_sol": F._solve_poisson_fft(rr, (eg, nl)),
                    f"{tag}_div": F._compute_divergence_periodic(ua, va, dxq, dyq),
                    f"{tag}_gx": gxp, f"{tag}_gy": gyp,
                    f"{tag}_A_a": r1[0], f"{tag}_A_b": r1[1], f"{tag}_A_p": r1[2],
                    f"{tag}_B_a": r2[0], f"{tag}_B_b": r2[1], f"{tag}_B_p": r2[2]})
    per["dt"] = 2e-3
    save("periodic", **per)

    # ------------------------------------------------------------- time step
    ts = []
    for args in ((0.2, 1e-3, 0.1, 1.0, 0.0, 1.0, 0.01, 0.01, 0.0),
                 (0.2, 1e-2, 0.0, 0.0, 0.0, 1.0, 0.01, 0.0, 0.0),
                 (0.3, 1.0, 2.0, 1.2, 0.07, 0.8, 1e-3, 0.0, 1.5)):
        CFL, cap, mu_s, rho_s, gam, rho_f, mu_f, eta, kap = args
        ts.append(list(args) + [F.compute_timestep(u, v, dx, dy, CFL, cap, mu_s, rho_s, gam, rho_f,
                                                   mu_f=mu_f, eta_s=eta, kappa=kap)])
    phib = f.copy()
    save("timestep", u=u, v=v, dx=dx, dy=dy, table=np.array(ts),
         phibc_in=f, phibc_out=F.apply_phi_BCs(phib))

    # ------------------------------------------------------------- full FSI steps
    # config-2 physics on a 40x40 grid, three consecutive steps with each scheme;
    # every step stores its complete input state so each can be replayed alone.
    N = 40
    Xs, Ys, dxs, dys = F.create_grid(N, N, 1.0, 1.0)
    x0, y0, R = 0.6, 0.5, 0.2
    phi0 = lambda Xq, Yq: C.initialize_disc(Xq, Yq, x0, y0, R)
    eigs = F._precompute_poisson_eigenvalues(N, N, dxs, dys)
    prm = dict(mu_s=0.1, kappa=0.0, rho_s=1.0, eta_s=0.01, mu_f=0.01, rho_f=1.0, w_t=2 * dxs,
               layers=3, CFL=0.2, cap=1e-3)
    fs = dict(X=Xs, Y=Ys, dx=dxs, dy=dys, eig=eigs, disc=np.array([x0, y0, R]),
              prm=np.array([prm[k] for k in ("mu_s", "kappa", "rho_s", "eta_s", "mu_f", "rho_f",
                                             "w_t", "layers", "CFL", "cap")]))
    for sch in ("semilagrangian", "weno5", "central2"):
        phis = F.apply_phi_BCs(phi0(Xs, Ys))
        msk = (phis <= 0).astype(float)
        S1, S2 = F.extrapolate_reference_map(Xs * msk, Ys * msk, phis, dxs, dys, 3)
        aa = np.zeros((N, N)); bb = np.zeros((N, N)); pp_ = np.zeros((N, N))
        nsteps, keep = 60, (0, 30, 59)
        for n in range(nsteps):
            dts = F.compute_timestep(aa, bb, dxs, dys, 0.2, 1e-3, 0.1, 1.0, 0.0, 1.0,
                                     mu_f=0.01, eta_s=0.01, kappa=0.0)
            if n in keep:
                for nm, arr in (("a", aa), ("b", bb), ("p", pp_), ("X1", S1), ("X2", S2)):
                    fs[f"{sch}_{n}_in_{nm}"] = arr.copy()
                fs[f"{sch}_{n}_dt"] = dts
            phis = F.rebuild_phi_from_reference_map(S1, S2, phi0)
            msk = (phis <= 0).astype(float)
            S1 = F.advect_reference_map(S1, aa, bb, Xs, Ys, dts, dxs, dys, phis, sch, 0.0) * msk
            S2 = F.advect_reference_map(S2, aa, bb, Xs, Ys, dts, dxs, dys, phis, sch, 0.0) * msk
            S1, S2 = F.extrapolate_reference_map(S1, S2, phis, dxs, dys, 3)
            phis = F.rebuild_phi_from_reference_map(S1, S2, phi0)
            a_s, b_s, sxx, sxy, syy, J = F.momentum_step_rk4(
                aa, bb, pp_, S1, S2, lid, 0.1, 0.0, 0.01, dxs, dys, dts, 1.0, 1.0, phis, 0.01,
                2 * dxs, 0.0)
            Hs = F.smoothed_heaviside(phis, 2 * dxs)
            rl = (1 - Hs) * 1.0 + Hs * 1.0
            aa, bb, pp_, _, _ = F.pressure_projection_amg(a_s, b_s, dxs, dys, dts, rl, lid,
                                                          p_prev=pp_, eigenvalues=eigs)
            if n in keep:
                for nm, arr in (("a", aa), ("b", bb), ("p", pp_), ("X1", S1), ("X2", S2),
                                ("phi", phis), ("J", J), ("sxx", sxx)):
                    fs[f"{sch}_{n}_out_{nm}"] = arr.copy()
    save("fsi_steps", **fs)

    for k, nb in out.items():
        print(f"{k:18s} {nb / 1024:8.1f} KiB (uncompressed)")


def _rim(f):
    m = np.zeros(f.shape, dtype=bool)
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = True
    return m


if __name__ == "__main__":
    main()

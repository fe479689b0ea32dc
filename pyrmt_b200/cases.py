"""The three CPU-runnable validation cases of the reference (BASELINE.json configs[0..2]),
driven through the drop-in operator API on device-resident tensors.

Each function is the time loop of the corresponding upstream driver with the plotting
and file output removed:
  lid_driven_cavity      benchmarks/lid_driven_cavity.py:25-80
  soft_disc_in_lid       benchmarks/soft_disc_in_lid_driven.py:42-127
  disc_in_taylor_green   benchmarks/disc_in_taylor_green.py:39-117
Diagnostics (centroid, energies) are evaluated on the device with torch reductions; they
are off the timed path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import functions as F
from . import output as O
from .bc import free_slip_box_bc
from .driver import LidBC
from .levelset import DiscSDF


def _up(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def lid_driven_cavity(Re=100.0, N=129, max_steps=60000, steady_tol=2e-5, log=None):
    """Pure-fluid lid-driven cavity to steady state.  Returns dict(step, t, y, u_line)."""
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    U_lid, rho_f = 1.0, 1.0
    mu_f = rho_f * U_lid * 1.0 / Re
    mu_s = kappa = rho_s = eta_s = 0.0
    w_t = 2.0 * dx
    phi = torch.ones((N, N), dtype=torch.float64, device="cuda")
    X1, X2 = _up(X), _up(Y)
    bc = LidBC(U_lid)
    a0, b0 = bc(np.zeros((N, N)), np.zeros((N, N)))
    a, b, p = _up(a0), _up(b0), _up(np.zeros((N, N)))
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    t, step = 0.0, 0
    for step in range(1, max_steps + 1):
        dt = F.compute_timestep(a, b, dx, dy, 0.2, 1e-2, mu_s, rho_s, 0.0, rho_f, mu_f=mu_f)
        a_prev = a
        a_s, b_s, *_ = F.momentum_step_rk4(a, b, p, X1, X2, bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f,
                                          phi, mu_f, w_t, 0.0)
        a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_f, bc, p_prev=p, eigenvalues=eig)
        if step % 200 == 0 or step == 1:
            res = float((a - a_prev).abs().max().item()) / dt
            if log:
                log("  step %6d  t=%7.3f  resid=%.2e" % (step, t + dt, res))
            if step > 1 and res < steady_tol:
                break
        t += dt
    i_mid = N // 2
    return dict(step=step, t=t, y=Y[:, i_mid].copy(), u_line=a[:, i_mid].cpu().numpy(),
                a=a, b=b, p=p)


def ghia_rms(y, u_line, ghia_csv):
    gd = np.loadtxt(ghia_csv, delimiter=",", skiprows=1)
    yg, ug = gd[:, 0], gd[:, 1]
    return float(np.sqrt(np.mean((np.interp(yg, y, u_line) - ug) ** 2)))


def _disc_case(N, x0, y0, R):
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    sdf = DiscSDF([x0], [y0], [R])
    Xd, Yd = _up(X), _up(Y)
    phi = F.apply_phi_BCs(sdf(Xd, Yd))
    X1, X2 = F.mask_solid(Xd, phi), F.mask_solid(Yd, phi)
    return X, Y, dx, dy, sdf, Xd, Yd, phi, X1, X2


def _centroid(phi, Xd, Yd):
    from .output import disc_centroid
    return disc_centroid(phi, Xd, Yd)


def soft_disc_in_lid(N=128, scheme="semilagrangian", t_end=8.0, sample_times=(1, 2, 3, 4, 5, 6, 7, 8),
                     log=None):
    """Soft neo-Hookean disc in a lid-driven cavity (Sugiyama benchmark).  Returns the centroid
    at the first step with t >= each sample time, and the extents of the whole trajectory."""
    X, Y, dx, dy, sdf, Xd, Yd, phi, X1, X2 = _disc_case(N, 0.6, 0.5, 0.2)
    bc = LidBC(1.0)
    mu_s, kappa, rho_s, eta_s, mu_f, rho_f = 0.1, 0.0, 1.0, 0.01, 0.01, 1.0
    w_t = 2.0 * dx
    layers = max(3, int(np.ceil(w_t / dx)) + 1)
    X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
    z = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    a, b, p = z.clone(), z.clone(), z.clone()
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    t, step, k = 0.0, 0, 0
    samples, xs, ys, Jmin, Jmax = [], [], [], np.inf, -np.inf
    while t < t_end:
        step += 1
        dt = F.compute_timestep(a, b, dx, dy, 0.2, 1e-3, mu_s, rho_s, 0.0, rho_f, mu_f=mu_f, eta_s=eta_s,
                                kappa=kappa)
        phi = F.rebuild_phi_from_reference_map(X1, X2, sdf)
        X1 = F.mask_solid(F.advect_reference_map(X1, a, b, Xd, Yd, dt, dx, dy, phi, scheme, 0.0), phi)
        X2 = F.mask_solid(F.advect_reference_map(X2, a, b, Xd, Yd, dt, dx, dy, phi, scheme, 0.0), phi)
        X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
        phi = F.rebuild_phi_from_reference_map(X1, X2, sdf)
        a_s, b_s, sxx, sxy, syy, J = F.momentum_step_rk4(a, b, p, X1, X2, bc, mu_s, kappa, eta_s, dx, dy, dt,
                                                         rho_s, rho_f, phi, mu_f, w_t, 0.0)
        _, rho_local = F.heaviside_and_density(phi, w_t, rho_s, rho_f)
        a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, bc, p_prev=p,
                                                  eigenvalues=eig)
        t += dt
        if step % 20 == 0 or (k < len(sample_times) and t >= sample_times[k]) or t >= t_end:
            cx, cy = _centroid(phi, Xd, Yd)
            xs.append(cx); ys.append(cy)
            Jmin, Jmax = min(Jmin, float(J.min().item())), max(Jmax, float(J.max().item()))
            while k < len(sample_times) and t >= sample_times[k]:
                samples.append((sample_times[k], t, cx, cy))
                k += 1
                if log:
                    log("  step %6d t=%6.3f centroid=(%.4f, %.4f)" % (step, t, cx, cy))
    return dict(steps=step, samples=samples, x_range=(min(xs), max(xs)), y_range=(min(ys), max(ys)),
                minJ=Jmin, maxJ=Jmax)


def disc_in_taylor_green(N=128, scheme="semilagrangian", t_end=1.0, log=None):
    """Soft disc in a Taylor-Green vortex with free-slip walls (the reference's own pairing).
    Returns the energy history summary: E0, E1, drift %, KE/SE at the end."""
    X, Y, dx, dy, sdf, Xd, Yd, phi, X1, X2 = _disc_case(N, 0.5, 0.5, 0.2)
    mu_s, kappa, rho_s, eta_s, mu_f, rho_f = 1.0, 0.0, 1.0, 0.0, 1.0e-3, 1.0
    w_t = 2.0 * dx
    layers = max(3, int(np.ceil(w_t / dx)) + 1)
    X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
    k = 2.0 * np.pi
    a0 = 0.05 * k * np.sin(k * X) * np.cos(k * Y)
    b0 = -0.05 * k * np.cos(k * X) * np.sin(k * Y)
    a0, b0 = free_slip_box_bc(a0, b0)
    a, b, p = _up(a0), _up(b0), _up(np.zeros((N, N)))
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    t, step, integ = 0.0, 0, 0.0
    E0 = E1 = ke = se = None
    while t < t_end:
        step += 1
        dt = F.compute_timestep(a, b, dx, dy, 0.2, 1e-4, mu_s, rho_s, 0.0, rho_f, mu_f=mu_f, eta_s=eta_s,
                                kappa=kappa)
        if t + dt > t_end:
            dt = t_end - t
        phi = F.rebuild_phi_from_reference_map(X1, X2, sdf)
        X1 = F.mask_solid(F.advect_reference_map(X1, a, b, Xd, Yd, dt, dx, dy, phi, scheme, 0.0), phi)
        X2 = F.mask_solid(F.advect_reference_map(X2, a, b, Xd, Yd, dt, dx, dy, phi, scheme, 0.0), phi)
        X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, layers)
        phi = F.rebuild_phi_from_reference_map(X1, X2, sdf)
        a_s, b_s, sxx, sxy, syy, J = F.momentum_step_rk4(a, b, p, X1, X2, free_slip_box_bc, mu_s, kappa, eta_s,
                                                         dx, dy, dt, rho_s, rho_f, phi, mu_f, w_t, gamma=0.0)
        _, rho_local = F.heaviside_and_density(phi, w_t, rho_s, rho_f)
        a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, free_slip_box_bc, p_prev=p,
                                                  eigenvalues=eig)
        diss = O.compute_viscous_dissipation(a, b, mu_f, phi, w_t, dx, dy, eta_s)
        integ += diss * dt
        t += dt
        if step == 1 or t >= t_end or step % 500 == 0:
            ke = O.compute_kinetic_energy(a, b, rho_f, rho_s, phi, w_t, dx, dy)
            se = O.compute_strain_energy(X1, X2, phi, mu_s, dx, dy, kappa=kappa)
            E1 = ke + se + integ
            if step == 1:
                E0 = E1
            if log:
                log("  step %5d t=%5.3f KE=%.4e SE=%.4e E=%.4e" % (step, t, ke, se, E1))
    return dict(steps=step, E0=E0, E1=E1, drift_pct=(E1 - E0) / max(abs(E0), 1e-30) * 100.0, KE=ke, SE=se,
                integrated_dissipation=integ)

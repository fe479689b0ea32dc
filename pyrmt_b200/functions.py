"""Drop-in operator API of the collocated RMT timestep -- same names, positional
orders, defaults and error behaviour as upstream ``pyRMT/functions.py`` -- with
every array operation executed by the sm_100a kernels of librmt_b200.so.

Array type policy: type-preserving.  NumPy arrays in -> NumPy arrays
out (host<->device copies inside the call; this is how the reference's
``benchmarks/*.py`` run unchanged); fp64 CUDA tensors in -> CUDA tensors out
(zero-copy; what the 4096^2 / 16384^2 drivers use).  Inputs are never modified
(except ``apply_phi_BCs``, in place upstream too).  There is no CPU fallback:
without the CUDA library or a device every operator raises.

Each function cites the upstream lines it replaces (paths relative to the
reference repository root).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._runtime import (F64, ctx, finite_cache, is_np, profiler, ptr, shape2, stream, to_dev, to_user)
from .bc import apply_bc_
from .interpolators import bilinear_interpolate, bicubic_interpolate  # noqa: F401 (re-export)
from .utils import (grad_central_x_2nd, grad_central_y_2nd, grad_central_x_4th,  # noqa: F401
                    grad_central_y_4th, diff_upwind_3rd, lap_2nd, fast_solve_3x3)

_SCHEMES = {"central2": 0, "weno5": 1, "conservative": 2}


def _chk(code, what):
    _lib.check(code, what)


# --------------------------------------------------------------------------
# grid, level set, time step
# --------------------------------------------------------------------------
def create_grid(Nx, Ny, Lx, Ly):
    """pyRMT/functions.py:25-31 -- node-based linspace grid (host arrays, as upstream)."""
    x = np.linspace(0, Lx, Nx)
    y = np.linspace(0, Ly, Ny)
    X, Y = np.meshgrid(x, y)
    return X, Y, x[1] - x[0], y[1] - y[0]


def apply_phi_BCs(phi):
    """pyRMT/functions.py:33-46 -- in-place 3-cell periodic wrap (setup only)."""
    if isinstance(phi, torch.Tensor):
        phi[0:3, :] = phi[-6:-3, :].clone()
        phi[-3:, :] = phi[3:6, :].clone()
        phi[:, 0:3] = phi[:, -6:-3].clone()
        phi[:, -3:] = phi[:, 3:6].clone()
        return phi
    phi[0:3, :] = phi[-6:-3, :]
    phi[-3:, :] = phi[3:6, :]
    phi[:, 0:3] = phi[:, -6:-3]
    phi[:, -3:] = phi[:, 3:6]
    return phi


def rebuild_phi_from_reference_map(X1, X2, phi_init_func):
    """pyRMT/functions.py:1366-1367 -- calls the user's phi0(X1, X2).  Use
    ``pyrmt_b200.levelset.DiscSDF`` for a device-side disc family."""
    return phi_init_func(X1, X2)


def compute_timestep(a, b, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f=0.0, eta_s=0.0,
                     kappa=0.0):
    """pyRMT/functions.py:165-192.  The max|u| reduction runs on the device; the
    result is a Python float (one host sync, as the loop's control flow needs)."""
    c = ctx()
    ad, bd = to_dev(a), to_dev(b)
    out = c.max_speed(ad, bd).cpu()
    speed, bad = float(out[0]), float(out[1])
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        finite_cache.put(a, b, bad == 0.0)
    if bad:
        speed = float("nan")
    return timestep_from_speed(speed, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f, eta_s, kappa)


def compute_timestep_begin(a, b):
    """First half of `compute_timestep` for device tensors: the reduction is queued and its result travels to
    pinned host memory asynchronously.  Kernels that do not need dt (the level set of the current map) can be
    queued before `compute_timestep_end` waits for it, so the GPU is not idle during the host round trip."""
    host, ev = ctx().max_speed_async(to_dev(a), to_dev(b))
    return (a, b, host, ev)


def compute_timestep_end(handle, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f=0.0, eta_s=0.0, kappa=0.0):
    a, b, host, ev = handle
    ev.synchronize()
    speed, bad = float(host[0]), float(host[1])
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        finite_cache.put(a, b, bad == 0.0)
    if bad:
        speed = float("nan")
    return timestep_from_speed(speed, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f, eta_s, kappa)


def timestep_from_speed(speed, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f=0.0, eta_s=0.0, kappa=0.0):
    """The host formula of compute_timestep (functions.py:177-192) for a given max sqrt(a^2 + b^2)
    (a slab-decomposed run reduces the speed over the ranks first)."""
    wave = np.sqrt((kappa + mu_s * 4.0 / 3.0) / (rho_s + 1e-12))
    dt_solid = CFL * dx / (wave + 1e-14)
    dt_fluid = CFL * dx / (speed + 1e-6)
    dt_cap = 1.0
    if gamma > 1e-12:
        rho_avg = 0.5 * (rho_s + rho_f)
        dt_cap = np.sqrt((rho_avg * dx**3) / (2 * np.pi * gamma)) * 0.5
    dt_visc = 1.0
    mu_max = max(mu_f, eta_s)
    rho_min = min(rho_s, rho_f)
    if mu_max > 1e-12 and rho_min > 1e-12:
        dt_visc = CFL * rho_min * dx**2 / (4.0 * mu_max)
    return min(dt_solid, dt_fluid, dt_cap, dt_visc, dt_min_cap)


# --------------------------------------------------------------------------
# reference-map advection
# --------------------------------------------------------------------------
def _sl(q, a, b, X, Y, dt, dx, dy, cubic):
    as_np = is_np(q)
    qd, ad, bd, Xd, Yd = (to_dev(t) for t in (q, a, b, X, Y))
    Ny, Nx = shape2(qd)
    out = torch.empty_like(qd)
    _chk(ctx().lib.rmt_advect_sl_rk4(ptr(qd), None, ptr(ad), ptr(bd), ptr(Xd), ptr(Yd), ptr(out), None,
                                     Ny, Nx, float(dt), float(dx), float(dy), cubic, stream()),
         "rmt_advect_sl_rk4")
    return to_user(out, as_np)


def advect_semilagrangian_rk4(q, a, b, X, Y, dt, dx, dy):
    """pyRMT/functions.py:194-227 (RK4 backtrace + bilinear sample)."""
    return _sl(q, a, b, X, Y, dt, dx, dy, 0)


def advect_semilagrangian_cubic_rk4(q, a, b, X, Y, dt, dx, dy):
    """pyRMT/functions.py:230-251 (RK4 backtrace + monotone bicubic sample)."""
    return _sl(q, a, b, X, Y, dt, dx, dy, 1)


def advect_semilagrangian_pair(q0, q1, a, b, X, Y, dt, dx, dy, cubic=False, slab=None):
    """Both reference-map components with ONE shared backtrace (the reference
    recomputes it per component, soft_disc_in_lid_driven.py:88-91).  Results are
    identical to two advect_semilagrangian_rk4 calls.  ``slab=(Ny_global, row_offset)``: the
    arrays are a row slab of a larger grid (pyrmt_b200/slab.py)."""
    as_np = is_np(q0)
    q0d, q1d, ad, bd, Xd, Yd = (to_dev(t) for t in (q0, q1, a, b, X, Y))
    Ny, Nx = shape2(q0d)
    Nyg, joff = (Ny, 0) if slab is None else (int(slab[0]), int(slab[1]))
    o0, o1 = torch.empty_like(q0d), torch.empty_like(q1d)
    _chk(ctx().lib.rmt_advect_sl_rk4_rows(ptr(q0d), ptr(q1d), ptr(ad), ptr(bd), ptr(Xd), ptr(Yd), ptr(o0),
                                          ptr(o1), Ny, Nx, Nyg, joff, float(dt), float(dx), float(dy),
                                          int(bool(cubic)), stream()), "rmt_advect_sl_rk4_rows")
    return to_user(o0, as_np), to_user(o1, as_np)


def _euler_rk3(q, a, b, dx, dy, dt, phi, w_cut, scheme):
    as_np = is_np(q)
    qd, ad, bd, pd = (to_dev(t) for t in (q, a, b, phi))
    Ny, Nx = shape2(qd)
    out, w1, w2 = torch.empty_like(qd), torch.empty_like(qd), torch.empty_like(qd)
    _chk(ctx().lib.rmt_advect_euler_rk3(ptr(qd), ptr(ad), ptr(bd), ptr(pd), ptr(out), ptr(w1), ptr(w2),
                                        Ny, Nx, float(dx), float(dy), float(dt), float(w_cut), scheme,
                                        stream()), "rmt_advect_euler_rk3")
    return to_user(out, as_np)


def advect_weno5_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):
    """pyRMT/functions.py:396-415."""
    return _euler_rk3(q, a, b, dx, dy, dt, phi, w_cut, 1)


def advect_central2_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):
    """pyRMT/functions.py:443-459."""
    return _euler_rk3(q, a, b, dx, dy, dt, phi, w_cut, 0)


def advect_conservative_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):
    """pyRMT/functions.py:489-496."""
    return _euler_rk3(q, a, b, dx, dy, dt, phi, w_cut, 2)


def advect_reference_map_pair(q0, q1, a, b, X, Y, dt, dx, dy, phi, scheme='semilagrangian', w_cut=0.0,
                              mask_solid=False, slab=None):
    """Both reference-map components in one call -- bitwise identical to two
    ``advect_reference_map`` calls (the drivers advect xi1 and xi2 with the same a, b, phi, dt:
    benchmarks/soft_disc_in_lid_driven.py:88-91) but a, b, phi are read once per stage and the
    semi-Lagrangian backtrace is shared.  ``mask_solid=True`` also applies the driver's
    ``* (phi <= 0)`` to both results."""
    if not _velocity_is_finite(a, b):
        raise FloatingPointError(
            "advect_reference_map: non-finite velocity (the simulation diverged)")
    if scheme in ('semilagrangian', 'semilagrangian_cubic'):
        r0, r1 = advect_semilagrangian_pair(q0, q1, a, b, X, Y, dt, dx, dy, cubic=scheme.endswith('cubic'),
                                            slab=slab)
        return (mask_solid_(r0, phi), mask_solid_(r1, phi)) if mask_solid else (r0, r1)
    if scheme not in _SCHEMES:
        raise ValueError("Unknown advection scheme %r (expected 'semilagrangian', "
                         "'central2', 'weno5' or 'conservative')" % (scheme,))
    as_np = is_np(q0)
    q0d, q1d, ad, bd, pd = (to_dev(t) for t in (q0, q1, a, b, phi))
    Ny, Nx = shape2(q0d)
    o0, o1 = torch.empty_like(q0d), torch.empty_like(q0d)
    work = torch.empty((4, Ny, Nx), dtype=F64, device=q0d.device)
    _chk(ctx().lib.rmt_advect_euler_rk3_pair(ptr(q0d), ptr(q1d), ptr(ad), ptr(bd), ptr(pd), ptr(o0), ptr(o1),
                                             ptr(work), Ny, Nx, float(dx), float(dy), float(dt), float(w_cut),
                                             _SCHEMES[scheme], 1 if mask_solid else 0, stream()),
         "rmt_advect_euler_rk3_pair")
    return to_user(o0, as_np), to_user(o1, as_np)


def _euler_rhs(q, a, b, dx, dy, phi, w_cut, scheme):
    as_np = is_np(q)
    qd, ad, bd, pd = (to_dev(t) for t in (q, a, b, phi))
    Ny, Nx = shape2(qd)
    out = torch.empty_like(qd)
    _chk(ctx().lib.rmt_euler_rhs(ptr(qd), ptr(ad), ptr(bd), ptr(pd), ptr(out), Ny, Nx, float(dx),
                                 float(dy), float(w_cut), scheme, stream()), "rmt_euler_rhs")
    return to_user(out, as_np)


def _weno5_rhs(q, a, b, dx, dy, phi, w_cut):
    """pyRMT/functions.py:321-393."""
    return _euler_rhs(q, a, b, dx, dy, phi, w_cut, 1)


def _central2_rhs(q, a, b, dx, dy, phi, w_cut):
    """pyRMT/functions.py:420-440."""
    return _euler_rhs(q, a, b, dx, dy, phi, w_cut, 0)


def _conservative_rhs(q, a, b, dx, dy, phi, w_cut):
    """pyRMT/functions.py:462-486."""
    return _euler_rhs(q, a, b, dx, dy, phi, w_cut, 2)


def _velocity_is_finite(a, b):
    hit = finite_cache.get(a, b) if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) else None
    if hit is not None:
        return hit
    ad, bd = to_dev(a), to_dev(b)
    bad = float(ctx().max_speed(ad, bd)[1].item())
    ok = bad == 0.0
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        finite_cache.put(a, b, ok)
    return ok


def advect_reference_map(q, a, b, X, Y, dt, dx, dy, phi, scheme='semilagrangian', w_cut=0.0):
    """pyRMT/functions.py:501-542 -- finite guard + scheme dispatch."""
    if not _velocity_is_finite(a, b):
        raise FloatingPointError(
            "advect_reference_map: non-finite velocity (the simulation diverged)")
    if scheme == 'semilagrangian':
        return advect_semilagrangian_rk4(q, a, b, X, Y, dt, dx, dy)
    if scheme == 'semilagrangian_cubic':
        return advect_semilagrangian_cubic_rk4(q, a, b, X, Y, dt, dx, dy)
    if scheme == 'central2':
        return advect_central2_rk3(q, a, b, dx, dy, dt, phi, w_cut)
    if scheme == 'weno5':
        return advect_weno5_rk3(q, a, b, dx, dy, dt, phi, w_cut)
    if scheme == 'conservative':
        return advect_conservative_rk3(q, a, b, dx, dy, dt, phi, w_cut)
    raise ValueError("Unknown advection scheme %r (expected 'semilagrangian', "
                     "'central2', 'weno5' or 'conservative')" % (scheme,))


# --------------------------------------------------------------------------
# narrow-band extrapolation
# --------------------------------------------------------------------------
def extrapolate_reference_map(X1, X2, phi, dx, dy, max_layers, row_offset=0, inplace=False):
    """pyRMT/functions.py:48-163 -- serial-order-faithful least-squares extrapolation.
    ``row_offset`` (not upstream; default 0): the arrays are rows [row_offset, ...) of a taller
    grid -- the fit then uses the global y coordinates, which the slab decomposition needs.
    ``inplace`` (not upstream; device tensors only): X1, X2 are temporaries of the caller and are
    extrapolated in place instead of into the copies of functions.py:69-70 (saves one pass over both)."""
    as_np = is_np(X1)
    x1, x2, ph = to_dev(X1), to_dev(X2), to_dev(phi)
    Ny, Nx = shape2(x1)
    c = ctx()
    if inplace and not as_np and x1 is X1 and x2 is X2 and x1.data_ptr() != x2.data_ptr():
        o1, o2 = x1, x2
    else:
        o1, o2 = torch.empty_like(x1), torch.empty_like(x2)
    ws = c.extrap_workspace(Ny, Nx)
    profiler.launches += 5 * int(max_layers)
    _chk(c.lib.rmt_extrapolate_rows(ptr(x1), ptr(x2), ptr(ph), ptr(o1), ptr(o2), Ny, Nx, int(row_offset),
                                    float(dx), float(dy), int(max_layers), ptr(ws), stream()),
         "rmt_extrapolate_rows")
    return to_user(o1, as_np), to_user(o2, as_np)


_EXT_VARIANTS = {"auto": -1, "per_layer": 0, "body": 1, "fused8": 8, "fused16": 16}


def _extrapolate_set_mode(variant="auto", rows=0, cap=0):
    """Not upstream: pick the sweep variant of extrapolate_reference_map (all are bit-identical to the
    reference's serial sweep).  'auto' = chosen on the device per call; 'per_layer', 'fused16', 'fused8',
    'body' force one (a variant that cannot run on the given bands falls back to the device's choice).
    ``cap`` > 0 limits the prepared-record capacity of the per-layer sweep (exercises inline phase A)."""
    v = _EXT_VARIANTS[variant] if isinstance(variant, str) else int(variant)
    _chk(ctx().lib.rmt_extrapolate_set_mode(v, int(rows), int(cap)), "rmt_extrapolate_set_mode")


def _extrapolate_last_mode(Ny, Nx):
    """(variant, rows) the last extrapolate_reference_map call on an (Ny, Nx) grid ran with."""
    import ctypes as C
    c = ctx()
    out = (C.c_int * 3)()
    _chk(c.lib.rmt_extrapolate_last_mode(ptr(c.extrap_workspace(Ny, Nx)), Ny, Nx, out, stream()),
         "rmt_extrapolate_last_mode")
    names = {v: k for k, v in _EXT_VARIANTS.items()}
    return names.get(out[0], out[0]), int(out[1])


# --------------------------------------------------------------------------
# stress, Heaviside, momentum predictor
# --------------------------------------------------------------------------
def solid_cauchy_stress(X1, X2, dx, dy, mu_s, kappa, phi, w_cut=0.0, detg_clamp=0.0, isochoric=False):
    """pyRMT/functions.py:545-658."""
    as_np = is_np(X1)
    x1, x2, ph = to_dev(X1), to_dev(X2), to_dev(phi)
    Ny, Nx = shape2(x1)
    sxx, sxy, syy, J = (torch.empty_like(x1) for _ in range(4))
    _chk(ctx().lib.rmt_solid_stress(ptr(x1), ptr(x2), ptr(ph), ptr(sxx), ptr(sxy), ptr(syy), ptr(J), Ny, Nx,
                                    float(dx), float(dy), float(mu_s), float(kappa), float(w_cut),
                                    float(detg_clamp), 1 if isochoric else 0, stream()),
         "rmt_solid_stress")
    return tuple(to_user(t, as_np) for t in (sxx, sxy, syy, J))


def smoothed_heaviside(x, w_t):
    """pyRMT/functions.py:660-671 (the sin form)."""
    as_np = is_np(x)
    xd = to_dev(x)
    H = torch.empty_like(xd)
    _chk(ctx().lib.rmt_heaviside(ptr(xd), ptr(H), xd.numel(), float(w_t), stream()), "rmt_heaviside")
    return to_user(H, as_np)


def heaviside_and_density(phi, w_t, rho_s, rho_f):
    """H = smoothed_heaviside(phi, w_t) and rho_local = (1-H)*rho_s + H*rho_f in one
    pass -- the two lines of glue every driver runs before the projection
    (benchmarks/soft_disc_in_lid_driven.py:102-103)."""
    as_np = is_np(phi)
    pd = to_dev(phi)
    H, rho = torch.empty_like(pd), torch.empty_like(pd)
    _chk(ctx().lib.rmt_heaviside_rho(ptr(pd), ptr(H), ptr(rho), pd.numel(), float(w_t), float(rho_s),
                                     float(rho_f), stream()), "rmt_heaviside_rho")
    return to_user(H, as_np), to_user(rho, as_np)


def rebuild_phi_and_stress(X1, X2, phi_init, dx, dy, mu_s, kappa, w_t, stress_band=False, detg_clamp=3.0):
    """``rebuild_phi_from_reference_map`` (functions.py:1366-1367) followed by the ``solid_cauchy_stress`` call
    of ``momentum_step_rk4`` (:693-700) -- the drivers always run the two back to back on the extrapolated map.
    With a ``DiscSDF`` level set and CUDA tensors this is ONE kernel (xi read once, phi not re-read);
    anything else falls back to the two operators.  -> (phi, (sxx, sxy, syy, J))."""
    from .levelset import DiscSDF
    w_cut_stress = float(w_t) if stress_band else 0.0
    clamp = float(detg_clamp) if stress_band else 0.0
    if isinstance(phi_init, DiscSDF) and not is_np(X1):
        phi, sxx, sxy, syy, J = phi_init.with_stress(X1, X2, dx, dy, mu_s, kappa, w_cut_stress, clamp)
        return phi, (sxx, sxy, syy, J)
    phi = rebuild_phi_from_reference_map(X1, X2, phi_init)
    return phi, solid_cauchy_stress(X1, X2, dx, dy, mu_s, kappa, phi, w_cut=w_cut_stress, detg_clamp=clamp)


def mask_solid_(q, phi):
    return mask_solid(q, phi)


def mask_solid(q, phi):
    """q * (phi <= 0) -- the ``* solid_mask`` glue of soft_disc_in_lid_driven.py:88-91."""
    as_np = is_np(q)
    qd, pd = to_dev(q), to_dev(phi)
    out = torch.empty_like(qd)
    _chk(ctx().lib.rmt_mask_mul(ptr(qd), ptr(pd), ptr(out), qd.numel(), stream()), "rmt_mask_mul")
    return to_user(out, as_np)


def compute_curvature(phi, dx, dy):
    """pyRMT/functions.py:837-861."""
    as_np = is_np(phi)
    pd = to_dev(phi)
    Ny, Nx = shape2(pd)
    out = torch.empty_like(pd)
    _chk(ctx().lib.rmt_curvature(ptr(pd), ptr(out), Ny, Nx, float(dx), float(dy), stream()),
         "rmt_curvature")
    return to_user(out, as_np)


def apply_velocity_BCs(bc, u, v):
    """pyRMT/functions.py:946-947."""
    if is_np(u):
        return bc(u, v)
    uu, vv = u.clone(), v.clone()
    return apply_bc_(bc, uu, vv)


def _field_or_none(x, like):
    """st_force arguments may be arrays or the scalar 0.0 upstream (functions.py:706-707)."""
    if isinstance(x, (int, float)):
        if x == 0:
            return None
        return torch.full_like(like, float(x))
    return to_dev(x)


def velocity_rhs_blended_optimized(u, v, p, sigma_sxx, sigma_sxy, sigma_syy, dx, dy, phi, mu_f, H, dH_dx,
                                   dH_dy, rho_local, st_force_x, st_force_y):
    """pyRMT/functions.py:897-944 -- one fused kernel instead of ~15 NumPy temporaries.
    (phi, dH_dx, dH_dy are accepted and unused, as upstream.)"""
    as_np = is_np(u)
    ud, vd, pd, sxx, sxy, syy, Hd, rd = (to_dev(t) for t in
                                         (u, v, p, sigma_sxx, sigma_sxy, sigma_syy, H, rho_local))
    Ny, Nx = shape2(ud)
    fx, fy = _field_or_none(st_force_x, ud), _field_or_none(st_force_y, ud)
    ru, rv = torch.empty_like(ud), torch.empty_like(ud)
    _chk(ctx().lib.rmt_velocity_rhs(ptr(ud), ptr(vd), ptr(pd), ptr(sxx), ptr(sxy), ptr(syy), ptr(Hd), ptr(rd),
                                    ptr(fx), ptr(fy), ptr(ru), ptr(rv), Ny, Nx, float(dx), float(dy),
                                    float(mu_f), stream()), "rmt_velocity_rhs")
    return to_user(ru, as_np), to_user(rv, as_np)


def momentum_step_rk4(u, v, p, X1, X2, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f, phi, mu_f,
                      w_t, gamma=0.0, stress_band=False, detg_clamp=3.0):
    """pyRMT/functions.py:673-762 -- classical RK4 momentum predictor (see ``momentum_step_rk4_with_stress``)."""
    return momentum_step_rk4_with_stress(u, v, p, X1, X2, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f,
                                         phi, mu_f, w_t, gamma, stress_band, detg_clamp, None)


def momentum_step_rk4_with_stress(u, v, p, X1, X2, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f, phi,
                                  mu_f, w_t, gamma=0.0, stress_band=False, detg_clamp=3.0, stress=None):
    """pyRMT/functions.py:673-762 -- classical RK4 momentum predictor.

    One stress kernel, then four fused stage kernels (RHS + Kelvin-Voigt term +
    stage update) with the BC table applied between them; H and rho_local are
    evaluated from phi inside the stage kernel.  ``stress`` (not upstream): the (sxx, sxy, syy, J) of
    ``rebuild_phi_and_stress`` for the same (X1, X2, phi, stress_band, detg_clamp), to skip the stress kernel."""
    as_np = is_np(u)
    ud, vd, pd, x1, x2, ph = (to_dev(t) for t in (u, v, p, X1, X2, phi))
    Ny, Nx = shape2(ud)
    lib = ctx().lib
    st = stream()
    dx, dy, dt = float(dx), float(dy), float(dt)

    w_cut_stress = float(w_t) if stress_band else 0.0
    clamp = float(detg_clamp) if stress_band else 0.0
    if stress is not None:
        sxx, sxy, syy, J = (to_dev(t) for t in stress)
    else:
        sxx, sxy, syy, J = (torch.empty_like(ud) for _ in range(4))
        _chk(lib.rmt_solid_stress(ptr(x1), ptr(x2), ptr(ph), ptr(sxx), ptr(sxy), ptr(syy), ptr(J), Ny, Nx, dx, dy,
                                  float(mu_s), float(kappa), w_cut_stress, clamp, 0, st), "rmt_solid_stress")
    fsx = fsy = None
    if gamma > 1e-12:
        fsx, fsy = torch.empty_like(ud), torch.empty_like(ud)
        _chk(lib.rmt_surface_tension(ptr(ph), ptr(fsx), ptr(fsy), Ny, Nx, dx, dy, float(w_t), float(gamma),
                                     st), "rmt_surface_tension")

    # stage-1 state: BC applied to a private copy (the update uses the raw u, v)
    sa_u, sa_v = apply_bc_(velocity_bc, ud.clone(), vd.clone())
    sb_u, sb_v = torch.empty_like(ud), torch.empty_like(ud)
    acc_u, acc_v = torch.empty_like(ud), torch.empty_like(ud)
    un, vn = torch.empty_like(ud), torch.empty_like(ud)

    def stage(k, iu, iv, ou, ov):
        _chk(lib.rmt_momentum_stage(ptr(iu), ptr(iv), ptr(pd), ptr(sxx), ptr(sxy), ptr(syy), ptr(ph),
                                    ptr(fsx), ptr(fsy), ptr(ud), ptr(vd), ptr(acc_u), ptr(acc_v),
                                    ptr(ou), ptr(ov), Ny, Nx, dx, dy, dt, float(mu_f), float(eta_s),
                                    float(w_t), float(rho_s), float(rho_f), k, st), "rmt_momentum_stage")
        return apply_bc_(velocity_bc, ou, ov)

    sb_u, sb_v = stage(1, sa_u, sa_v, sb_u, sb_v)
    sa_u, sa_v = stage(2, sb_u, sb_v, sa_u, sa_v)
    sb_u, sb_v = stage(3, sa_u, sa_v, sb_u, sb_v)
    un, vn = stage(4, sb_u, sb_v, un, vn)
    return tuple(to_user(t, as_np) for t in (un, vn, sxx, sxy, syy, J))


def momentum_step_rk4_2solids(u, v, p, X1a, X2a, X1b, X2b, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt,
                              rho_s, rho_f, phi_a, phi_b, mu_f, w_t, k_rep=0.0, w_c=None, detg_clamp=4.0):
    """pyRMT/functions.py:765-835 -- RK4 predictor for TWO neo-Hookean solids sharing one velocity
    field, with the repulsive contact force of compute_contact_force as a body force.

    Two stress kernels (detG clamped), the contact-force kernel, then four fused stage kernels
    (n = 2 mixture stress + stage update; Ha, Hb and rho_local evaluated from the level sets
    in-kernel) with the BC table applied between them.  eta_s is accepted and unused, as upstream.
    Returns (u_new, v_new, min(Ja, Jb))."""
    as_np = is_np(u)
    ud, vd, pd, xa1, xa2, xb1, xb2, pa, pb = (to_dev(t) for t in (u, v, p, X1a, X2a, X1b, X2b, phi_a, phi_b))
    Ny, Nx = shape2(ud)
    lib, st = ctx().lib, stream()
    dx, dy, dt = float(dx), float(dy), float(dt)
    if w_c is None:
        w_c = 2.0 * w_t
    sA = [torch.empty_like(ud) for _ in range(4)]
    sB = [torch.empty_like(ud) for _ in range(4)]
    for (x1, x2, ph, out) in ((xa1, xa2, pa, sA), (xb1, xb2, pb, sB)):
        _chk(lib.rmt_solid_stress(ptr(x1), ptr(x2), ptr(ph), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
                                  Ny, Nx, dx, dy, float(mu_s), float(kappa), 0.0, float(detg_clamp), 0, st),
             "rmt_solid_stress")
    fcx = fcy = None
    if k_rep > 0.0:
        fcx, fcy = torch.empty_like(ud), torch.empty_like(ud)
        _chk(lib.rmt_contact_force(ptr(pa), ptr(pb), ptr(fcx), ptr(fcy), Ny, Nx, dx, dy, float(k_rep), float(w_c),
                                   st), "rmt_contact_force")
    sa_u, sa_v = apply_bc_(velocity_bc, ud.clone(), vd.clone())
    sb_u, sb_v = torch.empty_like(ud), torch.empty_like(ud)
    acc_u, acc_v = torch.empty_like(ud), torch.empty_like(ud)
    un, vn = torch.empty_like(ud), torch.empty_like(ud)

    def stage(k, iu, iv, ou, ov):
        _chk(lib.rmt_momentum_stage_2solids(ptr(iu), ptr(iv), ptr(pd), ptr(sA[0]), ptr(sA[1]), ptr(sA[2]),
                                            ptr(sB[0]), ptr(sB[1]), ptr(sB[2]), ptr(pa), ptr(pb), ptr(fcx),
                                            ptr(fcy), ptr(ud), ptr(vd), ptr(acc_u), ptr(acc_v), ptr(ou), ptr(ov),
                                            Ny, Nx, dx, dy, dt, float(mu_f), float(w_t), float(rho_s),
                                            float(rho_f), k, st), "rmt_momentum_stage_2solids")
        return apply_bc_(velocity_bc, ou, ov)

    sb_u, sb_v = stage(1, sa_u, sa_v, sb_u, sb_v)
    sa_u, sa_v = stage(2, sb_u, sb_v, sa_u, sa_v)
    sb_u, sb_v = stage(3, sa_u, sa_v, sb_u, sb_v)
    un, vn = stage(4, sb_u, sb_v, un, vn)
    Jmin = torch.empty_like(ud)
    _chk(lib.rmt_min2(ptr(sA[3]), ptr(sB[3]), ptr(Jmin), Jmin.numel(), st), "rmt_min2")
    return tuple(to_user(t, as_np) for t in (un, vn, Jmin))


def compute_contact_force(phi1, phi2, k_rep, w_c, dx, dy):
    """pyRMT/functions.py:864-895 -- repulsive solid-solid contact body force (fx, fy)."""
    as_np = is_np(phi1)
    p1, p2 = to_dev(phi1), to_dev(phi2)
    Ny, Nx = shape2(p1)
    fx, fy = torch.empty_like(p1), torch.empty_like(p1)
    _chk(ctx().lib.rmt_contact_force(ptr(p1), ptr(p2), ptr(fx), ptr(fy), Ny, Nx, float(dx), float(dy),
                                     float(k_rep), float(w_c), stream()), "rmt_contact_force")
    return to_user(fx, as_np), to_user(fy, as_np)


# --------------------------------------------------------------------------
# pressure projection
# --------------------------------------------------------------------------
class _LazyPoissonMatrix:
    """Placeholder returned by build_poisson_matrix: the 5-point matrix is never
    touched when ``eigenvalues`` is passed (every in-scope driver), and building it
    upstream is an O(N^2) Python loop (functions.py:949-1003)."""

    def __init__(self, Nx, Ny, dx, dy):
        self.shape = (Nx * Ny, Nx * Ny)
        self.grid = (Nx, Ny, dx, dy)

    def __repr__(self):
        return "<lazy 5-point Poisson matrix %dx%d (AMG path is out of scope)>" % self.shape


def build_poisson_matrix(Nx, Ny, dx, dy):
    """pyRMT/functions.py:949-1003 -- returns a lazy placeholder (AMG is out of scope)."""
    return _LazyPoissonMatrix(Nx, Ny, dx, dy)


def _compute_divergence(a_star, b_star, dx, dy):
    """pyRMT/functions.py:1005-1014."""
    as_np = is_np(a_star)
    ad, bd = to_dev(a_star), to_dev(b_star)
    Ny, Nx = shape2(ad)
    out = torch.empty_like(ad)
    _chk(ctx().lib.rmt_divergence(ptr(ad), ptr(bd), ptr(out), Ny, Nx, float(dx), float(dy), stream()),
         "rmt_divergence")
    return to_user(out, as_np)


def _rho_device(rho, like, need_ptp):
    """-> (rho tensor or None, scalar, device sum tensor).  Raises for the
    variable-density branch (np.ptp(rho) > 1e-10, functions.py:1298), which is
    out of scope and must not silently differ."""
    c = ctx()
    if isinstance(rho, (np.ndarray, torch.Tensor)) and getattr(rho, "ndim", 0) == 2:
        rd = to_dev(rho)
        st4 = c.stats(rd)
        if need_ptp:
            h = st4.cpu()
            if float(h[2]) - float(h[1]) > 1e-10:
                raise NotImplementedError(
                    "variable-density projection (np.ptp(rho) > 1e-10) is out of scope for the B200 path")
        return rd, 0.0, st4
    r = float(rho)
    s = torch.full((4,), r * like.numel(), dtype=F64, device=like.device)
    return None, r, s


def _compute_divergence_rc(a_star, b_star, p_prev, dt, rho, dx, dy):
    """pyRMT/functions.py:1016-1071 (constant-density branch)."""
    as_np = is_np(a_star)
    ad, bd, pd = to_dev(a_star), to_dev(b_star), to_dev(p_prev)
    Ny, Nx = shape2(ad)
    _, _, rsum = _rho_device(rho, ad, True)
    out = torch.empty_like(ad)
    _chk(ctx().lib.rmt_divergence_rc(ptr(ad), ptr(bd), ptr(pd), ptr(rsum), ptr(out), Ny, Nx, float(dx),
                                     float(dy), float(dt), stream()), "rmt_divergence_rc")
    return to_user(out, as_np)


def _compute_pressure_gradient(p, dx, dy):
    """pyRMT/functions.py:1073-1089."""
    as_np = is_np(p)
    pd = to_dev(p)
    Ny, Nx = shape2(pd)
    gx, gy = torch.empty_like(pd), torch.empty_like(pd)
    _chk(ctx().lib.rmt_pressure_gradient(ptr(pd), ptr(gx), ptr(gy), Ny, Nx, float(dx), float(dy), stream()),
         "rmt_pressure_gradient")
    return to_user(gx, as_np), to_user(gy, as_np)


def _precompute_poisson_eigenvalues(Nx, Ny, dx, dy):
    """pyRMT/functions.py:1091-1104 -- host table, same NumPy expressions as upstream."""
    kx = np.arange(Nx)
    ky = np.arange(Ny)
    lam_x = -2.0 * (1.0 - np.cos(np.pi * kx / (Nx - 1))) / dx**2
    lam_y = -2.0 * (1.0 - np.cos(np.pi * ky / (Ny - 1))) / dy**2
    eigenvalues = lam_x[np.newaxis, :] + lam_y[:, np.newaxis]
    eigenvalues[0, 0] = 1.0
    return eigenvalues


def _solve_poisson_dct(rhs_2d, eigenvalues):
    """pyRMT/functions.py:1107-1119 -- hand-written DCT-I solve (no scipy, no cuFFT)."""
    as_np = is_np(rhs_2d)
    rd = to_dev(rhs_2d)
    Ny, Nx = shape2(rd)
    c = ctx()
    if tuple(np.shape(eigenvalues)) != (Ny, Nx):
        raise ValueError("eigenvalues shape %s does not match the grid %s" % (tuple(np.shape(eigenvalues)), (Ny, Nx)))
    plan, (eig,) = c.plan_with_tables(Ny, Nx, 0, (eigenvalues,), (F64,))
    sol = torch.empty_like(rd)
    _chk(c.lib.rmt_poisson_solve_dct(plan, ptr(rd), ptr(eig), ptr(sol), None, stream()),
         "rmt_poisson_solve_dct")
    return to_user(sol, as_np)


def _precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy):
    """pyRMT/functions.py:1177-1202 -- host tables (eig, null_mask)."""
    mx, my = Nx - 1, Ny - 1
    kx = np.arange(mx)
    ky = np.arange(my)
    lam_x = -(np.sin(2.0 * np.pi * kx / mx) / dx) ** 2
    lam_y = -(np.sin(2.0 * np.pi * ky / my) / dy) ** 2
    eig = lam_x[np.newaxis, :] + lam_y[:, np.newaxis]
    null_mask = np.abs(eig) < 1e-12
    eig = eig.copy()
    eig[null_mask] = 1.0
    return eig, null_mask


def _tile_overlap(field_reduced, Ny, Nx):
    """pyRMT/functions.py:1205-1213."""
    if isinstance(field_reduced, torch.Tensor):
        out = torch.empty((Ny, Nx), dtype=field_reduced.dtype, device=field_reduced.device)
    else:
        out = np.empty((Ny, Nx), dtype=np.asarray(field_reduced).dtype)
    out[:-1, :-1] = field_reduced
    out[-1, :-1] = field_reduced[0, :]
    out[:-1, -1] = field_reduced[:, 0]
    out[-1, -1] = field_reduced[0, 0]
    return out


def _solve_poisson_fft(rhs_full, eigenvalues_periodic):
    """pyRMT/functions.py:1216-1233 -- hand-written FFT solve on the reduced grid."""
    as_np = is_np(rhs_full)
    rd = to_dev(rhs_full)
    Ny, Nx = shape2(rd)
    c = ctx()
    eig_h, null_h = eigenvalues_periodic
    if tuple(np.shape(eig_h)) != (Ny - 1, Nx - 1):
        raise ValueError("periodic eigenvalues shape %s does not match the reduced grid %s"
                         % (tuple(np.shape(eig_h)), (Ny - 1, Nx - 1)))
    plan, (eig, null) = c.plan_with_tables(Ny, Nx, 1, (eig_h, null_h), (F64, torch.uint8))
    sol = torch.empty_like(rd)
    _chk(c.lib.rmt_poisson_solve_fft(plan, ptr(rd), ptr(eig), ptr(null), ptr(sol), None,
                                     stream()), "rmt_poisson_solve_fft")
    return to_user(sol, as_np)


def _compute_divergence_periodic(a_star, b_star, dx, dy):
    """pyRMT/functions.py:1236-1243."""
    as_np = is_np(a_star)
    ad, bd = to_dev(a_star), to_dev(b_star)
    Ny, Nx = shape2(ad)
    out = torch.empty_like(ad)
    _chk(ctx().lib.rmt_divergence_periodic(ptr(ad), ptr(bd), ptr(out), Ny, Nx, float(dx), float(dy), stream()),
         "rmt_divergence_periodic")
    return to_user(out, as_np)


def _compute_pressure_gradient_periodic(p, dx, dy):
    """pyRMT/functions.py:1246-1252."""
    as_np = is_np(p)
    pd = to_dev(p)
    Ny, Nx = shape2(pd)
    gx, gy = torch.empty_like(pd), torch.empty_like(pd)
    _chk(ctx().lib.rmt_pressure_gradient_periodic(ptr(pd), ptr(gx), ptr(gy), Ny, Nx, float(dx), float(dy),
                                                  stream()), "rmt_pressure_gradient_periodic")
    return to_user(gx, as_np), to_user(gy, as_np)


def pressure_projection_amg(a_star, b_star, dx, dy, dt, rho, velocity_bc, A=None, ml=None, p_prev=None,
                            eigenvalues=None, bc_type='neumann'):
    """pyRMT/functions.py:1255-1364, constant-density branches.

    Neumann: Rhie-Chow (or plain, when p_prev is None) divergence -> rho*div/dt ->
    DCT-I solve -> correction -> BC -> incremental pressure, mean removed.
    Periodic: wide-central divergence on the reduced grid -> FFT solve -> the same
    back end.  Variable density and the AMG fallback raise NotImplementedError.
    Returns (a, b, p, A, ml) with A and ml passed through."""
    as_np = is_np(a_star)
    ad, bd = to_dev(a_star), to_dev(b_star)
    Ny, Nx = shape2(ad)
    c = ctx()
    lib, st = c.lib, stream()
    dx, dy, dt = float(dx), float(dy), float(dt)
    periodic = 1 if bc_type == 'periodic' else 0
    pp = to_dev(p_prev) if p_prev is not None else None

    if periodic:
        if eigenvalues is None:
            eigenvalues = _precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy)
        rd, rscalar, rsum = _rho_device(rho, ad, False)
    else:
        rd, rscalar, rsum = _rho_device(rho, ad, True)
        if eigenvalues is None:
            raise NotImplementedError(
                "pressure_projection_amg without `eigenvalues` falls back to AMG upstream; the AMG "
                "path is out of scope -- pass _precompute_poisson_eigenvalues(...)")

    rhs = torch.empty_like(ad)
    _chk(lib.rmt_projection_rhs(ptr(ad), ptr(bd), None if periodic else ptr(pp), ptr(rd), rscalar, ptr(rsum),
                                ptr(rhs), Ny, Nx, dx, dy, dt, periodic, st), "rmt_projection_rhs")
    sol = torch.empty_like(ad)
    ssum = torch.empty(4, dtype=F64, device=ad.device)
    if periodic:
        eig_h, null_h = eigenvalues
        if tuple(eig_h.shape if isinstance(eig_h, torch.Tensor) else np.shape(eig_h)) != (Ny - 1, Nx - 1):
            raise ValueError("periodic eigenvalues do not match the reduced grid")
        plan, (eig, null) = c.plan_with_tables(Ny, Nx, 1, (eig_h, null_h), (F64, torch.uint8))
        _chk(lib.rmt_poisson_solve_fft(plan, ptr(rhs), ptr(eig), ptr(null), ptr(sol), ptr(ssum),
                                       st), "rmt_poisson_solve_fft")
    else:
        if tuple(np.shape(eigenvalues)) != (Ny, Nx):
            raise ValueError("eigenvalues shape %s does not match the grid %s"
                             % (tuple(np.shape(eigenvalues)), (Ny, Nx)))
        plan, (eig,) = c.plan_with_tables(Ny, Nx, 0, (eigenvalues,), (F64,))
        _chk(lib.rmt_poisson_solve_dct(plan, ptr(rhs), ptr(eig), ptr(sol), ptr(ssum), st),
             "rmt_poisson_solve_dct")
    a, b, p = torch.empty_like(ad), torch.empty_like(ad), torch.empty_like(ad)
    # p -= mean(p) (:1361) inside the back end: mean(p_prev + pc) is mean(p_prev) to rounding, and sum(p_prev)
    # was produced by the call that made p_prev (carried on the tensor; anything else is summed here)
    pp_sum = None
    if pp is not None:
        carried = getattr(pp, "_rmt_sum", None)
        pp_sum = carried[0] if carried is not None and carried[1] == pp._version else c.stats(pp)
    partial = c.scratch("proj_partial", int(lib.rmt_projection_partials(Ny, Nx)))
    psum = torch.empty(1, dtype=F64, device=ad.device)
    _chk(lib.rmt_projection_correct_centered(ptr(sol), ptr(ssum), ptr(ad), ptr(bd), ptr(rd), rscalar, ptr(pp),
                                             ptr(pp_sum), ptr(a), ptr(b), ptr(p), ptr(partial), ptr(psum), Ny, Nx,
                                             dx, dy, dt, periodic, st), "rmt_projection_correct_centered")
    p._rmt_sum = (psum, p._version)
    a, b = apply_bc_(velocity_bc, a, b)
    return to_user(a, as_np), to_user(b, as_np), to_user(p, as_np), A, ml


# --------------------------------------------------------------------------
# level-set reinitialisation (only the default 'none' is on the hot path)
# --------------------------------------------------------------------------
def reinitialize_phi_PDE(phi_in, dx, dy, num_iters, apply_phi_BCs_func, dt_reinit_factor=0.5):
    """pyRMT/functions.py:1369-1411 -- Sussman-Smereka-Osher pseudo-time reinitialisation: one sign
    kernel, then one Godunov-upwind step kernel per iteration (ping-pong buffers).  The optional
    phi BC callback runs between steps as upstream: this module's apply_phi_BCs works on the device
    tensor in place; any other callable receives what the caller passed in (an ndarray costs a
    host round trip per iteration)."""
    as_np = is_np(phi_in)
    cur = to_dev(phi_in).clone()
    Ny, Nx = shape2(cur)
    lib, st = ctx().lib, stream()
    dx, dy = float(dx), float(dy)
    s0 = torch.empty_like(cur)
    _chk(lib.rmt_reinit_sign(ptr(cur), ptr(s0), cur.numel(), dx, st), "rmt_reinit_sign")
    dtau = float(dt_reinit_factor) * min(dx, dy)
    nxt = torch.empty_like(cur)
    for _ in range(int(num_iters)):
        _chk(lib.rmt_reinit_step(ptr(cur), ptr(s0), ptr(nxt), Ny, Nx, dx, dy, dtau, st), "rmt_reinit_step")
        cur, nxt = nxt, cur
        if apply_phi_BCs_func is None:
            continue
        if apply_phi_BCs_func is apply_phi_BCs or not as_np:
            cur = to_dev(apply_phi_BCs_func(cur))
        else:
            cur = to_dev(np.ascontiguousarray(apply_phi_BCs_func(cur.cpu().numpy()), dtype=np.float64))
        if cur.data_ptr() == nxt.data_ptr():
            nxt = torch.empty_like(cur)
    return to_user(cur, as_np)


def reinitialize_phi_fmm(phi, dx, dy):
    """pyRMT/functions.py:1414-1433 -- needs scikit-fmm upstream (third party)."""
    raise ImportError("reinitialize_phi_fmm requires scikit-fmm (pip install scikit-fmm)")


def reinitialize_level_set(phi, dx, dy, method='none', num_iters=20, dt_reinit_factor=0.2,
                           apply_phi_BCs_func=None):
    """pyRMT/functions.py:1436-1456."""
    if method == 'none':
        return phi
    if method == 'pde':
        return reinitialize_phi_PDE(phi, dx, dy, num_iters, apply_phi_BCs_func, dt_reinit_factor)
    if method == 'fmm':
        return reinitialize_phi_fmm(phi, dx, dy)
    raise ValueError("Unknown reinit method %r (expected 'none', 'pde' or 'fmm')" % (method,))


# Deprecated aliases kept upstream for notebooks (functions.py:1462-1466)
velocity_RK4 = momentum_step_rk4
heaviside_smooth_alt = smoothed_heaviside
compute_solid_stress = solid_cauchy_stress
extrapolate_transverse_layers_2field = extrapolate_reference_map
advect_semi_lagrangian_rk4 = advect_semilagrangian_rk4

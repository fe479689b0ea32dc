"""CPU oracle for the collocated RMT timestep -- TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithm of the reference's hot path
(upstream ``pyRMT/functions.py``, ``interpolators.py``, ``utils.py``).  It is
the checker the CUDA path in ``pyrmt_b200`` is compared against; it is never
part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Layout
------
* ``rmt_oracle.c`` holds the loops the reference compiles with Numba (C, built
  with ``-ffp-contract=off`` so no FMA contraction happens; the four kernels the
  reference runs ``parallel=True`` use OpenMP).
* This file holds the NumPy-level operators in the same operation order as the
  reference, plus the third-party transforms the reference itself calls
  (``scipy.fft.dctn/idctn(type=1)`` at functions.py:1115-1117 and
  ``numpy.fft.fft2/ifft2`` at :1227-1230 -- pocketfft, unpinned upstream; here
  scipy 1.18.1 / numpy 2.3.5).

Parity is PINNED: ``tests/test_oracle_golden.py`` compares every operator with
outputs recorded from the real reference (``tests/golden/*.npz``, produced by
``tests/golden/make_golden.py`` importing /root/reference in the build
container).  Citations are ``file:line`` in the upstream repo.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np
from scipy.fft import dctn as _dctn, idctn as _idctn

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "rmt_oracle.c")
_SO = os.path.join(_HERE, "_build", "librmt_oracle.so")
_lock = threading.Lock()
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile rmt_oracle.c into oracle/_build/librmt_oracle.so (gcc only)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            _lib = ctypes.CDLL(build())
            _lib.orc_extrapolate.restype = ctypes.c_int
            _lib.orc_cubic_convolution.restype = ctypes.c_double
            _lib.orc_cubic_convolution.argtypes = [ctypes.c_double] * 5
            _lib.orc_weno5_left.restype = ctypes.c_double
            _lib.orc_weno5_left.argtypes = [ctypes.c_double] * 5
            _lib.orc_weno5_right.restype = ctypes.c_double
            _lib.orc_weno5_right.argtypes = [ctypes.c_double] * 5
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(c_dp)


_d = ctypes.c_double
_i = ctypes.c_int
_l = ctypes.c_long


# --------------------------------------------------------------------------
# utils.py
# --------------------------------------------------------------------------
def _stencil(name, f, h):
    f = _f64(f)
    out = np.zeros_like(f)
    Ny, Nx = f.shape
    getattr(lib(), name)(_p(f), _p(out), _i(Ny), _i(Nx), _d(h))
    return out


def grad_central_x_2nd(f, dx):      # utils.py:4-14
    return _stencil("orc_grad_x_2nd", f, dx)


def grad_central_y_2nd(f, dy):      # utils.py:16-25
    return _stencil("orc_grad_y_2nd", f, dy)


def grad_central_x_4th(f, dx):      # utils.py:27-42
    return _stencil("orc_grad_x_4th", f, dx)


def grad_central_y_4th(f, dy):      # utils.py:44-59
    return _stencil("orc_grad_y_4th", f, dy)


def diff_upwind_3rd(f, u, h, axis):  # utils.py:61-114
    f = _f64(f)
    u = _f64(u)
    out = np.empty_like(f)
    Ny, Nx = f.shape
    lib().orc_diff_upwind_3rd(_p(f), _p(u), _p(out), _i(Ny), _i(Nx), _d(h), _i(int(axis)))
    return out


def lap_2nd(f, dx, dy):             # utils.py:116-131
    f = _f64(f)
    out = np.empty_like(f)
    Ny, Nx = f.shape
    lib().orc_lap_2nd(_p(f), _p(out), _i(Ny), _i(Nx), _d(dx), _d(dy))
    return out


def fast_solve_3x3(A, b):           # utils.py:134-167
    A = _f64(A)
    b = _f64(b)
    x = np.zeros(3)
    lib().orc_fast_solve_3x3(_p(A), _p(b), _p(x))
    return x


# --------------------------------------------------------------------------
# interpolators.py
# --------------------------------------------------------------------------
def _sample(name, u, xq, yq, dx, dy, Nx, Ny):
    u = _f64(u)
    xq = _f64(xq)
    yq = _f64(yq)
    out = np.zeros_like(xq)
    getattr(lib(), name)(_p(u), _p(xq), _p(yq), _p(out), _l(xq.size), _d(dx), _d(dy), _i(Nx),
                         _i(Ny))
    return out


def bilinear_interpolate(u, xq, yq, dx, dy, Nx, Ny):   # interpolators.py:4-62
    return _sample("orc_bilinear", u, xq, yq, dx, dy, Nx, Ny)


def bicubic_interpolate(u, xq, yq, dx, dy, Nx, Ny):    # interpolators.py:64-141
    return _sample("orc_bicubic", u, xq, yq, dx, dy, Nx, Ny)


def cubic_convolution(v0, v1, v2, v3, x):              # interpolators.py:143-154
    return lib().orc_cubic_convolution(v0, v1, v2, v3, x)


# --------------------------------------------------------------------------
# functions.py -- grid, level set, time step
# --------------------------------------------------------------------------
def create_grid(Nx, Ny, Lx, Ly):                       # functions.py:25-31
    xs = np.linspace(0, Lx, Nx)
    ys = np.linspace(0, Ly, Ny)
    X, Y = np.meshgrid(xs, ys)
    return X, Y, xs[1] - xs[0], ys[1] - ys[0]


def apply_phi_BCs(phi):                                # functions.py:33-46 (in place)
    phi[0:3, :] = phi[-6:-3, :]
    phi[-3:, :] = phi[3:6, :]
    phi[:, 0:3] = phi[:, -6:-3]
    phi[:, -3:] = phi[:, 3:6]
    return phi


def rebuild_phi_from_reference_map(X1, X2, phi_init_func):   # functions.py:1366-1367
    return phi_init_func(X1, X2)


def disc_sdf(X1, X2, centres_x, centres_y, radii):
    """phi0 = min_k(|xi - c_k| - R_k); one disc == benchmarks/common.py:55-57."""
    X1 = _f64(X1)
    X2 = _f64(X2)
    cx = _f64(np.atleast_1d(centres_x))
    cy = _f64(np.atleast_1d(centres_y))
    R = _f64(np.atleast_1d(radii))
    out = np.empty_like(X1)
    lib().orc_disc_sdf(_p(X1), _p(X2), _p(out), _l(X1.size), _p(cx), _p(cy), _p(R), _i(cx.size))
    return out


def compute_timestep(a, b, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f=0.0,
                     eta_s=0.0, kappa=0.0):            # functions.py:165-192
    wave = np.sqrt((kappa + mu_s * 4.0 / 3.0) / (rho_s + 1e-12))
    dt_solid = CFL * dx / (wave + 1e-14)
    speed = np.max(np.sqrt(a**2 + b**2))
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
# functions.py -- reference-map advection
# --------------------------------------------------------------------------
def _sl(q, a, b, X, Y, dt, dx, dy, cubic):
    q, a, b, X, Y = map(_f64, (q, a, b, X, Y))
    Ny, Nx = q.shape
    out = np.empty_like(q)
    lib().orc_advect_sl_rk4(_p(q), _p(a), _p(b), _p(X), _p(Y), _p(out), _i(Ny), _i(Nx), _d(dt),
                            _d(dx), _d(dy), _i(cubic))
    return out


def advect_semilagrangian_rk4(q, a, b, X, Y, dt, dx, dy):        # functions.py:194-227
    return _sl(q, a, b, X, Y, dt, dx, dy, 0)


def advect_semilagrangian_cubic_rk4(q, a, b, X, Y, dt, dx, dy):  # functions.py:230-251
    return _sl(q, a, b, X, Y, dt, dx, dy, 1)


def _euler_rhs(name, q, a, b, dx, dy, phi, w_cut):
    q, a, b, phi = map(_f64, (q, a, b, phi))
    Ny, Nx = q.shape
    rhs = np.empty_like(q)
    getattr(lib(), name)(_p(q), _p(a), _p(b), _p(phi), _p(rhs), _i(Ny), _i(Nx), _d(dx), _d(dy),
                         _d(w_cut))
    return rhs


def _weno5_rhs(q, a, b, dx, dy, phi, w_cut):           # functions.py:321-393
    return _euler_rhs("orc_weno5_rhs", q, a, b, dx, dy, phi, w_cut)


def _central2_rhs(q, a, b, dx, dy, phi, w_cut):        # functions.py:420-440
    return _euler_rhs("orc_central2_rhs", q, a, b, dx, dy, phi, w_cut)


def _conservative_rhs(q, a, b, dx, dy, phi, w_cut):    # functions.py:462-486
    return _euler_rhs("orc_conservative_rhs", q, a, b, dx, dy, phi, w_cut)


def _ssp_rk3(rhs_fn, q, a, b, dx, dy, dt, phi, w_cut):
    """Shu-Osher SSP-RK3 exactly as functions.py:406-415 / :450-459 / :493-496."""
    s1 = q + dt * rhs_fn(q, a, b, dx, dy, phi, w_cut)
    s2 = 0.75 * q + 0.25 * (s1 + dt * rhs_fn(s1, a, b, dx, dy, phi, w_cut))
    return (1.0 / 3.0) * q + (2.0 / 3.0) * (s2 + dt * rhs_fn(s2, a, b, dx, dy, phi, w_cut))


def advect_weno5_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):          # functions.py:396-415
    return _ssp_rk3(_weno5_rhs, q, a, b, dx, dy, dt, phi, w_cut)


def advect_central2_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):       # functions.py:443-459
    return _ssp_rk3(_central2_rhs, q, a, b, dx, dy, dt, phi, w_cut)


def advect_conservative_rk3(q, a, b, dx, dy, dt, phi, w_cut=0.0):   # functions.py:489-496
    return _ssp_rk3(_conservative_rhs, q, a, b, dx, dy, dt, phi, w_cut)


def advect_reference_map(q, a, b, X, Y, dt, dx, dy, phi, scheme='semilagrangian', w_cut=0.0):
    """Dispatcher + finite guard, functions.py:501-542."""
    if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
        raise FloatingPointError(
            "advect_reference_map: non-finite velocity (the simulation diverged)")
    table = {
        'semilagrangian': lambda: advect_semilagrangian_rk4(q, a, b, X, Y, dt, dx, dy),
        'semilagrangian_cubic': lambda: advect_semilagrangian_cubic_rk4(q, a, b, X, Y, dt, dx, dy),
        'central2': lambda: advect_central2_rk3(q, a, b, dx, dy, dt, phi, w_cut),
        'weno5': lambda: advect_weno5_rk3(q, a, b, dx, dy, dt, phi, w_cut),
        'conservative': lambda: advect_conservative_rk3(q, a, b, dx, dy, dt, phi, w_cut),
    }
    if scheme not in table:
        raise ValueError("Unknown advection scheme %r (expected 'semilagrangian', "
                         "'central2', 'weno5' or 'conservative')" % (scheme,))
    return table[scheme]()


# --------------------------------------------------------------------------
# functions.py -- extrapolation
# --------------------------------------------------------------------------
def extrapolate_reference_map(X1, X2, phi, dx, dy, max_layers):    # functions.py:48-163
    X1e = _f64(X1).copy()
    X2e = _f64(X2).copy()
    phi = _f64(phi)
    Ny, Nx = X1e.shape
    lib().orc_extrapolate(_p(X1e), _p(X2e), _p(phi), _i(Ny), _i(Nx), _d(dx), _d(dy),
                          _i(int(max_layers)))
    return X1e, X2e


def libm_exp(x):
    """glibc exp() on an array -- what Numba's scalar np.exp lowers to (H2)."""
    x = _f64(x)
    y = np.empty_like(x)
    lib().orc_exp_array(_p(x), _p(y), _l(x.size))
    return y


# --------------------------------------------------------------------------
# functions.py -- stress, Heaviside, momentum predictor
# --------------------------------------------------------------------------
def solid_cauchy_stress(X1, X2, dx, dy, mu_s, kappa, phi, w_cut=0.0, detg_clamp=0.0,
                        isochoric=False):              # functions.py:545-658
    X1, X2, phi = map(_f64, (X1, X2, phi))
    Ny, Nx = X1.shape
    sxx, sxy, syy, J = (np.empty_like(X1) for _ in range(4))
    lib().orc_solid_stress(_p(X1), _p(X2), _p(phi), _p(sxx), _p(sxy), _p(syy), _p(J), _i(Ny),
                           _i(Nx), _d(dx), _d(dy), _d(mu_s), _d(kappa), _d(w_cut), _d(detg_clamp),
                           _i(1 if isochoric else 0))
    return sxx, sxy, syy, J


def smoothed_heaviside(x, w_t):                        # functions.py:660-671
    inv_w = 1.0 / w_t
    inv_pi = 1.0 / np.pi
    H = 0.5 * (1.0 + x * inv_w + inv_pi * np.sin(np.pi * x * inv_w))
    H = np.where(x > w_t, 1.0, H)
    return np.where(x < -w_t, 0.0, H)


def velocity_rhs_blended_optimized(u, v, p, sxx_s, sxy_s, syy_s, dx, dy, phi, mu_f, H, dH_dx,
                                   dH_dy, rho_local, st_force_x, st_force_y):
    """Blended-stress momentum RHS, functions.py:897-944 (same op order)."""
    ux = grad_central_x_2nd(u, dx)
    vy = grad_central_y_2nd(v, dy)
    uy = grad_central_y_2nd(u, dy)
    vx = grad_central_x_2nd(v, dx)
    f_xx = 2 * mu_f * ux
    f_yy = 2 * mu_f * vy
    f_xy = mu_f * (uy + vx)
    t_xx = H * f_xx + (1 - H) * sxx_s
    t_yy = H * f_yy + (1 - H) * syy_s
    t_xy = H * f_xy + (1 - H) * sxy_s
    div_x = grad_central_x_2nd(t_xx, dx) + grad_central_y_2nd(t_xy, dy)
    div_y = grad_central_x_2nd(t_xy, dx) + grad_central_y_2nd(t_yy, dy)
    adv_u = -u * diff_upwind_3rd(u, u, dx, 1) - v * diff_upwind_3rd(u, v, dy, 0)
    adv_v = -u * diff_upwind_3rd(v, u, dx, 1) - v * diff_upwind_3rd(v, v, dy, 0)
    px = grad_central_x_2nd(p, dx)
    py = grad_central_y_2nd(p, dy)
    ru = adv_u + (div_x + st_force_x - px) / (rho_local + 1e-12)
    rv = adv_v + (div_y + st_force_y - py) / (rho_local + 1e-12)
    return ru, rv


def compute_curvature(phi, dx, dy):                    # functions.py:837-861
    gx = grad_central_x_2nd(phi, dx)
    gy = grad_central_y_2nd(phi, dy)
    mag = np.sqrt(gx * gx + gy * gy) + 1e-12
    return grad_central_x_2nd(gx / mag, dx) + grad_central_y_2nd(gy / mag, dy)


def apply_velocity_BCs(bc, u, v):                      # functions.py:946-947
    return bc(u, v)


def momentum_step_rk4(u, v, p, X1, X2, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f,
                      phi, mu_f, w_t, gamma=0.0, stress_band=False, detg_clamp=3.0):
    """Classical RK4 momentum predictor, functions.py:673-762."""
    band_cut = w_t if stress_band else 0.0
    clamp = detg_clamp if stress_band else 0.0
    exx, exy, eyy, J = solid_cauchy_stress(X1, X2, dx, dy, mu_s, kappa, phi, w_cut=band_cut,
                                           detg_clamp=clamp)
    H = smoothed_heaviside(phi, w_t)
    dHx = grad_central_x_2nd(H, dx)
    dHy = grad_central_y_2nd(H, dy)
    rho_local = (1 - H) * rho_s + H * rho_f
    if gamma > 1e-12:
        curv = compute_curvature(phi, dx, dy)
        fsx = -gamma * curv * dHx
        fsy = -gamma * curv * dHy
    else:
        fsx = fsy = 0.0
    solid = phi <= 0.0

    def slope(us, vs):
        us, vs = velocity_bc(us, vs)
        if eta_s > 0.0 and np.any(solid):
            txx, txy, tyy = exx.copy(), exy.copy(), eyy.copy()
            ux = grad_central_x_2nd(us, dx)
            vy = grad_central_y_2nd(vs, dy)
            uy = grad_central_y_2nd(us, dy)
            vx = grad_central_x_2nd(vs, dx)
            txx[solid] += eta_s * ux[solid]
            tyy[solid] += eta_s * vy[solid]
            txy[solid] += eta_s * 0.5 * (uy[solid] + vx[solid])
        else:
            txx, txy, tyy = exx, exy, eyy
        return velocity_rhs_blended_optimized(us, vs, p, txx, txy, tyy, dx, dy, phi, mu_f, H, dHx,
                                              dHy, rho_local, fsx, fsy)

    k1u, k1v = slope(u, v)
    k2u, k2v = slope(u + 0.5 * dt * k1u, v + 0.5 * dt * k1v)
    k3u, k3v = slope(u + 0.5 * dt * k2u, v + 0.5 * dt * k2v)
    k4u, k4v = slope(u + dt * k3u, v + dt * k3v)
    un = u + (dt / 6.0) * (k1u + 2 * k2u + 2 * k3u + k4u)
    vn = v + (dt / 6.0) * (k1v + 2 * k2v + 2 * k3v + k4v)
    un, vn = velocity_bc(un, vn)
    return un, vn, exx, exy, eyy, J


def compute_contact_force(phi1, phi2, k_rep, w_c, dx, dy):      # functions.py:864-895
    """Repulsion across the mid-surface phi12 = (phi1 - phi2)/2 inside either solid."""
    phi12 = 0.5 * (phi1 - phi2)
    aphi = np.abs(phi12)
    delta = np.where(aphi < w_c, (1.0 + np.cos(np.pi * phi12 / w_c)) / (2.0 * w_c), 0.0)
    g12x = grad_central_x_2nd(phi12, dx)
    g12y = grad_central_y_2nd(phi12, dy)
    gmag = np.sqrt(g12x ** 2 + g12y ** 2) + 1e-12
    n12x = g12x / gmag
    n12y = g12y / gmag
    active = ((phi1 < 0.0) | (phi2 < 0.0)).astype(float)
    sgn = np.sign(phi12)
    return k_rep * delta * sgn * n12x * active, k_rep * delta * sgn * n12y * active


def momentum_step_rk4_2solids(u, v, p, X1a, X2a, X1b, X2b, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt,
                              rho_s, rho_f, phi_a, phi_b, mu_f, w_t, k_rep=0.0, w_c=None,
                              detg_clamp=4.0):
    """Two neo-Hookean solids in one velocity field, functions.py:765-835 (eta_s is accepted and
    unused upstream as well).  Returns (u_new, v_new, min(Ja, Jb))."""
    if w_c is None:
        w_c = 2.0 * w_t
    sAxx, sAxy, sAyy, Ja = solid_cauchy_stress(X1a, X2a, dx, dy, mu_s, kappa, phi_a, detg_clamp=detg_clamp)
    sBxx, sBxy, sByy, Jb = solid_cauchy_stress(X1b, X2b, dx, dy, mu_s, kappa, phi_b, detg_clamp=detg_clamp)
    Ha = smoothed_heaviside(phi_a, w_t)
    Hb = smoothed_heaviside(phi_b, w_t)
    Hf = Ha + Hb - 1.0
    rho_local = Hf * rho_f + (1.0 - Ha) * rho_s + (1.0 - Hb) * rho_s
    if k_rep > 0.0:
        fcx, fcy = compute_contact_force(phi_a, phi_b, k_rep, w_c, dx, dy)
    else:
        fcx = fcy = 0.0

    def rhs(us, vs):
        us, vs = velocity_bc(us, vs)
        ux = grad_central_x_2nd(us, dx)
        vy = grad_central_y_2nd(vs, dy)
        uy = grad_central_y_2nd(us, dy)
        vx = grad_central_x_2nd(vs, dx)
        sfxx = 2.0 * mu_f * ux
        sfyy = 2.0 * mu_f * vy
        sfxy = mu_f * (uy + vx)
        sig_xx = Hf * sfxx + (1.0 - Ha) * sAxx + (1.0 - Hb) * sBxx
        sig_yy = Hf * sfyy + (1.0 - Ha) * sAyy + (1.0 - Hb) * sByy
        sig_xy = Hf * sfxy + (1.0 - Ha) * sAxy + (1.0 - Hb) * sBxy
        div_x = grad_central_x_2nd(sig_xx, dx) + grad_central_y_2nd(sig_xy, dy)
        div_y = grad_central_x_2nd(sig_xy, dx) + grad_central_y_2nd(sig_yy, dy)
        u_adv = -us * diff_upwind_3rd(us, us, dx, 1) - vs * diff_upwind_3rd(us, vs, dy, 0)
        v_adv = -us * diff_upwind_3rd(vs, us, dx, 1) - vs * diff_upwind_3rd(vs, vs, dy, 0)
        px = grad_central_x_2nd(p, dx)
        py = grad_central_y_2nd(p, dy)
        return (u_adv + (div_x + fcx - px) / (rho_local + 1e-12),
                v_adv + (div_y + fcy - py) / (rho_local + 1e-12))

    k1u, k1v = rhs(u, v)
    k2u, k2v = rhs(u + 0.5 * dt * k1u, v + 0.5 * dt * k1v)
    k3u, k3v = rhs(u + 0.5 * dt * k2u, v + 0.5 * dt * k2v)
    k4u, k4v = rhs(u + dt * k3u, v + dt * k3v)
    un = u + (dt / 6.0) * (k1u + 2 * k2u + 2 * k3u + k4u)
    vn = v + (dt / 6.0) * (k1v + 2 * k2v + 2 * k3v + k4v)
    un, vn = velocity_bc(un, vn)
    return un, vn, np.minimum(Ja, Jb)


# --------------------------------------------------------------------------
# functions.py -- pressure projection (constant-density branches)
# --------------------------------------------------------------------------
def _compute_divergence(a_star, b_star, dx, dy):       # functions.py:1005-1014
    d = np.zeros_like(a_star)
    d[1:-1, 1:-1] = ((a_star[1:-1, 2:] - a_star[1:-1, :-2]) / (2 * dx)
                     + (b_star[2:, 1:-1] - b_star[:-2, 1:-1]) / (2 * dy))
    return d


def _compute_pressure_gradient(p, dx, dy):             # functions.py:1073-1089
    gx = np.zeros_like(p)
    gy = np.zeros_like(p)
    gx[1:-1, 1:-1] = (p[1:-1, 2:] - p[1:-1, :-2]) / (2 * dx)
    gy[1:-1, 1:-1] = (p[2:, 1:-1] - p[:-2, 1:-1]) / (2 * dy)
    gx[:, 0] = (-3.0 * p[:, 0] + 4.0 * p[:, 1] - p[:, 2]) / (2.0 * dx)
    gx[:, -1] = (3.0 * p[:, -1] - 4.0 * p[:, -2] + p[:, -3]) / (2.0 * dx)
    gy[0, :] = (-3.0 * p[0, :] + 4.0 * p[1, :] - p[2, :]) / (2.0 * dy)
    gy[-1, :] = (3.0 * p[-1, :] - 4.0 * p[-2, :] + p[-3, :]) / (2.0 * dy)
    return gx, gy


def _compute_divergence_rc(a_star, b_star, p_prev, dt, rho, dx, dy):   # functions.py:1016-1071
    Ny, Nx = a_star.shape
    if isinstance(rho, np.ndarray) and rho.ndim == 2 and np.ptp(rho) > 1e-10:
        raise NotImplementedError("variable-density Rhie-Chow is out of scope (SURVEY 8a a17)")
    gx = np.zeros((Ny, Nx))
    gy = np.zeros((Ny, Nx))
    gx[:, 1:-1] = (p_prev[:, 2:] - p_prev[:, :-2]) / (2.0 * dx)
    gx[:, 0] = (-3.0 * p_prev[:, 0] + 4.0 * p_prev[:, 1] - p_prev[:, 2]) / (2.0 * dx)
    gx[:, -1] = (3.0 * p_prev[:, -1] - 4.0 * p_prev[:, -2] + p_prev[:, -3]) / (2.0 * dx)
    gy[1:-1, :] = (p_prev[2:, :] - p_prev[:-2, :]) / (2.0 * dy)
    gy[0, :] = (-3.0 * p_prev[0, :] + 4.0 * p_prev[1, :] - p_prev[2, :]) / (2.0 * dy)
    gy[-1, :] = (3.0 * p_prev[-1, :] - 4.0 * p_prev[-2, :] + p_prev[-3, :]) / (2.0 * dy)
    d_f = dt / float(np.mean(rho))
    uf = 0.5 * (a_star[:, :-1] + a_star[:, 1:])
    uf = uf - d_f * ((p_prev[:, 1:] - p_prev[:, :-1]) / dx - 0.5 * (gx[:, :-1] + gx[:, 1:]))
    vf = 0.5 * (b_star[:-1, :] + b_star[1:, :])
    vf = vf - d_f * ((p_prev[1:, :] - p_prev[:-1, :]) / dy - 0.5 * (gy[:-1, :] + gy[1:, :]))
    d = np.zeros((Ny, Nx))
    d[1:-1, 1:-1] = ((uf[1:-1, 1:] - uf[1:-1, :-1]) / dx + (vf[1:, 1:-1] - vf[:-1, 1:-1]) / dy)
    return d


def _precompute_poisson_eigenvalues(Nx, Ny, dx, dy):   # functions.py:1091-1104
    lam_x = -2.0 * (1.0 - np.cos(np.pi * np.arange(Nx) / (Nx - 1))) / dx**2
    lam_y = -2.0 * (1.0 - np.cos(np.pi * np.arange(Ny) / (Ny - 1))) / dy**2
    eig = lam_x[np.newaxis, :] + lam_y[:, np.newaxis]
    eig[0, 0] = 1.0
    return eig


def _solve_poisson_dct(rhs_2d, eigenvalues):           # functions.py:1107-1119
    sol = _idctn(_dctn(rhs_2d, type=1) / eigenvalues, type=1)
    sol -= np.mean(sol)
    return sol


def _precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy):   # functions.py:1177-1202
    mx, my = Nx - 1, Ny - 1
    lam_x = -(np.sin(2.0 * np.pi * np.arange(mx) / mx) / dx) ** 2
    lam_y = -(np.sin(2.0 * np.pi * np.arange(my) / my) / dy) ** 2
    eig = lam_x[np.newaxis, :] + lam_y[:, np.newaxis]
    null = np.abs(eig) < 1e-12
    eig = eig.copy()
    eig[null] = 1.0
    return eig, null


def _tile_overlap(red, Ny, Nx):                        # functions.py:1205-1213
    out = np.empty((Ny, Nx))
    out[:-1, :-1] = red
    out[-1, :-1] = red[0, :]
    out[:-1, -1] = red[:, 0]
    out[-1, -1] = red[0, 0]
    return out


def _solve_poisson_fft(rhs_full, eigenvalues_periodic):   # functions.py:1216-1233
    eig, null = eigenvalues_periodic
    Ny, Nx = rhs_full.shape
    r = rhs_full[:-1, :-1].copy()
    r -= np.mean(r)
    spec = np.fft.fft2(r) / eig
    spec[null] = 0.0
    sol = _tile_overlap(np.real(np.fft.ifft2(spec)), Ny, Nx)
    sol -= np.mean(sol)
    return sol


def _compute_divergence_periodic(a_star, b_star, dx, dy):   # functions.py:1236-1243
    Ny, Nx = a_star.shape
    au = a_star[:-1, :-1]
    bv = b_star[:-1, :-1]
    dudx = (np.roll(au, -1, axis=1) - np.roll(au, 1, axis=1)) / (2.0 * dx)
    dvdy = (np.roll(bv, -1, axis=0) - np.roll(bv, 1, axis=0)) / (2.0 * dy)
    return _tile_overlap(dudx + dvdy, Ny, Nx)


def _compute_pressure_gradient_periodic(p, dx, dy):    # functions.py:1246-1252
    Ny, Nx = p.shape
    pr = p[:-1, :-1]
    gx = (np.roll(pr, -1, axis=1) - np.roll(pr, 1, axis=1)) / (2.0 * dx)
    gy = (np.roll(pr, -1, axis=0) - np.roll(pr, 1, axis=0)) / (2.0 * dy)
    return _tile_overlap(gx, Ny, Nx), _tile_overlap(gy, Ny, Nx)


def pressure_projection_amg(a_star, b_star, dx, dy, dt, rho, velocity_bc, A=None, ml=None,
                            p_prev=None, eigenvalues=None, bc_type='neumann'):
    """Constant-density projection, functions.py:1255-1364 (AMG / variable
    density branches are out of scope and raise)."""
    Ny, Nx = a_star.shape
    if bc_type == 'periodic':
        if eigenvalues is None:
            eigenvalues = _precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy)
        div = _compute_divergence_periodic(a_star, b_star, dx, dy)
        rho_bar = float(np.mean(rho)) if isinstance(rho, np.ndarray) else float(rho)
        corr = _solve_poisson_fft(rho_bar * div / dt, eigenvalues)
        gx, gy = _compute_pressure_gradient_periodic(corr, dx, dy)
        a = a_star - (dt / rho) * gx
        b = b_star - (dt / rho) * gy
        a, b = velocity_bc(a, b)
        p = (p_prev + corr) if p_prev is not None else corr
        p -= np.mean(p)
        return a, b, p, A, ml
    if p_prev is not None:
        div = _compute_divergence_rc(a_star, b_star, p_prev, dt, rho, dx, dy)
    else:
        div = _compute_divergence(a_star, b_star, dx, dy)
    if isinstance(rho, np.ndarray) and rho.ndim == 2 and np.ptp(rho) > 1e-10:
        raise NotImplementedError("variable-density projection is out of scope (SURVEY 8a a17)")
    if eigenvalues is None:
        raise NotImplementedError("AMG fallback is out of scope; pass eigenvalues")
    corr = _solve_poisson_dct(rho * div / dt, eigenvalues)
    gx, gy = _compute_pressure_gradient(corr, dx, dy)
    a = a_star - (dt / rho) * gx
    b = b_star - (dt / rho) * gy
    a, b = velocity_bc(a, b)
    p = (p_prev + corr) if p_prev is not None else corr
    p -= np.mean(p)
    return a, b, p, A, ml


def reinitialize_phi_PDE(phi_in, dx, dy, num_iters, apply_phi_BCs_func, dt_reinit_factor=0.5):
    """Sussman-Smereka-Osher pseudo-time reinitialisation, functions.py:1369-1411: forward Euler on
    phi_tau = -S0 (|grad phi|_upwind - 1), S0 = phi0 / sqrt(phi0^2 + dx^2), one-sided differences with
    the edge value repeated outside the grid, Godunov choice by the sign of S0 (zero where S0 == 0)."""
    phi = np.array(phi_in, dtype=np.float64, copy=True)
    s0 = phi_in / np.sqrt(phi_in ** 2 + dx ** 2)
    dtau = dt_reinit_factor * min(dx, dy)
    zero = np.zeros_like(phi)
    for _ in range(num_iters):
        left = np.concatenate([phi[:, :1], phi[:, :-1]], axis=1)
        right = np.concatenate([phi[:, 1:], phi[:, -1:]], axis=1)
        down = np.concatenate([phi[:1, :], phi[:-1, :]], axis=0)
        up = np.concatenate([phi[1:, :], phi[-1:, :]], axis=0)
        bx, fx = (phi - left) / dx, (right - phi) / dx
        by, fy = (phi - down) / dy, (up - phi) / dy
        pos_x = np.maximum(np.maximum(bx, 0.0) ** 2, np.minimum(fx, 0.0) ** 2)
        pos_y = np.maximum(np.maximum(by, 0.0) ** 2, np.minimum(fy, 0.0) ** 2)
        neg_x = np.maximum(np.minimum(bx, 0.0) ** 2, np.maximum(fx, 0.0) ** 2)
        neg_y = np.maximum(np.minimum(by, 0.0) ** 2, np.maximum(fy, 0.0) ** 2)
        gx2 = np.where(s0 > 0, pos_x, np.where(s0 < 0, neg_x, zero))
        gy2 = np.where(s0 > 0, pos_y, np.where(s0 < 0, neg_y, zero))
        phi = phi - dtau * (s0 * (np.sqrt(gx2 + gy2) - 1.0))
        if apply_phi_BCs_func is not None:
            phi = apply_phi_BCs_func(phi)
    return phi


# --------------------------------------------------------------------------
# output.py -- energy diagnostics; common.py:110-115 -- centroid
# --------------------------------------------------------------------------
def compute_kinetic_energy(a, b, rho_f, rho_s, phi, w_t, dx, dy):           # output.py:6-39
    H = smoothed_heaviside(phi, w_t)
    rho = (1 - H) * rho_s + H * rho_f
    return float(np.sum(0.5 * rho * (a ** 2 + b ** 2)) * dx * dy)


def compute_strain_energy(X1, X2, phi, mu_s, dx, dy, kappa=0.0):             # output.py:41-134
    """Central differences of the reference map with the edge value repeated outside the grid
    (np.pad mode='edge', width 4, cropped again), F = G^-1 on solid cells with |det G| > 1e-10."""
    def cdiff(f, axis, h):
        up = np.concatenate([np.take(f, range(1, f.shape[axis]), axis), np.take(f, [-1], axis)], axis)
        dn = np.concatenate([np.take(f, [0], axis), np.take(f, range(0, f.shape[axis] - 1), axis)], axis)
        return (up - dn) / (2.0 * h)
    G11, G12 = cdiff(X1, 1, dx), cdiff(X1, 0, dy)
    G21, G22 = cdiff(X2, 1, dx), cdiff(X2, 0, dy)
    det = G11 * G22 - G12 * G21
    good = (np.abs(det) > 1e-10) & (phi <= 0.0)
    d = np.where(good, det, 1.0)
    F11, F12, F21, F22 = G22 / d, -G12 / d, -G21 / d, G11 / d
    I1 = (F11 ** 2 + F21 ** 2) + (F12 ** 2 + F22 ** 2)
    dens = 0.5 * mu_s * (I1 - 2.0) + 0.5 * kappa * (1.0 / d - 1.0) ** 2
    return float(np.sum(np.where(good, dens, 0.0)) * dx * dy)


def compute_viscous_dissipation(a, b, mu_f, phi, w_t, dx, dy, eta_s=0.0):    # output.py:136-193
    Dxx = grad_central_x_2nd(a, dx)
    Dyy = grad_central_y_2nd(b, dy)
    Dxy = 0.5 * (grad_central_y_2nd(a, dy) + grad_central_x_2nd(b, dx))
    H = smoothed_heaviside(phi, w_t)
    mu = H * mu_f + (1 - H) * eta_s
    return float(np.sum(2.0 * mu * (Dxx ** 2 + Dyy ** 2 + 2.0 * Dxy ** 2)) * dx * dy)


def disc_centroid(phi, X, Y):                                               # benchmarks/common.py:110-115
    m = phi <= 0.0
    if not np.any(m):
        return float("nan"), float("nan")
    return float(X[m].mean()), float(Y[m].mean())


# --------------------------------------------------------------------------
# caller-side contract (benchmarks/common.py) used by drivers and tests
# --------------------------------------------------------------------------
def no_slip_lid_bc(u, v, lid_speed=1.0):               # benchmarks/common.py:27-37
    u = u.copy()
    v = v.copy()
    for f in (u, v):
        f[:, 0] = 0.0
        f[:, -1] = 0.0
        f[0, :] = 0.0
    u[-1, :] = lid_speed
    v[-1, :] = 0.0
    for f in (u, v):
        f[0, 0] = f[0, -1] = f[-1, 0] = f[-1, -1] = 0.0
    return u, v


def free_slip_box_bc(u, v):                            # benchmarks/common.py:40-50
    u = u.copy()
    v = v.copy()
    u[:, 0] = 0.0
    u[:, -1] = 0.0
    v[:, 0] = v[:, 1]
    v[:, -1] = v[:, -2]
    v[0, :] = 0.0
    v[-1, :] = 0.0
    u[0, :] = u[1, :]
    u[-1, :] = u[-2, :]
    return u, v


def periodic_bc(u, v):                                 # tests/test_poisson.py:60-64
    u = u.copy()
    v = v.copy()
    u[:, -1] = u[:, 0]
    v[:, -1] = v[:, 0]
    u[-1, :] = u[0, :]
    v[-1, :] = v[0, :]
    return u, v


def wall_bc(u, v):                                     # tests/test_poisson.py:39-43
    u = u.copy()
    v = v.copy()
    for f in (u, v):
        f[:, 0] = f[:, -1] = f[0, :] = f[-1, :] = 0.0
    return u, v


def initialize_disc(X, Y, x0, y0, R):                  # benchmarks/common.py:55-57
    return np.sqrt((X - x0) ** 2 + (Y - y0) ** 2) - R


def taylor_green_velocity(X, Y, U0=1.0):               # benchmarks/common.py:60-65
    k = 2.0 * np.pi
    return U0 * k * np.sin(k * X) * np.cos(k * Y), -U0 * k * np.cos(k * X) * np.sin(k * Y)


def disc_centroid(phi, X, Y):                          # benchmarks/common.py:110-115
    m = phi <= 0.0
    if not np.any(m):
        return np.nan, np.nan
    return X[m].mean(), Y[m].mean()


# --------------------------------------------------------------------------
# one full FSI step exactly as benchmarks/soft_disc_in_lid_driven.py:78-106
# --------------------------------------------------------------------------
def fsi_step(state, prm, dt=None):
    """Advance (a, b, p, X1, X2) by one step.  ``prm`` carries the scalars of
    the driver loop, the BC callable, the level-set callable, X, Y and eig.
    Returns (new_state, dt, extras)."""
    a, b, p, X1, X2 = state
    dx, dy = prm['dx'], prm['dy']
    if dt is None:
        dt = compute_timestep(a, b, dx, dy, prm['CFL'], prm['dt_cap'], prm['mu_s'], prm['rho_s'],
                              prm.get('gamma', 0.0), prm['rho_f'], mu_f=prm['mu_f'],
                              eta_s=prm['eta_s'], kappa=prm['kappa'])
    phi = rebuild_phi_from_reference_map(X1, X2, prm['phi_init'])
    solid = (phi <= 0).astype(float)
    sch, wc = prm['scheme'], prm.get('w_cut', 0.0)
    X1 = advect_reference_map(X1, a, b, prm['X'], prm['Y'], dt, dx, dy, phi, sch, wc) * solid
    X2 = advect_reference_map(X2, a, b, prm['X'], prm['Y'], dt, dx, dy, phi, sch, wc) * solid
    X1, X2 = extrapolate_reference_map(X1, X2, phi, dx, dy, prm['layers'])
    phi = rebuild_phi_from_reference_map(X1, X2, prm['phi_init'])
    a_s, b_s, sxx, sxy, syy, J = momentum_step_rk4(
        a, b, p, X1, X2, prm['bc'], prm['mu_s'], prm['kappa'], prm['eta_s'], dx, dy, dt,
        prm['rho_s'], prm['rho_f'], phi, prm['mu_f'], prm['w_t'], prm.get('gamma', 0.0))
    H = smoothed_heaviside(phi, prm['w_t'])
    rho_local = (1 - H) * prm['rho_s'] + H * prm['rho_f']
    a, b, p, _, _ = pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, prm['bc'],
                                            p_prev=p, eigenvalues=prm['eig'],
                                            bc_type=prm.get('bc_type', 'neumann'))
    return (a, b, p, X1, X2), dt, dict(phi=phi, sxx=sxx, sxy=sxy, syy=syy, J=J)

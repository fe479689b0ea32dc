"""Energy diagnostics of upstream ``pyRMT/output.py`` (compute_kinetic_energy
:6-39, compute_strain_energy :41-134, compute_viscous_dissipation :136-193) and the
solid centroid of ``benchmarks/common.py:110-115``.

They sit beside the timestep in every driver.  One device
kernel (rmt_diagnostics) forms every density and reduces it in a single pass over
the grid; only six doubles come back to the host.  ``diagnostics`` returns all of them
from one launch; the upstream-named functions are thin views of it.  HDF5/CSV writers
are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._runtime import F64, ctx, ptr, stream, to_dev

_work = {}


def _run(a, b, X1, X2, phi, X, Y, dx, dy, rho_f=1.0, rho_s=1.0, mu_f=0.0, mu_s=0.0, kappa=0.0, eta_s=0.0,
         w_t=None):
    ph = to_dev(phi)
    Ny, Nx = ph.shape
    dev = [None if t is None else to_dev(t) for t in (a, b, X1, X2, X, Y)]
    lib = ctx().lib
    key = ph.device.index
    if key not in _work:
        _work[key] = torch.empty(lib.rmt_diagnostics_workspace_doubles(), dtype=F64, device=ph.device)
    out = torch.empty(6, dtype=F64, device=ph.device)
    w = float(w_t) if w_t is not None else 2.0 * float(dx)
    _lib.check(lib.rmt_diagnostics(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), ptr(ph), ptr(dev[4]),
                                   ptr(dev[5]), Ny, Nx, float(dx), float(dy), float(rho_f), float(rho_s),
                                   float(mu_f), float(mu_s), float(kappa), float(eta_s), w, ptr(_work[key]),
                                   ptr(out), stream()), "rmt_diagnostics")
    return out.cpu().numpy()


def diagnostics(a, b, X1, X2, phi, dx, dy, rho_f, rho_s, mu_f, mu_s, w_t, kappa=0.0, eta_s=0.0, X=None, Y=None):
    """All diagnostics of one state from ONE kernel launch: dict with kinetic_energy, strain_energy,
    viscous_dissipation, solid_cells, centroid (x, y) (NaN without solid cells)."""
    s = _run(a, b, X1, X2, phi, X, Y, dx, dy, rho_f, rho_s, mu_f, mu_s, kappa, eta_s, w_t)
    cell = float(dx) * float(dy)
    n = s[3]
    cen = (s[4] / n, s[5] / n) if n > 0 else (float("nan"), float("nan"))
    return dict(kinetic_energy=float(s[0] * cell), strain_energy=float(s[1] * cell),
                viscous_dissipation=float(s[2] * cell), solid_cells=int(n), centroid=cen)


def compute_kinetic_energy(a, b, rho_f, rho_s, phi, w_t, dx, dy):
    """KE = sum(0.5 * rho_local * (a^2 + b^2)) dx dy  (output.py:6-39)."""
    return float(_run(a, b, None, None, phi, None, None, dx, dy, rho_f=rho_f, rho_s=rho_s, w_t=w_t)[0] * dx * dy)


def compute_strain_energy(X1, X2, phi, mu_s, dx, dy, kappa=0.0):
    """SE = sum over solid cells of 0.5 mu_s (I1 - 2) + 0.5 kappa (J - 1)^2  (output.py:41-134):
    central gradients of the edge-padded reference map, F = G^-1, I1 = tr(F^T F)."""
    return float(_run(None, None, X1, X2, phi, None, None, dx, dy, mu_s=mu_s, kappa=kappa)[1] * dx * dy)


def compute_viscous_dissipation(a, b, mu_f, phi, w_t, dx, dy, eta_s=0.0):
    """eps = sum(2 mu_local (Dxx^2 + Dyy^2 + 2 Dxy^2)) dx dy  (output.py:136-193)."""
    return float(_run(a, b, None, None, phi, None, None, dx, dy, mu_f=mu_f, eta_s=eta_s, w_t=w_t)[2] * dx * dy)


def disc_centroid(phi, X, Y):
    """benchmarks/common.py:110-115 -- mean node coordinates of the cells with phi <= 0."""
    dx = 1.0
    s = _run(None, None, None, None, phi, X, Y, dx, dx)
    if s[3] <= 0:
        return float("nan"), float("nan")
    return float(s[4] / s[3]), float(s[5] / s[3])


def output_simulation_data(*args, **kwargs):
    """output.py:213-321 (HDF5/CSV dumps) -- out of scope (needs h5py)."""
    raise NotImplementedError("HDF5/CSV snapshot output is out of scope for the B200 path")

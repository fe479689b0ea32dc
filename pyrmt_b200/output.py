"""Energy diagnostics of upstream ``pyRMT/output.py`` (compute_kinetic_energy
:6-39, compute_strain_energy :41-134, compute_viscous_dissipation :136-193).

They sit beside the timestep in every driver but are OFF the timed path (SURVEY
2 #14, 8f rank 2): gradients and the Heaviside come from the device operators,
the final elementwise products and sums are evaluated with torch/NumPy on
whatever array type the caller uses.  HDF5/CSV writers are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from .functions import smoothed_heaviside
from .utils import grad_central_x_2nd, grad_central_y_2nd


def _sum(x):
    return float(x.sum().item()) if isinstance(x, torch.Tensor) else float(np.sum(x))


def compute_kinetic_energy(a, b, rho_f, rho_s, phi, w_t, dx, dy):
    """KE = sum(0.5 * rho_local * (a^2 + b^2)) dx dy  (output.py:6-39)."""
    H = smoothed_heaviside(phi, w_t)
    rho_local = (1 - H) * rho_s + H * rho_f
    return _sum(0.5 * rho_local * (a**2 + b**2)) * dx * dy


def _edge_pad(x, w):
    if isinstance(x, torch.Tensor):
        return torch.nn.functional.pad(x[None, None], (w, w, w, w), mode="replicate")[0, 0].contiguous()
    return np.pad(x, w, mode="edge")


def compute_strain_energy(X1, X2, phi, mu_s, dx, dy, kappa=0.0):
    """SE = sum over solid cells of 0.5 mu_s (I1 - 2) + 0.5 kappa (J - 1)^2  (output.py:41-134):
    central gradients of the edge-padded reference map, F = G^-1, I1 = tr(F^T F)."""
    w = 4
    g = lambda f, fn, h: fn(_edge_pad(f, w), h)[w:-w, w:-w]
    G11, G12 = g(X1, grad_central_x_2nd, dx), g(X1, grad_central_y_2nd, dy)
    G21, G22 = g(X2, grad_central_x_2nd, dx), g(X2, grad_central_y_2nd, dy)
    detG = G11 * G22 - G12 * G21
    xp = torch if isinstance(phi, torch.Tensor) else np
    good = (xp.abs(detG) > 1e-10) & (phi <= 0.0)
    safe = xp.where(good, detG, xp.ones_like(detG))
    F11, F12, F21, F22 = G22 / safe, -G12 / safe, -G21 / safe, G11 / safe
    I1 = (F11**2 + F21**2) + (F12**2 + F22**2)
    dens = 0.5 * mu_s * (I1 - 2.0) + 0.5 * kappa * (1.0 / safe - 1.0) ** 2
    return _sum(xp.where(good, dens, xp.zeros_like(dens))) * dx * dy


def compute_viscous_dissipation(a, b, mu_f, phi, w_t, dx, dy, eta_s=0.0):
    """eps = sum(2 mu_local (Dxx^2 + Dyy^2 + 2 Dxy^2)) dx dy  (output.py:136-193)."""
    Dxx, Dyy = grad_central_x_2nd(a, dx), grad_central_y_2nd(b, dy)
    Dxy = 0.5 * (grad_central_y_2nd(a, dy) + grad_central_x_2nd(b, dx))
    H = smoothed_heaviside(phi, w_t)
    mu_local = H * mu_f + (1 - H) * eta_s
    return _sum(2.0 * mu_local * (Dxx**2 + Dyy**2 + 2.0 * Dxy**2)) * dx * dy


def output_simulation_data(*args, **kwargs):
    """output.py:213-321 (HDF5/CSV dumps) -- out of scope (needs h5py)."""
    raise NotImplementedError("HDF5/CSV snapshot output is out of scope for the B200 path")

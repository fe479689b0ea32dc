"""Energy diagnostics of upstream ``pyRMT/output.py`` (compute_kinetic_energy
:6-39, compute_strain_energy :41-134, compute_viscous_dissipation :136-193) and the
solid centroid of ``benchmarks/common.py:110-115``.

They sit beside the timestep in every driver.  One device
kernel (rmt_diagnostics) forms every density and reduces it in a single pass over
the grid; only six doubles come back to the host.  ``diagnostics`` returns all of them
from one launch; the upstream-named functions are thin views of it.
``output_simulation_data`` is upstream's snapshot writer on top of it (HDF5 when h5py is
installed, .npz with the same names otherwise).
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


def divergence_2d_interior(u, v, dx, dy, pad=3):
    """output.py:195-211 -- central divergence with `pad` rim cells left at zero; returns (full field,
    interior view).  Device: rmt_divergence (same formula on every interior node) + rim zeroing."""
    from .functions import _compute_divergence
    full = _compute_divergence(u, v, dx, dy)
    if isinstance(full, torch.Tensor):
        out = torch.zeros_like(full)
    else:
        out = np.zeros_like(full)
    out[pad:-pad, pad:-pad] = full[pad:-pad, pad:-pad]
    return out, out[pad:-pad, pad:-pad]


def _host(t):
    """Device tensor -> ndarray through a pinned staging buffer (asynchronous copy, one sync by the caller)."""
    if not isinstance(t, torch.Tensor):
        return np.asarray(t)
    if t.device.type != "cuda":
        return t.numpy()
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    return buf


def output_simulation_data(dx, dy, phi, solid_mask, X1, X2, a, b, p, vis_output_freq, directory_name, step, dt,
                           sigma_sxx, sigma_sxy, sigma_syy, J, mu_s=0.0, mu_f=0.0, rho_s=1.0, rho_f=1.0, w_t=None,
                           eta_s=0.0, kappa=0.0, time=0.0, integrated_dissipation=0.0):
    """output.py:213-321 -- every `vis_output_freq` steps (and at step 1): the log line, one row of
    outputs/<directory_name>/energy_history.csv and a snapshot of the fields.

    The energies come from ONE device kernel (rmt_diagnostics); the fields go to the host through pinned
    staging buffers with asynchronous copies (a single synchronisation) and are written as
    data_<step>.h5 with upstream's dataset / attribute names when h5py is importable, else as
    data_<step>.npz with the same names (attributes as 0-d arrays)."""
    import csv
    import os
    if w_t is None:
        w_t = 2.0 * dx
    if not (step % vis_output_freq == 0 or step == 1):
        return integrated_dissipation
    d = diagnostics(a, b, X1, X2, phi, dx, dy, rho_f, rho_s, mu_f, mu_s, w_t, kappa=kappa, eta_s=eta_s)
    ke, se, diss = d["kinetic_energy"], d["strain_energy"], d["viscous_dissipation"]
    total = ke + se + integrated_dissipation
    c = ctx()
    ad, bd = to_dev(a), to_dev(b)
    vmax = float(c.max_speed(ad, bd)[0].item())
    div_field, div_int = divergence_2d_interior(ad, bd, dx, dy, pad=4)
    st = c.stats(div_int.contiguous()).cpu()
    div_max = max(abs(float(st[1])), abs(float(st[2])))
    Jmin = float(c.stats(to_dev(J))[1].item())
    sx, sq, sy = to_dev(sigma_sxx), to_dev(sigma_sxy), to_dev(sigma_syy)
    smax = float(torch.sqrt(sx * sx + sy * sy + 2.0 * sq * sq).max().item())     # log line only
    print(f"[Step {step:05d}] t={time:.3f}, dt={dt:.2e}, max|v|={vmax:.3f}, KE={ke:.4e}, SE={se:.4e}, "
          f"\u03b5={diss:.4e}, E_tot={total:.4e}, min(J)={Jmin:.3f}, max|\u03c3|={smax:.2f}, max|div|={div_max:.2e}")
    out_dir = os.path.join("outputs", directory_name)
    os.makedirs(out_dir, exist_ok=True)
    energy_file = os.path.join(out_dir, "energy_history.csv")
    exists = os.path.isfile(energy_file)
    with open(energy_file, "a", newline="") as fh:
        names = ["step", "time", "dt", "kinetic_energy", "strain_energy", "dissipation_rate",
                 "integrated_dissipation", "total_energy"]
        w = csv.DictWriter(fh, fieldnames=names)
        if not exists or step == 1:
            w.writeheader()
        w.writerow(dict(zip(names, (step, time, dt, ke, se, diss, integrated_dissipation, total))))
    fields = dict(phi=phi, X1=X1, X2=X2, J=J, a=a, b=b, p=p, sigma_xx=sigma_sxx, sigma_yy=sigma_syy,
                  sigma_xy=sigma_sxy, div_vel=div_field)
    staged = {k: _host(v) for k, v in fields.items()}
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    arrays = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in staged.items()}
    attrs = dict(time=time, kinetic_energy=ke, strain_energy=se, dissipation_rate=diss,
                 integrated_dissipation=integrated_dissipation, total_energy=total)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None and hasattr(h5py, "File"):
        with h5py.File(os.path.join(out_dir, f"data_{step:06d}.h5"), "w") as f:
            for k, v in arrays.items():
                f.create_dataset(k, data=v)
            for k, v in attrs.items():
                f.attrs[k] = v
    else:
        np.savez(os.path.join(out_dir, f"data_{step:06d}.npz"), **arrays, **{k: np.asarray(v) for k, v in attrs.items()})
    return integrated_dissipation

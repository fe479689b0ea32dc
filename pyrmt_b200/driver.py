"""The FSI time loop of the reference's drivers, written against the drop-in
operator API, plus the synthetic multi-disc cases of SURVEY 8(d).

``fsi_step`` is benchmarks/soft_disc_in_lid_driven.py:78-106 (identical in
disc_in_taylor_green.py:78-106) line for line; it works on ndarrays (host
round trips inside every operator) or on CUDA tensors (device resident).
"""
from __future__ import annotations

import numpy as np
import torch

from . import functions as F
from .bc import free_slip_box_bc, no_slip_lid_bc, periodic_bc
from .levelset import DiscSDF


def fsi_step(state, prm, dt=None):
    """Advance (a, b, p, X1, X2) by one step.  Returns (state, dt, extras)."""
    a, b, p, X1, X2 = state
    dx, dy = prm["dx"], prm["dy"]
    # compute_timestep needs a host value: its reduction is queued first, the level set (which does not need
    # dt) behind it, and only then the host waits -- the GPU works through the round trip
    pending = F.compute_timestep_begin(a, b) if dt is None else None
    phi = F.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"])
    phi = F.reinitialize_level_set(phi, dx, dy, "none")
    if pending is not None:
        dt = F.compute_timestep_end(pending, dx, dy, prm["CFL"], prm["dt_cap"], prm["mu_s"], prm["rho_s"],
                                    prm.get("gamma", 0.0), prm["rho_f"], mu_f=prm["mu_f"], eta_s=prm["eta_s"],
                                    kappa=prm["kappa"])
    sch, wc = prm["scheme"], prm.get("w_cut", 0.0)
    if prm.get("fuse_pair", True):      # xi1, xi2 and the "* solid_mask" in one pass (bitwise identical)
        X1, X2 = F.advect_reference_map_pair(X1, X2, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, sch, wc,
                                             mask_solid=True)
    else:                               # the reference's loop, operator by operator
        X1 = F.mask_solid(F.advect_reference_map(X1, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, sch, wc), phi)
        X2 = F.mask_solid(F.advect_reference_map(X2, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, sch, wc), phi)
    X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, prm["layers"], inplace=bool(prm.get("fuse_pair", True)))
    # level set of the extrapolated map + the stress the predictor needs, in one pass (two operators upstream)
    phi, stress = F.rebuild_phi_and_stress(X1, X2, prm["phi_init"], dx, dy, prm["mu_s"], prm["kappa"], prm["w_t"])
    a_s, b_s, sxx, sxy, syy, J = F.momentum_step_rk4_with_stress(
        a, b, p, X1, X2, prm["bc"], prm["mu_s"], prm["kappa"], prm["eta_s"], dx, dy, dt, prm["rho_s"],
        prm["rho_f"], phi, prm["mu_f"], prm["w_t"], prm.get("gamma", 0.0), stress=stress)
    rho_local = _density(phi, prm)
    a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, prm["bc"], p_prev=p,
                                              eigenvalues=prm["eig"], bc_type=prm.get("bc_type", "neumann"))
    return (a, b, p, X1, X2), dt, dict(phi=phi, sxx=sxx, sxy=sxy, syy=syy, J=J)


def _density(phi, prm):
    """rho_local = (1 - H) rho_s + H rho_f (soft_disc_in_lid_driven.py:102-103).  With equal densities the
    field is the constant rho_f to rounding and np.ptp(rho) <= 1e-10 sends the projection down its
    constant-density branch (functions.py:1298) either way: the scalar is passed instead of a field."""
    if float(prm["rho_s"]) == float(prm["rho_f"]):
        return float(prm["rho_f"])
    return F.heaviside_and_density(phi, prm["w_t"], prm["rho_s"], prm["rho_f"])[1]


_copy_streams = {}


def fsi_step_host(host_state, prm, dt=None):
    """One step with the state living in pinned HOST tensors -- the end-to-end path bench.py times.
    The same operator sequence as `fsi_step`, with the PCIe traffic overlapped: the five fields go up
    on a copy stream in the order the step consumes them (a, b for the time step; xi1, xi2 for the
    advection; p only before the predictor), and xi1, xi2 start their way back on a second copy
    stream as soon as the extrapolation is done, while the predictor and the projection still run."""
    dev = torch.device("cuda", torch.cuda.current_device())
    key = dev.index
    if key not in _copy_streams:
        _copy_streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    s_in, s_out = _copy_streams[key]
    main = torch.cuda.current_stream(dev)
    ha, hb, hp, h1, h2 = host_state
    s_in.wait_stream(main)
    with torch.cuda.stream(s_in):
        a, b = ha.to(dev, non_blocking=True), hb.to(dev, non_blocking=True)
        ev_ab = s_in.record_event()
        X1, X2 = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)
        ev_x = s_in.record_event()
        p = hp.to(dev, non_blocking=True)
        ev_p = s_in.record_event()
    for t in (a, b, p, X1, X2):
        t.record_stream(main)
    dx, dy = prm["dx"], prm["dy"]
    main.wait_event(ev_ab)
    pending = F.compute_timestep_begin(a, b) if dt is None else None
    main.wait_event(ev_x)
    phi = F.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"])
    if pending is not None:
        dt = F.compute_timestep_end(pending, dx, dy, prm["CFL"], prm["dt_cap"], prm["mu_s"], prm["rho_s"],
                                    prm.get("gamma", 0.0), prm["rho_f"], mu_f=prm["mu_f"], eta_s=prm["eta_s"],
                                    kappa=prm["kappa"])
    X1, X2 = F.advect_reference_map_pair(X1, X2, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, prm["scheme"],
                                         prm.get("w_cut", 0.0), mask_solid=True)
    X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, prm["layers"], inplace=True)   # fresh pair tensors
    ev_xi = main.record_event()
    out = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in host_state]
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_xi)
        out[3].copy_(X1, non_blocking=True)
        out[4].copy_(X2, non_blocking=True)
    X1.record_stream(s_out)
    X2.record_stream(s_out)
    phi, stress = F.rebuild_phi_and_stress(X1, X2, prm["phi_init"], dx, dy, prm["mu_s"], prm["kappa"], prm["w_t"])
    main.wait_event(ev_p)
    a_s, b_s, *_ = F.momentum_step_rk4_with_stress(a, b, p, X1, X2, prm["bc"], prm["mu_s"], prm["kappa"], prm["eta_s"], dx, dy,
                                       dt, prm["rho_s"], prm["rho_f"], phi, prm["mu_f"], prm["w_t"],
                                       prm.get("gamma", 0.0), stress=stress)
    rho_local = _density(phi, prm)
    a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, prm["bc"], p_prev=p,
                                              eigenvalues=prm["eig"], bc_type=prm.get("bc_type", "neumann"))
    ev_done = main.record_event()
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_done)
        for h, d in zip(out[:3], (a, b, p)):
            h.copy_(d, non_blocking=True)
    for t in (a, b, p):
        t.record_stream(s_out)
    s_out.synchronize()
    main.wait_stream(s_out)
    return tuple(out)


class LidBC:
    """no_slip_lid_bc with a fixed lid speed (a picklable callable)."""

    def __init__(self, lid_speed=1.0):
        self.lid_speed = float(lid_speed)

    def __call__(self, u, v):
        return no_slip_lid_bc(u, v, self.lid_speed)

    def rmt_table(self, Ny, Nx):
        """The gather table of this BC written down directly (what bc.classify would find by
        probing; saves the host-side probe arrays on 8193^2 / 16385^2 grids)."""
        from .bc import BCTable, _FIELD_BIT
        j = np.concatenate([np.zeros(Nx, np.int64), np.full(Nx, Ny - 1, np.int64),
                            np.arange(1, Ny - 1), np.arange(1, Ny - 1)])
        i = np.concatenate([np.arange(Nx), np.arange(Nx), np.zeros(Ny - 2, np.int64),
                            np.full(Ny - 2, Nx - 1, np.int64)])
        cell = j * Nx + i
        lid = (j == Ny - 1) & (i > 0) & (i < Nx - 1)
        dst = np.concatenate([cell, cell | _FIELD_BIT]).astype(np.int64)
        cb = np.concatenate([np.where(lid, self.lid_speed, 0.0), np.zeros(cell.size)])
        return BCTable(dst, np.full(dst.size, -1, np.int64), np.zeros(dst.size), cb.astype(np.float64), (Ny, Nx))


class PeriodicBC:
    """tests/test_poisson.py:60-64 (`_periodic_bc`) as a callable that also writes its gather table down
    directly (bc.classify probes with full host arrays: too much at 16385^2)."""

    def __call__(self, u, v):
        from .bc import periodic_bc
        return periodic_bc(u, v)

    def rmt_table(self, Ny, Nx):
        from .bc import BCTable, _FIELD_BIT
        j = np.arange(Ny - 1)
        i = np.arange(Nx - 1)
        dst = np.concatenate([j * Nx + (Nx - 1), (Ny - 1) * Nx + i, [(Ny - 1) * Nx + Nx - 1]]).astype(np.int64)
        src = np.concatenate([j * Nx, i, [0]]).astype(np.int64)
        dst = np.concatenate([dst, dst | _FIELD_BIT])
        src = np.concatenate([src, src | _FIELD_BIT])
        return BCTable(dst, src, np.ones(dst.size), np.zeros(dst.size), (Ny, Nx))


def disc_lattice(k_side, L, R_frac, seed=20240607, jitter=0.01):
    """K = k_side^2 discs on a jittered lattice (jitter 0.01 L on the 8 x 8
    lattice; finer lattices scale it with the spacing so the discs stay disjoint)."""
    rng = np.random.default_rng(seed)
    m, n = np.meshgrid(np.arange(k_side), np.arange(k_side))
    jx = rng.uniform(-jitter, jitter, size=m.shape)
    jy = rng.uniform(-jitter, jitter, size=m.shape)
    cx = ((m + 0.5) / k_side + jx) * L
    cy = ((n + 0.5) / k_side + jy) * L
    return cx.ravel(), cy.ravel(), np.full(cx.size, R_frac * L)


def make_case(N, L=1.0, k_side=8, R_frac=0.04, scheme="weno5", bc_kind="lid", device=None,
              mu_s=0.1, kappa=0.0, rho=1.0, eta_s=0.01, mu_f=0.01, CFL=0.2, dt_cap=1e-3, layers=3,
              taylor_green_U0=None, as_numpy=False):
    """Synthetic multi-disc FSI case on an N x N node grid.

    Returns (state, prm) with device-resident fp64 tensors (or ndarrays)."""
    X, Y, dx, dy = F.create_grid(N, N, L, L)
    cx, cy, R = disc_lattice(k_side, L, R_frac)
    sdf = DiscSDF(cx, cy, R, domain=(L, L))
    if bc_kind == "lid":
        bc, bc_type = LidBC(1.0), "neumann"
        eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    elif bc_kind == "free_slip":
        bc, bc_type = free_slip_box_bc, "neumann"
        eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    elif bc_kind == "periodic":
        bc, bc_type = periodic_bc, "periodic"
        eig = F._precompute_poisson_eigenvalues_periodic(N, N, dx, dy)
    else:
        raise ValueError("bc_kind must be 'lid', 'free_slip' or 'periodic'")
    if as_numpy:
        up = lambda arr: np.ascontiguousarray(arr, dtype=np.float64)
    else:
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        up = lambda arr: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)
    Xd, Yd = up(X), up(Y)
    phi0 = sdf(Xd, Yd)
    X1 = F.mask_solid(Xd, phi0)
    X2 = F.mask_solid(Yd, phi0)
    X1, X2 = F.extrapolate_reference_map(X1, X2, phi0, dx, dy, layers)
    if taylor_green_U0 is not None:
        k = 2.0 * np.pi / L
        a0 = taylor_green_U0 * k * np.sin(k * X) * np.cos(k * Y)
        b0 = -taylor_green_U0 * k * np.cos(k * X) * np.sin(k * Y)
    else:
        a0, b0 = np.zeros_like(X), np.zeros_like(X)
    a, b = bc(a0, b0)
    state = (up(a), up(b), up(np.zeros_like(X)), X1, X2)
    prm = dict(dx=dx, dy=dy, CFL=CFL, dt_cap=dt_cap, mu_s=mu_s, kappa=kappa, rho_s=rho, rho_f=rho,
               eta_s=eta_s, mu_f=mu_f, w_t=2.0 * dx, layers=layers, scheme=scheme, w_cut=0.0,
               phi_init=sdf, bc=bc, bc_type=bc_type, eig=eig, X=Xd, Y=Yd, N=N, L=L,
               discs=(cx, cy, R))
    return state, prm

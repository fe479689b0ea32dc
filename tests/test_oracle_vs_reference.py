"""Live cross-check of the oracle against the reference itself.

Only runs where the upstream sources are mounted (/root/reference in the build
container); on the GPU box the reference does not exist and this file skips --
there the committed golden vectors (test_oracle_golden.py) carry the pin.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

REF = os.environ.get("PYRMT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyRMT")),
                                reason="reference sources not mounted")


@pytest.fixture(scope="module")
def ref():
    stubs = tempfile.mkdtemp(prefix="pyrmt_stubs_")
    for name in ("pyamg", "h5py"):
        os.makedirs(os.path.join(stubs, name))
        open(os.path.join(stubs, name, "__init__.py"), "w").close()
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    sys.path[:0] = [stubs, REF]
    import pyRMT.functions as F
    yield F
    for p in (stubs, REF):
        sys.path.remove(p)


def test_extrapolation_bitwise_n256(ref):
    from oracle import rmt_oracle as O
    N = 256
    X, Y, dx, dy = ref.create_grid(N, N, 1.0, 1.0)
    phi = np.sqrt((X - 0.55) ** 2 + (Y - 0.45) ** 2) - 0.21
    m = (phi <= 0).astype(float)
    X1 = (X + 0.02 * np.sin(3 * X + Y)) * m
    X2 = (Y - 0.015 * np.cos(2 * X - 3 * Y)) * m
    r1, r2 = ref.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
    o1, o2 = O.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
    assert np.array_equal(r1, o1) and np.array_equal(r2, o2)


def test_libm_exp_is_numba_exp(ref):
    """H2: Numba's scalar np.exp is libm exp, i.e. what the oracle calls."""
    import numba
    from oracle import rmt_oracle as O

    @numba.njit
    def nexp(x):
        y = np.empty_like(x)
        for k in range(x.size):
            y[k] = np.exp(x[k])
        return y
    x = -np.random.default_rng(1).random(200000)
    assert np.array_equal(nexp(x), O.libm_exp(x))


def test_fsi_steps_n128(ref):
    """Five steps of the config-2 loop at N=128: identical state trajectories."""
    from oracle import rmt_oracle as O
    N = 128
    X, Y, dx, dy = ref.create_grid(N, N, 1.0, 1.0)
    phi0 = lambda A, B: np.sqrt((A - 0.6) ** 2 + (B - 0.5) ** 2) - 0.2
    lid = lambda u, v: O.no_slip_lid_bc(u, v, 1.0)
    eig = ref._precompute_poisson_eigenvalues(N, N, dx, dy)
    phi = ref.apply_phi_BCs(phi0(X, Y))
    m = (phi <= 0).astype(float)
    X1, X2 = ref.extrapolate_reference_map(X * m, Y * m, phi, dx, dy, 3)
    a = np.zeros((N, N)); b = np.zeros((N, N)); p = np.zeros((N, N))
    prm = dict(dx=dx, dy=dy, X=X, Y=Y, eig=eig, mu_s=0.1, kappa=0.0, rho_s=1.0, eta_s=0.01,
               mu_f=0.01, rho_f=1.0, w_t=2 * dx, layers=3, CFL=0.2, dt_cap=1e-3,
               scheme='semilagrangian', bc=lid, phi_init=phi0)
    so = (a, b, p, X1, X2)
    for n in range(5):
        dt = ref.compute_timestep(a, b, dx, dy, 0.2, 1e-3, 0.1, 1.0, 0.0, 1.0, mu_f=0.01,
                                  eta_s=0.01, kappa=0.0)
        phi = ref.rebuild_phi_from_reference_map(X1, X2, phi0)
        m = (phi <= 0).astype(float)
        X1 = ref.advect_reference_map(X1, a, b, X, Y, dt, dx, dy, phi, 'semilagrangian', 0.0) * m
        X2 = ref.advect_reference_map(X2, a, b, X, Y, dt, dx, dy, phi, 'semilagrangian', 0.0) * m
        X1, X2 = ref.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
        phi = ref.rebuild_phi_from_reference_map(X1, X2, phi0)
        a_s, b_s, *_ = ref.momentum_step_rk4(a, b, p, X1, X2, lid, 0.1, 0.0, 0.01, dx, dy, dt,
                                             1.0, 1.0, phi, 0.01, 2 * dx, 0.0)
        H = ref.smoothed_heaviside(phi, 2 * dx)
        a, b, p, _, _ = ref.pressure_projection_amg(a_s, b_s, dx, dy, dt, (1 - H) + H, lid,
                                                    p_prev=p, eigenvalues=eig)
        so, dto, _ = O.fsi_step(so, prm)
        assert dto == dt
        for x, y in zip(so, (a, b, p, X1, X2)):
            assert np.array_equal(x, y)


def test_next_rows_against_live_reference(ref):
    """SURVEY 8f rows on seeded inputs larger than the golden fixtures: two-solid step + contact force
    (functions.py:765-895), PDE reinitialisation (:1369-1411), energy diagnostics (output.py:6-193)."""
    import pyRMT.output as OUT
    from oracle import rmt_oracle as O
    N = 96
    X, Y, dx, dy = ref.create_grid(N, N, 1.0, 1.0)
    disc = lambda x0, y0, R: np.sqrt((X - x0) ** 2 + (Y - y0) ** 2) - R
    pa, pb = disc(0.36, 0.5, 0.15), disc(0.655, 0.52, 0.15)
    ma, mb = (pa <= 0).astype(float), (pb <= 0).astype(float)
    X1a, X2a = ref.extrapolate_reference_map((X + 0.01 * np.sin(4 * Y)) * ma, Y * 0.98 * ma, pa, dx, dy, 3)
    X1b, X2b = ref.extrapolate_reference_map(X * 1.02 * mb, (Y - 0.01 * X) * mb, pb, dx, dy, 3)
    rng = np.random.default_rng(3)
    u = 0.3 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + 0.01 * rng.standard_normal((N, N))
    v = -0.2 * np.cos(np.pi * X) * np.sin(2 * np.pi * Y)
    p = 0.1 * np.cos(np.pi * X * Y)
    bc = lambda a, b: O.no_slip_lid_bc(a, b, 1.0)
    fr, fo = ref.compute_contact_force(pa, pb, 2.0, 3 * dx, dx, dy), O.compute_contact_force(pa, pb, 2.0, 3 * dx, dx, dy)
    assert np.abs(fr[0]).max() > 0 and np.allclose(fr[0], fo[0], rtol=0, atol=1e-13 * np.abs(fr[0]).max())
    assert np.allclose(fr[1], fo[1], rtol=0, atol=1e-13 * np.abs(fr[1]).max())
    args = (u, v, p, X1a, X2a, X1b, X2b, bc, 0.9, 0.4, 0.0, dx, dy, 1e-3, 1.2, 1.0, pa, pb, 0.02, 2 * dx)
    rr = ref.momentum_step_rk4_2solids(*args, k_rep=2.0, w_c=3 * dx)
    ro = O.momentum_step_rk4_2solids(*args, k_rep=2.0, w_c=3 * dx)
    for x, y in zip(rr, ro):
        assert np.max(np.abs(x - y)) <= 1e-13 * np.max(np.abs(x))
    phi = pa * (1.0 + 0.5 * np.sin(6 * X) * np.cos(5 * Y))
    assert np.array_equal(ref.reinitialize_phi_PDE(phi, dx, dy, 15, ref.apply_phi_BCs, 0.3),
                          O.reinitialize_phi_PDE(phi, dx, dy, 15, O.apply_phi_BCs, 0.3))
    w_t = 2 * dx
    rel = lambda a, b: abs(a - b) / abs(a)
    assert rel(OUT.compute_kinetic_energy(u, v, 1.0, 1.3, pa, w_t, dx, dy),
               O.compute_kinetic_energy(u, v, 1.0, 1.3, pa, w_t, dx, dy)) < 1e-13
    assert rel(OUT.compute_strain_energy(X1a, X2a, pa, 0.7, dx, dy, kappa=0.2),
               O.compute_strain_energy(X1a, X2a, pa, 0.7, dx, dy, kappa=0.2)) < 1e-12
    assert rel(OUT.compute_viscous_dissipation(u, v, 0.02, pa, w_t, dx, dy, eta_s=0.03),
               O.compute_viscous_dissipation(u, v, 0.02, pa, w_t, dx, dy, eta_s=0.03)) < 1e-13

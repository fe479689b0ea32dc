"""CPU-only tests: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), and the host-side logic of the
drop-in layer (BC classification, disc binning, signatures, error behaviour).
"""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from pyrmt_b200 import build, _lib
    build.build()                      # nvcc cross-compiles without a GPU
    return _lib.load()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "rmt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rmt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    from pyrmt_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.rmt_abi_version() == 1            # pure host call
    assert built_lib.rmt_reduce_workspace_doubles() > 0


def test_library_is_sm100a_only(built_lib):
    import subprocess
    from pyrmt_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyrmt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("SURVEY", ""), f


def test_signatures_match_reference_names():
    """Names / positional order / defaults of the drop-in API.  The
    expected strings were transcribed from the reference's def lines."""
    import pyrmt_b200.functions as F
    import pyrmt_b200.interpolators as I
    import pyrmt_b200.utils as U
    expect = {
        F.create_grid: "(Nx, Ny, Lx, Ly)",
        F.apply_phi_BCs: "(phi)",
        F.extrapolate_reference_map: "(X1, X2, phi, dx, dy, max_layers, row_offset=0, inplace=False)",
        F.compute_timestep: "(a, b, dx, dy, CFL, dt_min_cap, mu_s, rho_s, gamma, rho_f, mu_f=0.0, eta_s=0.0, kappa=0.0)",
        F.advect_semilagrangian_rk4: "(q, a, b, X, Y, dt, dx, dy)",
        F.advect_semilagrangian_cubic_rk4: "(q, a, b, X, Y, dt, dx, dy)",
        F.advect_weno5_rk3: "(q, a, b, dx, dy, dt, phi, w_cut=0.0)",
        F.advect_central2_rk3: "(q, a, b, dx, dy, dt, phi, w_cut=0.0)",
        F.advect_conservative_rk3: "(q, a, b, dx, dy, dt, phi, w_cut=0.0)",
        F.advect_reference_map: "(q, a, b, X, Y, dt, dx, dy, phi, scheme='semilagrangian', w_cut=0.0)",
        F.solid_cauchy_stress: "(X1, X2, dx, dy, mu_s, kappa, phi, w_cut=0.0, detg_clamp=0.0, isochoric=False)",
        F.smoothed_heaviside: "(x, w_t)",
        F.momentum_step_rk4: "(u, v, p, X1, X2, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f, phi, mu_f, w_t, gamma=0.0, stress_band=False, detg_clamp=3.0)",
        F.compute_curvature: "(phi, dx, dy)",
        F.momentum_step_rk4_2solids: "(u, v, p, X1a, X2a, X1b, X2b, velocity_bc, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f, phi_a, phi_b, mu_f, w_t, k_rep=0.0, w_c=None, detg_clamp=4.0)",
        F.compute_contact_force: "(phi1, phi2, k_rep, w_c, dx, dy)",
        F.velocity_rhs_blended_optimized: "(u, v, p, sigma_sxx, sigma_sxy, sigma_syy, dx, dy, phi, mu_f, H, dH_dx, dH_dy, rho_local, st_force_x, st_force_y)",
        F.apply_velocity_BCs: "(bc, u, v)",
        F.build_poisson_matrix: "(Nx, Ny, dx, dy)",
        F._compute_divergence: "(a_star, b_star, dx, dy)",
        F._compute_divergence_rc: "(a_star, b_star, p_prev, dt, rho, dx, dy)",
        F._compute_pressure_gradient: "(p, dx, dy)",
        F._precompute_poisson_eigenvalues: "(Nx, Ny, dx, dy)",
        F._solve_poisson_dct: "(rhs_2d, eigenvalues)",
        F._precompute_poisson_eigenvalues_periodic: "(Nx, Ny, dx, dy)",
        F._solve_poisson_fft: "(rhs_full, eigenvalues_periodic)",
        F.pressure_projection_amg: "(a_star, b_star, dx, dy, dt, rho, velocity_bc, A=None, ml=None, p_prev=None, eigenvalues=None, bc_type='neumann')",
        F.rebuild_phi_from_reference_map: "(X1, X2, phi_init_func)",
        F.reinitialize_phi_PDE: "(phi_in, dx, dy, num_iters, apply_phi_BCs_func, dt_reinit_factor=0.5)",
        F.reinitialize_level_set: "(phi, dx, dy, method='none', num_iters=20, dt_reinit_factor=0.2, apply_phi_BCs_func=None)",
        I.bilinear_interpolate: "(u, xq, yq, dx, dy, Nx, Ny)",
        I.bicubic_interpolate: "(u, xq, yq, dx, dy, Nx, Ny)",
        I.cubic_convolution: "(v0, v1, v2, v3, x)",
        U.grad_central_x_2nd: "(f, dx)",
        U.grad_central_y_2nd: "(f, dy)",
        U.diff_upwind_3rd: "(f, u, h, axis)",
        U.lap_2nd: "(f, dx, dy)",
        U.fast_solve_3x3: "(A, b)",
    }
    for fn, sig in expect.items():
        assert str(inspect.signature(fn)) == sig, fn.__name__


def test_host_tables_match_oracle(golden):
    import pyrmt_b200.functions as F
    from oracle import rmt_oracle as O
    for nx, ny, dx, dy in ((36, 28, 1.2 / 35, 0.9 / 27), (129, 129, 1 / 128, 1 / 128)):
        assert np.array_equal(F._precompute_poisson_eigenvalues(nx, ny, dx, dy),
                              O._precompute_poisson_eigenvalues(nx, ny, dx, dy))
        e1, n1 = F._precompute_poisson_eigenvalues_periodic(nx, ny, dx, dy)
        e2, n2 = O._precompute_poisson_eigenvalues_periodic(nx, ny, dx, dy)
        assert np.array_equal(e1, e2) and np.array_equal(n1, n2)
    assert np.array_equal(F.fast_solve_3x3(golden("utils")["A"], golden("utils")["b"]), golden("utils")["x"])
    X, Y, dx, dy = F.create_grid(36, 28, 1.2, 0.9)
    X2, Y2, dx2, dy2 = O.create_grid(36, 28, 1.2, 0.9)
    assert np.array_equal(X, X2) and np.array_equal(Y, Y2) and dx == dx2 and dy == dy2


def test_timestep_formula_matches_oracle(golden):
    """functions.timestep_from_speed is compute_timestep's host formula (functions.py:177-192)."""
    import pyrmt_b200.functions as F
    from oracle import rmt_oracle as O
    g = golden("timestep")
    u, v, dx, dy = g["u"], g["v"], float(g["dx"]), float(g["dy"])
    speed = float(np.max(np.sqrt(u ** 2 + v ** 2)))
    for args in ((0.2, 1e-3, 0.1, 1.0, 0.0, 1.0, 0.01, 0.01, 0.0), (0.5, 1.0, 2.0, 1.3, 0.07, 0.9, 0.0, 0.0, 0.4)):
        CFL, cap, mu_s, rho_s, gamma, rho_f, mu_f, eta_s, kappa = args
        ref = O.compute_timestep(u, v, dx, dy, CFL, cap, mu_s, rho_s, gamma, rho_f, mu_f=mu_f, eta_s=eta_s, kappa=kappa)
        got = F.timestep_from_speed(speed, dx, dy, CFL, cap, mu_s, rho_s, gamma, rho_f, mu_f, eta_s, kappa)
        assert got == ref


def test_reinit_and_unbuilt_rows():
    import pyrmt_b200.functions as F
    phi = np.ones((8, 8))
    assert F.reinitialize_level_set(phi, 0.1, 0.1, "none") is phi
    with pytest.raises(ValueError):
        F.reinitialize_level_set(phi, 0.1, 0.1, "bogus")
    with pytest.raises(ImportError):
        F.reinitialize_level_set(phi, 0.1, 0.1, "fmm")
    A = F.build_poisson_matrix(129, 129, 0.1, 0.1)         # lazy placeholder
    assert A.shape == (129 * 129, 129 * 129)


# ----------------------------------------------------------------- BC classifier
def _bcs():
    from pyrmt_b200.bc import free_slip_box_bc, no_slip_lid_bc, periodic_bc, wall_bc
    return {"lid": lambda u, v: no_slip_lid_bc(u, v, 1.0), "lid2": lambda u, v: no_slip_lid_bc(u, v, 2.5),
            "slip": free_slip_box_bc, "periodic": periodic_bc, "wall": wall_bc,
            "identity": lambda u, v: (u.copy(), v.copy()),
            "odd": lambda u, v: (np.concatenate([-u[:, 1:2], u[:, 1:]], axis=1), v.copy())}


@pytest.mark.parametrize("name", ["lid", "lid2", "slip", "periodic", "wall", "identity", "odd"])
def test_bc_classifier_reproduces_callable(name):
    from pyrmt_b200.bc import classify
    bc = _bcs()[name]
    for Ny, Nx in ((9, 13), (28, 36)):
        t = classify(bc, Ny, Nx)
        assert t is not None, name
        rng = np.random.default_rng(Ny)
        u, v = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
        ru, rv = bc(u.copy(), v.copy())
        tu, tv = t.apply_host(u, v)
        assert np.array_equal(tu, ru) and np.array_equal(tv, rv)
        if name == "identity":
            assert t.n == 0
        else:
            assert 0 < t.n <= 4 * (Ny + Nx)


def test_bc_classifier_rejects_non_gather():
    from pyrmt_b200.bc import classify
    assert classify(lambda u, v: (u * 2.0, v), 8, 8) is None               # scaling, not a gather
    assert classify(lambda u, v: (u + v, v), 8, 8) is None                  # mixes two cells

    def shifted(u, v):                                                      # sources are rewritten too
        u = u.copy()
        u[:, 0], u[:, 1] = u[:, 1].copy(), u[:, 0].copy()
        return u, v.copy()
    assert classify(shifted, 8, 8) is None


def test_disc_bins_are_a_superset():
    """The per-bin candidate lists must contain the exact minimiser for any point."""
    from pyrmt_b200.driver import disc_lattice
    from pyrmt_b200.levelset import DiscSDF
    L = 1.0
    cx, cy, R = disc_lattice(8, L, 0.04)
    sdf = DiscSDF(cx, cy, R, domain=(L, L))
    gb, start, cand, Lx, Ly = sdf._bins
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, L, size=(20000, 2))
    pts[:50] = 0.0
    d = np.hypot(pts[:, :1] - cx[None, :], pts[:, 1:] - cy[None, :]) - R[None, :]
    best = d.argmin(axis=1)
    bx = np.minimum((pts[:, 0] * gb / Lx).astype(int), gb - 1)
    by = np.minimum((pts[:, 1] * gb / Ly).astype(int), gb - 1)
    for k in range(pts.shape[0]):
        b = by[k] * gb + bx[k]
        assert best[k] in cand[start[b]:start[b + 1]]
    assert start[-1] < 12 * gb * gb          # the lists stay short


def test_bc_table_cache_keys_closures_by_value(monkeypatch):
    """a lambda re-created every step must not be re-probed every step, a closure whose captured
    value changed must get a new table, and `rmt_dynamic = True` opts out of the table altogether."""
    from pyrmt_b200 import bc as B
    calls = {"n": 0}
    real = B.classify

    def counting(f, Ny, Nx):
        calls["n"] += 1
        return real(f, Ny, Nx)
    monkeypatch.setattr(B, "classify", counting)
    B._cache.clear()

    def make(speed):
        return lambda u, v: B.no_slip_lid_bc(u, v, speed)
    t1 = B.table_for(make(1.0), 17, 19)
    for _ in range(5):                              # "per-step lambdas": same code, same captured value
        assert B.table_for(make(1.0), 17, 19) is t1
    assert calls["n"] == 1
    t2 = B.table_for(make(2.0), 17, 19)             # a ramped lid speed: new value, new table
    assert calls["n"] == 2 and t2 is not t1
    u = np.zeros((17, 19)); v = np.zeros((17, 19))
    assert t2.apply_host(u, v)[0][-1, 5] == 2.0 and t1.apply_host(u, v)[0][-1, 5] == 1.0
    dyn = make(3.0)
    dyn.rmt_dynamic = True
    assert B.table_for(dyn, 17, 19) is None and calls["n"] == 2
    for k in range(B._CACHE_MAX + 10):              # bounded
        B.table_for(make(10.0 + k), 9, 9)
    assert len(B._cache) <= B._CACHE_MAX


REFERENCE = "/root/reference"
IN_SCOPE_DRIVERS = ["common.py", "lid_driven_cavity.py", "soft_disc_in_lid_driven.py", "disc_in_taylor_green.py",
                    "convergence_taylor_green.py", "two_disc_contact.py", "two_disc_tg_collision.py",
                    "surface_tension_drop.py"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "benchmarks")),
                    reason="the reference tree only exists in the build container")
def test_compat_shim_resolves_every_import_of_the_in_scope_drivers():
    """every `from pyRMT... import name` of the reference's collocated-grid drivers (the MAC /
    viscoelastic ones are out of scope) must resolve against `compat.install()` -- that is what
    'the benchmarks/ scripts run unchanged' rests on."""
    import ast
    import importlib
    from pyrmt_b200 import compat
    compat.install()
    try:
        seen = 0
        for name in IN_SCOPE_DRIVERS:
            tree = ast.parse(open(os.path.join(REFERENCE, "benchmarks", name)).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "pyRMT":
                    mod = importlib.import_module(node.module)
                    for alias in node.names:
                        assert hasattr(mod, alias.name), "%s: from %s import %s" % (name, node.module, alias.name)
                        seen += 1
                elif isinstance(node, ast.Import):
                    for alias in node.names:
                        if alias.name.split(".")[0] == "pyRMT":
                            importlib.import_module(alias.name)
                            seen += 1
        assert seen >= 30
        import pyRMT                                          # the package namespace itself (pyRMT/__init__.py)
        init = ast.parse(open(os.path.join(REFERENCE, "pyRMT", "__init__.py")).read())
        for node in ast.walk(init):
            if isinstance(node, ast.ImportFrom) and node.level == 1 and node.module in ("functions", "interpolators", "utils", "output"):
                for alias in node.names:
                    if alias.name != "*":
                        assert hasattr(pyRMT, alias.asname or alias.name), alias.name
    finally:
        compat.uninstall()

"""Pin the CPU oracle against golden vectors recorded from the real reference.

tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports
the upstream pyRMT package and stores its inputs/outputs.  The oracle restates
the same arithmetic operation by operation, so almost everything is required
to be BIT-EXACT; the only tolerance (1e-13) is on paths through NumPy's
vectorised transcendental / FFT routines.
"""
import numpy as np
import pytest

from conftest import rel_linf
from oracle import rmt_oracle as O


def same(x, ref):
    return np.array_equal(np.asarray(x), np.asarray(ref), equal_nan=True)


def test_utils_bit_exact(golden):
    g = golden("utils")
    f, w, dx, dy = g["f"], g["w"], float(g["dx"]), float(g["dy"])
    assert same(O.grad_central_x_2nd(f, dx), g["gx2"])
    assert same(O.grad_central_y_2nd(f, dy), g["gy2"])
    assert same(O.grad_central_x_4th(f, dx), g["gx4"])
    assert same(O.grad_central_y_4th(f, dy), g["gy4"])
    assert same(O.diff_upwind_3rd(f, w, dx, 1), g["up_x"])
    assert same(O.diff_upwind_3rd(f, w, dy, 0), g["up_y"])
    assert same(O.lap_2nd(f, dx, dy), g["lap"])
    assert same(O.fast_solve_3x3(g["A"], g["b"]), g["x"])
    assert same(O.fast_solve_3x3(g["As"], g["bs"]), g["xs"])      # singular -> zeros


def test_interpolators_bit_exact(golden):
    g = golden("interp")
    args = (g["u"], g["xq"], g["yq"], float(g["dx"]), float(g["dy"]), int(g["Nx"]), int(g["Ny"]))
    assert same(O.bilinear_interpolate(*args), g["bil"])
    assert same(O.bicubic_interpolate(*args), g["bic"])
    assert O.cubic_convolution(*g["cc_in"]) == g["cc"][0]


def test_extrapolation_bit_exact(golden):
    g = golden("extrap")
    dx, dy = float(g["dx"]), float(g["dy"])
    for a, b, ph, ea, eb, L in (("X1", "X2", "phi", "X1e", "X2e", "layers"),
                               ("Z1", "Z2", "phi2", "Z1e", "Z2e", "layers2"),
                               ("X1", "X2", "phi3", "W1e", "W2e", "layers3")):
        r1, r2 = O.extrapolate_reference_map(g[a], g[b], g[ph], dx, dy, int(g[L]))
        assert same(r1, g[ea]) and same(r2, g[eb])


@pytest.mark.parametrize("scheme", ["semilagrangian", "semilagrangian_cubic", "central2", "weno5",
                                    "conservative"])
def test_advection_bit_exact(golden, scheme):
    g = golden("advect")
    dx, dy, dt = float(g["dx"]), float(g["dy"]), float(g["dt"])
    r = O.advect_reference_map(g["q"], g["a"], g["b"], g["X"], g["Y"], dt, dx, dy, g["phi"],
                               scheme, 0.0)
    assert same(r, g["out_" + scheme])
    r = O.advect_reference_map(g["q2"], g["a"], g["b"], g["X"], g["Y"], dt, dx, dy, g["phi"],
                               scheme, 1.5 * dx)
    assert same(r, g["out2_" + scheme])


def test_advection_rim_and_clamp(golden):
    g = golden("advect")
    dx, dy, dt = float(g["dx"]), float(g["dy"]), float(g["dt"])
    phin = -np.ones_like(g["phi"])
    for sch in ("weno5", "central2", "conservative"):
        r = O.advect_reference_map(g["full_q"], g["a"], g["b"], g["X"], g["Y"], dt, dx, dy, phin,
                                   sch, 0.0)
        assert same(r, g["full_" + sch]), sch
    assert same(O._weno5_rhs(g["full_q"], g["a"], g["b"], dx, dy, phin, 0.0), g["rhs_weno5"])
    r = O.advect_reference_map(g["full_q"], g["a"], g["b"], g["X"], g["Y"], float(g["far_dt"]),
                               dx, dy, g["phi"], "semilagrangian", 0.0)
    assert same(r, g["far_sl"])


def test_advection_errors():
    z = np.zeros((8, 8))
    bad = z.copy()
    bad[2, 2] = np.inf
    with pytest.raises(FloatingPointError):
        O.advect_reference_map(z, bad, z, z, z, 0.1, 0.1, 0.1, z, 'semilagrangian')
    with pytest.raises(ValueError):
        O.advect_reference_map(z, z, z, z, z, 0.1, 0.1, 0.1, z, 'upwind1')


def test_stress_and_heaviside(golden):
    g = golden("stress")
    dx, dy = float(g["dx"]), float(g["dy"])
    kws = {"legacy": dict(mu_s=0.7, kappa=0.0), "kappa": dict(mu_s=0.7, kappa=2.0),
           "band": dict(mu_s=1.3, kappa=0.5, w_cut=2 * dx, detg_clamp=3.0),
           "iso": dict(mu_s=1.1, kappa=0.4, isochoric=True)}
    for tag, kw in kws.items():
        r = O.solid_cauchy_stress(g["X1"], g["X2"], dx, dy, phi=g["phi"], **kw)
        for nm, arr in zip(("sxx", "sxy", "syy", "J"), r):
            assert same(arr, g[f"{tag}_{nm}"]), (tag, nm)
    assert same(O.smoothed_heaviside(g["hv_x"], float(g["hv_w"])), g["hv"])
    assert same(O.smoothed_heaviside(g["phi"], float(g["hv_w"])), g["hv_phi"])


def _bc(name):
    if name == "lid":
        return lambda u, v: O.no_slip_lid_bc(u, v, 1.0)
    if name == "slip":
        return O.free_slip_box_bc
    return O.wall_bc


MOM_CASES = {"lid_eta": "lid", "lid_noeta": "lid", "slip_eta": "slip", "lid_band": "lid",
             "lid_gamma": "lid", "wall_fluid": "wall"}


@pytest.mark.parametrize("case", sorted(MOM_CASES))
def test_momentum_step(golden, case):
    g = golden("momentum")
    dx, dy, dt, w_t = (float(g[k]) for k in ("dx", "dy", "dt", "w_t"))
    mu_s, kap, eta, rs, rf, muf, gam, band, clamp = g[case + "_prm"]
    phi = np.ones_like(g["phi"]) if case == "wall_fluid" else g["phi"]
    r = O.momentum_step_rk4(g["u"], g["v"], g["p"], g["X1"], g["X2"], _bc(MOM_CASES[case]), mu_s,
                            kap, eta, dx, dy, dt, rs, rf, phi, muf, w_t, gam,
                            stress_band=bool(band), detg_clamp=clamp)
    for k, arr in zip(("un", "vn", "sxx", "sxy", "syy", "J"), r):
        assert same(arr, g[f"{case}_{k}"]), k


def test_velocity_rhs_and_curvature(golden):
    g = golden("momentum")
    dx, dy, w_t = float(g["dx"]), float(g["dy"]), float(g["w_t"])
    phi = g["phi"]
    H = O.smoothed_heaviside(phi, w_t)
    rho = (1 - H) * float(g["rhs_rho_s"]) + H * float(g["rhs_rho_f"])
    exx, exy, eyy, _ = O.solid_cauchy_stress(g["X1"], g["X2"], dx, dy, float(g["rhs_mu_s"]), 0.0, phi)
    ru, rv = O.velocity_rhs_blended_optimized(
        g["u"], g["v"], g["p"], exx, exy, eyy, dx, dy, phi, float(g["rhs_mu_f"]), H,
        O.grad_central_x_2nd(H, dx), O.grad_central_y_2nd(H, dy), rho, 0.0, 0.0)
    assert same(ru, g["rhs_u"]) and same(rv, g["rhs_v"])
    assert same(O.compute_curvature(phi, dx, dy), g["curv"])


CONTACT_CASES = {"m_lid": ("lid", 2.0, 3.0, 4.0), "m_free": ("identity", 0.0, None, 4.0),
                 "m_clamp": ("lid", 1.5, None, 1.2)}        # bc, k_rep, w_c / dx, detg_clamp


def contact_args(g, case):
    from oracle.rmt_oracle import no_slip_lid_bc
    bcn, k_rep, wc, clamp = CONTACT_CASES[case]
    bc = (lambda u, v: no_slip_lid_bc(u, v, 1.0)) if bcn == "lid" else (lambda u, v: (u.copy(), v.copy()))
    dx, dy = float(g["dx"]), float(g["dy"])
    mu_s, kappa, eta_s, dt, rho_s, rho_f, mu_f, w_t = (float(x) for x in g["prm"])
    pos = (g["u"], g["v"], g["p"], g["X1a"], g["X2a"], g["X1b"], g["X2b"], bc, mu_s, kappa, eta_s, dx, dy, dt,
           rho_s, rho_f, g["phi_a"], g["phi_b"], mu_f, w_t)
    return pos, dict(k_rep=k_rep, w_c=None if wc is None else wc * dx, detg_clamp=clamp)


def test_contact_force_and_two_solid_step(golden):
    """functions.py:765-895 against the vectors of tests/golden/make_golden_contact.py."""
    g = golden("contact")
    dx, dy = float(g["dx"]), float(g["dy"])
    for tag in ("c1", "c2"):
        fx, fy = O.compute_contact_force(g["phi_a"], g["phi_b"], float(g[tag + "_k"]), float(g[tag + "_w"]), dx, dy)
        assert rel_linf(fx, g[tag + "_fx"]) < 1e-13 and rel_linf(fy, g[tag + "_fy"]) < 1e-13
        assert np.abs(g[tag + "_fx"]).max() > 0
    for case in CONTACT_CASES:
        pos, kw = contact_args(g, case)
        un, vn, Jm = O.momentum_step_rk4_2solids(*pos, **kw)
        assert rel_linf(un, g[case + "_u"]) < 1e-13, case
        assert rel_linf(vn, g[case + "_v"]) < 1e-13, case
        assert same(Jm, g[case + "_J"]), case


def test_reinitialize_phi_pde(golden):
    """functions.py:1369-1411 against tests/golden/make_golden_reinit.py (bit-exact: IEEE ops only)."""
    g = golden("reinit")
    dx, dy = float(g["dx"]), float(g["dy"])
    assert same(O.reinitialize_phi_PDE(g["phi"], dx, dy, 7, None, 0.5), g["r_none_7"])
    assert same(O.reinitialize_phi_PDE(g["phi"], dx, dy, 20, O.apply_phi_BCs, 0.2), g["r_bc_20"])
    assert same(O.reinitialize_phi_PDE(g["phi"], dx, dy, 5, None, 0.3), g["r_level_set"])


def test_energy_diagnostics(golden):
    """output.py:6-193 + common.py:110-115 against tests/golden/make_golden_diag.py."""
    g = golden("diag")
    dx, dy, w_t = float(g["dx"]), float(g["dy"]), float(g["w_t"])
    rho_f, rho_s, mu_f, mu_s, kappa, eta_s = (float(x) for x in g["prm"])
    rel = lambda x, r: abs(x - float(r)) / abs(float(r))
    assert rel(O.compute_kinetic_energy(g["a"], g["b"], rho_f, rho_s, g["phi"], w_t, dx, dy), g["ke"]) < 1e-13
    assert rel(O.compute_strain_energy(g["X1"], g["X2"], g["phi"], mu_s, dx, dy, kappa=kappa), g["se"]) < 1e-12
    assert rel(O.compute_strain_energy(g["X1"], g["X2"], g["phi"], mu_s, dx, dy), g["se0"]) < 1e-12
    assert rel(O.compute_viscous_dissipation(g["a"], g["b"], mu_f, g["phi"], w_t, dx, dy, eta_s=eta_s), g["diss"]) < 1e-13
    assert rel(O.compute_viscous_dissipation(g["a"], g["b"], mu_f, g["phi"], w_t, dx, dy), g["diss0"]) < 1e-13
    assert np.allclose(O.disc_centroid(g["phi"], g["X"], g["Y"]), g["centroid"], rtol=1e-14, atol=0)


def test_projection_neumann(golden):
    g = golden("projection")
    dx, dy, dt = float(g["dx"]), float(g["dy"]), float(g["dt"])
    assert same(O._precompute_poisson_eigenvalues(g["a"].shape[1], g["a"].shape[0], dx, dy), g["eig"])
    assert same(O._compute_divergence(g["a"], g["b"], dx, dy), g["div"])
    assert same(O._compute_divergence_rc(g["a"], g["b"], g["p"], dt, 1.0, dx, dy), g["div_rc"])
    gx, gy = O._compute_pressure_gradient(g["p"], dx, dy)
    assert same(gx, g["gx"]) and same(gy, g["gy"])
    assert same(O._solve_poisson_dct(g["rhs"], g["eig"]), g["sol"])
    lid = _bc("lid")
    r = O.pressure_projection_amg(g["a"], g["b"], dx, dy, dt, g["rho_arr"], lid, p_prev=g["p"],
                                  eigenvalues=g["eig"])
    assert same(r[0], g["A_a"]) and same(r[1], g["A_b"]) and same(r[2], g["A_p"])
    r = O.pressure_projection_amg(g["a"], g["b"], dx, dy, dt, 1.0, O.free_slip_box_bc,
                                  p_prev=None, eigenvalues=g["eig"])
    assert same(r[0], g["B_a"]) and same(r[1], g["B_b"]) and same(r[2], g["B_p"])
    r = O.pressure_projection_amg(g["a"], g["b"], dx, dy, dt, 0.8, lid, p_prev=g["p"],
                                  eigenvalues=g["eig"])
    assert same(r[0], g["C_a"]) and same(r[1], g["C_b"]) and same(r[2], g["C_p"])
    g2 = golden("projection_pow2")
    assert same(O._solve_poisson_dct(g2["rhs"], g2["eig"]), g2["sol"])
    r = O.pressure_projection_amg(g2["a"], g2["b"], float(g2["dx"]), float(g2["dy"]),
                                  float(g2["dt"]), 1.0, lid, p_prev=g2["p"], eigenvalues=g2["eig"])
    assert same(r[0], g2["A_a"]) and same(r[1], g2["A_b"]) and same(r[2], g2["A_p"])


@pytest.mark.parametrize("tag", ["odd", "pow2"])
def test_projection_periodic(golden, tag):
    g = golden("periodic")
    dx, dy, dt = float(g[tag + "_dx"]), float(g[tag + "_dy"]), float(g["dt"])
    a, b, p = g[tag + "_a"], g[tag + "_b"], g[tag + "_p"]
    Ny, Nx = a.shape
    eig, null = O._precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy)
    assert same(eig, g[tag + "_eig"]) and same(null, g[tag + "_null"])
    assert same(O._solve_poisson_fft(g[tag + "_rhs"], (eig, null)), g[tag + "_sol"])
    assert same(O._compute_divergence_periodic(a, b, dx, dy), g[tag + "_div"])
    gx, gy = O._compute_pressure_gradient_periodic(p, dx, dy)
    assert same(gx, g[tag + "_gx"]) and same(gy, g[tag + "_gy"])
    r = O.pressure_projection_amg(a, b, dx, dy, dt, np.ones_like(a), O.periodic_bc, p_prev=p,
                                  eigenvalues=(eig, null), bc_type='periodic')
    assert same(r[0], g[tag + "_A_a"]) and same(r[1], g[tag + "_A_b"]) and same(r[2], g[tag + "_A_p"])
    r = O.pressure_projection_amg(a, b, dx, dy, dt, 1.0, O.periodic_bc, p_prev=None,
                                  eigenvalues=None, bc_type='periodic')
    assert same(r[0], g[tag + "_B_a"]) and same(r[1], g[tag + "_B_b"]) and same(r[2], g[tag + "_B_p"])


def test_timestep_and_phi_bcs(golden):
    g = golden("timestep")
    dx, dy = float(g["dx"]), float(g["dy"])
    for row in g["table"]:
        CFL, cap, mu_s, rho_s, gam, rho_f, mu_f, eta, kap, want = row
        got = O.compute_timestep(g["u"], g["v"], dx, dy, CFL, cap, mu_s, rho_s, gam, rho_f,
                                 mu_f=mu_f, eta_s=eta, kappa=kap)
        assert got == want
    assert same(O.apply_phi_BCs(g["phibc_in"].copy()), g["phibc_out"])
    X, Y, ddx, ddy = O.create_grid(36, 28, 1.2, 0.9)
    u = golden("utils")
    assert ddx == float(u["dx"]) and ddy == float(u["dy"])


@pytest.mark.parametrize("scheme", ["semilagrangian", "weno5", "central2"])
@pytest.mark.parametrize("step", [0, 30, 59])
def test_full_fsi_step(golden, scheme, step):
    """One complete step of the driver loop (soft_disc_in_lid_driven.py:78-106)."""
    g = golden("fsi_steps")
    x0, y0, R = g["disc"]
    mu_s, kappa, rho_s, eta_s, mu_f, rho_f, w_t, layers, CFL, cap = g["prm"]
    prm = dict(dx=float(g["dx"]), dy=float(g["dy"]), X=g["X"], Y=g["Y"], eig=g["eig"],
               mu_s=mu_s, kappa=kappa, rho_s=rho_s, eta_s=eta_s, mu_f=mu_f, rho_f=rho_f, w_t=w_t,
               layers=int(layers), CFL=CFL, dt_cap=cap, scheme=scheme,
               bc=lambda u, v: O.no_slip_lid_bc(u, v, 1.0),
               phi_init=lambda A, B: O.initialize_disc(A, B, x0, y0, R))
    key = f"{scheme}_{step}_"
    state = tuple(g[key + "in_" + k] for k in ("a", "b", "p", "X1", "X2"))
    new, dt, ex = O.fsi_step(state, prm)
    assert dt == float(g[key + "dt"])
    for k, arr in zip(("a", "b", "p", "X1", "X2"), new):
        assert rel_linf(arr, g[key + "out_" + k]) <= 1e-13, k
    assert same(ex["phi"], g[key + "out_phi"])
    assert same(ex["J"], g[key + "out_J"])

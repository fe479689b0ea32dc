"""CUDA-event times of the single operators of one FSI step on the config-4 state (4097^2 by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrmt_b200 import functions as F
from pyrmt_b200.driver import make_case, fsi_step

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
state, prm = make_case(N, scheme="weno5")
for _ in range(6):
    state, dt, ex = fsi_step(state, prm)
a, b, p, X1, X2 = state
dx, dy = prm["dx"], prm["dy"]
phi = ex["phi"]


def t(name, fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-34s %.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)


t("rebuild_phi (disc sdf)", lambda: F.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"]))
t("solid_cauchy_stress", lambda: F.solid_cauchy_stress(X1, X2, dx, dy, prm["mu_s"], prm["kappa"], phi))
t("rebuild_phi_and_stress (fused)", lambda: F.rebuild_phi_and_stress(X1, X2, prm["phi_init"], dx, dy, prm["mu_s"], prm["kappa"], prm["w_t"]))
t("advect pair weno5", lambda: F.advect_reference_map_pair(X1, X2, a, b, None, None, dt, dx, dy, phi, "weno5", 0.0, mask_solid=True))
Xa, Xb = F.advect_reference_map_pair(X1, X2, a, b, None, None, dt, dx, dy, phi, "weno5", 0.0, mask_solid=True)
t("extrapolate", lambda: F.extrapolate_reference_map(Xa, Xb, phi, dx, dy, 3))
ph2, st = F.rebuild_phi_and_stress(X1, X2, prm["phi_init"], dx, dy, prm["mu_s"], prm["kappa"], prm["w_t"])
t("momentum_step_rk4 (given stress)", lambda: F.momentum_step_rk4_with_stress(a, b, p, X1, X2, prm["bc"], prm["mu_s"], prm["kappa"], prm["eta_s"], dx, dy, dt, 1.0, 1.0, ph2, prm["mu_f"], prm["w_t"], stress=st))
t("pressure_projection_amg", lambda: F.pressure_projection_amg(a, b, dx, dy, dt, 1.0, prm["bc"], p_prev=p, eigenvalues=prm["eig"]))
t("compute_timestep", lambda: F.compute_timestep(a, b, dx, dy, prm["CFL"], prm["dt_cap"], prm["mu_s"], 1.0, 0.0, 1.0, mu_f=prm["mu_f"], eta_s=prm["eta_s"], kappa=prm["kappa"]))
t("fsi_step", lambda: fsi_step(state, prm))

"""Run BASELINE.json configs[0..2] on the GPU and compare with the numbers the reference
produced in the build container.  Writes gpurun_out/validation.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pyrmt_b200 import cases

out = {}
log = lambda s: print(s, flush=True)
which = sys.argv[1:] or ["1", "2", "3"]
if "1" in which:
    t0 = time.time()
    r = cases.lid_driven_cavity(100.0, 129, log=None)
    rms = cases.ghia_rms(r["y"], r["u_line"], os.path.join(ROOT, "tests", "golden", "ghia_re100_u_y.csv"))
    out["config1_lid_driven_Re100_N129"] = dict(steady_step=r["step"], rms_vs_ghia=rms,
                                                reference=dict(steady_step=10600, rms_vs_ghia=1.6890e-3),
                                                wall_s=time.time() - t0)
    log("config 1: steady at step %d, RMS vs Ghia %.4e (reference 1.6890e-03 at step 10600), %.1f s"
        % (r["step"], rms, time.time() - t0))
if "2" in which:
    t0 = time.time()
    r = cases.soft_disc_in_lid(128, "semilagrangian", 8.0, log=log)
    ref = [(0.5317, 0.4925), (0.4136, 0.5295), (0.3140, 0.6448), (0.3018, 0.7983), (0.4594, 0.8706),
           (0.6880, 0.8082), (0.6718, 0.6404), (0.5608, 0.5708)]
    out["config2_soft_disc_lid_N128_SL"] = dict(steps=r["steps"], samples=r["samples"], x_range=r["x_range"],
                                               y_range=r["y_range"], minJ=r["minJ"], maxJ=r["maxJ"],
                                               reference=dict(steps=25807, centroid_t1_8=ref,
                                                              x_range=(0.2896, 0.7053), y_range=(0.4922, 0.8709),
                                                              minJ=0.449, maxJ=5.47),
                                               wall_s=time.time() - t0)
    log("config 2: %d steps, x in [%.4f, %.4f], y in [%.4f, %.4f], %.1f s"
        % (r["steps"], *r["x_range"], *r["y_range"], time.time() - t0))
if "3" in which:
    t0 = time.time()
    r = cases.disc_in_taylor_green(128, "semilagrangian", 1.0, log=None)
    out["config3_disc_TG_N128_SL_freeslip"] = dict(**r, reference=dict(steps=10000, E0=2.50607e-2, E1=2.44811e-2,
                                                                     drift_pct=-2.31, KE=1.2212e-2, SE=4.2942e-3,
                                                                     integrated_dissipation=7.9754e-3),
                                                   wall_s=time.time() - t0)
    log("config 3: %d steps, E0=%.5e E1=%.5e drift %.2f%% (reference -2.31%%), %.1f s"
        % (r["steps"], r["E0"], r["E1"], r["drift_pct"], time.time() - t0))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "validation.json"), "w") as f:
    json.dump(out, f, indent=1, default=float)
print(json.dumps(out, default=float))

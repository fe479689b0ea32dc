#!/usr/bin/env python
"""bench.py -- Mcell-steps/s of the full collocated FSI step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3]): synthetic multi-disc
lid-driven FSI on a 4097 x 4097 node grid (4096^2 cells), 64 neo-Hookean discs,
WENO5 + SSP-RK3 reference-map advection, 3 extrapolation layers, RK4 momentum
predictor, Rhie-Chow + DCT-I projection; fp64 throughout.  One "step" is one pass
of the loop of benchmarks/soft_disc_in_lid_driven.py:78-106 through the drop-in
operator API (pyrmt_b200.driver.fsi_step).

  value         device-resident loop, state already in HBM, CUDA-event timed
  e2e           the same step called with HOST state (pyrmt_b200.driver.fsi_step_host):
                the five state fields are copied from pinned host memory to the
                device and the five results back, every step, inside the timed
                region (uploads ordered by first use, xi downloads overlapped with
                the predictor / projection on copy streams)
  roofline      the dominant entry point (by summed CUDA-event time in the timed
                region): algorithmic bytes per launch / mean launch duration;
                kernels_roofline lists every entry point with its ncu DRAM traffic
  gpu_launches  kernels launched in the timed region, counted inside
                librmt_b200.so (rmt_launch_count), per rank
  cpu_baseline  the CPU oracle (oracle/, a port of the reference's Numba/NumPy
                path) on a bounded sample of the same workload, on this box's cores

--impl reference times that CPU oracle alone (the reference itself is pure
Python + Numba and cannot travel to the GPU box).
N > 1: the SAME 4097^2 problem is slab-decomposed over the N GPUs (pyrmt_b200/slab.py: y-slabs,
halo exchange and the transpose + all-to-all of the DCT solve as kernels storing into the peers' memory
over NVLink (pyrmt_b200/slab.py PeerComm; RMT_SLAB_COMM=nccl: NCCL calls), overlap-swept extrapolation) --
strong scaling; the results equal the single-GPU step (xi bit for bit).  Beside it: the fluid
half of the step on one 8193^2 grid (`slab_fluid_step`, Neumann/DCT) and BASELINE configs[4]
(`config5_periodic`: periodic Taylor-Green FSI at 8193^2 and the periodic fluid step at 16385^2,
distributed Hartley/FFT solve).  With N > 1 the step is short enough to be host-bound, so the
timed region runs without per-call CUDA events and the kernel breakdown comes from a few extra
profiled steps.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_CELL_STEP = {"semilagrangian": 496.0, "weno5": 512.0, "central2": 512.0}

# Algorithmic (compulsory) bytes per cell and launch of each entry point: every input
# field read once + every output written once, 8 B each ("Kernels").
ALG_BYTES_PER_CELL_LAUNCH = {
    "rmt_momentum_stage": 8.0 * 12.0,     # stages 1..4: 11, 13, 13, 11 fields (sigma_s is read in the solid third only)
    "rmt_advect_euler_rk3": 8.0 * 17.0,   # 3 stage kernels per call: 5 + 6 + 6 fields
    "rmt_poisson_solve_dct": 8.0 * 7.0,   # rows 2, columns (+eig) 3, rows 2
    "rmt_solid_stress": 8.0 * 7.0,
    "rmt_projection_rhs": 8.0 * 4.0,      # a*, b*, p in, rhs out (scalar density)
    "rmt_projection_correct": 8.0 * 8.0,
    "rmt_projection_correct_centered": 8.0 * 7.0,   # sol, a*, b*, p in; a, b, p out
    "rmt_disc_sdf_stress": 8.0 * 7.0,     # xi1, xi2 in; phi, sxx, sxy, syy, J out
    "rmt_extrapolate": 8.0 * 5.0,        # a dependency-chain (latency) kernel, not a bandwidth one
    "rmt_extrapolate_rows": 8.0 * 1.0,   # in place (the step drivers): phi in, state bytes + band cells out
    "rmt_advect_euler_rk3_pair": 8.0 * 7.0,    # operator boundary: xi1, xi2, a, b, phi in, xi1', xi2' out (the three
                                               # SSP-RK3 stages move more: see `traffic`)
    "rmt_dct_lines": 8.0 * 2.0,
    "rmt_disc_sdf": 8.0 * 3.0,
    "rmt_mask_mul": 8.0 * 3.0,
    "rmt_heaviside_rho": 8.0 * 3.0,
    "rmt_advect_sl_rk4": 8.0 * 6.0,
    "rmt_max_speed": 8.0 * 2.0,
    "rmt_field_stats": 8.0 * 1.0,
    "rmt_subtract_mean": 8.0 * 2.0,
    "rmt_apply_bc": 0.0,
}


# DRAM bytes per call (dram__bytes_read.sum + dram__bytes_write.sum) from the `ncu --set full` capture of
# this workload at 4097^2 on one B200 (profiles/r0*_ncu_full_summary.txt, one per round); per-call = sum over the
# kernels the entry point launches.  Only reported for N = 1 at the default size.
NCU_TRAFFIC_BYTES_4097 = {                 # profiles/r02h_ncu_full_summary.txt (one steady step, end of round 2)
    "rmt_extrapolate_rows": 0.185e9,       # in place: layer-0 count + state bytes 0.148 GB (phi once, 1 B/cell out),
                                           # k_ext_body 0.036 GB (it works out of L1/L2 and shared memory)
    "rmt_momentum_stage": 1.594e9,         # mean of the four stages (1.45 / 1.72 / 1.72 / 1.47 GB)
    "rmt_advect_euler_rk3_pair": 1.84e9,   # classify 0.14 + three stage kernels 0.38 + 0.50 + 0.81 GB
    "rmt_poisson_solve_dct": 1.238e9,      # 2 x lines<0> (0.22 + 0.21) + lines<1> (0.36) + 2 x transpose (0.22)
    "rmt_projection_correct_centered": 0.903e9,
    "rmt_projection_rhs": 0.520e9,
    "rmt_disc_sdf_stress": 0.887e9,
    "rmt_disc_sdf": 0.377e9,
    "rmt_max_speed": 0.275e9,
}


def workload_config(N, scheme, world):
    cells = N * N
    return {"workload": "synthetic 64-disc lid-driven FSI at %dx%d nodes (%d^2 cells), %s + SSP-RK3 ref-map "
                        "advection, 3-layer extrapolation, RK4 momentum, Rhie-Chow + DCT-I projection"
                        % (N, N, N - 1, scheme),
            "grid": [N, N], "discs": 64, "scheme": scheme,
            "cache": "working set ~%.1f GB per step >> 126 MB L2 (no flush needed)" % (25 * cells * 8 / 1e9),
            "parallelism": ("one grid in %d y-slabs: halo rows and the fused transpose + all-to-all of the DCT solve stored "
                            "straight into the peers' memory over NVLink (CUDA IPC arenas, on-stream barriers), "
                            "overlap-swept extrapolation (strong scaling)" % world) if world > 1 else "single GPU"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def _host_threads():
    """Use every host core for the CPU arm, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1).
    Must run before numpy / numba / the oracle's OpenMP library are loaded."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["NUMBA_NUM_THREADS"] = str(n)
    return n


def _omp_threads():
    import ctypes
    try:
        return int(ctypes.CDLL("libgomp.so.1").omp_get_max_threads())
    except Exception:
        return None


def cpu_fsi_step(M, state, prm):
    """One step of benchmarks/soft_disc_in_lid_driven.py:78-106 with the functions of module M -- the
    staged reference (pyRMT.functions) or the oracle port (same names, same signatures)."""
    a, b, p, X1, X2 = state
    dx, dy = prm["dx"], prm["dy"]
    dt = M.compute_timestep(a, b, dx, dy, prm["CFL"], prm["dt_cap"], prm["mu_s"], prm["rho_s"], 0.0, prm["rho_f"],
                            mu_f=prm["mu_f"], eta_s=prm["eta_s"], kappa=prm["kappa"])
    phi = M.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"])
    solid_mask = (phi <= 0).astype(float)
    X1 = M.advect_reference_map(X1, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, prm["scheme"], 0.0) * solid_mask
    X2 = M.advect_reference_map(X2, a, b, prm["X"], prm["Y"], dt, dx, dy, phi, prm["scheme"], 0.0) * solid_mask
    X1, X2 = M.extrapolate_reference_map(X1, X2, phi, dx, dy, prm["layers"])
    phi = M.rebuild_phi_from_reference_map(X1, X2, prm["phi_init"])
    a_s, b_s, *_ = M.momentum_step_rk4(a, b, p, X1, X2, prm["bc"], prm["mu_s"], prm["kappa"], prm["eta_s"], dx, dy,
                                       dt, prm["rho_s"], prm["rho_f"], phi, prm["mu_f"], prm["w_t"], 0.0)
    H = M.smoothed_heaviside(phi, prm["w_t"])
    rho_local = (1 - H) * prm["rho_s"] + H * prm["rho_f"]
    a, b, p, _, _ = M.pressure_projection_amg(a_s, b_s, dx, dy, dt, rho_local, velocity_bc=prm["bc"], A=None,
                                              ml=None, p_prev=p, eigenvalues=prm["eig"], bc_type="neumann")
    return (a, b, p, X1, X2)


def cpu_case(N, scheme):
    """The bench workload (64 discs, lid-driven) on an N x N node grid as host arrays: (params, state)."""
    import numpy as np
    from oracle import rmt_oracle as O
    from pyrmt_b200.driver import disc_lattice
    X, Y, dx, dy = O.create_grid(N, N, 1.0, 1.0)
    cx, cy, R = disc_lattice(8, 1.0, 0.04)
    phi0 = lambda A, B: O.disc_sdf(A, B, cx, cy, R)          # the user-side phi0 callable (C helper: fast)
    lid = lambda u, v: O.no_slip_lid_bc(u, v, 1.0)           # benchmarks/common.py:27-37
    prm = dict(dx=dx, dy=dy, CFL=0.2, dt_cap=1e-3, mu_s=0.1, kappa=0.0, rho_s=1.0, rho_f=1.0, eta_s=0.01,
               mu_f=0.01, w_t=2 * dx, layers=3, scheme=scheme, w_cut=0.0, bc=lid, X=X, Y=Y, phi_init=phi0,
               eig=O._precompute_poisson_eigenvalues(N, N, dx, dy), discs=(cx, cy, R))
    ph = phi0(X, Y)
    m = (ph <= 0).astype(float)
    X1, X2 = O.extrapolate_reference_map(X * m, Y * m, ph, dx, dy, 3)
    z = np.zeros_like(X)
    a, b = lid(z, z)
    return prm, (a, b, z.copy(), X1, X2)


def cpu_rate(n_nodes, steps, warmup, scheme, want_reference=True):
    """Mcell-steps/s of the CPU arm on a bounded sample (same physics, fewer cells): the UNMODIFIED reference
    (oracle/_ref, staged by oracle/make_ref.py; Numba on all host cores) when it is there and numba imports,
    else the oracle port (C/OpenMP + NumPy/pocketfft).  Returns (rate, ms/step, kind, threads, final state)."""
    M, kind, threads = None, "port", None
    if want_reference:
        from oracle import make_ref
        M = make_ref.load()
        if M is not None:
            import numba
            kind, threads = "reference", int(numba.get_num_threads())
    from oracle import rmt_oracle as O
    if M is None:
        M, threads = O, _omp_threads()
    prm, state = cpu_case(n_nodes, scheme)
    for _ in range(warmup):
        state = cpu_fsi_step(M, state, prm)
    t0 = time.perf_counter()
    for _ in range(steps):
        state = cpu_fsi_step(M, state, prm)
    el = time.perf_counter() - t0
    return n_nodes * n_nodes * steps / el / 1e6, el / steps * 1e3, kind, threads, (prm, state)


def slab_fluid_rate(N, steps, rank, world):
    """Momentum predictor + DCT projection on y-slabs over all ranks (peer-memory halo exchange +
    all-to-all transposes, pyrmt_b200/slab.py): ms per step, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pyrmt_b200 import functions as F
    from pyrmt_b200.driver import LidBC
    from pyrmt_b200.slab import SlabFluidSolver, SlabLayout
    lay = SlabLayout(N, N, world, rank, halo=4)
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    solver = SlabFluidSolver(lay, LidBC(1.0), F._precompute_poisson_eigenvalues(N, N, dx, dy))
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    z = torch.zeros((lay.nl, N), dtype=torch.float64, device="cuda")
    a, b, p = z.clone(), z.clone(), z.clone()
    if rank == world - 1:
        a[lay.o1 - 1, 1:-1] = 1.0
    X1, X2, phi = up(lay.take(X)), up(lay.take(Y)), torch.ones_like(z)
    del X, Y
    prm = dict(dx=dx, dy=dy, mu_s=0.0, kappa=0.0, eta_s=0.0, rho_s=1.0, rho_f=1.0, mu_f=0.01, w_t=2 * dx)
    dt = 0.2 * dx * dx / (4 * 0.01)
    for _ in range(3):
        a, b, p = solver.fluid_step(a, b, p, X1, X2, phi, prm, dt)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        a, b, p = solver.fluid_step(a, b, p, X1, X2, phi, prm, dt)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"what": "momentum_step_rk4 + Neumann/DCT projection, y-slabs, peer-memory halo exchange + fused transpose / all-to-all "
                    "transposes (strong scaling: one %dx%d grid over %d GPUs)" % (N, N, world),
            "grid": [N, N], "ms_per_step": ms, "value": N * N / ms / 1e3, "unit": "Mcell-steps/s",
            "finite": bool(torch.isfinite(a).all().item())}


CPU_KIND_TEXT = {
    "reference": "the UNMODIFIED reference (pyRMT.functions staged under oracle/_ref): Numba kernels on %d threads "
                 "(only the 4 parallel=True kernels are multithreaded upstream), NumPy / pocketfft single-threaded",
    "port": "CPU oracle = C/OpenMP port of the reference's Numba kernels + NumPy/pocketfft, OpenMP kernels on %d "
            "threads, NumPy parts single-threaded as upstream",
}


def slab_parity_vs_1gpu(N, scheme, rank, world, nsteps=3):
    """N > 1: the slab-decomposed FSI step against the single-GPU step on the same N x N problem (config-4
    geometry: bodies straddle the cuts).  Every rank runs the single-GPU step redundantly and compares its
    own rows; the verdict is all-reduced.  Untimed."""
    import torch
    import torch.distributed as dist
    from pyrmt_b200.driver import fsi_step, make_case
    from pyrmt_b200.slab import SlabFSISolver, SlabLayout
    state, prm = make_case(N, L=1.0, k_side=8, R_frac=0.04, scheme=scheme, bc_kind="lid")
    lay = SlabLayout(N, N, world, rank, halo=12)
    # overlap: one slab height at most (the solver refuses slabs thinner than the overlap); the 41-cell discs of
    # this lattice (128-cell pitch) fit in it down to 128-row slabs (8 ranks)
    solver = SlabFSISolver(lay, prm["bc"], prm["eig"], prm["phi_init"], overlap=min(512, N // world),
                           layers=prm["layers"])
    sstate = tuple(lay.take(t).contiguous() for t in state)
    sprm = dict(prm, X=None, Y=None)
    worst = torch.zeros(5, dtype=torch.float64, device="cuda")
    xi_equal = True
    for _ in range(nsteps):
        state, dt, _ = fsi_step(state, prm)
        sstate = solver.fsi_step(sstate, sprm, dt, check_guard=True)
        for k, (ref, got) in enumerate(zip(state, sstate)):
            r, g = ref[lay.r0:lay.r1], lay.owned(got)
            worst[k] = torch.maximum(worst[k], (r - g).abs().max() / ref.abs().max().clamp_min(1e-300))
            if k >= 3:
                xi_equal = xi_equal and bool(torch.equal(r, g))
    solver.comm.allreduce(worst, "max")
    flag = torch.tensor([1.0 if xi_equal else 0.0], dtype=torch.float64, device="cuda")
    solver.comm.allreduce(flag, "min")
    rel = dict(zip(("u", "v", "p", "xi1", "xi2"), (float(x) for x in worst.tolist())))
    return {"against": "the single-GPU step of the same %dx%d problem, %d steps, own rows of every rank" % (N, N, nsteps),
            "grid": [N, N], "rel_linf": rel, "max_rel_linf": max(rel.values()), "xi_bit_exact": bool(flag.item() > 0.5),
            "tolerance": 1e-12, "ok": bool(max(rel.values()) <= 1e-12)}


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the step on this box's host cores.  The
    timed workload is a BOUNDED SAMPLE of the b200 arm's config -- the same 64-disc lid-driven case on a
    --ref-size grid -- and the line says so: `config.grid` / `config.workload` are what actually ran."""
    if rank != 0:
        return
    cores = _host_threads()
    n = args.ref_size
    warm = max(1, min(args.warmup, 2))             # the first step carries the Numba JIT
    rate, ms, kind, threads, _ = cpu_rate(n, args.steps, warm, args.scheme)
    threads = threads or cores
    cfg = workload_config(n, args.scheme, 1)
    cfg["workload"] = ("bounded sample of the b200 arm's workload (4097x4097): " + cfg["workload"]
                       + "; per-cell cost of the CPU path is size-independent (SURVEY 6.2), so Mcell-steps/s compares")
    cfg["sample_of"] = [args.size, args.size]
    cfg.pop("cache", None)
    cfg["parallelism"] = "host CPU, %d threads" % threads
    sample = ("%d steps of the same 64-disc lid-driven %s step on a %dx%d node grid (1/%d of the cells of %dx%d); "
              % (args.steps, args.scheme, n, n, round((args.size / n) ** 2), args.size, args.size)
              + CPU_KIND_TEXT[kind] % threads)
    line = {"impl": "reference", "metric": "Mcell-steps/s full FSI step", "value": rate,
            "unit": "Mcell-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": "Mcell-steps/s", "cores": threads, "host_cores": cores,
                             "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "Mcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4097, help="nodes per side (4097 = 4096^2 cells)")
    ap.add_argument("--scheme", default="weno5")
    ap.add_argument("--ref-size", type=int, default=1025)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--overlap", type=int, default=512, help="rows above a slab re-swept by the extrapolation")
    ap.add_argument("--slab-size", type=int, default=8193, help="nodes per side of the sharded fluid-step timing")
    ap.add_argument("--cfg5-fsi-size", type=int, default=8193, help="nodes per side of the periodic FSI timing (N > 1)")
    ap.add_argument("--cfg5-fluid-size", type=int, default=16385, help="nodes per side of the periodic fluid timing (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from pyrmt_b200._runtime import profiler
    from pyrmt_b200.driver import fsi_step, fsi_step_host, make_case

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    N = args.size
    state, prm = make_case(N, L=1.0, k_side=8, R_frac=0.04, scheme=args.scheme, bc_kind="lid")
    if world == 1:
        def step(st):
            return fsi_step(st, prm)[0]
    else:
        # ONE grid over all ranks: y-slabs, peer-memory halo exchange, fused transpose + all-to-all in the DCT solve,
        # overlap-swept extrapolation (pyrmt_b200/slab.py); strong scaling of the N = 1 workload
        from pyrmt_b200 import functions as Fn
        from pyrmt_b200.slab import SlabFSISolver, SlabLayout
        lay = SlabLayout(N, N, world, rank, halo=12)
        solver = SlabFSISolver(lay, prm["bc"], prm["eig"], prm["phi_init"], overlap=args.overlap, layers=prm["layers"])
        state = tuple(lay.take(t).contiguous() for t in state)
        sprm = dict(prm, X=None, Y=None)

        def step(st):
            return solver.fsi_step(st, sprm, None, check_guard=False)      # dt = compute_timestep of the whole grid
        solver.fsi_step(state, sprm, 1e-7, check_guard=True)      # validates the overlap once (raises if too small)
    for _ in range(args.warmup):
        state = step(state)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # The timed region runs clean: per-entry-point CUDA events (two per call, ~100 per step) cost 0.15 ms per
    # step at N = 1 (measured: 6.32 ms with them, 6.17 without) and more when the step is host-bound (N > 1).
    # The kernel breakdown and the roofline come from extra profiled steps right after the timed region,
    # same state, same stream.
    profiler.reset(timing=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from pyrmt_b200 import _lib as _rmt_lib
    barrier()
    launches0 = _rmt_lib.load().rmt_launch_count()
    e0.record()
    for _ in range(args.steps):
        state = step(state)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(_rmt_lib.load().rmt_launch_count() - launches0)   # counted inside librmt_b200.so
    prof_steps = min(5 if world > 1 else 10, args.steps)
    profiler.reset(timing=True)
    for _ in range(prof_steps):
        state = step(state)
    torch.cuda.synchronize()
    per_kernel = profiler.summary()
    profiler.reset(timing=False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    finite = bool(torch.isfinite(state[0]).all().item())

    cells = N * N
    value = cells * args.steps / (ms * 1e-3) / 1e6          # one N x N grid, whatever the number of GPUs
    peak, peak_kind = peaks()

    # ---- dominant kernel -> roofline --------------------------------------
    top = max(per_kernel.items(), key=lambda kv: kv[1][1])
    kname, (kcalls, ktotal) = top
    alg_launch = ALG_BYTES_PER_CELL_LAUNCH.get(kname, 0.0) * (state[0].numel() if world > 1 else cells)
    kavg_ms = ktotal / kcalls
    achieved = alg_launch / (kavg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_kind": peak_kind,
                "launches_timed": kcalls, "avg_launch_ms": kavg_ms,
                "share_of_step": (ktotal / prof_steps) / (ms / args.steps),
                "alg_bytes_per_launch": alg_launch}
    ncell_rank = state[0].numel() if world > 1 else cells
    if world == 1 and N == 4097:
        roofline["traffic"] = NCU_TRAFFIC_BYTES_4097.get(kname)
        if roofline["traffic"]:
            roofline["frac_by_traffic"] = roofline["traffic"] / (kavg_ms * 1e-3) / 1e9 / peak
    roofline["measured_over"] = "%d profiled steps right after the timed region (the timed region itself runs without per-call events)" % prof_steps
    roofline["note"] = ("dominant entry point by summed CUDA-event time; rmt_extrapolate* is bound by the serial "
                        "dependency chain of the reference's raster sweep (latency), not by HBM -- see DESIGN.md 5; "
                        "`kernels_roofline` lists every entry point")
    kernels_roofline = []
    for k, (c, t) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        ab = ALG_BYTES_PER_CELL_LAUNCH.get(k, 0.0) * ncell_rank
        kernels_roofline.append({"kernel": k, "share_of_step": (t / prof_steps) / (ms / args.steps),
                                 "avg_launch_ms": t / c,
                                 "achieved_GBps": ab / (t / c * 1e-3) / 1e9 if t > 0 else None,
                                 "frac": ab / (t / c * 1e-3) / 1e9 / peak if t > 0 else None,
                                 "alg_bytes_per_launch": ab,
                                 "traffic": NCU_TRAFFIC_BYTES_4097.get(k) if (world == 1 and N == 4097) else None,
                                 # the same fraction from the DRAM bytes ncu measured for this entry point
                                 "frac_by_traffic": (NCU_TRAFFIC_BYTES_4097[k] / (t / c * 1e-3) / 1e9 / peak)
                                 if (world == 1 and N == 4097 and k in NCU_TRAFFIC_BYTES_4097 and t > 0) else None})
    step_gbs = cells * ALG_BYTES_PER_CELL_STEP.get(args.scheme, 512.0) * args.steps / (ms * 1e-3) / 1e9
    breakdown = {k: {"calls": c, "ms_per_step": t / prof_steps} for k, (c, t) in
                 sorted(per_kernel.items(), key=lambda kv: -kv[1][1])}

    # ---- end to end: host state in, host state out, every step --------------
    e2e = None
    if not args.no_e2e:
        def host_step(hs):
            if world == 1:                    # the package's host-state entry point (copies overlapped with the step)
                return fsi_step_host(hs, prm)
            return solver.fsi_step_host(hs, sprm)   # the slab counterpart: same overlap of the PCIe traffic, per rank
        hstate = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in state)
        hstate = host_step(hstate)            # warm-up
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.e2e_steps):
            hstate = host_step(hstate)
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        nbytes = 5 * state[0].numel() * 8 * world
        e2e = {"value": cells * args.e2e_steps / (ems * 1e-3) / 1e6, "unit": "Mcell-steps/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": args.e2e_steps,
               "ms_per_step": ems / args.e2e_steps,
               "call": ("pyrmt_b200.driver.fsi_step_host: the five state fields pinned host -> device -> fsi step -> "
                        "pinned host every step, uploads ordered by first use and xi downloads overlapped with "
                        "the predictor/projection on copy streams") if world == 1 else
                       "pyrmt_b200.slab.SlabFSISolver.fsi_step_host: the five state fields (per rank: its slab) pinned host -> "
                       "device -> slab fsi step -> pinned host every step, copies overlapped like the single-GPU call"}

    # ---- N > 1: the part of the step that IS slab-decomposed (momentum + projection) ----
    # ---- the sharded sides of the path: the Neumann fluid half at 8193^2 (N > 1) and BASELINE configs[4]
    #      (periodic Taylor-Green: FSI at 8193^2, fluid half at 16385^2) -- at EVERY N including 1, so that
    #      the driver's scaling run carries its own 1 -> 8 curve for them
    slab = cfg5 = also_L32 = None
    del state
    torch.cuda.empty_cache()
    if world == 1 and not args.no_slab:
        # "run L = 1 as the headline and also L = 32 (dx = 1/128)" -- the same grid, discs and
        # kernels on the larger box, where the reference's absolute |det| > 1e-10 gate (functions.py:155) is wide open
        st32, prm32 = make_case(N, L=32.0, k_side=8, R_frac=0.04, scheme=args.scheme, bc_kind="lid")
        for _ in range(3):
            st32 = fsi_step(st32, prm32)[0]
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(5):
            st32 = fsi_step(st32, prm32)[0]
        g1.record()
        torch.cuda.synchronize()
        also_L32 = {"what": "the same step on the L = 32 box (dx = 1/128), 5 steps after 3", "grid": [N, N], "L": 32.0,
                    "ms_per_step": g0.elapsed_time(g1) / 5, "value": cells * 5 / (g0.elapsed_time(g1) * 1e-3) / 1e6,
                    "unit": "Mcell-steps/s", "finite": bool(torch.isfinite(st32[0]).all().item())}
        del st32, prm32
        torch.cuda.empty_cache()
    if not args.no_slab:
        if world > 1:
            slab = slab_fluid_rate(args.slab_size, 10, rank, world)
        # SURVEY 8d config 5 (periodic Taylor-Green, distributed FFT solve): the full FSI step at the largest
        # size where the reference's absolute-coordinate LSQ is still meaningful, the fluid half at 16385^2
        from pyrmt_b200.slab import time_periodic_fluid, time_periodic_fsi
        torch.cuda.empty_cache()
        cfg5 = {"fsi_step": time_periodic_fsi(args.cfg5_fsi_size, world, rank, steps=5, warmup=3)}
        torch.cuda.empty_cache()
        cfg5["fsi_step_semilagrangian"] = time_periodic_fsi(args.cfg5_fsi_size, world, rank, steps=5, warmup=3,
                                                            scheme="semilagrangian")
        torch.cuda.empty_cache()
        cfg5["fluid_step"] = time_periodic_fluid(args.cfg5_fluid_size, world, rank, steps=5, warmup=3)
        torch.cuda.empty_cache()
        if world > 1:        # P GPUs against one at the full size of config 5 (untimed; every rank holds the whole grid too)
            from pyrmt_b200.slab import periodic_fluid_parity
            try:
                cfg5["fluid_step_parity_vs_1gpu"] = periodic_fluid_parity(args.cfg5_fluid_size, world, rank, steps=2)
            except (ValueError, RuntimeError, NotImplementedError) as e:
                cfg5["fluid_step_parity_vs_1gpu"] = {"error": repr(e)[:300], "ok": False}
            torch.cuda.empty_cache()
        cfg5["note"] = ("full FSI at 16385^2 is not runnable with the reference's own algorithm (absolute-coordinate "
                        "normal equations lose all significance at index ~16384: pinned by "
                        "tests/test_gpu_parity.py::test_extrapolation_large_row_offset_matches_oracle, DESIGN.md 6): "
                        "8193^2 is the FSI size, 16385^2 the fluid-half size")

    # ---- parity, untimed: N = 1 -> one GPU step from the CPU arm's state against the CPU arm's own next step
    #      (rel L-inf on u, v, p, xi; xi bit for bit);  N > 1 -> the slab-decomposed step against the
    #      single-GPU step of the same 1025^2 problem (every rank runs the single-GPU step redundantly)
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = args.ref_size
        rate, cms, kind, threads, (cprm, cstate) = cpu_rate(n, args.cpu_steps, 1, args.scheme)
        threads = threads or (os.cpu_count() or 1)
        cpu = {"value": rate, "unit": "Mcell-steps/s", "cores": threads, "host_cores": os.cpu_count() or 1,
               "kind": kind,
               "sample": "%d steps of the same workload on a %dx%d node grid (1/%d of the cells), %.0f ms/step; "
                         % (args.cpu_steps, n, n, round((N / n) ** 2), cms) + CPU_KIND_TEXT[kind] % threads}
        from oracle import make_ref, rmt_oracle as O
        M = make_ref.load() if kind == "reference" else O
        ref_next = cpu_fsi_step(M, cstate, cprm)
        _, gprm = make_case(n, L=1.0, k_side=8, R_frac=0.04, scheme=args.scheme, bc_kind="lid")
        up = lambda t: torch.from_numpy(np.ascontiguousarray(t)).cuda()
        got = fsi_step(tuple(up(t) for t in cstate), dict(gprm, X=None if args.scheme == "weno5" else gprm["X"]))[0]
        rel, exact = {}, True
        for nm, g, r in zip(("u", "v", "p", "xi1", "xi2"), got, ref_next):
            g = g.cpu().numpy()
            rel[nm] = float(np.max(np.abs(g - r)) / max(np.max(np.abs(r)), 1e-300))
            if nm.startswith("xi"):
                exact = exact and bool(np.array_equal(g, r))
        parity = {"against": "the CPU arm (%s) on the same input state" % kind, "grid": [n, n], "rel_linf": rel,
                  "max_rel_linf": max(rel.values()), "xi_bit_exact": exact, "tolerance": 1e-10,
                  "ok": bool(max(rel.values()) <= 1e-10)}
    elif world > 1 and not args.no_slab:
        try:
            parity = slab_parity_vs_1gpu(1025, args.scheme, rank, world)
        except (ValueError, RuntimeError) as e:         # reported in the line, never at the cost of the line
            parity = {"error": repr(e)[:300], "ok": False}

    comm_kind = None
    if world > 1:
        from pyrmt_b200.slab import PeerComm, default_comm
        c = default_comm()
        comm_kind = type(c).__name__
        if isinstance(c, PeerComm):                    # did any on-stream barrier ever time out?  (say so; keep the line)
            try:
                c.check()
            except RuntimeError as e:
                comm_kind = "PeerComm (WATCHDOG: %s)" % e
    if rank == 0:
        line = {"metric": "Mcell-steps/s full FSI step", "value": value, "unit": "Mcell-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": workload_config(N, args.scheme, world),
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "step_roofline": {"alg_bytes_per_cell_step": ALG_BYTES_PER_CELL_STEP.get(args.scheme, 512.0),
                                  "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak},
                "cpu_baseline": cpu, "parity": parity, "kernels": breakdown, "kernels_roofline": kernels_roofline[:8],
                "finite": finite, "also_L32": also_L32, "slab_comm": comm_kind, "slab_fluid_step": slab,
                "config5_periodic": cfg5}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-GPU check + timing of the slab-decomposed fluid step (torchrun, one rank per GPU).

  torchrun --nproc-per-node P scripts/slab_check.py --check 1025 --time 8193 --steps 10

--check N : every rank also runs the single-GPU operators on the full N x N grid and compares
            its owned rows of (a, b, p) after 3 steps (lid-driven cavity + frozen discs).
--time  N : times the slab step at N x N (CUDA events, max over ranks)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--check", type=int, default=1025)
ap.add_argument("--time", type=int, default=0)
ap.add_argument("--fsi", type=int, default=0, help="check the full slab FSI step against the single-GPU step")
ap.add_argument("--fsi-time", type=int, default=0, help="time the full slab FSI step at N x N")
ap.add_argument("--pfsi", type=int, default=0, help="check the periodic (config 5) slab FSI step against the single-GPU step")
ap.add_argument("--pfsi-time", type=int, default=0, help="time the periodic Taylor-Green multi-disc slab FSI step at N x N")
ap.add_argument("--pfluid-time", type=int, default=0, help="time the periodic pure-fluid slab step (momentum + FFT projection) at N x N")
ap.add_argument("--pfsi-k", type=int, default=3, help="lattice side of the --pfsi check (discs of radius 0.24/k of the box)")
ap.add_argument("--L", type=float, default=0.0, help="domain side for --pfsi-time (default (N-1)/128, i.e. dx = 1/128)")
ap.add_argument("--scheme", default="weno5", help="advection scheme of the --fsi / --pfsi checks")
ap.add_argument("--overlap", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
SAME_GPU = os.environ.get("RMT_SAME_GPU") == "1"      # all ranks on cuda:0, gloo for the set-up (PeerComm carries the data)
torch.cuda.set_device(0 if SAME_GPU else local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if SAME_GPU:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def allmax(t):
    """MAX over the ranks of a small CUDA tensor (through the host when the group is gloo)."""
    if world > 1:
        if SAME_GPU:
            h = t.cpu()
            dist.all_reduce(h, op=dist.ReduceOp.MAX)
            return h.to(t.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t

from pyrmt_b200 import functions as F
from pyrmt_b200.driver import LidBC, disc_lattice
from pyrmt_b200.levelset import DiscSDF
from pyrmt_b200.slab import Comm, SlabFluidSolver, SlabLayout

up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def case(N, discs):
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    bc = LidBC(1.0)
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    Xd, Yd = up(X), up(Y)
    if discs:
        cx, cy, R = disc_lattice(4, 1.0, 0.08)
        phi = DiscSDF(cx, cy, R, domain=(1.0, 1.0))(Xd, Yd)
        X1 = F.mask_solid(Xd * 1.01, phi)       # a slightly stretched map: non-zero solid stress
        X2 = F.mask_solid(Yd * 0.99, phi)
        X1, X2 = F.extrapolate_reference_map(X1, X2, phi, dx, dy, 3)
    else:
        phi, X1, X2 = torch.ones_like(Xd), Xd.clone(), Yd.clone()
    a0, b0 = bc(np.zeros((N, N)), np.zeros((N, N)))
    prm = dict(dx=dx, dy=dy, mu_s=0.1 if discs else 0.0, kappa=0.0, eta_s=0.01 if discs else 0.0, rho_s=1.0,
               rho_f=1.0, mu_f=0.01, w_t=2 * dx)
    return bc, eig, (up(a0), up(b0), up(np.zeros((N, N)))), (X1, X2, phi), prm


out = {"world": world}
if args.check:
    N = args.check
    bc, eig, (a, b, p), (X1, X2, phi), prm = case(N, discs=True)
    lay = SlabLayout(N, N, world, rank, halo=4)
    solver = SlabFluidSolver(lay, bc, eig)
    sa, sb, sp = (lay.take(t).contiguous() for t in (a, b, p))
    s1, s2, sph = (lay.take(t).contiguous() for t in (X1, X2, phi))
    dt = 2e-4
    worst = 0.0
    for n in range(3):
        a_s, b_s, *_ = F.momentum_step_rk4(a, b, p, X1, X2, bc, prm["mu_s"], 0.0, prm["eta_s"], prm["dx"], prm["dy"],
                                          dt, 1.0, 1.0, phi, prm["mu_f"], prm["w_t"])
        a, b, p, _, _ = F.pressure_projection_amg(a_s, b_s, prm["dx"], prm["dy"], dt, 1.0, bc, p_prev=p, eigenvalues=eig)
        sa, sb, sp = solver.fluid_step(sa, sb, sp, s1, s2, sph, prm, dt)
        for ref, got in ((a, sa), (b, sb), (p, sp)):
            r = ref[lay.r0:lay.r1]
            err = float(((lay.owned(got) - r).abs().max() / ref.abs().max()).item())
            worst = max(worst, err)
        # halos must be valid too (they feed the next step)
        for ref, got in ((a, sa), (b, sb), (p, sp)):
            worst = max(worst, float(((got - ref[lay.e0:lay.e1]).abs().max() / ref.abs().max()).item()))
    t = torch.tensor([worst], device="cuda")
    t = allmax(t)
    out["check"] = {"N": N, "rel_linf_vs_single_gpu": float(t.item())}
    del a, b, p, X1, X2, phi, solver
    torch.cuda.empty_cache()

if args.fsi:
    from pyrmt_b200.driver import fsi_step
    from pyrmt_b200.slab import SlabFSISolver
    N = args.fsi
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    cx, cy, R = disc_lattice(3, 1.0, 0.08)        # the middle lattice row straddles the 2-rank cut
    sdf = DiscSDF(cx, cy, R, domain=(1.0, 1.0))
    bc = LidBC(1.0)
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    Xd, Yd = up(X), up(Y)
    phi0 = sdf(Xd, Yd)
    X1, X2 = F.extrapolate_reference_map(F.mask_solid(Xd, phi0), F.mask_solid(Yd, phi0), phi0, dx, dy, 3)
    a0, b0 = bc(np.zeros((N, N)), np.zeros((N, N)))
    state = (up(a0), up(b0), up(np.zeros((N, N))), X1, X2)
    prm = dict(dx=dx, dy=dy, CFL=0.2, dt_cap=1e-3, mu_s=0.1, kappa=0.0, rho_s=1.0, rho_f=1.0, eta_s=0.01,
               mu_f=0.01, w_t=2 * dx, layers=3, scheme=args.scheme, w_cut=0.0, phi_init=sdf, bc=bc, eig=eig,
               X=Xd, Y=Yd)
    lay = SlabLayout(N, N, world, rank, halo=12)
    solver = SlabFSISolver(lay, bc, eig, sdf, overlap=args.overlap, layers=3)
    sstate = tuple(lay.take(t).contiguous() for t in state)
    sprm = dict(prm, X=lay.take(Xd).contiguous(), Y=lay.take(Yd).contiguous())
    worst = {}
    for n in range(4):
        state, dt, _ = fsi_step(state, prm)
        sstate = solver.fsi_step(sstate, sprm, dt)
        for nm, ref, got in zip(("a", "b", "p", "X1", "X2"), state, sstate):
            err = float(((got - ref[lay.e0:lay.e1]).abs().max() / ref.abs().max()).item())   # owned rows + halos
            worst[nm] = max(worst.get(nm, 0.0), err)
    t = torch.tensor([worst[k] for k in ("a", "b", "p", "X1", "X2")], device="cuda")
    t = allmax(t)
    out["fsi_check"] = {"N": N, "steps": 4, "scheme": args.scheme, "rel_linf_vs_single_gpu": dict(zip(("a", "b", "p", "X1", "X2"), t.tolist()))}
    del state, sstate, solver, X1, X2, Xd, Yd, phi0
    torch.cuda.empty_cache()

if args.fsi_time:
    from pyrmt_b200.slab import SlabFSISolver
    N = args.fsi_time
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    k_side = 8 * max(1, (N - 1) // 4096)
    cx, cy, R = disc_lattice(k_side, 1.0, 0.04 * 4096.0 / (N - 1))
    sdf = DiscSDF(cx, cy, R, domain=(1.0, 1.0))
    bc = LidBC(1.0)
    lay = SlabLayout(N, N, world, rank, halo=12)
    solver = SlabFSISolver(lay, bc, F._precompute_poisson_eigenvalues(N, N, dx, dy), sdf, overlap=512, layers=3)
    Xs, Ys = up(lay.take(X)), up(lay.take(Y))
    del X, Y
    phi0 = sdf(Xs, Ys)
    # initial map: identity in the solid, extrapolated locally (every disc of the lattice lies inside one slab
    # or is re-swept through the overlap during the steps; the initial halo rows are exact copies)
    X1, X2 = F.mask_solid(Xs, phi0), F.mask_solid(Ys, phi0)
    B1, B2, Bp = solver._gather_big((X1, X2, phi0))
    E1, E2 = F.extrapolate_reference_map(B1, B2, Bp, dx, dy, 3, row_offset=lay.r0 - solver.top)
    n_own = lay.r1 - lay.r0
    X1, X2 = torch.zeros_like(Xs), torch.zeros_like(Xs)
    lay.owned(X1).copy_(E1[solver.top:solver.top + n_own]); lay.owned(X2).copy_(E2[solver.top:solver.top + n_own])
    solver.comm.halo_exchange(lay, (X1, X2))
    z = torch.zeros_like(Xs)
    a = z.clone()
    if rank == world - 1:
        a[lay.o1 - 1, 1:-1] = 1.0
    sstate = (a, z.clone(), z.clone(), X1, X2)
    prm = dict(dx=dx, dy=dy, mu_s=0.1, kappa=0.0, rho_s=1.0, rho_f=1.0, eta_s=0.01, mu_f=0.01, w_t=2 * dx,
               scheme="weno5", w_cut=0.0, X=None, Y=None)
    dt = 0.2 * dx * dx / (4 * 0.01)
    for _ in range(3):
        sstate = solver.fsi_step(sstate, prm, dt)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sstate = solver.fsi_step(sstate, prm, dt, check_guard=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    ms = allmax(ms)
    ms = float(ms.item()) / args.steps
    out["fsi_time"] = {"N": N, "discs": int(cx.size), "ms_per_step": ms, "Mcell_steps_per_s": N * N / ms / 1e3,
                       "finite": bool(torch.isfinite(sstate[0]).all().item())}
    del sstate, solver
    torch.cuda.empty_cache()

from pyrmt_b200.slab import taylor_green as tg_velocity


if args.pfsi:
    from pyrmt_b200.driver import PeriodicBC, fsi_step
    from pyrmt_b200.slab import SlabFSISolver, slab_initial_state
    N, L = args.pfsi, 8.0
    X, Y, dx, dy = F.create_grid(N, N, L, L)
    cx, cy, R = disc_lattice(args.pfsi_k, L, 0.24 / args.pfsi_k, jitter=0.03 / args.pfsi_k)   # odd k: a lattice row straddles the middle cut
    sdf, bc = DiscSDF(cx, cy, R, domain=(L, L)), PeriodicBC()
    eig = F._precompute_poisson_eigenvalues_periodic(N, N, dx, dy)
    Xd, Yd = up(X), up(Y)
    phi0 = sdf(Xd, Yd)
    X1, X2 = F.extrapolate_reference_map(F.mask_solid(Xd, phi0), F.mask_solid(Yd, phi0), phi0, dx, dy, 3)
    a0, b0 = bc(*tg_velocity(L)(X, Y))
    state = (up(a0), up(b0), up(np.zeros((N, N))), X1, X2)
    prm = dict(dx=dx, dy=dy, CFL=0.2, dt_cap=1e-3, mu_s=0.1, kappa=0.0, rho_s=1.0, rho_f=1.0, eta_s=0.01,
               mu_f=0.01, w_t=2 * dx, layers=3, scheme=args.scheme, w_cut=0.0, phi_init=sdf, bc=bc, eig=eig,
               bc_type="periodic", X=Xd, Y=Yd)
    lay = SlabLayout(N, N, world, rank, halo=12, periodic=True)
    solver = SlabFSISolver(lay, bc, None, sdf, overlap=args.overlap, layers=3, spacing=(dx, dy))
    sstate, _, _ = slab_initial_state(solver, L, sdf, tg_velocity(L))
    worst = {}
    for nm, ref, got in zip(("a", "b", "p", "X1", "X2"), state, sstate):
        worst[nm] = float(((got - ref[lay.e0:lay.e1]).abs().max() / max(float(ref.abs().max()), 1e-300)).item())
    sprm = dict(prm, X=solver.coords[0], Y=solver.coords[1])
    for n in range(4):
        state, dt, _ = fsi_step(state, prm)
        sstate = solver.fsi_step(sstate, sprm, dt)
        for nm, ref, got in zip(("a", "b", "p", "X1", "X2"), state, sstate):
            err = float(((got - ref[lay.e0:lay.e1]).abs().max() / ref.abs().max()).item())   # owned rows + halos
            worst[nm] = max(worst.get(nm, 0.0), err)
            if os.environ.get("RMT_CHECK_DEBUG") and nm in ("X1", "X2") and err > 0:
                d = (got - ref[lay.e0:lay.e1]).abs()
                k = int(d.argmax()); jj, ii = k // N, k % N
                print("[rank %d] step %d %s differs at stored row %d (global %d, owned [%d,%d)) col %d: %r vs %r; "
                      "#cells %d" % (rank, n, nm, jj, jj + lay.e0, lay.r0, lay.r1, ii, float(got[jj, ii]),
                                     float(ref[jj + lay.e0, ii]), int((d > 0).sum())), flush=True)
    t = torch.tensor([worst[k] for k in ("a", "b", "p", "X1", "X2")], device="cuda")
    t = allmax(t)
    out["periodic_fsi_check"] = {"N": N, "steps": 4, "scheme": args.scheme,
                                 "rel_linf_vs_single_gpu": dict(zip(("a", "b", "p", "X1", "X2"), t.tolist()))}
    del state, sstate, solver, X1, X2, Xd, Yd, phi0, eig, X, Y
    torch.cuda.empty_cache()

if args.pfsi_time:
    from pyrmt_b200.slab import time_periodic_fsi
    out["periodic_fsi_time"] = time_periodic_fsi(args.pfsi_time, world, rank, steps=args.steps, L=args.L or None)
    torch.cuda.empty_cache()

if args.pfluid_time:
    from pyrmt_b200.slab import time_periodic_fluid
    out["periodic_fluid_time"] = time_periodic_fluid(args.pfluid_time, world, rank, steps=args.steps)
    torch.cuda.empty_cache()

if args.time:
    N = args.time
    lay = SlabLayout(N, N, world, rank, halo=4)
    X, Y, dx, dy = F.create_grid(N, N, 1.0, 1.0)
    bc = LidBC(1.0)
    eig = F._precompute_poisson_eigenvalues(N, N, dx, dy)
    solver = SlabFluidSolver(lay, bc, eig)
    del eig
    z = torch.zeros((lay.nl, N), dtype=torch.float64, device="cuda")
    sa, sb, sp = z.clone(), z.clone(), z.clone()
    if rank == world - 1:
        sa[lay.o1 - 1, 1:-1] = 1.0
    s1, s2, sph = up(lay.take(X)), up(lay.take(Y)), torch.ones_like(z)
    del X, Y
    prm = dict(dx=dx, dy=dy, mu_s=0.0, kappa=0.0, eta_s=0.0, rho_s=1.0, rho_f=1.0, mu_f=0.01, w_t=2 * dx)
    dt = 0.2 * dx * dx / (4 * 0.01)
    for _ in range(3):
        sa, sb, sp = solver.fluid_step(sa, sb, sp, s1, s2, sph, prm, dt)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sa, sb, sp = solver.fluid_step(sa, sb, sp, s1, s2, sph, prm, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    ms = allmax(ms)
    ms = float(ms.item()) / args.steps
    out["time"] = {"N": N, "ms_per_fluid_step": ms, "Mcell_steps_per_s": N * N / ms / 1e3,
                   "finite": bool(torch.isfinite(sa).all().item())}
from pyrmt_b200.slab import PeerComm, default_comm
comm = default_comm()
out["comm"] = type(comm).__name__
if isinstance(comm, PeerComm):
    comm.check()                     # no barrier ever timed out
    out["peer_barriers"] = comm.epoch
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

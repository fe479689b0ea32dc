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
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["check"] = {"N": N, "rel_linf_vs_single_gpu": float(t.item())}
    del a, b, p, X1, X2, phi, solver
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
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / args.steps
    out["time"] = {"N": N, "ms_per_fluid_step": ms, "Mcell_steps_per_s": N * N / ms / 1e3,
                   "finite": bool(torch.isfinite(sa).all().item())}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

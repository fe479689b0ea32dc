"""Slab decomposition over the GPUs of one node: y-slabs with halo rows,
nearest-neighbour halo exchange for the stencil kernels, and the distributed DCT-I
Poisson solve (local row transforms, all-to-all transpose, local column solve,
all-to-all back).  One process per GPU, torch.distributed (NCCL over NVLink) for the
exchanges, the same C-ABI kernels as the single-GPU path on every rank.

What is sharded in this round: the momentum predictor (`momentum_step_rk4`) and the
constant-density Neumann projection (`pressure_projection_amg`) -- i.e. the pure-fluid
step of benchmarks/lid_driven_cavity.py:58-80, and the fluid half of the FSI step with
the solid fields supplied as slabs.  Reference-map advection and extrapolation across
slab boundaries are the next rows.

Layout: rank k owns rows [r0, r1) of the (Ny, Nx) grid and stores rows
[e0, e1) = [r0 - H, r1 + H) clipped to the grid ("extended slab").  Kernels run on the
extended slab as if it were a whole grid: their one-sided rim stencils then only
contaminate halo rows next to an interior cut, which the next halo exchange overwrites;
at the true domain edge the slab edge IS the grid edge, so the rim handling is exact.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._runtime import F64, ctx, ptr, stream


def split(n, parts):
    """Contiguous near-even partition of range(n): list of (start, stop)."""
    base, extra = divmod(n, parts)
    out, s = [], 0
    for k in range(parts):
        e = s + base + (1 if k < extra else 0)
        out.append((s, e))
        s = e
    return out


class SlabLayout:
    def __init__(self, Ny, Nx, world, rank, halo=4):
        self.Ny, self.Nx, self.world, self.rank, self.H = Ny, Nx, world, rank, halo
        self.rows = split(Ny, world)
        self.cols = split(Nx, world)
        self.r0, self.r1 = self.rows[rank]
        self.c0, self.c1 = self.cols[rank]
        if min(e - s for s, e in self.rows) < halo:
            raise ValueError("slabs of %d rows are thinner than the halo (%d)" % (Ny // world, halo))
        self.e0, self.e1 = max(self.r0 - halo, 0), min(self.r1 + halo, Ny)
        self.nl = self.e1 - self.e0                 # rows stored
        self.o0, self.o1 = self.r0 - self.e0, self.r1 - self.e0   # owned rows inside the slab

    def take(self, full):
        """Extended slab of a full (Ny, Nx) host/device array."""
        return full[self.e0:self.e1]

    def owned(self, slab):
        return slab[self.o0:self.o1]


# ---------------------------------------------------------------------------- comms
class Comm:
    """The two exchange patterns, on top of torch.distributed (NCCL on GPUs; the same
    code runs on gloo/CPU in the tests, where all_to_all_single does not exist)."""

    def __init__(self, group=None):
        self.on = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.nccl = self.on and dist.get_backend(group) == "nccl"

    def halo_exchange(self, lay, fields, width=None):
        """Fill the halo rows of every extended slab in `fields` from the neighbours' owned rows."""
        if self.world == 1:
            return
        H = lay.H if width is None else width
        ops = []
        for f in fields:
            if lay.rank + 1 < lay.world:                      # upper neighbour
                ops.append(dist.P2POp(dist.isend, f[lay.o1 - H:lay.o1], self.rank + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, f[lay.o1:lay.o1 + H], self.rank + 1, self.group))
            if lay.rank > 0:                                  # lower neighbour
                ops.append(dist.P2POp(dist.isend, f[lay.o0:lay.o0 + H], self.rank - 1, self.group))
                ops.append(dist.P2POp(dist.irecv, f[lay.o0 - H:lay.o0], self.rank - 1, self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def all_to_all(self, send, send_counts, recv, recv_counts):
        """Variable-size all-to-all of flat fp64 buffers (element counts per peer)."""
        if self.world == 1:
            recv.copy_(send)
            return
        if self.nccl:
            dist.all_to_all_single(recv, send, list(recv_counts), list(send_counts), group=self.group)
            return
        so = np.concatenate([[0], np.cumsum(send_counts)])
        ro = np.concatenate([[0], np.cumsum(recv_counts)])
        ops = []
        for q in range(self.world):
            if q == self.rank:
                recv[ro[q]:ro[q + 1]].copy_(send[so[q]:so[q + 1]])
                continue
            ops.append(dist.P2POp(dist.isend, send[so[q]:so[q + 1]], q, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[ro[q]:ro[q + 1]], q, self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def allreduce(self, t, op="sum"):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)
        return t


# ------------------------------------------------------------------- device line ops
class CudaOps:
    """The three device primitives of the distributed transform (C ABI)."""

    def dct_lines(self, x, eig=None, scale=1.0):
        nrows, N = x.shape
        _lib.check(ctx().lib.rmt_dct_lines(ptr(x), ptr(x), ptr(eig), nrows, N, float(scale), stream()),
                   "rmt_dct_lines")
        return x

    def transpose(self, x):
        R, C = x.shape
        out = torch.empty((C, R), dtype=x.dtype, device=x.device)
        _lib.check(ctx().lib.rmt_transpose(ptr(x), ptr(out), R, C, stream()), "rmt_transpose")
        return out

    def copy2d(self, src, dst):
        """dst[:, :] = src[:, :] for 2-D views with unit column stride."""
        rows, cols = src.shape
        _lib.check(ctx().lib.rmt_copy2d(src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0), dst.stride(0),
                                        stream()), "rmt_copy2d")


class DistPoissonDCT:
    """_solve_poisson_dct (functions.py:1107-1119) over row slabs.

        rows DCT (local) -> transpose -> all-to-all -> columns: DCT, x 1/(4 Mx My eig), DCT (local)
        -> all-to-all back -> transpose -> rows DCT (local);  sum(sol) by all-reduce.

    Per rank and solve the all-to-all moves (nr x Nx) doubles out and in, twice."""

    def __init__(self, lay, eig, comm=None, ops=None, device=None):
        self.lay, self.comm, self.ops = lay, comm or Comm(), ops or CudaOps()
        Ny, Nx = lay.Ny, lay.Nx
        eig = np.asarray(eig, dtype=np.float64)
        if eig.shape != (Ny, Nx):
            raise ValueError("eigenvalues shape %s does not match the grid %s" % (eig.shape, (Ny, Nx)))
        eT = np.ascontiguousarray(eig[:, lay.c0:lay.c1].T)          # (nc_me, Ny): my columns as lines
        self.eigT = torch.from_numpy(eT).to(device) if device is not None else torch.from_numpy(eT)
        self.scale = 1.0 / (4.0 * (Nx - 1) * (Ny - 1))
        self.nr = [e - s for s, e in lay.rows]
        self.nc = [e - s for s, e in lay.cols]

    def solve(self, rhs_owned):
        """rhs_owned: (nr_me, Nx) contiguous.  Returns (sol_owned, global sum of sol as a 1-element tensor)."""
        lay, ops, comm = self.lay, self.ops, self.comm
        me, P = lay.rank, lay.world
        nr_me, nc_me, Ny, Nx = self.nr[me], self.nc[me], lay.Ny, lay.Nx
        A = ops.dct_lines(rhs_owned.clone())                          # rows, local
        At = ops.transpose(A)                                          # (Nx, nr_me): column blocks contiguous
        send_counts = [self.nc[q] * nr_me for q in range(P)]
        recv_counts = [nc_me * self.nr[q] for q in range(P)]
        rbuf = torch.empty(nc_me * Ny, dtype=At.dtype, device=At.device)
        comm.all_to_all(At.reshape(-1), send_counts, rbuf, recv_counts)
        B = torch.empty((nc_me, Ny), dtype=At.dtype, device=At.device)
        off = 0
        for q, (s, e) in enumerate(lay.rows):                          # unpack: block q is (nc_me, nr_q)
            ops.copy2d(rbuf[off:off + nc_me * (e - s)].view(nc_me, e - s), B[:, s:e])
            off += nc_me * (e - s)
        ops.dct_lines(B, eig=self.eigT, scale=self.scale)              # columns: fwd, 1/eig, inverse
        sbuf = torch.empty(nc_me * Ny, dtype=At.dtype, device=At.device)
        off = 0
        for q, (s, e) in enumerate(lay.rows):                          # pack
            ops.copy2d(B[:, s:e], sbuf[off:off + nc_me * (e - s)].view(nc_me, e - s))
            off += nc_me * (e - s)
        At2 = torch.empty(Nx * nr_me, dtype=At.dtype, device=At.device)
        comm.all_to_all(sbuf, recv_counts, At2, send_counts)          # blocks (nc_q, nr_me) = rows of At
        sol = ops.dct_lines(ops.transpose(At2.view(Nx, nr_me)))        # (nr_me, Nx), rows again
        total = comm.allreduce(sol.sum().reshape(1))
        return sol, total


# ------------------------------------------------------------------ BC table per slab
def local_bc_table(bc, lay):
    """The global gather table of a BC callable restricted to the rows this rank owns,
    re-indexed into the extended slab.  Sources must live in the same slab (true for
    wall / lid / free-slip BCs; a periodic wrap across ranks is not supported here)."""
    from .bc import BCTable, _FIELD_BIT, table_for
    full = table_for(bc, lay.Ny, lay.Nx)
    if full is None:
        raise NotImplementedError("slab decomposition needs a rim-gather BC callable")
    dst, src, ca, cb = full.host
    Nx = lay.Nx
    cell = lambda k: k & (_FIELD_BIT - 1)
    keep = (cell(dst) // Nx >= lay.r0) & (cell(dst) // Nx < lay.r1)
    dst, src, ca, cb = dst[keep], src[keep], ca[keep], cb[keep]
    has = src >= 0
    srow = cell(src[has]) // Nx
    if np.any((srow < lay.e0) | (srow >= lay.e1)):
        raise NotImplementedError("BC copies across slab boundaries (e.g. a periodic wrap) are not supported")
    shift = lay.e0 * Nx
    rel = lambda k: (k & _FIELD_BIT) | (cell(k) - shift)
    src2 = src.copy()
    src2[has] = rel(src[has])
    return BCTable(rel(dst).astype(np.int64), src2.astype(np.int64), ca, cb, (lay.nl, Nx))


class SlabFluidSolver:
    """momentum_step_rk4 + pressure_projection_amg (Neumann, constant density) on row slabs.
    All field arguments and results are extended slabs (lay.nl, Nx) with valid halos."""

    def __init__(self, lay, bc, eig, comm=None):
        self.lay, self.comm = lay, comm or Comm()
        self.table = local_bc_table(bc, lay)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.poisson = DistPoissonDCT(lay, eig, self.comm, device=dev)
        self.lib = ctx().lib

    # -- helpers -----------------------------------------------------------------
    def _bc_and_halo(self, u, v):
        self.table.apply_(u, v)
        self.comm.halo_exchange(self.lay, (u, v))

    def max_speed(self, a, b):
        lay = self.lay
        out = ctx().max_speed(lay.owned(a), lay.owned(b))
        return self.comm.allreduce(out[:1].clone(), "max")

    # -- functions.py:673-762 ----------------------------------------------------------
    def momentum_step(self, u, v, p, X1, X2, phi, mu_s, kappa, eta_s, dx, dy, dt, rho_s, rho_f, mu_f, w_t):
        lay, lib, st = self.lay, self.lib, stream()
        nl, Nx = lay.nl, lay.Nx
        sxx, sxy, syy, J = (torch.empty_like(u) for _ in range(4))
        _lib.check(lib.rmt_solid_stress(ptr(X1), ptr(X2), ptr(phi), ptr(sxx), ptr(sxy), ptr(syy), ptr(J), nl, Nx,
                                        dx, dy, mu_s, kappa, 0.0, 0.0, 0, st), "rmt_solid_stress")
        sa_u, sa_v = u.clone(), v.clone()
        self._bc_and_halo(sa_u, sa_v)
        sb_u, sb_v = torch.empty_like(u), torch.empty_like(u)
        acc_u, acc_v = torch.empty_like(u), torch.empty_like(u)
        un, vn = torch.empty_like(u), torch.empty_like(u)

        def stage(k, iu, iv, ou, ov):
            _lib.check(lib.rmt_momentum_stage(ptr(iu), ptr(iv), ptr(p), ptr(sxx), ptr(sxy), ptr(syy), ptr(phi),
                                              None, None, ptr(u), ptr(v), ptr(acc_u), ptr(acc_v), ptr(ou),
                                              ptr(ov), nl, Nx, dx, dy, dt, mu_f, eta_s, w_t, rho_s, rho_f, k, st),
                       "rmt_momentum_stage")
            self._bc_and_halo(ou, ov)

        stage(1, sa_u, sa_v, sb_u, sb_v)
        stage(2, sb_u, sb_v, sa_u, sa_v)
        stage(3, sa_u, sa_v, sb_u, sb_v)
        stage(4, sb_u, sb_v, un, vn)
        return un, vn, sxx, sxy, syy, J

    # -- functions.py:1255-1364 (Neumann, constant density) ---------------------------------
    def projection(self, a_star, b_star, p_prev, rho, dx, dy, dt):
        lay, lib, st, comm = self.lay, self.lib, stream(), self.comm
        nl, Nx, ncell = lay.nl, lay.Nx, lay.Ny * lay.Nx
        if isinstance(rho, torch.Tensor):
            rsum = comm.allreduce(lay.owned(rho).sum().reshape(1))
            rd, rscalar = rho, 0.0
        else:
            rd, rscalar = None, float(rho)
            rsum = torch.full((1,), rscalar * ncell, dtype=F64, device=a_star.device)
        # the kernels divide the device sum by THEIR grid size: rescale to the local one
        rsum_local = rsum * (float(nl * Nx) / float(ncell))
        rhs = torch.empty_like(a_star)
        _lib.check(lib.rmt_projection_rhs(ptr(a_star), ptr(b_star), ptr(p_prev), ptr(rd), rscalar,
                                          ptr(rsum_local), ptr(rhs), nl, Nx, dx, dy, dt, 0, st),
                   "rmt_projection_rhs")
        sol_o, total = self.poisson.solve(lay.owned(rhs).contiguous())
        sol = torch.zeros_like(a_star)
        lay.owned(sol).copy_(sol_o)
        comm.halo_exchange(lay, (sol,))
        ssum_local = total * (float(nl * Nx) / float(ncell))
        a, b, p = torch.empty_like(a_star), torch.empty_like(a_star), torch.empty_like(a_star)
        _lib.check(lib.rmt_projection_correct(ptr(sol), ptr(ssum_local), ptr(a_star), ptr(b_star), ptr(rd),
                                              rscalar, ptr(p_prev), ptr(a), ptr(b), ptr(p), nl, Nx, dx, dy, dt,
                                              0, st), "rmt_projection_correct")
        self.table.apply_(a, b)
        psum = comm.allreduce(lay.owned(p).sum().reshape(1)) * (float(nl * Nx) / float(ncell))
        _lib.check(lib.rmt_subtract_mean(ptr(p), ptr(psum), p.numel(), st), "rmt_subtract_mean")
        comm.halo_exchange(lay, (a, b, p))
        return a, b, p

    def fluid_step(self, a, b, p, X1, X2, phi, prm, dt):
        """One pure-fluid / frozen-solid step: momentum predictor + projection."""
        a_s, b_s, *_ = self.momentum_step(a, b, p, X1, X2, phi, prm["mu_s"], prm["kappa"], prm["eta_s"],
                                          prm["dx"], prm["dy"], dt, prm["rho_s"], prm["rho_f"], prm["mu_f"],
                                          prm["w_t"])
        return self.projection(a_s, b_s, p, prm["rho_f"], prm["dx"], prm["dy"], dt)

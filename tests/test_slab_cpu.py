"""CPU tests of the slab-decomposition host logic: partition arithmetic, the
per-slab BC table, and -- with two gloo processes -- the halo exchange and the all-to-all
bookkeeping of the distributed DCT solve (device kernels replaced by SciPy test doubles;
the same Comm / DistPoissonDCT code runs on NCCL on the GPUs)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_split_and_layout():
    from pyrmt_b200.slab import SlabLayout, split
    assert split(10, 3) == [(0, 4), (4, 7), (7, 10)]
    for Ny, P in ((129, 2), (4097, 8), (33, 4)):
        rows = split(Ny, P)
        assert rows[0][0] == 0 and rows[-1][1] == Ny and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        for r in range(P):
            L = SlabLayout(Ny, 17, P, r, halo=4)
            assert L.e0 == max(L.r0 - 4, 0) and L.e1 == min(L.r1 + 4, Ny)
            assert L.o1 - L.o0 == L.r1 - L.r0 and L.nl == L.e1 - L.e0
    with pytest.raises(ValueError):
        SlabLayout(16, 16, 8, 0, halo=4)


def test_lid_table_matches_classifier():
    from pyrmt_b200.bc import classify
    from pyrmt_b200.driver import LidBC
    bc = LidBC(1.5)
    for Ny, Nx in ((9, 13), (28, 36)):
        t1, t2 = bc.rmt_table(Ny, Nx), classify(lambda u, v: bc(u, v), Ny, Nx)
        rng = np.random.default_rng(0)
        u, v = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
        a1, a2, ref = t1.apply_host(u, v), t2.apply_host(u, v), bc(u, v)
        assert np.array_equal(a1[0], ref[0]) and np.array_equal(a1[1], ref[1])
        assert np.array_equal(a2[0], ref[0]) and np.array_equal(a2[1], ref[1])


def test_periodic_table_matches_callable():
    from pyrmt_b200.bc import periodic_bc
    from pyrmt_b200.driver import PeriodicBC
    for Ny, Nx in ((9, 13), (28, 36)):
        t = PeriodicBC().rmt_table(Ny, Nx)
        rng = np.random.default_rng(2)
        u, v = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
        got, ref = t.apply_host(u, v), periodic_bc(u, v)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_local_bc_table_reproduces_global_bc():
    from pyrmt_b200.bc import free_slip_box_bc
    from pyrmt_b200.driver import LidBC
    from pyrmt_b200.slab import SlabLayout, local_bc_table
    Ny, Nx, P = 40, 23, 3
    rng = np.random.default_rng(1)
    u, v = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
    for bc in (LidBC(2.0), free_slip_box_bc):
        ru, rv = bc(u, v)
        for r in range(P):
            L = SlabLayout(Ny, Nx, P, r, halo=4)
            t, rem = local_bc_table(bc, L)
            assert rem is None
            su, sv = t.apply_host(L.take(u).copy(), L.take(v).copy())
            assert np.array_equal(L.owned(su), ru[L.r0:L.r1]) and np.array_equal(L.owned(sv), rv[L.r0:L.r1])
            # extended table (stages that compute their halo rows locally): the BC on EVERY stored row
            te, rem = local_bc_table(bc, L, extended=True)
            assert rem is None
            eu, ev = te.apply_host(L.take(u).copy(), L.take(v).copy())
            assert np.array_equal(eu, ru[L.e0:L.e1]) and np.array_equal(ev, rv[L.e0:L.e1])


class SciPyOps:
    """Test doubles of the three device primitives."""

    def dct_lines(self, x, eig=None, scale=1.0):
        from scipy.fft import dct
        y = dct(x.numpy(), type=1, axis=1)
        if eig is not None:
            y = dct(y * (scale / eig.numpy()), type=1, axis=1)
        else:
            y = y * scale
        x.copy_(torch.from_numpy(np.ascontiguousarray(y)))
        return x

    def dht_lines(self, x, out, m, mul=None, scale=1.0):
        Fx = np.fft.fft(x.numpy()[:, :m], axis=1)
        h = Fx.real - Fx.imag
        h = h * (mul.numpy() if mul is not None else scale)
        out[:, :m].copy_(torch.from_numpy(np.ascontiguousarray(h)))
        return out

    def sum(self, x):
        return x.sum().reshape(1)

    def transpose(self, x):
        return x.t().contiguous()

    def copy2d(self, src, dst):
        dst.copy_(src)


def _worker(rank, world, port, Ny, Nx, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rmt_oracle as O
        from pyrmt_b200.slab import Comm, DistPoissonDCT, SlabLayout
        lay = SlabLayout(Ny, Nx, world, rank, halo=4)
        comm = Comm()
        # halo exchange: every rank fills its slab with global row ids, halos poisoned
        f = torch.full((lay.nl, Nx), -1.0, dtype=torch.float64)
        g = torch.full((lay.nl, Nx), -1.0, dtype=torch.float64)
        rows = torch.arange(lay.r0, lay.r1, dtype=torch.float64)[:, None]
        lay.owned(f).copy_(rows.expand(-1, Nx))
        lay.owned(g).copy_(1000.0 + rows.expand(-1, Nx))
        comm.halo_exchange(lay, (f, g))
        want = torch.arange(lay.e0, lay.e1, dtype=torch.float64)[:, None].expand(-1, Nx)
        ok_halo = bool(torch.equal(f, want) and torch.equal(g, want + 1000.0))
        # distributed DCT solve against the serial oracle
        X, Y, dx, dy = O.create_grid(Nx, Ny, 1.3, 0.7)
        eig = O._precompute_poisson_eigenvalues(Nx, Ny, dx, dy)
        rng = np.random.default_rng(7)
        rhs = rng.standard_normal((Ny, Nx))
        ref = O._solve_poisson_dct(rhs, eig)
        solver = DistPoissonDCT(lay, eig, comm, ops=SciPyOps())
        sol, total = solver.solve(torch.from_numpy(np.ascontiguousarray(rhs[lay.r0:lay.r1])))
        sol = sol.numpy() - float(total) / (Ny * Nx)
        err = float(np.max(np.abs(sol - ref[lay.r0:lay.r1])) / np.max(np.abs(ref)))
        # round 2: the solution written straight into a caller's slab with the partial sum all-reduced on the
        # barrier of the next halo exchange (one row wide), and the overlap gather of the extrapolation
        ext = torch.full((lay.nl, Nx), -5.0, dtype=torch.float64)
        part = solver.solve(torch.from_numpy(np.ascontiguousarray(rhs[lay.r0:lay.r1])), out=lay.owned(ext),
                            reduce=False)
        comm.halo_exchange(lay, (ext,), width=1, reduce=((part, "sum"),))
        lo, hi = (lay.o0 - 1 if rank > 0 else lay.o0), (lay.o1 + 1 if rank + 1 < world else lay.o1)
        got = ext[lo:hi].numpy() - float(part) / (Ny * Nx)
        err = max(err, float(np.max(np.abs(got - ref[lay.e0 + lo:lay.e0 + hi])) / np.max(np.abs(ref))))
        err = max(err, abs(float(part) - float(total)) / max(abs(float(total)), 1e-300))
        glob = torch.from_numpy(rng.standard_normal((Ny, Nx)))
        top_of = lambda q: min(5, lay.rows[q][0])
        bot_of = lambda q: min(3, Ny - lay.rows[q][1])
        (big,) = comm.gather_overlap(lay, (lay.take(glob).clone(),), top_of(rank), bot_of(rank), top_of, bot_of)
        ok_halo = ok_halo and bool(torch.equal(big, glob[lay.r0 - top_of(rank):lay.r1 + bot_of(rank)]))
        mx = comm.allreduce(torch.tensor([float(rank + 3)]), "max")
        # collective verdicts: one rank's failure must become every rank's verdict
        agree = comm.all_agree([True, rank != world - 1, rank != 0])
        mn = comm.allreduce(torch.tensor([float(rank + 3)]), "min")
        out[rank] = (ok_halo, err, float(mx), agree, float(mn))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Ny,Nx", [(2, 33, 17), (3, 40, 29)])
def test_halo_exchange_and_distributed_dct_gloo(world, Ny, Nx):
    ctx_ = mp.get_context("spawn")
    with ctx_.Manager() as m:
        out = m.dict()
        port = 29500 + (os.getpid() % 2000) + world
        procs = [ctx_.Process(target=_worker, args=(r, world, port, Ny, Nx, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        for r in range(world):
            ok_halo, err, mx, agree, mn = out[r]
            assert ok_halo, "halo exchange wrong on rank %d" % r
            assert err < 1e-12, (r, err)
            assert mx == world + 2 and mn == 3.0
            assert agree == [True, False, False], "rank %d would not raise with the others" % r


def test_periodic_layout():
    from pyrmt_b200.slab import SlabLayout
    for Ny, Nx, P in ((33, 17, 2), (65, 33, 4), (41, 29, 3)):
        for r in range(P):
            L = SlabLayout(Ny, Nx, P, r, halo=4, periodic=True)
            assert L.rows[-1][1] == Ny and L.red_rows[-1][1] == Ny - 1 and L.cols[-1][1] == Nx - 1
            assert all(a == b for a, b in zip(L.rows[:-1], L.red_rows[:-1]))


def _periodic_worker(rank, world, port, Ny, Nx, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rmt_oracle as O
        from pyrmt_b200.bc import periodic_bc
        from pyrmt_b200.slab import Comm, DistPoissonFFT, SlabLayout, local_bc_table
        lay = SlabLayout(Ny, Nx, world, rank, halo=4, periodic=True)
        comm = Comm()
        # the periodic BC: column wrap from the local table, row wrap (rank 0 -> last rank) by message
        rng = np.random.default_rng(11)
        u, v = rng.standard_normal((Ny, Nx)), rng.standard_normal((Ny, Nx))
        ru, rv = periodic_bc(u, v)
        table, rem = local_bc_table(periodic_bc, lay)
        su, sv = table.apply_host(lay.take(u).copy(), lay.take(v).copy())
        su, sv = torch.from_numpy(su.copy()), torch.from_numpy(sv.copy())
        assert (rem is not None) == (world > 1)
        if rem is not None:
            rem.apply_(su, sv, comm)
        ok_bc = bool(np.array_equal(lay.owned(su).numpy(), ru[lay.r0:lay.r1])
                     and np.array_equal(lay.owned(sv).numpy(), rv[lay.r0:lay.r1]))
        # ring exchange: global row ids travel with wrap-around
        n_red = lay.red_rows[rank][1] - lay.red_rows[rank][0]
        E = torch.full((n_red + 3, 3), -1.0, dtype=torch.float64)
        E[1:1 + n_red] = torch.arange(lay.red_rows[rank][0], lay.red_rows[rank][1], dtype=torch.float64)[:, None]
        comm.ring_exchange([(E[n_red:n_red + 1], E[1:3], E[0:1], E[1 + n_red:3 + n_red])])
        my = Ny - 1
        want = [(lay.red_rows[rank][0] - 1) % my] + list(range(lay.red_rows[rank][0], lay.red_rows[rank][1])) + \
               [lay.red_rows[rank][1] % my, (lay.red_rows[rank][1] + 1) % my]
        ok_ring = bool(torch.equal(E[:, 0], torch.tensor(want, dtype=torch.float64)))
        # distributed Hartley solve against numpy's fft2 through the oracle, table and separable set-up
        X, Y, dx, dy = O.create_grid(Nx, Ny, 1.3, 0.7)
        eig = O._precompute_poisson_eigenvalues_periodic(Nx, Ny, dx, dy)
        rhs = rng.standard_normal((Ny, Nx))
        ref = O._solve_poisson_fft(rhs, eig)
        errs = []
        for kw in (dict(eig=eig), dict(spacing=(dx, dy))):
            solver = DistPoissonFFT(lay, comm=comm, ops=SciPyOps(), **kw)
            r0, r1 = lay.red_rows[rank]
            sol = torch.zeros((r1 - r0, Nx), dtype=torch.float64)
            total = solver.solve(torch.from_numpy(np.ascontiguousarray(rhs[r0:r1])), sol)
            got = sol.numpy() - float(total) / (Ny * Nx)
            errs.append(float(np.max(np.abs(got - ref[r0:r1])) / np.max(np.abs(ref))))
        out[rank] = (ok_bc, ok_ring, max(errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Ny,Nx", [(2, 33, 17), (3, 40, 29), (1, 17, 33)])
def test_periodic_wrap_and_distributed_fft_gloo(world, Ny, Nx):
    ctx_ = mp.get_context("spawn")
    with ctx_.Manager() as m:
        out = m.dict()
        port = 31500 + (os.getpid() % 2000) + world
        procs = [ctx_.Process(target=_periodic_worker, args=(r, world, port, Ny, Nx, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        for r in range(world):
            ok_bc, ok_ring, err = out[r]
            assert ok_bc, "periodic BC wrong on rank %d" % r
            assert ok_ring, "ring exchange wrong on rank %d" % r
            assert err < 1e-12, (r, err)

"""Device plumbing shared by the operator modules: array marshalling, streams,
per-device workspaces, plan and table caches.  PyTorch is used only to own
device memory and streams; all arithmetic happens in librmt_b200.so.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib

F64 = torch.float64


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("pyrmt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


def is_np(x):
    return not isinstance(x, torch.Tensor)


def to_dev(x):
    """ndarray / tensor -> contiguous fp64 CUDA tensor (no copy when already one)."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.to(device())
        if x.dtype != F64:
            x = x.to(F64)
        return x.contiguous()
    arr = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(arr).to(device())


def to_user(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def ptr(t):
    return None if t is None else t.data_ptr()


def empty_like(t):
    return torch.empty_like(t)


def shape2(t):
    if t.dim() != 2:
        raise ValueError("expected a 2-D (Ny, Nx) field, got shape %s" % (tuple(t.shape),))
    return int(t.shape[0]), int(t.shape[1])


# Kernels launched by one call of each entry point (for bench.py's gpu_launches;
# rmt_extrapolate launches 1 + 4 per layer: its caller adds the 4 per layer).
_LAUNCHES = {"rmt_max_speed": 2, "rmt_field_stats": 2, "rmt_advect_euler_rk3": 5, "rmt_advect_euler_rk3_pair": 5, "rmt_extrapolate": 1, "rmt_extrapolate_rows": 1,
             "rmt_poisson_solve_dct": 4, "rmt_poisson_solve_fft": 16, "rmt_abi_version": 0, "rmt_launch_count": 0, "rmt_diagnostics_workspace_doubles": 0, "rmt_diagnostics": 2,
             "rmt_reduce_workspace_doubles": 0, "rmt_projection_partials": 0, "rmt_projection_correct_centered": 2, "rmt_extrapolate_workspace_bytes": 0, "rmt_extrapolate_set_mode": 0, "rmt_extrapolate_last_mode": 0,
             "rmt_peer_alloc": 0, "rmt_peer_free": 0, "rmt_peer_export": 0, "rmt_peer_import": 0, "rmt_peer_release": 0,
             "rmt_poisson_plan_create": 0, "rmt_poisson_plan_destroy": 0, "rmt_poisson_plan_is_fast": 0, "rmt_poisson_plan_invalidate": 0}


def check_separable_periodic_symbol(eig, null):
    """The fast periodic solve (csrc/fft.cu: separable Hartley pair, no mean removal before the transform)
    is the reference's fft2 -> /eig -> ifft2 (functions.py:1216-1233) only for a symbol that is even in each
    wavenumber separately and whose (0, 0) mode is in the null mask -- what
    _precompute_poisson_eigenvalues_periodic builds.  The reference accepts any (eig, null) pair, so a table
    that breaks either assumption is refused here instead of being solved differently (once per table)."""
    e = torch.as_tensor(eig)
    n = torch.as_tensor(null).to(torch.bool)
    fx = lambda t: torch.roll(torch.flip(t, dims=(1,)), 1, dims=1)       # k -> (m - k) mod m along x
    fy = lambda t: torch.roll(torch.flip(t, dims=(0,)), 1, dims=0)
    ok = bool(n[0, 0]) and all(bool(torch.equal(n, f(n))) for f in (fx, fy))
    if ok:
        m = ~n
        tol = 1e-12 * float(e[m].abs().max()) if bool(m.any()) else 0.0
        ok = all(float(((e - f(e)).abs() * m).max()) <= tol for f in (fx, fy))
    if not ok:
        raise ValueError("the fast periodic Poisson solve needs a symbol that is even in each wavenumber with the "
                         "(0, 0) mode in the null mask (as _precompute_poisson_eigenvalues_periodic builds it); "
                         "this (eig, null_mask) table is not")


class Profiler:
    """Optional per-entry-point CUDA-event timing and launch counting (bench.py)."""

    def __init__(self):
        self.timing = False
        self.launches = 0
        self.events = {}

    def reset(self, timing=False):
        self.timing = timing
        self.launches = 0
        self.events = {}

    def summary(self):
        """name -> (calls, total ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.events.items()}


profiler = Profiler()


class _LibProxy:
    """Forwards to the ctypes library, counting launches and (optionally) timing
    each entry point with CUDA events on the launching stream."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        n_launch = _LAUNCHES.get(name, 1)
        prof = profiler

        def call(*args):
            prof.launches += n_launch
            if prof.timing and n_launch:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*args)
                e1.record()
                prof.events.setdefault(name, []).append((e0, e1))
                return r
            return fn(*args)

        self.__dict__[name] = call
        return call


class _Ctx:
    """Per-device scratch: reduction workspace, Poisson plans, extrapolation
    workspace, device copies of eigenvalue tables."""

    def __init__(self, dev):
        self.dev = dev
        self.lib = _LibProxy(_lib.load())
        n = self.lib.rmt_reduce_workspace_doubles()
        self.red = torch.empty(n, dtype=F64, device=dev)
        self.plans = {}
        self.plan_tables = {}  # plan key -> (identity, device tensors, the caller's objects kept alive)
        self.extrap_ws = {}
        self.tables = {}      # id(ndarray) -> (weakref/obj, device tensor)
        self.finite_cache = None

    def plan(self, Ny, Nx, kind):
        key = (Ny, Nx, kind)
        p = self.plans.get(key)
        if p is None:
            h = C.c_void_p()
            _lib.check(self.lib.rmt_poisson_plan_create(Ny, Nx, kind, C.byref(h)),
                       "rmt_poisson_plan_create(%d,%d,%d)" % key)
            p = h
            self.plans[key] = p
        return p

    @staticmethod
    def _table_identity(t):
        """What makes two eigenvalue tables "the same table" for the plan's derived data: a raw pointer is
        not enough (allocators recycle addresses, ndarrays can be edited in place)."""
        if isinstance(t, torch.Tensor):
            return ("t", t.data_ptr(), t._version, tuple(t.shape), str(t.device), str(t.dtype))
        a = np.asarray(t)
        flat = a.reshape(-1)
        step = max(1, flat.size // 2048)
        sample = flat[::step]
        return ("a", id(t), a.shape, a.strides, str(a.dtype), float(np.sum(sample, dtype=np.float64)),
                float(flat[0]) if flat.size else 0.0, float(flat[-1]) if flat.size else 0.0)

    def plan_with_tables(self, Ny, Nx, kind, tables, dtypes):
        """The Poisson plan of (Ny, Nx, kind) bound to the caller's eigenvalue table(s): device copies are
        made once per table identity and kept alive with the plan (their addresses cannot be recycled while
        the plan's derived tables refer to them); a new identity invalidates the derived tables."""
        plan = self.plan(Ny, Nx, kind)
        key = (Ny, Nx, kind)
        ident = tuple(self._table_identity(t) for t in tables)
        ent = self.plan_tables.get(key)
        if ent is None or ent[0] != ident:
            devs = []
            for t, dt in zip(tables, dtypes):
                if isinstance(t, torch.Tensor):
                    d = t.to(self.dev)
                    d = (d if d.dtype == dt else d.to(dt)).contiguous()
                else:
                    npdt = np.float64 if dt == F64 else np.uint8
                    d = torch.from_numpy(np.ascontiguousarray(t, dtype=npdt)).to(self.dev)
                devs.append(d)
            _lib.check(self.lib.rmt_poisson_plan_invalidate(plan), "rmt_poisson_plan_invalidate")
            if kind == 1 and self.lib.rmt_poisson_plan_is_fast(plan) == 1:
                check_separable_periodic_symbol(devs[0], devs[1])
            ent = (ident, tuple(devs), tuple(tables))
            self.plan_tables[key] = ent
        return plan, ent[1]

    def scratch(self, name, ndoubles):
        """A named fp64 scratch buffer that is reused across calls (grown on demand)."""
        d = self.__dict__.setdefault("_scratch", {})
        t = d.get(name)
        if t is None or t.numel() < ndoubles:
            t = torch.empty(int(ndoubles), dtype=F64, device=self.dev)
            d[name] = t
        return t

    def extrap_workspace(self, Ny, Nx):
        key = (Ny, Nx)
        w = self.extrap_ws.get(key)
        if w is None:
            nbytes = self.lib.rmt_extrapolate_workspace_bytes(Ny, Nx)
            w = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)
            self.extrap_ws[key] = w
        return w

    def table(self, arr, dtype=F64):
        """Device copy of a host table (eigenvalues, masks), cached by identity."""
        if isinstance(arr, torch.Tensor):
            t = arr.to(self.dev)
            return (t if t.dtype == dtype else t.to(dtype)).contiguous()
        key = (id(arr), dtype)
        hit = self.tables.get(key)
        if hit is not None and hit[0] is arr:
            return hit[1]
        npdt = np.float64 if dtype == F64 else np.uint8
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=npdt)).to(self.dev)
        if len(self.tables) > 64:
            self.tables.clear()
        self.tables[key] = (arr, t)
        return t

    # --- reductions -------------------------------------------------------
    def stats(self, t):
        """[sum, min, max, #non-finite] of a device tensor, as a 4-vector on the device."""
        out = torch.empty(4, dtype=F64, device=self.dev)
        _lib.check(self.lib.rmt_field_stats(ptr(t), t.numel(), ptr(self.red), ptr(out), stream()),
                   "rmt_field_stats")
        return out

    def max_speed(self, a, b):
        out = torch.empty(2, dtype=F64, device=self.dev)
        _lib.check(self.lib.rmt_max_speed(ptr(a), ptr(b), a.numel(), ptr(self.red), ptr(out), stream()),
                   "rmt_max_speed")
        return out

    def max_speed_async(self, a, b):
        """max_speed with the two numbers on their way to pinned host memory: returns (host tensor, event);
        the caller queues whatever does not need them, then `event.synchronize()`."""
        out = self.max_speed(a, b)
        ring = self.__dict__.get("_pinned2")
        if ring is None:                         # a small ring of pinned landing buffers, made once
            ring = self.__dict__["_pinned2"] = [torch.empty(2, dtype=F64, pin_memory=True) for _ in range(8)]
        host = ring.pop(0)
        ring.append(host)
        host.copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev


_ctxs = {}


def ctx():
    dev = device()
    c = _ctxs.get(dev.index)
    if c is None:
        c = _Ctx(dev)
        _ctxs[dev.index] = c
    return c


class FiniteCache:
    """Remembers that a velocity pair was already checked finite (the guard of
    functions.py:524 runs twice per step on the same arrays, and
    compute_timestep reduces them just before)."""

    def __init__(self):
        self.ref = None

    def put(self, a, b, ok):
        # identity = the tensor objects, their torch versions AND their storage addresses; kernels that write
        # through raw pointers (the BC table) do not bump _version, so those call invalidate(), and a verdict
        # is only served for the guards of ONE step (two advect calls, or one pair call)
        self.ref = [weakref.ref(a), a._version, weakref.ref(b), b._version, ok, a.data_ptr(), b.data_ptr(), 2]

    def get(self, a, b):
        r = self.ref
        if r is None:
            return None
        if (r[0]() is a and r[2]() is b and r[1] == a._version and r[3] == b._version
                and r[5] == a.data_ptr() and r[6] == b.data_ptr() and r[7] > 0):
            r[7] -= 1
            return r[4]
        return None

    def invalidate(self):
        self.ref = None


finite_cache = FiniteCache()

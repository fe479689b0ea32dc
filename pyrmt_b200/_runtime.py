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
_LAUNCHES = {"rmt_max_speed": 2, "rmt_field_stats": 2, "rmt_advect_euler_rk3": 3, "rmt_advect_euler_rk3_pair": 3, "rmt_extrapolate": 1, "rmt_extrapolate_rows": 1,
             "rmt_poisson_solve_dct": 4, "rmt_poisson_solve_fft": 16, "rmt_abi_version": 0, "rmt_launch_count": 0, "rmt_diagnostics_workspace_doubles": 0, "rmt_diagnostics": 2,
             "rmt_reduce_workspace_doubles": 0, "rmt_extrapolate_workspace_bytes": 0, "rmt_extrapolate_set_mode": 0, "rmt_extrapolate_last_mode": 0,
             "rmt_poisson_plan_create": 0, "rmt_poisson_plan_destroy": 0, "rmt_poisson_plan_is_fast": 0}


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
        self.ref = (weakref.ref(a), a._version, weakref.ref(b), b._version, ok)

    def get(self, a, b):
        r = self.ref
        if r is None:
            return None
        if r[0]() is a and r[2]() is b and r[1] == a._version and r[3] == b._version:
            return r[4]
        return None


finite_cache = FiniteCache()

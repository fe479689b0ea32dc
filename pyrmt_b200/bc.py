"""Velocity boundary-condition callables on the device.

The reference invokes an arbitrary Python callable ``bc(u, v) -> (u, v)`` on
every RK4 stage, on the predictor result and on the projected velocity
(pyRMT/functions.py:714,760,1287,1355).  Every BC in
scope (benchmarks/common.py:27-50, tests/test_poisson.py:39-64,
tests/test_contact.py:58) only rewrites rim cells with a constant, a copy of
another cell, or a sign-flipped copy.  Such a callable is classified ONCE per
grid shape by probing it on the host with sentinel arrays, turned into a gather
table, verified against the callable on random data, and from then on applied
in place on the device by one small kernel (rmt_apply_bc).  A callable that
cannot be expressed that way is still honoured exactly, through a device->host->
callable->device round trip per application (slow path).
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from . import _lib
from ._runtime import F64, ctx, device, ptr, stream

_FIELD_BIT = 1 << 62


class BCTable:
    """out_field[cell] = ca * in_field[src_cell] + cb for the listed cells."""

    def __init__(self, dst, src, ca, cb, shape):
        self.shape = shape
        self.n = int(dst.size)
        self.host = (dst, src, ca, cb)
        self._dev = {}

    def _on(self, dev):
        t = self._dev.get(dev.index)
        if t is None:
            dst, src, ca, cb = self.host
            t = (torch.from_numpy(dst).to(dev), torch.from_numpy(src).to(dev),
                 torch.from_numpy(ca).to(dev), torch.from_numpy(cb).to(dev))
            self._dev[dev.index] = t
        return t

    def apply_(self, u, v):
        """In place on two private device tensors."""
        if self.n == 0:
            return
        dst, src, ca, cb = self._on(u.device)
        _lib.check(ctx().lib.rmt_apply_bc(ptr(u), ptr(v), ptr(dst), ptr(src), ptr(ca), ptr(cb),
                                          self.n, stream()), "rmt_apply_bc")

    def apply_host(self, u, v):
        """NumPy emulation of the table (used to verify the classification)."""
        dst, src, ca, cb = self.host
        n = u.size
        flat = lambda k: np.where(k & _FIELD_BIT, (k & (_FIELD_BIT - 1)) + n, k)
        fin = np.concatenate([u.ravel(), v.ravel()])
        out = fin.copy()
        vals = cb.copy()
        has = src >= 0
        vals[has] = vals[has] + ca[has] * fin[flat(src[has])]
        out[flat(dst)] = vals
        return out[:n].reshape(u.shape), out[n:].reshape(v.shape)


def _probe(bc, base, Ny, Nx):
    n = Ny * Nx
    u = (base + np.arange(n, dtype=np.float64)).reshape(Ny, Nx)
    v = (base + n + np.arange(n, dtype=np.float64)).reshape(Ny, Nx)
    ou, ov = bc(u.copy(), v.copy())
    ou = np.asarray(ou, dtype=np.float64)
    ov = np.asarray(ov, dtype=np.float64)
    if ou.shape != (Ny, Nx) or ov.shape != (Ny, Nx):
        raise ValueError("BC callable changed the field shape")
    return np.concatenate([ou.ravel(), ov.ravel()])


def _decode(vals, base, n2):
    """sentinel value -> (source index in [0, 2n), sign) or -1."""
    src = np.full(vals.shape, -1, dtype=np.int64)
    sign = np.ones(vals.shape)
    for sg in (1.0, -1.0):
        k = sg * vals - base
        ok = np.isfinite(k) & (k >= 0) & (k < n2) & (k == np.floor(k)) & (src < 0)
        src[ok] = k[ok].astype(np.int64)
        sign[ok] = sg
    return src, sign


def classify(bc, Ny, Nx):
    """Return a BCTable for ``bc`` on an (Ny, Nx) grid, or None if the callable
    is not a rim gather (then the slow path is used)."""
    n = Ny * Nx
    try:
        b1, b2 = 1048576.5, 7340032.25          # exactly representable, non-integer sentinels
        o1 = _probe(bc, b1, Ny, Nx)
        o2 = _probe(bc, b2, Ny, Nx)
    except Exception:
        return None
    s1, g1 = _decode(o1, b1, 2 * n)
    s2, g2 = _decode(o2, b2, 2 * n)
    ident = np.arange(2 * n)
    copy = (s1 >= 0) & (s1 == s2) & (g1 == g2)
    const = (s1 < 0) & (s2 < 0) & (o1 == o2)
    if not np.all(copy | const):
        return None
    changed = ~(copy & (s1 == ident) & (g1 > 0))
    idx = np.nonzero(changed)[0]
    pack = lambda k: np.where(k >= n, (k - n) | _FIELD_BIT, k).astype(np.int64)
    dst = pack(idx)
    is_copy = copy[idx]
    src = np.where(is_copy, pack(np.where(is_copy, s1[idx], 0)), -1).astype(np.int64)
    ca = np.where(is_copy, g1[idx], 0.0).astype(np.float64)
    cb = np.where(is_copy, 0.0, o1[idx]).astype(np.float64)
    # in-place application needs sources that are not themselves rewritten
    if np.intersect1d(dst, src[src >= 0]).size:
        return None
    table = BCTable(dst, src, ca, cb, (Ny, Nx))
    # verify on random data (exact equality: copies and constants only)
    rng = np.random.default_rng(12345)
    u = rng.standard_normal((Ny, Nx))
    v = rng.standard_normal((Ny, Nx))
    try:
        ru, rv = bc(u.copy(), v.copy())
    except Exception:
        return None
    tu, tv = table.apply_host(u, v)
    if not (np.array_equal(tu, ru) and np.array_equal(tv, rv)):
        return None
    return table


_cache = collections.OrderedDict()      # LRU, bounded: key -> (keep-alive objects, table)
_CACHE_MAX = 64


def _callable_key(bc):
    """What identifies a BC callable for the table cache.

    The reference calls ``bc(u, v)`` on every stage; here the callable is probed ONCE per grid and replayed as
    a gather table, so it must be a PURE function of (u, v): the same rim copies and constants on every call.
    Two things follow.  (1) A plain function or lambda is keyed by its code object and the VALUES of its
    closure cells and defaults, not by ``id``: a lambda re-created inside the time loop hits the cache
    instead of being re-probed every step, and a closure whose captured lid speed has changed gets a new
    table.  (2) State the key cannot see (globals, attributes mutated in place, time read from a clock) needs
    the opt-out ``bc.rmt_dynamic = True``, which sends every application down the host round-trip path.
    Returns (key, keep_alive) or (None, None) for the dynamic opt-out."""
    if getattr(bc, "rmt_dynamic", False):
        return None, None
    if hasattr(bc, "__self__"):                      # bound method: the object it is bound to is the identity
        return ("self", id(bc.__self__), getattr(bc.__func__, "__code__", None)), (bc.__self__,)
    code = getattr(bc, "__code__", None)
    if code is not None and not hasattr(bc, "rmt_table"):
        try:
            vals = tuple(c.cell_contents for c in (bc.__closure__ or ())) + tuple(bc.__defaults__ or ())
            import types
            const = (int, float, bool, str, type(None), types.ModuleType, types.FunctionType,
                     types.BuiltinFunctionType, type)        # values, or objects that stand for themselves
            if all(isinstance(v, const) for v in vals):
                return ("code", code, tuple(v if isinstance(v, (int, float, bool, str, type(None))) else id(v)
                                            for v in vals)), (code, vals)
        except ValueError:                           # an empty closure cell
            pass
    return ("id", id(bc)), (bc,)


def table_for(bc, Ny, Nx):
    """The gather table of ``bc`` on an (Ny, Nx) grid, or None (slow host path).  See `_callable_key` for
    the purity requirement and the ``rmt_dynamic`` opt-out."""
    ck, keep = _callable_key(bc)
    if ck is None:
        return None
    key = (ck, Ny, Nx)
    hit = _cache.get(key)
    if hit is not None:
        _cache.move_to_end(key)
        return hit[1]
    t = getattr(bc, "rmt_table", None)
    table = t(Ny, Nx) if callable(t) else classify(bc, Ny, Nx)
    _cache[key] = (keep, table)
    while len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return table


def apply_bc_(bc, u, v):
    """Apply ``bc`` to two PRIVATE device tensors; returns the (possibly new) pair."""
    Ny, Nx = u.shape
    table = table_for(bc, Ny, Nx)
    if table is not None:
        from ._runtime import finite_cache
        finite_cache.invalidate()        # in-place write through raw pointers: torch's _version does not see it
        table.apply_(u, v)
        return u, v
    # slow path: honour the arbitrary callable on the host
    ru, rv = bc(u.cpu().numpy(), v.cpu().numpy())
    dev = u.device
    return (torch.from_numpy(np.ascontiguousarray(ru, dtype=np.float64)).to(dev),
            torch.from_numpy(np.ascontiguousarray(rv, dtype=np.float64)).to(dev))


# ---- ready-made callables (work on ndarrays like the reference's; carry no state
# beyond their parameters, so the classifier handles them like any other) -------
def no_slip_lid_bc(u, v, lid_speed=1.0):
    """benchmarks/common.py:27-37 (walls at rest, lid moving in +x, corners zero)."""
    u = u.copy()
    v = v.copy()
    for f in (u, v):
        f[:, 0] = 0.0
        f[:, -1] = 0.0
        f[0, :] = 0.0
    u[-1, :] = lid_speed
    v[-1, :] = 0.0
    for f in (u, v):
        f[0, 0] = f[0, -1] = f[-1, 0] = f[-1, -1] = 0.0
    return u, v


def free_slip_box_bc(u, v):
    """benchmarks/common.py:40-50."""
    u = u.copy()
    v = v.copy()
    u[:, 0] = 0.0
    u[:, -1] = 0.0
    v[:, 0] = v[:, 1]
    v[:, -1] = v[:, -2]
    v[0, :] = 0.0
    v[-1, :] = 0.0
    u[0, :] = u[1, :]
    u[-1, :] = u[-2, :]
    return u, v


def periodic_bc(u, v):
    """tests/test_poisson.py:60-64."""
    u = u.copy()
    v = v.copy()
    u[:, -1] = u[:, 0]
    v[:, -1] = v[:, 0]
    u[-1, :] = u[0, :]
    v[-1, :] = v[0, :]
    return u, v


def wall_bc(u, v):
    """tests/test_poisson.py:39-43."""
    u = u.copy()
    v = v.copy()
    for f in (u, v):
        f[:, 0] = f[:, -1] = f[0, :] = f[-1, :] = 0.0
    return u, v

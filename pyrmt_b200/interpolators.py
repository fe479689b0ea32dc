"""Drop-in for upstream ``pyRMT/interpolators.py`` -- sampling on the device."""
from __future__ import annotations

import torch

from . import _lib
from ._runtime import ctx, is_np, ptr, stream, to_dev, to_user


def _sample(u, xq, yq, dx, dy, Nx, Ny, cubic):
    as_np = is_np(u)
    ud, xd, yd = to_dev(u), to_dev(xq), to_dev(yq)
    if tuple(ud.shape) != (int(Ny), int(Nx)):
        raise ValueError("u has shape %s, expected (Ny, Nx) = (%d, %d)" % (tuple(ud.shape), Ny, Nx))
    if xd.shape != yd.shape:
        raise ValueError("xq and yq must have the same shape")
    out = torch.empty_like(xd)
    _lib.check(ctx().lib.rmt_sample(ptr(ud), ptr(xd), ptr(yd), ptr(out), xd.numel(), float(dx), float(dy),
                                    int(Nx), int(Ny), cubic, stream()), "rmt_sample")
    return to_user(out, as_np)


def bilinear_interpolate(u, xq, yq, dx, dy, Nx, Ny):
    """pyRMT/interpolators.py:4-62 -- clamped bilinear sampling; non-finite query
    coordinates give NaN (same guards as upstream)."""
    return _sample(u, xq, yq, dx, dy, Nx, Ny, 0)


def bicubic_interpolate(u, xq, yq, dx, dy, Nx, Ny):
    """pyRMT/interpolators.py:64-141 -- monotone (clamped) Catmull-Rom sampling."""
    return _sample(u, xq, yq, dx, dy, Nx, Ny, 1)


def cubic_convolution(v0, v1, v2, v3, x):
    """pyRMT/interpolators.py:143-154 -- scalar Catmull-Rom segment (host helper;
    the device kernels carry their own copy)."""
    a0 = -0.5 * v0 + 1.5 * v1 - 1.5 * v2 + 0.5 * v3
    a1 = v0 - 2.5 * v1 + 2.0 * v2 - 0.5 * v3
    a2 = -0.5 * v0 + 0.5 * v2
    a3 = v1
    return a0 * (x * x * x) + a1 * (x * x) + a2 * x + a3

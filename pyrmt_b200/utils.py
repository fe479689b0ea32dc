"""Drop-in for upstream ``pyRMT/utils.py`` -- finite-difference stencils on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._runtime import ctx, is_np, ptr, shape2, stream, to_dev, to_user


def _stencil(f, hx, hy, op):
    as_np = is_np(f)
    fd = to_dev(f)
    Ny, Nx = shape2(fd)
    out = torch.empty_like(fd)
    _lib.check(ctx().lib.rmt_stencil_op(ptr(fd), ptr(out), Ny, Nx, float(hx), float(hy), op, stream()),
               "rmt_stencil_op")
    return to_user(out, as_np)


def grad_central_x_2nd(f, dx):
    """pyRMT/utils.py:4-14."""
    return _stencil(f, dx, 1.0, 0)


def grad_central_y_2nd(f, dy):
    """pyRMT/utils.py:16-25."""
    return _stencil(f, 1.0, dy, 1)


def grad_central_x_4th(f, dx):
    """pyRMT/utils.py:27-42."""
    return _stencil(f, dx, 1.0, 2)


def grad_central_y_4th(f, dy):
    """pyRMT/utils.py:44-59."""
    return _stencil(f, 1.0, dy, 3)


def lap_2nd(f, dx, dy):
    """pyRMT/utils.py:116-131."""
    return _stencil(f, dx, dy, 4)


def diff_upwind_3rd(f, u, h, axis):
    """pyRMT/utils.py:61-114 (axis 1 = x, axis 0 = y)."""
    as_np = is_np(f)
    fd, ud = to_dev(f), to_dev(u)
    Ny, Nx = shape2(fd)
    out = torch.empty_like(fd)
    _lib.check(ctx().lib.rmt_diff_upwind_3rd(ptr(fd), ptr(ud), ptr(out), Ny, Nx, float(h), int(axis),
                                             stream()), "rmt_diff_upwind_3rd")
    return to_user(out, as_np)


def fast_solve_3x3(A, b):
    """pyRMT/utils.py:134-167 -- Cramer's rule for one 3x3 system (zeros when
    |det| < 1e-15).  A scalar host helper upstream too; the extrapolation kernel
    carries the same expression trees on the device."""
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    detA = (A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1])
            - A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0])
            + A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
    x = np.zeros(3)
    if abs(detA) < 1e-15:
        return x
    inv = 1.0 / detA
    x[0] = (b[0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1])
            - A[0, 1] * (b[1] * A[2, 2] - A[1, 2] * b[2])
            + A[0, 2] * (b[1] * A[2, 1] - A[1, 1] * b[2])) * inv
    x[1] = (A[0, 0] * (b[1] * A[2, 2] - A[1, 2] * b[2])
            - b[0] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0])
            + A[0, 2] * (A[1, 0] * b[2] - b[1] * A[2, 0])) * inv
    x[2] = (A[0, 0] * (A[1, 1] * b[2] - b[1] * A[2, 1])
            - A[0, 1] * (A[1, 0] * b[2] - b[1] * A[2, 0])
            + b[0] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0])) * inv
    return x

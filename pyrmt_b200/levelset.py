"""Level-set callables for ``rebuild_phi_from_reference_map``.

The reference hands an arbitrary Python callable phi0(X1, X2) to
``rebuild_phi_from_reference_map`` (pyRMT/functions.py:1366-1367); every
in-scope driver uses the disc signed distance of benchmarks/common.py:55-57.
``DiscSDF`` is that family as a descriptor object: called with ndarrays or CUDA
tensors it evaluates  phi0(xi) = min_k(|xi - c_k| - R_k)  on the device
(rmt_disc_sdf), one disc reproducing ``initialize_disc`` exactly
(sqrt(ex*ex + ey*ey) - R, no contraction across the sqrt).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._runtime import ctx, is_np, ptr, stream, to_dev, to_user


class DiscSDF:
    """phi0(xi) = min over discs of (|xi - c_k| - R_k).

    centres_x / centres_y / radii : scalars or 1-D sequences (reference space).
    domain : (Lx, Ly) of reference space; with more than 8 discs a uniform bin
        grid over the domain restricts each point to the discs that can attain
        the minimum (the result is identical to the exhaustive minimum).
    """

    def __init__(self, centres_x, centres_y, radii, domain=None):
        self.cx = np.atleast_1d(np.asarray(centres_x, dtype=np.float64)).copy()
        self.cy = np.atleast_1d(np.asarray(centres_y, dtype=np.float64)).copy()
        r = np.atleast_1d(np.asarray(radii, dtype=np.float64))
        self.R = np.broadcast_to(r, self.cx.shape).copy()
        if not (self.cx.shape == self.cy.shape == self.R.shape) or self.cx.ndim != 1:
            raise ValueError("centres_x, centres_y and radii must be 1-D and of equal length")
        self.domain = domain
        self._bins = None
        self._dev = {}
        if domain is not None and self.cx.size > 8:
            self._build_bins()

    # -- host: conservative candidate lists per bin ----------------------------
    def _build_bins(self):
        Lx, Ly = map(float, self.domain)
        nd = self.cx.size
        gb = int(min(256, max(2, round(8.0 * np.sqrt(nd)))))
        bw_x, bw_y = Lx / gb, Ly / gb
        diag = np.hypot(bw_x, bw_y)
        start = np.zeros(gb * gb + 1, dtype=np.int32)
        cand = []
        for by in range(gb):
            ymid = (by + 0.5) * bw_y
            for bx in range(gb):
                xmid = (bx + 0.5) * bw_x
                d = np.hypot(self.cx - xmid, self.cy - ymid) - self.R     # value at the bin centre
                # |phi_k(x) - phi_k(mid)| <= |x - mid| <= diag/2 inside the bin, for every disc
                keep = np.nonzero(d <= d.min() + diag * (1.0 + 1e-9) + 1e-300)[0]
                cand.append(keep.astype(np.int32))
                start[by * gb + bx + 1] = start[by * gb + bx] + keep.size
        self._bins = (gb, start, np.concatenate(cand) if cand else np.zeros(0, np.int32), Lx, Ly)

    def _on(self, dev):
        t = self._dev.get(dev.index)
        if t is None:
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            t = {"cx": up(self.cx), "cy": up(self.cy), "R": up(self.R)}
            if self._bins is not None:
                t["start"] = up(self._bins[1])
                t["cand"] = up(self._bins[2])
            self._dev[dev.index] = t
        return t

    def __call__(self, X1, X2):
        as_np = is_np(X1)
        x1, x2 = to_dev(X1), to_dev(X2)
        if x1.shape != x2.shape:
            raise ValueError("X1 and X2 must have the same shape")
        t = self._on(x1.device)
        phi = torch.empty_like(x1)
        if self._bins is not None:
            gb, _, _, Lx, Ly = self._bins
            st, cd = ptr(t["start"]), ptr(t["cand"])
        else:
            gb, Lx, Ly, st, cd = 0, 1.0, 1.0, None, None
        _lib.check(ctx().lib.rmt_disc_sdf(ptr(x1), ptr(x2), ptr(phi), x1.numel(), ptr(t["cx"]),
                                          ptr(t["cy"]), ptr(t["R"]), int(self.cx.size), st, cd, gb,
                                          Lx, Ly, stream()), "rmt_disc_sdf")
        return to_user(phi, as_np)


    def with_stress(self, X1, X2, dx, dy, mu_s, kappa, w_cut, detg_clamp, isochoric=0):
        """phi0(X1, X2) and the solid Cauchy stress of (X1, X2, phi) in one pass (rmt_disc_sdf_stress):
        -> (phi, sxx, sxy, syy, J), all CUDA tensors.  phi is bitwise what ``__call__`` returns."""
        x1, x2 = to_dev(X1), to_dev(X2)
        if x1.shape != x2.shape or x1.dim() != 2:
            raise ValueError("X1 and X2 must be 2-D and of the same shape")
        Ny, Nx = int(x1.shape[0]), int(x1.shape[1])
        t = self._on(x1.device)
        phi, sxx, sxy, syy, J = (torch.empty_like(x1) for _ in range(5))
        if self._bins is not None:
            gb, _, _, Lx, Ly = self._bins
            st, cd = ptr(t["start"]), ptr(t["cand"])
        else:
            gb, Lx, Ly, st, cd = 0, 1.0, 1.0, None, None
        _lib.check(ctx().lib.rmt_disc_sdf_stress(ptr(x1), ptr(x2), ptr(phi), ptr(sxx), ptr(sxy), ptr(syy), ptr(J),
                                                 Ny, Nx, float(dx), float(dy), float(mu_s), float(kappa),
                                                 float(w_cut), float(detg_clamp), int(isochoric), ptr(t["cx"]),
                                                 ptr(t["cy"]), ptr(t["R"]), int(self.cx.size), st, cd, gb, Lx, Ly,
                                                 stream()), "rmt_disc_sdf_stress")
        return phi, sxx, sxy, syy, J


def initialize_disc(X, Y, x0, y0, R):
    """benchmarks/common.py:55-57 on the device."""
    return DiscSDF([x0], [y0], [R])(X, Y)

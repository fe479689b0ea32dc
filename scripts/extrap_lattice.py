"""Extrapolation time on the bench's jittered 8x8 lattice (driver.disc_lattice) at 4097^2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from extrap_probe import case
from pyrmt_b200.driver import disc_lattice
cx, cy, R = disc_lattice(8, 1.0, 0.04)
print("jittered 8x8: %.3f ms" % case(4097, list(zip(cx, cy)), 0.04 * 4096))
import numpy as np
rows = np.sort(np.round(cy * 4096).astype(int).reshape(8, 8), axis=1)
print("centre rows per lattice row (min..max):", [(int(r.min()), int(r.max())) for r in np.round(cy * 4096).astype(int).reshape(8, 8)])
print("centre cols (min..max) per lattice col:", [(int(c.min()), int(c.max())) for c in np.round(cx * 4096).astype(int).reshape(8, 8).T])

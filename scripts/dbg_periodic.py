import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrmt_b200 import functions as F
for N in (int(a) for a in sys.argv[1:]):
    L = (N - 1) / 128.0
    dx = L / (N - 1)
    eig = F._precompute_poisson_eigenvalues_periodic(N, N, dx, dx)
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal((N, N))
    t = time.time()
    r = rhs[:-1, :-1] - rhs[:-1, :-1].mean()
    spec = np.fft.fft2(r) / eig[0]
    spec[eig[1]] = 0
    ref = np.fft.ifft2(spec).real
    ref = F._tile_overlap(ref, N, N)
    ref -= ref.mean()
    print(N, "numpy", time.time() - t, flush=True)
    got = F._solve_poisson_fft(rhs, eig)
    print(N, "finite", np.isfinite(got).all(), "rel err", np.abs(got - ref).max() / np.abs(ref).max(), flush=True)

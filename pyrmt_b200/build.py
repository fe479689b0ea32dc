"""Build librmt_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pyrmt_b200.build [--force]

Every translation unit is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo``.  The units on the
reference-map path (``extrap.cu``, ``advect.cu``, ``stencil_ops.cu``) are built
with ``-fmad=false``: the narrow-band extrapolation amplifies 1-ulp differences
by up to 1e6, so it and everything feeding it (advected xi, the level set, the
solid mask) reproduce the reference's IEEE operation order bit for bit (SURVEY
Appendix A, H2); explicit ``fma()`` calls are unaffected by that flag.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "librmt_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE]
UNITS = {
    "stencil_ops.cu": ["-fmad=false"],   # disc SDF / mask feed the bit-exact xi path
    "advect.cu": ["-fmad=false"],        # advected xi must equal the reference bit for bit
    "momentum.cu": [],
    "projection.cu": [],
    "fft.cu": [],
    "extrap.cu": ["-fmad=false"],
    "peer.cu": [],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps(src):
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "rmt_b200.h"), __file__]
    if src.endswith("extrap.cu"):
        deps.append(os.path.join(CSRC, "exp_table.inc"))
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, _deps(src)):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(7, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

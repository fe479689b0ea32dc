"""Stage the UNMODIFIED reference package under oracle/_ref so that `bench.py --impl reference` can time
the reference's own Numba path on the GPU box's host cores (where /root/reference does not exist).

    python -m oracle.make_ref [--ref /root/reference]

This is test / measurement infrastructure, not product code: oracle/_ref/ is git-ignored (the reference's
sources never enter this repository's history) but travels with the work tree like the built .so files.
What is staged: the `pyRMT` package directory exactly as upstream ships it, two EMPTY stub packages for
the imports the hot path never reaches (`pyamg`: AMG fallback, `h5py`: snapshot writer), and a
manifest with the sha256 of every staged file.  Nothing is edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make(ref: str = "/root/reference", force: bool = False):
    src = os.path.join(ref, "pyRMT")
    if not os.path.isdir(src):
        return None                                   # not in the build container: keep whatever is staged
    man = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(man) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src, os.path.join(DST, "pyRMT"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.nbi", "*.nbc"))
    for stub in ("pyamg", "h5py"):
        os.makedirs(os.path.join(DST, "_stubs", stub))
        with open(os.path.join(DST, "_stubs", stub, "__init__.py"), "w") as f:
            f.write("# empty stub: the collocated hot path never reaches this import (SURVEY 8c)\n")
    files = {}
    for root, _, names in os.walk(os.path.join(DST, "pyRMT")):
        for n in sorted(names):
            p = os.path.join(root, n)
            with open(p, "rb") as f:
                files[os.path.relpath(p, DST)] = hashlib.sha256(f.read()).hexdigest()
    with open(man, "w") as f:
        json.dump({"source": src, "files": files}, f, indent=1)
    return DST


def load():
    """Import the staged reference (pyRMT.functions) or return None.  Needs numba; the JIT cache goes to
    a scratch directory (the staged tree stays untouched)."""
    if not os.path.isdir(os.path.join(DST, "pyRMT")):
        return None
    try:
        import numba  # noqa: F401
    except Exception:
        return None
    import tempfile
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    for p in (os.path.join(DST, "_stubs"), DST):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import pyRMT.functions as RF
        return RF
    except Exception:
        return None


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--ref") + 1] if "--ref" in sys.argv else "/root/reference"
    print(make(ref, force="--force" in sys.argv))

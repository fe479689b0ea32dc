"""Run the reference's own scripts against the B200 path, unchanged and uncopied.

The reference's drivers do ``sys.path.insert(0, <repo root>)`` and then
``from pyRMT.functions import ...`` (benchmarks/lid_driven_cavity.py:17-23).
``sys.modules`` wins over ``sys.path``, so registering this package under the
names ``pyRMT``, ``pyRMT.functions``, ``pyRMT.interpolators``, ``pyRMT.utils`` and
``pyRMT.output`` before the script is loaded makes every operator call land in
librmt_b200.so (ndarray in -> ndarray out).

    import pyrmt_b200.compat as compat
    compat.install()
    compat.run_script("/path/to/pyRMT/benchmarks/lid_driven_cavity.py", ["100", "129"])
"""
from __future__ import annotations

import runpy
import sys
import types


def install():
    """Register pyrmt_b200 as ``pyRMT`` in sys.modules (idempotent)."""
    from . import functions, interpolators, output, utils
    import pyrmt_b200 as pkg
    top = types.ModuleType("pyRMT")
    top.__doc__ = "pyRMT operator API served by pyrmt_b200 (B200 / sm_100a)"
    top.__path__ = []          # a package, with no files of its own
    for name in dir(pkg):
        if not name.startswith("__"):
            setattr(top, name, getattr(pkg, name))
    for name in ("compute_kinetic_energy", "compute_strain_energy", "compute_viscous_dissipation",
                 "output_simulation_data"):
        setattr(top, name, getattr(output, name))
    top.functions, top.interpolators, top.utils, top.output = functions, interpolators, utils, output
    sys.modules["pyRMT"] = top
    sys.modules["pyRMT.functions"] = functions
    sys.modules["pyRMT.interpolators"] = interpolators
    sys.modules["pyRMT.utils"] = utils
    sys.modules["pyRMT.output"] = output
    return top


def uninstall():
    for name in ("pyRMT", "pyRMT.functions", "pyRMT.interpolators", "pyRMT.utils", "pyRMT.output"):
        sys.modules.pop(name, None)


def run_script(path, argv=()):
    """runpy a reference driver with the shim installed."""
    install()
    old = sys.argv
    sys.argv = [path, *map(str, argv)]
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old

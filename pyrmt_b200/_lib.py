"""ctypes binding of librmt_b200.so (the C ABI declared in include/rmt_b200.h).

The library is the product: if it is missing or a CUDA device is absent the
operators raise -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librmt_b200.so")

_lock = threading.Lock()
_lib = None

vp, dbl, i32, i64 = C.c_void_p, C.c_double, C.c_int, C.c_long

# name -> argtypes   (restype is int unless listed in _RESTYPE)
_SIG = {
    "rmt_abi_version": [],
    "rmt_launch_count": [],
    "rmt_stencil_op": [vp, vp, i32, i32, dbl, dbl, i32, vp],
    "rmt_diff_upwind_3rd": [vp, vp, vp, i32, i32, dbl, i32, vp],
    "rmt_heaviside": [vp, vp, i64, dbl, vp],
    "rmt_heaviside_rho": [vp, vp, vp, i64, dbl, dbl, dbl, vp],
    "rmt_reinit_sign": [vp, vp, i64, dbl, vp],
    "rmt_reinit_step": [vp, vp, vp, i32, i32, dbl, dbl, dbl, vp],
    "rmt_mask_mul": [vp, vp, vp, i64, vp],
    "rmt_reduce_workspace_doubles": [],
    "rmt_max_speed": [vp, vp, i64, vp, vp, vp],
    "rmt_field_stats": [vp, i64, vp, vp, vp],
    "rmt_apply_bc": [vp, vp, vp, vp, vp, vp, i32, vp],
    "rmt_disc_sdf": [vp, vp, vp, i64, vp, vp, vp, i32, vp, vp, i32, dbl, dbl, vp],
    "rmt_disc_sdf_stress": [vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, dbl, dbl, i32, vp, vp, vp, i32, vp, vp,
                            i32, dbl, dbl, vp],
    "rmt_sample": [vp, vp, vp, vp, i64, dbl, dbl, i32, i32, i32, vp],
    "rmt_advect_sl_rk4": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_advect_sl_rk4_rows": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_advect_euler_rk3": [vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, i32, vp],
    "rmt_advect_euler_rk3_pair": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, i32, i32, vp],
    "rmt_euler_rhs": [vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_extrapolate_workspace_bytes": [i32, i32],
    "rmt_extrapolate": [vp, vp, vp, vp, vp, i32, i32, dbl, dbl, i32, vp, vp],
    "rmt_extrapolate_rows": [vp, vp, vp, vp, vp, i32, i32, i32, dbl, dbl, i32, vp, vp],
    "rmt_extrapolate_set_mode": [i32, i32, i64],
    "rmt_extrapolate_last_mode": [vp, i32, i32, C.POINTER(i32), vp],
    "rmt_exp_probe": [vp, vp, i64, vp],
    "rmt_weno_div_probe": [vp, vp, vp, i64, i32, vp],
    "rmt_solid_stress": [vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, dbl, dbl, i32, vp],
    "rmt_curvature": [vp, vp, i32, i32, dbl, dbl, vp],
    "rmt_surface_tension": [vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, vp],
    "rmt_velocity_rhs": [vp] * 12 + [i32, i32, dbl, dbl, dbl, vp],
    "rmt_momentum_stage": [vp] * 15 + [i32, i32] + [dbl] * 8 + [i32, vp],
    "rmt_momentum_stage_2solids": [vp] * 19 + [i32, i32] + [dbl] * 7 + [i32, vp],
    "rmt_contact_force": [vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, vp],
    "rmt_min2": [vp, vp, vp, i64, vp],
    "rmt_diagnostics_workspace_doubles": [],
    "rmt_diagnostics": [vp] * 7 + [i32, i32] + [dbl] * 9 + [vp, vp, vp],
    "rmt_divergence": [vp, vp, vp, i32, i32, dbl, dbl, vp],
    "rmt_divergence_rc": [vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, vp],
    "rmt_pressure_gradient": [vp, vp, vp, i32, i32, dbl, dbl, vp],
    "rmt_divergence_periodic": [vp, vp, vp, i32, i32, dbl, dbl, vp],
    "rmt_pressure_gradient_periodic": [vp, vp, vp, i32, i32, dbl, dbl, vp],
    "rmt_projection_rhs": [vp, vp, vp, vp, dbl, vp, vp, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_projection_correct": [vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_subtract_mean": [vp, vp, i64, vp],
    "rmt_projection_partials": [i32, i32],
    "rmt_projection_correct_centered": [vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, i32, vp],
    "rmt_poisson_plan_create": [i32, i32, i32, C.POINTER(vp)],
    "rmt_poisson_plan_destroy": [vp],
    "rmt_poisson_plan_is_fast": [vp],
    "rmt_poisson_plan_invalidate": [vp],
    "rmt_poisson_solve_dct": [vp, vp, vp, vp, vp, vp],
    "rmt_poisson_solve_fft": [vp, vp, vp, vp, vp, vp, vp],
    "rmt_dct_lines": [vp, vp, vp, i32, i32, dbl, vp],
    "rmt_dht_lines": [vp, vp, vp, i32, i32, i64, i64, dbl, vp],
    "rmt_transpose": [vp, vp, i32, i32, vp],
    "rmt_copy2d": [vp, vp, i32, i32, i64, i64, vp],
    "rmt_peer_alloc": [C.c_size_t, C.POINTER(vp)],
    "rmt_peer_free": [vp],
    "rmt_peer_export": [vp, vp],
    "rmt_peer_import": [vp, C.POINTER(vp)],
    "rmt_peer_release": [vp],
    "rmt_peer_put2d": [vp, i32, vp],
    "rmt_transpose_scatter": [vp, i32, i32, i64, i32, vp, vp, vp, vp],
    "rmt_peer_barrier": [vp, i32, i32, vp, dbl, vp, vp],
    "rmt_peer_reduce": [vp, i32, i64, i32, i32, vp, vp],
}
_RESTYPE = {"rmt_launch_count": C.c_ulonglong, "rmt_extrapolate_workspace_bytes": i64, "rmt_projection_partials": i64, "rmt_poisson_plan_destroy": None}

EXPORTS = tuple(_SIG)


class RmtError(RuntimeError):
    pass


def load():
    """Return the loaded library (ctypes.CDLL) with typed entry points."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    "pyrmt_b200: %s is missing -- build it with `python -m pyrmt_b200.build` "
                    "(there is no CPU fallback)" % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, args in _SIG.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = _RESTYPE.get(name, i32)
            _lib = lib
    return _lib


def check(code, what):
    if code == 0:
        return
    if code == -1:
        raise RmtError("%s: invalid argument" % what)
    if code == -2:
        raise NotImplementedError("%s: unsupported size for this build" % what)
    raise RmtError("%s: CUDA error %d" % (what, code))

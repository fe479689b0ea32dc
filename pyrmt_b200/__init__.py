"""pyrmt_b200 -- the collocated Reference-Map-Technique timestep of pyRMT,
rebuilt for NVIDIA B200 (sm_100a).

Same operator names and signatures as upstream ``pyRMT`` (functions.py,
interpolators.py, utils.py); every array operation runs in hand-written CUDA
kernels (librmt_b200.so, C ABI in include/rmt_b200.h).  No CPU fallback.
"""
from .functions import (  # noqa: F401
    create_grid, apply_phi_BCs, extrapolate_reference_map, compute_timestep,
    advect_semilagrangian_rk4, advect_semilagrangian_cubic_rk4, advect_weno5_rk3,
    advect_central2_rk3, advect_conservative_rk3, advect_reference_map, solid_cauchy_stress,
    smoothed_heaviside, momentum_step_rk4, momentum_step_rk4_2solids, compute_curvature,
    compute_contact_force, velocity_rhs_blended_optimized, apply_velocity_BCs, build_poisson_matrix,
    pressure_projection_amg, _precompute_poisson_eigenvalues,
    _precompute_poisson_eigenvalues_periodic, rebuild_phi_from_reference_map, reinitialize_phi_PDE,
    reinitialize_phi_fmm, reinitialize_level_set,
    velocity_RK4, heaviside_smooth_alt, compute_solid_stress, extrapolate_transverse_layers_2field,
    advect_semi_lagrangian_rk4,
)
from .interpolators import bilinear_interpolate, bicubic_interpolate  # noqa: F401
from .utils import (  # noqa: F401
    grad_central_x_2nd, grad_central_y_2nd, grad_central_x_4th, grad_central_y_4th,
    diff_upwind_3rd, lap_2nd, fast_solve_3x3,
)

__version__ = "0.1.0"

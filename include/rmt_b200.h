/*
 * rmt_b200.h -- C ABI of librmt_b200.so: the B200 (sm_100a) kernels behind the
 * collocated Reference-Map-Technique timestep of pyRMT.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is
 * pure Python + Numba); its operator API is the set of Python functions in
 * pyRMT/functions.py, pyRMT/interpolators.py and pyRMT/utils.py.  Each entry
 * point below names the reference function (file:line, relative to the
 * upstream repo root) it replaces; pyrmt_b200/functions.py binds them with
 * ctypes under the reference's own names and signatures.
 *
 * Conventions
 *   - Every array argument is a DEVICE pointer to C-contiguous fp64 data of
 *     shape (Ny, Nx), element [j*Nx + i] (x fastest) -- the layout of the
 *     reference's NumPy arrays -- unless stated otherwise.  Buffers are owned
 *     by the caller (PyTorch allocates them); no entry point allocates on the
 *     hot path (plans/workspaces are created once per grid).
 *   - `stream` is a cudaStream_t passed as void*.  All calls are asynchronous.
 *   - Return value: 0 = ok, -1 = invalid argument, -2 = unsupported size,
 *     >0 = cudaError_t of a failed launch.  Nothing throws across the ABI.
 *   - Input arrays are never modified unless the parameter says "in place".
 */
#ifndef RMT_B200_H
#define RMT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int rmt_abi_version(void);
/* Number of CUDA kernels this library has launched since it was loaded (bench.py: gpu_launches). */
unsigned long long rmt_launch_count(void);

/* ------------------------------------------------------------------ utils */
/* pyRMT/utils.py:4-59,116-131.  op: 0 grad_central_x_2nd, 1 grad_central_y_2nd,
 * 2 grad_central_x_4th, 3 grad_central_y_4th, 4 lap_2nd. */
int rmt_stencil_op(const double *f, double *out, int Ny, int Nx, double hx, double hy, int op,
                   void *stream);
/* pyRMT/utils.py:61-114 diff_upwind_3rd(f, u, h, axis); axis 1 = x, 0 = y. */
int rmt_diff_upwind_3rd(const double *f, const double *u, double *out, int Ny, int Nx, double h,
                        int axis, void *stream);
/* pyRMT/functions.py:660-671 smoothed_heaviside (the sin form). */
int rmt_heaviside(const double *x, double *H, long n, double w_t, void *stream);
/* H and rho_local=(1-H)*rho_s+H*rho_f in one pass (functions.py:695-698 and the
 * driver glue benchmarks/soft_disc_in_lid_driven.py:102-103).  H may be NULL. */
int rmt_heaviside_rho(const double *phi, double *H, double *rho, long n, double w_t, double rho_s,
                      double rho_f, void *stream);
/* pyRMT/functions.py:1369-1411 reinitialize_phi_PDE, one kernel per piece:
 *   rmt_reinit_sign: s0 = phi0 / sqrt(phi0^2 + dx^2)                       (:1371)
 *   rmt_reinit_step: out = phi - dtau * s0 * (|grad phi|_Godunov - 1)      (:1375-1404), out != phi.
 * The caller applies the optional phi BC callback between steps (:1406-1407). */
int rmt_reinit_sign(const double *phi0, double *s0, long n, double dx, void *stream);
int rmt_reinit_step(const double *phi, const double *s0, double *out, int Ny, int Nx, double dx, double dy,
                    double dtau, void *stream);
/* out = q * (phi <= 0): the "* solid_mask" glue, soft_disc_in_lid_driven.py:88-91. */
int rmt_mask_mul(const double *q, const double *phi, double *out, long n, void *stream);

/* ------------------------------------------------------------- reductions */
int rmt_reduce_workspace_doubles(void);
/* compute_timestep's reduction (functions.py:176) + the isfinite guard of
 * advect_reference_map (:524).  out2 = {max sqrt(a^2+b^2), #non-finite}. */
int rmt_max_speed(const double *a, const double *b, long n, double *work, double *out2, void *stream);
/* out4 = {sum, min, max, #non-finite} (np.mean / np.ptp call sites :1118,:1298,:1362). */
int rmt_field_stats(const double *x, long n, double *work, double *out4, void *stream);

/* ------------------------------------------------ boundary-condition table */
/* A BC callable (benchmarks/common.py:27-50, tests/test_poisson.py:39-64) is
 * classified once on the host into a gather table and applied in place:
 *   field(dst[e])[cell] = ca[e] * field(src[e])[cell] + cb[e]   (src<0: constant)
 * indices pack (field<<62 | cell), field 0 = u, 1 = v. */
int rmt_apply_bc(double *u, double *v, const long long *dst, const long long *src, const double *ca,
                 const double *cb, int n, void *stream);

/* ---------------------------------------------------------- level set (a2) */
/* rebuild_phi_from_reference_map (functions.py:1366) for the disc family of
 * benchmarks/common.py:55-57:  phi = min_k(|xi - c_k| - R_k).  bin_start/cand
 * describe a (gb x gb) candidate grid over [0,Lx]x[0,Ly]; gb = 0 -> exhaustive. */
int rmt_disc_sdf(const double *X1, const double *X2, double *phi, long n, const double *cx,
                 const double *cy, const double *R, int ndisc, const int *bin_start, const int *cand,
                 int gb, double Lx, double Ly, void *stream);
/* rmt_disc_sdf followed by rmt_solid_stress (functions.py:1366-1367 then :545-658) in one pass: the
 * drivers rebuild phi from the extrapolated map and hand both to momentum_step_rk4
 * (soft_disc_in_lid_driven.py:93-97).  phi is bitwise the rmt_disc_sdf result. */
int rmt_disc_sdf_stress(const double *X1, const double *X2, double *phi, double *sxx, double *sxy,
                        double *syy, double *J, int Ny, int Nx, double dx, double dy, double mu_s,
                        double kappa, double w_cut, double detg_clamp, int isochoric, const double *cx,
                        const double *cy, const double *R, int ndisc, const int *bin_start, const int *cand,
                        int gb, double Lx, double Ly, void *stream);

/* ------------------------------------------------------------ interpolators */
/* pyRMT/interpolators.py:4-62 (cubic=0) and :64-141 (cubic=1); nq query points. */
int rmt_sample(const double *u, const double *xq, const double *yq, double *out, long nq, double dx,
               double dy, int Nx, int Ny, int cubic, void *stream);

/* ---------------------------------------------------------------- advection */
/* pyRMT/functions.py:194-227 (cubic=0) / :230-251 (cubic=1).  q1/out1 may be
 * NULL; when given, a second component shares the backtrace. */
int rmt_advect_sl_rk4(const double *q0, const double *q1, const double *a, const double *b,
                      const double *X, const double *Y, double *out0, double *out1, int Ny, int Nx,
                      double dt, double dx, double dy, int cubic, void *stream);
/* The same on a row slab: the arrays hold rows [row_offset, row_offset + Ny_local) of a grid of Ny rows
 * (coordinates, clamping and cell indices stay global, so the result equals the full-grid call as long
 * as every departure point's stencil lies inside the stored rows -- the caller's halo must cover
 * max|b| dt / dy + 2 rows).  Used by pyrmt_b200/slab.py. */
int rmt_advect_sl_rk4_rows(const double *q0, const double *q1, const double *a, const double *b,
                           const double *X, const double *Y, double *out0, double *out1, int Ny_local, int Nx,
                           int Ny, int row_offset, double dt, double dx, double dy, int cubic, void *stream);
/* pyRMT/functions.py:396-415 (scheme 1 weno5), :443-459 (0 central2), :489-496
 * (2 conservative): SSP-RK3, RHS fused into each stage.  work1/work2: (Ny,Nx). */
int rmt_advect_euler_rk3(const double *q, const double *a, const double *b, const double *phi,
                         double *out, double *work1, double *work2, int Ny, int Nx, double dx,
                         double dy, double dt, double w_cut, int scheme, void *stream);
/* Both reference-map components in one pass per stage (the drivers advect xi1 and xi2 with the
 * same a, b, phi, dt: soft_disc_in_lid_driven.py:88-91); mask_solid != 0 also applies the
 * driver's "* (phi <= 0)" to the result.  Bitwise identical to two rmt_advect_euler_rk3 calls
 * (+ rmt_mask_mul).  work: 4 fields. */
int rmt_advect_euler_rk3_pair(const double *q0, const double *q1, const double *a, const double *b,
                              const double *phi, double *out0, double *out1, double *work, int Ny, int Nx,
                              double dx, double dy, double dt, double w_cut, int scheme, int mask_solid,
                              void *stream);
/* _central2_rhs :420-440 / _weno5_rhs :321-393 / _conservative_rhs :462-486. */
int rmt_euler_rhs(const double *q, const double *a, const double *b, const double *phi, double *out,
                  int Ny, int Nx, double dx, double dy, double w_cut, int scheme, void *stream);

/* ------------------------------------------------------------ extrapolation */
/* pyRMT/functions.py:48-163 extrapolate_reference_map + utils.py:134-167.  X1e, X2e receive the copies of
 * :69-70 with the band filled in; X1e == X1 and X2e == X2 (in place, no copy pass) is allowed. */
long rmt_extrapolate_workspace_bytes(int Ny, int Nx);
int rmt_extrapolate(const double *X1, const double *X2, const double *phi, double *X1e, double *X2e,
                    int Ny, int Nx, double dx, double dy, int max_layers, void *workspace,
                    void *stream);
/* The same on a block of rows [row_offset, row_offset + Ny) of a taller grid: the fit uses
 * the GLOBAL absolute coordinates y = dy * (j + row_offset), so a slab reproduces the bits of
 * the whole-grid sweep (pyrmt_b200/slab.py). */
int rmt_extrapolate_rows(const double *X1, const double *X2, const double *phi, double *X1e, double *X2e,
                         int Ny, int Nx, int row_offset, double dx, double dy, int max_layers,
                         void *workspace, void *stream);
/* Which sweep variant runs (all are bit-identical to the reference's serial raster sweep, functions.py:95-161):
 *   variant -1 = chosen on the device from the layer-0 band geometry (default), 0 = per-layer launches,
 *   16 / 8 = all-layers row-pipelined kernel with that many rows per block, 1 = all-layers kernel with one
 *   CTA per isolated tile (falls back to the device's choice when the bands are not isolated);
 *   rows = macro-tile height for 16 / 8 (0 = chosen on the device);
 *   cap  = prepared-record capacity of the per-layer sweep (0 = default; small values exercise the
 *          inline phase-A path).  Process-wide; a test / tuning hook. */
int rmt_extrapolate_set_mode(int variant, int rows, long cap);
/* out3 = {variant, rows or tile-size exponent, 0} the last rmt_extrapolate[_rows] call on this workspace ran
 * with (synchronises the stream). */
int rmt_extrapolate_last_mode(const void *workspace, int Ny, int Nx, int *out3, void *stream);
/* device exp() used for the weights, exposed so tests can prove bit-equality
 * with the host libm (functions.py:120). */
int rmt_exp_probe(const double *x, double *y, long n, void *stream);
/* the WENO5 kernels replace x / 6 and the three a_k / s of functions.py:256-318 by a correctly rounded
 * reciprocal-and-residual sequence; exposed so tests can prove bit-equality with IEEE division.
 * mode 0: out = a / 6 (b unused); mode 1: out = a / b. */
int rmt_weno_div_probe(const double *a, const double *b, double *out, long n, int mode, void *stream);

/* ------------------------------------------------- stress + momentum (a10-a13) */
/* pyRMT/functions.py:545-658 solid_cauchy_stress. */
int rmt_solid_stress(const double *X1, const double *X2, const double *phi, double *sxx, double *sxy,
                     double *syy, double *J, int Ny, int Nx, double dx, double dy, double mu_s,
                     double kappa, double w_cut, double detg_clamp, int isochoric, void *stream);
/* pyRMT/functions.py:837-861 compute_curvature. */
int rmt_curvature(const double *phi, double *curv, int Ny, int Nx, double dx, double dy, void *stream);
/* functions.py:700-704: st_force = -gamma * curvature(phi) * grad(H(phi)). */
int rmt_surface_tension(const double *phi, double *fsx, double *fsy, int Ny, int Nx, double dx,
                        double dy, double w_t, double gamma, void *stream);
/* pyRMT/functions.py:897-944 velocity_rhs_blended_optimized (H and rho_local
 * are arrays as in the reference; fsx/fsy may be NULL = 0.0). */
int rmt_velocity_rhs(const double *u, const double *v, const double *p, const double *sxx,
                     const double *sxy, const double *syy, const double *H, const double *rho,
                     const double *fsx, const double *fsy, double *ru, double *rv, int Ny, int Nx,
                     double dx, double dy, double mu_f, void *stream);
/* One classical-RK4 stage of momentum_step_rk4 (functions.py:711-758): the RHS
 * of the BC-applied stage state (us, vs) with the Kelvin-Voigt term (:717-730),
 * fused with the stage update.  H and rho_local are evaluated from phi in-kernel.
 *   stage 1: acc = k            out = u0 + (dt/2) k
 *   stage 2: acc += 2k          out = u0 + (dt/2) k
 *   stage 3: acc += 2k          out = u0 + dt k
 *   stage 4:                    out = u0 + (dt/6)(acc + k)
 * The caller applies the BC table to `out` between stages (rmt_apply_bc). */
int rmt_momentum_stage(const double *us, const double *vs, const double *p, const double *sxx,
                       const double *sxy, const double *syy, const double *phi, const double *fsx,
                       const double *fsy, const double *u0, const double *v0, double *acc_u,
                       double *acc_v, double *out_u, double *out_v, int Ny, int Nx, double dx,
                       double dy, double dt, double mu_f, double eta_s, double w_t, double rho_s,
                       double rho_f, int stage, void *stream);
/* One classical-RK4 stage of momentum_step_rk4_2solids (functions.py:765-835): two solid stresses and
 * two level sets, n = 2 mixture  sigma = (Ha+Hb-1) sigma_f + (1-Ha) sigma_A + (1-Hb) sigma_B,
 * rho = (Ha+Hb-1) rho_f + (1-Ha) rho_s + (1-Hb) rho_s, body force (fcx, fcy) (NULL = 0).  Stage
 * bookkeeping as in rmt_momentum_stage. */
int rmt_momentum_stage_2solids(const double *us, const double *vs, const double *p, const double *sAxx,
                               const double *sAxy, const double *sAyy, const double *sBxx, const double *sBxy,
                               const double *sByy, const double *phi_a, const double *phi_b, const double *fcx,
                               const double *fcy, const double *u0, const double *v0, double *acc_u,
                               double *acc_v, double *out_u, double *out_v, int Ny, int Nx, double dx, double dy,
                               double dt, double mu_f, double w_t, double rho_s, double rho_f, int stage,
                               void *stream);
/* pyRMT/functions.py:864-895 compute_contact_force: k_rep * delta_s(phi12) * sign(phi12) * n12 inside
 * either solid, phi12 = (phi1 - phi2)/2, delta_s(x) = (1 + cos(pi x / w_c)) / (2 w_c) for |x| < w_c. */
int rmt_contact_force(const double *phi1, const double *phi2, double *fx, double *fy, int Ny, int Nx, double dx,
                      double dy, double k_rep, double w_c, void *stream);
/* Energy diagnostics and solid centroid in one pass (pyRMT/output.py:6-39 compute_kinetic_energy,
 * :41-134 compute_strain_energy, :136-193 compute_viscous_dissipation; benchmarks/common.py:110-115
 * disc_centroid).  out6 = {sum 0.5 rho |u|^2, sum strain-energy density over solid cells,
 * sum 2 mu D:D, #cells with phi <= 0, sum x over them, sum y over them} -- the caller multiplies the
 * first three by dx*dy and divides the coordinate sums by the count.  (a, b) and (X1, X2) may be NULL
 * (their sums are then 0); Xc/Yc = node coordinates or NULL for i*dx, j*dy.
 * work: rmt_diagnostics_workspace_doubles() doubles. */
int rmt_diagnostics_workspace_doubles(void);
int rmt_diagnostics(const double *a, const double *b, const double *X1, const double *X2, const double *phi,
                    const double *Xc, const double *Yc, int Ny, int Nx, double dx, double dy, double rho_f,
                    double rho_s, double mu_f, double mu_s, double kappa, double eta_s, double w_t, double *work,
                    double *out6, void *stream);
/* out = min(a, b) elementwise (functions.py:835, np.minimum(Ja, Jb)). */
int rmt_min2(const double *a, const double *b, double *out, long n, void *stream);

/* ------------------------------------------------------ projection (a14-a17) */
/* _compute_divergence (functions.py:1005-1014). */
int rmt_divergence(const double *a, const double *b, double *div, int Ny, int Nx, double dx,
                   double dy, void *stream);
/* _compute_divergence_rc (:1016-1071), constant-density branch; d_f = dt/mean(rho)
 * is read from the device scalar rho_sum[0] / (Ny*Nx) (no host sync). */
int rmt_divergence_rc(const double *a, const double *b, const double *p_prev, const double *rho_sum,
                      double *div, int Ny, int Nx, double dx, double dy, double dt, void *stream);
/* _compute_pressure_gradient (:1073-1089). */
int rmt_pressure_gradient(const double *p, double *gx, double *gy, int Ny, int Nx, double dx,
                          double dy, void *stream);
/* periodic: _compute_divergence_periodic :1236-1243, _compute_pressure_gradient_periodic :1246-1252 */
int rmt_divergence_periodic(const double *a, const double *b, double *div, int Ny, int Nx, double dx,
                            double dy, void *stream);
int rmt_pressure_gradient_periodic(const double *p, double *gx, double *gy, int Ny, int Nx, double dx,
                                   double dy, void *stream);
/* Fused projection front end (:1292-1295,:1331 / :1286-1290):
 *   rhs = rho * div / dt      (Neumann: rho elementwise; rho==NULL -> rho_scalar)
 *   rhs = rho_bar * div / dt  (periodic: rho_bar = rho_sum[0]/(Ny*Nx))
 * div is Rhie-Chow when p_prev != NULL (Neumann only), else plain / periodic.
 * periodic: 0 Neumann; 1 periodic in x and y on the reduced (Ny-1, Nx-1) grid; 2 row slab of a
 * periodic grid: periodic in x, the y neighbours are the stored rows above / below (the caller has
 * exchanged the wrap rows; rows 0 and Ny-1 of the slab are halo rows whose outputs are meaningless).
 * The same values apply to rmt_projection_correct. */
int rmt_projection_rhs(const double *a, const double *b, const double *p_prev, const double *rho,
                       double rho_scalar, const double *rho_sum, double *rhs, int Ny, int Nx,
                       double dx, double dy, double dt, int periodic, void *stream);
/* Fused projection back end (:1350-1362 / :1284-1290):
 *   pc = sol - sol_sum[0]/(Ny*Nx)   (sol_sum: device scalar from the solver; NULL = already centred)
 *   a = a* - (dt/rho) dpc/dx ; b likewise ; p = (p_prev or 0) + pc
 * rho == NULL -> rho_scalar.  The caller removes mean(p) afterwards
 * (rmt_field_stats + rmt_subtract_mean) and applies the BC table to (a, b). */
int rmt_projection_correct(const double *sol, const double *sol_sum, const double *a_star,
                           const double *b_star, const double *rho, double rho_scalar,
                           const double *p_prev, double *a, double *b, double *p, int Ny, int Nx,
                           double dx, double dy, double dt, int periodic, void *stream);
/* The same back end with `p -= mean(p)` (:1361) folded in: p = p_prev + pc - p_prev_sum[0]/(Ny*Nx)
 * (mean(pc) is zero to rounding, so mean(p_prev) is mean(p_prev + pc) to rounding; p_prev_sum == NULL: 0),
 * and sum(p) of the field written here goes to p_sum_out (device scalar) for the next step -- p is
 * neither re-read for its mean nor rewritten.  partial: rmt_projection_partials(Ny, Nx) doubles. */
long rmt_projection_partials(int Ny, int Nx);
int rmt_projection_correct_centered(const double *sol, const double *sol_sum, const double *a_star,
                                    const double *b_star, const double *rho, double rho_scalar,
                                    const double *p_prev, const double *p_prev_sum, double *a, double *b,
                                    double *p, double *partial, double *p_sum_out, int Ny, int Nx, double dx,
                                    double dy, double dt, int periodic, void *stream);
/* x[k] = x[k] - s[0]/n  in place (the np.mean removals :1118,:1232,:1362). */
int rmt_subtract_mean(double *x, const double *sum, long n, void *stream);

/* -------------------------------------------------- spectral Poisson solves */
/* Plan for one grid.  kind 0: DCT-I / Neumann (_solve_poisson_dct :1107-1119,
 * scipy.fft.dctn/idctn type 1); kind 1: periodic FFT on the reduced
 * (Ny-1, Nx-1) grid (_solve_poisson_fft :1216-1233, numpy.fft.fft2/ifft2).
 * Power-of-two transform lengths run hand-written shared-memory FFTs; other
 * lengths (N <= RMT_DENSE_MAX) run dense O(N^2) transforms per line. */
#define RMT_DENSE_MAX 1536
typedef struct rmt_poisson_plan rmt_poisson_plan;
int rmt_poisson_plan_create(int Ny, int Nx, int kind, rmt_poisson_plan **plan);
void rmt_poisson_plan_destroy(rmt_poisson_plan *plan);
/* 1 if the plan uses the shared-memory FFT path in both directions, else 0. */
int rmt_poisson_plan_is_fast(const rmt_poisson_plan *plan);
/* The fast paths cache tables derived from `eig` (its transpose; the periodic inverse symbol) inside the
 * plan and rebuild them only when the eig POINTER changes.  A caller that passes a different table at a
 * recycled address (or edits the table in place) must call this first (pyrmt_b200/_runtime.py keys the
 * table by content and does). */
int rmt_poisson_plan_invalidate(rmt_poisson_plan *plan);
/* sol = idctn(dctn(rhs)/eig), then sol -= mean(sol).  eig: (Ny, Nx).
 * sum_out (device, 1 double): if non-NULL the mean is NOT removed; sum(sol) is
 * written there instead so the consumer (rmt_projection_correct) can fold the
 * shift into its own pass. */
int rmt_poisson_solve_dct(rmt_poisson_plan *plan, const double *rhs, const double *eig, double *sol,
                          double *sum_out, void *stream);
/* r = rhs[:-1,:-1]; r -= mean(r); spec = fft2(r)/eig; spec[null] = 0;
 * sol = tile_overlap(real(ifft2(spec))); sol -= mean(sol).
 * eig: (Ny-1, Nx-1) doubles; null: (Ny-1, Nx-1) bytes. */
int rmt_poisson_solve_fft(rmt_poisson_plan *plan, const double *rhs, const double *eig,
                          const unsigned char *null_mask, double *sol, double *sum_out, void *stream);

/* -------------------------------- building blocks of the slab-decomposed solve */
/* DCT-I (scipy.fft.dct type 1, the 1-D factor of functions.py:1115-1117) along the rows of
 * a (nrows, N) array, N - 1 a power of two in [8, 8192]:
 *   eig == NULL: out = scale * DCT-I(in)
 *   eig != NULL: out = DCT-I( DCT-I(in) * scale / eig ), eig laid out like in.
 * in == out is allowed.  Used by pyrmt_b200/slab.py between the all-to-all transposes. */
int rmt_dct_lines(const double *in, double *out, const double *eig, int nrows, int N, double scale,
                  void *stream);
/* Discrete Hartley transform H_k = sum_n x_n (cos + sin)(2 pi k n / m) along the rows (length m, a
 * power of two in [16, 16384], row strides ldi / ldo doubles) of a real array:
 *   out[r][k] = H(in[r])[k] * (mul ? mul[r*m + k] : scale).
 * The separable 2-D Hartley transform diagonalises the periodic Poisson operator of
 * functions.py:1177-1233 (its symbol is even in each wavenumber), so it stands in for
 * numpy.fft.fft2/ifft2 (:1227,:1230) on real fields.  in == out is allowed. */
int rmt_dht_lines(const double *in, double *out, const double *mul, int nrows, int m, long ldi, long ldo,
                  double scale, void *stream);
/* out (C, R) = in (R, C)^T */
int rmt_transpose(const double *in, double *out, int R, int C, void *stream);
/* dst[r*dst_ld + c] = src[r*src_ld + c], rows x cols block (all-to-all pack / unpack). */
int rmt_copy2d(const double *src, double *dst, int rows, int cols, long src_ld, long dst_ld, void *stream);

/* ------------------------------------------------ slab exchanges over NVLink peer memory */
/* The reference is one process on one grid.  Cut into row slabs (one process per GPU on one node), its
 * whole-grid stencils (pyRMT/utils.py:4-114), transforms (pyRMT/functions.py:1107-1119 dctn / idctn,
 * :1216-1233 fft2 / ifft2) and means (:1119, :1231, :1361) need a halo exchange, an all-to-all transpose
 * and an all-reduce.  These entry points do them with plain stores into the neighbour's memory (CUDA IPC
 * mapping; NVLink / NVSwitch underneath) -- pyrmt_b200/slab.py PeerComm is the caller. */
#define RMT_PEER_MAX 16           /* ranks of one node */
#define RMT_PUT_MAX 32            /* 2-D block copies per rmt_peer_put2d launch */
/* An arena other processes can map: cudaMalloc + zero fill (synchronous; set-up only). */
int rmt_peer_alloc(size_t bytes, void **ptr);
int rmt_peer_free(void *ptr);
/* 64-byte CUDA IPC handle of an arena / mapping of another process's arena into this one. */
int rmt_peer_export(void *ptr, unsigned char *handle64);
int rmt_peer_import(const unsigned char *handle64, void **ptr);
int rmt_peer_release(void *ptr);
/* dst[r*dst_ld + c] = src[r*src_ld + c] for n <= RMT_PUT_MAX blocks in ONE launch; dst (or src) may be
 * peer memory.  `desc` is a host array. */
typedef struct rmt_put2d {
    const double *src;
    double *dst;
    int rows, cols;
    long src_ld, dst_ld;
} rmt_put2d;
int rmt_peer_put2d(const rmt_put2d *desc, int n, void *stream);
/* The transpose AND the all-to-all of the distributed line transforms in one kernel:
 * in (R, C; row stride ldi);  for column c in [start[q], start[q+1]):  dst[q][(c - start[q]) * dst_ld[q] + r]
 * = in[r][c].  start (nparts + 1 entries, start[0] = 0, start[nparts] = C), dst and dst_ld are host arrays;
 * dst[q] is a device pointer valid in this process (own memory or a mapped peer arena). */
int rmt_transpose_scatter(const double *in, int R, int C, long ldi, int nparts, const int *start,
                          double *const *dst, const long *dst_ld, void *stream);
/* Barrier on `stream` between this rank and the peers q with epochs[q] != 0 (host array of `world` entries):
 * flags[q] = rank q's array of >= world uint64 counters as mapped in this process.  Rank r release-stores
 * epochs[q] into flags[q][r] (system scope, after a system fence: everything this stream did before is visible
 * to the peer first), then waits until flags[r][q] >= epochs[q].  The two sides of a pair must name each other
 * in the same barriers and count them identically (epochs[q] grows by one per barrier that involves q).
 * A peer that never arrives sets *err (device int) after timeout_s seconds instead of hanging the GPU; once
 * *err is set every later barrier falls through at once. */
int rmt_peer_barrier(void *const *flags, int rank, int world, const unsigned long long *epochs, double timeout_s,
                     int *err, void *stream);
/* out[k] = reduction over q = 0..world-1, in that order, of slots[q*stride + k]; op 0 sum, 1 max, 2 min
 * (the local half of the peer all-reduce: bitwise identical on every rank). */
int rmt_peer_reduce(const double *slots, int world, long stride, int n, int op, double *out, void *stream);


#ifdef __cplusplus
}
#endif
#endif /* RMT_B200_H */

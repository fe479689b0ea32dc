// advect.cu -- reference-map advection kernels.
//
// Replaces (reference file:line):
//   pyRMT/interpolators.py:4-62,64-141,143-154   bilinear / monotone bicubic sampling
//   pyRMT/functions.py:194-251                   semi-Lagrangian RK4 backtrace (bilinear, cubic)
//   pyRMT/functions.py:256-415                   WENO5 faces + RHS + SSP-RK3
//   pyRMT/functions.py:420-459                   central2 RHS + SSP-RK3
//   pyRMT/functions.py:462-496                   conservative RHS + SSP-RK3
//
// BIT-EXACT: the extrapolation downstream amplifies 1-ulp differences of its
// inputs by up to 1e6, so the advected map has to
// equal the reference's bit for bit.  Every expression tree below follows the
// reference's (functions.py / interpolators.py) operation by operation and this
// file is compiled with -fmad=false; fp64 division and sqrt are IEEE in CUDA.
//
// Semi-Lagrangian: one thread per node does the whole RK4 backtrace (8 velocity
// samples) and the final sample of q -- nine 4-tap gathers served by L1/L2
// (displacements are <= 1 cell at CFL 0.2), one coalesced store.
// Eulerian schemes: one launch per SSP-RK3 stage; the stage kernel fuses the
// RHS with the Shu-Osher combination so no RHS array ever reaches HBM.
#include "common.cuh"
#include "../../include/rmt_b200.h"

using namespace rmt;

namespace {

// ---------------------------------------------------------------- sampling
__device__ __forceinline__ bool clamp_coords(double xq, double yq, double inv_dx_unused, double dx,
                                             double dy, int Nx, int Ny, double &x, double &y)
{
    x = xq / dx;
    y = yq / dy;
    if (!(isfinite(x) && isfinite(y))) return false;
    if (x < 0.0) x = 0.0;
    else if (x > Nx - 1.0) x = Nx - 1.0;
    if (y < 0.0) y = 0.0;
    else if (y > Ny - 1.0) y = Ny - 1.0;
    return true;
}

// jlo, jhi: the rows [jlo, jhi) that exist in memory (a row slab addressed from the virtual row 0; the whole
// grid: 0, Ny).  Sample rows are clamped into them AFTER the weights are formed, so a node whose departure
// point leaves the slab (only halo nodes, whose results are discarded: the caller checks that the owned
// rows' stencils stay inside) reads valid memory, and every other result is unchanged bit for bit.
__device__ __forceinline__ double bilinear(const double *__restrict__ u, double xq, double yq,
                                           double dx, double dy, int Nx, int Ny, int jlo, int jhi)
{
    double x, y;
    if (!clamp_coords(xq, yq, 0.0, dx, dy, Nx, Ny, x, y)) return nan("");
    int ix = (int)floor(x), iy = (int)floor(y);
    if (ix >= Nx - 1) ix = Nx - 2;
    if (iy >= Ny - 1) iy = Ny - 2;
    double fx = x - ix, fy = y - iy;
    iy = min(max(iy, jlo), jhi - 2);
    const double *r0 = u + (size_t)iy * Nx + ix;
    const double *r1 = r0 + Nx;
    double v00 = __ldg(r0), v10 = __ldg(r0 + 1), v01 = __ldg(r1), v11 = __ldg(r1 + 1);
    return (1 - fx) * (1 - fy) * v00 + fx * (1 - fy) * v10 + (1 - fx) * fy * v01 + fx * fy * v11;
}

// Two fields sampled at the same point share the index/weight computation.
__device__ __forceinline__ void bilinear2(const double *__restrict__ a, const double *__restrict__ b,
                                          double xq, double yq, double dx, double dy, int Nx, int Ny,
                                          int jlo, int jhi, double &va, double &vb)
{
    double x, y;
    if (!clamp_coords(xq, yq, 0.0, dx, dy, Nx, Ny, x, y)) { va = vb = nan(""); return; }
    int ix = (int)floor(x), iy = (int)floor(y);
    if (ix >= Nx - 1) ix = Nx - 2;
    if (iy >= Ny - 1) iy = Ny - 2;
    double fx = x - ix, fy = y - iy;
    iy = min(max(iy, jlo), jhi - 2);
    double w00 = (1 - fx) * (1 - fy), w10 = fx * (1 - fy), w01 = (1 - fx) * fy, w11 = fx * fy;
    size_t o = (size_t)iy * Nx + ix;
    va = w00 * __ldg(a + o) + w10 * __ldg(a + o + 1) + w01 * __ldg(a + o + Nx) + w11 * __ldg(a + o + Nx + 1);
    vb = w00 * __ldg(b + o) + w10 * __ldg(b + o + 1) + w01 * __ldg(b + o + Nx) + w11 * __ldg(b + o + Nx + 1);
}

__device__ __forceinline__ double catmull_rom(double v0, double v1, double v2, double v3, double x)
{
    double a0 = -0.5 * v0 + 1.5 * v1 - 1.5 * v2 + 0.5 * v3;
    double a1 = v0 - 2.5 * v1 + 2.0 * v2 - 0.5 * v3;
    double a2 = -0.5 * v0 + 0.5 * v2;
    return a0 * (x * x * x) + a1 * (x * x) + a2 * x + v1;
}

__device__ __forceinline__ double bicubic(const double *__restrict__ u, double xq, double yq,
                                          double dx, double dy, int Nx, int Ny, int jlo, int jhi)
{
    double x, y;
    if (!clamp_coords(xq, yq, 0.0, dx, dy, Nx, Ny, x, y)) return nan("");
    int ix = (int)floor(x), iy = (int)floor(y);
    double fx = x - ix, fy = y - iy;
    double rows[4], lo = 1e18, hi = -1e18;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int yg = min(max(min(max(iy - 1 + m, 0), Ny - 1), jlo), jhi - 1);
        double c[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            int xg = min(max(ix - 1 + n, 0), Nx - 1);
            double v = __ldg(u + (size_t)yg * Nx + xg);
            c[n] = v;
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
        rows[m] = catmull_rom(c[0], c[1], c[2], c[3], fx);
    }
    double r = catmull_rom(rows[0], rows[1], rows[2], rows[3], fy);
    if (r < lo) r = lo;
    if (r > hi) r = hi;
    return r;
}

template <bool CUBIC>
__global__ void k_sample(const double *__restrict__ u, const double *__restrict__ xq,
                         const double *__restrict__ yq, double *__restrict__ out, long nq, double dx,
                         double dy, int Nx, int Ny)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < nq; k += (long)gridDim.x * blockDim.x)
        out[k] = CUBIC ? bicubic(u, xq[k], yq[k], dx, dy, Nx, Ny, 0, Ny)
                       : bilinear(u, xq[k], yq[k], dx, dy, Nx, Ny, 0, Ny);
}

// ------------------------------------------------- semi-Lagrangian RK4 backtrace
// NQ reference-map components share one backtrace (NQ=1 is the reference's
// per-component call; NQ=2 is the fused form the driver can use because X1 and
// X2 are advected with the same (a, b, dt)).
template <bool CUBIC, int NQ>
__global__ void k_advect_sl(const double *__restrict__ q0, const double *__restrict__ q1,
                            const double *__restrict__ a, const double *__restrict__ b,
                            const double *__restrict__ X, const double *__restrict__ Y,
                            double *__restrict__ o0, double *__restrict__ o1, int Ny, int Nx,
                            double dt, double dx, double dy, int Nyl, int joff)
{
    // Row slab (rmt_advect_sl_rk4_rows): the arrays hold rows [joff, joff + Nyl) of a grid of Ny rows.
    // Sampling works in global indices, so the sampled fields are addressed from the virtual row 0.
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Nyl) return;
    size_t c = (size_t)j * Nx + i;
    const size_t voff = (size_t)joff * Nx;
    const int jhi = joff + Nyl;
    a -= voff; b -= voff; q0 -= voff;
    if (NQ == 2) q1 -= voff;
    double x = X[c], y = Y[c];
    const double half_dt = 0.5 * dt, sixth_dt = dt / 6.0;
    double k1x, k1y, k2x, k2y, k3x, k3y, k4x, k4y;
    auto vel = [&](double xx, double yy, double &vx, double &vy) {
        if (CUBIC) {
            vx = bicubic(a, xx, yy, dx, dy, Nx, Ny, joff, jhi);
            vy = bicubic(b, xx, yy, dx, dy, Nx, Ny, joff, jhi);
        } else {
            bilinear2(a, b, xx, yy, dx, dy, Nx, Ny, joff, jhi, vx, vy);
        }
    };
    vel(x, y, k1x, k1y);
    vel(x - half_dt * k1x, y - half_dt * k1y, k2x, k2y);
    vel(x - half_dt * k2x, y - half_dt * k2y, k3x, k3y);
    vel(x - dt * k3x, y - dt * k3y, k4x, k4y);
    double xb = x - sixth_dt * (k1x + 2 * k2x + 2 * k3x + k4x);
    double yb = y - sixth_dt * (k1y + 2 * k2y + 2 * k3y + k4y);
    if (CUBIC) {
        o0[c] = bicubic(q0, xb, yb, dx, dy, Nx, Ny, joff, jhi);
        if (NQ == 2) o1[c] = bicubic(q1, xb, yb, dx, dy, Nx, Ny, joff, jhi);
    } else if (NQ == 2) {
        double r0, r1;
        bilinear2(q0, q1, xb, yb, dx, dy, Nx, Ny, joff, jhi, r0, r1);
        o0[c] = r0;
        o1[c] = r1;
    } else {
        o0[c] = bilinear(q0, xb, yb, dx, dy, Nx, Ny, joff, jhi);
    }
}

// ------------------------------------------------------------------- WENO5
__device__ __forceinline__ double sq(double v) { return v * v; }

// a / b from y = RN(1 / b), correctly rounded (IEEE a / b bit for bit) in five fused operations:
// q = RN(a y) is within two ulps, one residual step makes it faithful, and for a faithful q and a
// correctly rounded reciprocal the second step q + (a - b q) y rounds to RN(a / b) (Markstein 1990,
// the closing step of every Newton-Raphson divider; fma() is not subject to -fmad=false).  The
// residuals must be exact, which needs operands, reciprocal and quotient well inside the normal
// range: |a|, |b| in [1e-140, 1e140].  A reconstruction whose operands leave that range (zeros,
// subnormals, huge, non-finite) takes the plain IEEE version below, as a whole.
// rmt_weno_div_probe lets the tests compare this against `/` on tens of millions of operands.
__device__ __forceinline__ bool div_range_ok(double a) { const double m = fabs(a); return m >= 1e-140 && m <= 1e140; }
__device__ __forceinline__ double div_by_recip_raw(double a, double b, double y)
{
    const double q0 = a * y;
    const double r0 = fma(-b, q0, a);
    const double q1 = fma(r0, y, q0);
    const double r1 = fma(-b, q1, a);
    return fma(r1, y, q1);
}
__device__ __forceinline__ double div_by_recip(double a, double b, double y)    // guarded form (probe, tests)
{
    if (!(div_range_ok(a) && div_range_ok(b))) return a / b;
    return div_by_recip_raw(a, b, y);
}

// the reference's expression tree with IEEE divisions throughout (functions.py:256-286 / :289-318 after
// the stencil has been ordered): n0..n2 numerators of the three candidate values, b0..b2 smoothness
__device__ __noinline__ double weno_mix_ieee(double n0, double n1, double n2, double b0, double b1, double b2)
{
    const double eps = 1.0e-6;
    double r0 = n0 / 6.0, r1 = n1 / 6.0, r2 = n2 / 6.0;
    double a0 = 0.1 / sq(eps + b0), a1 = 0.6 / sq(eps + b1), a2 = 0.3 / sq(eps + b2);
    double s = a0 + a1 + a2;
    return (a0 / s) * r0 + (a1 / s) * r1 + (a2 / s) * r2;
}

__device__ __forceinline__ double weno_mix(double n0, double n1, double n2, double b0, double b1, double b2)
{
    const double eps = 1.0e-6;
    double a0 = 0.1 / sq(eps + b0), a1 = 0.6 / sq(eps + b1), a2 = 0.3 / sq(eps + b2);
    double s = a0 + a1 + a2;
    if (!(div_range_ok(n0) && div_range_ok(n1) && div_range_ok(n2) && div_range_ok(a0) && div_range_ok(a1) &&
          div_range_ok(a2) && div_range_ok(s)))
        return weno_mix_ieee(n0, n1, n2, b0, b1, b2);
    const double sixth = 1.0 / 6.0;
    const double r0 = div_by_recip_raw(n0, 6.0, sixth), r1 = div_by_recip_raw(n1, 6.0, sixth),
                 r2 = div_by_recip_raw(n2, 6.0, sixth);
    const double y = 1.0 / s;                  // one reciprocal serves the three quotients
    return div_by_recip_raw(a0, s, y) * r0 + div_by_recip_raw(a1, s, y) * r1 + div_by_recip_raw(a2, s, y) * r2;
}

// left-biased value at k+1/2 from (k-2..k+2), functions.py:256-286
__device__ __forceinline__ double weno_minus(double vm2, double vm1, double v0, double vp1, double vp2)
{
    double n0 = 2.0 * vm2 - 7.0 * vm1 + 11.0 * v0;
    double n1 = -vm1 + 5.0 * v0 + 2.0 * vp1;
    double n2 = 2.0 * v0 + 5.0 * vp1 - vp2;
    double b0 = (13.0 / 12.0) * sq(vm2 - 2.0 * vm1 + v0) + 0.25 * sq(vm2 - 4.0 * vm1 + 3.0 * v0);
    double b1 = (13.0 / 12.0) * sq(vm1 - 2.0 * v0 + vp1) + 0.25 * sq(vm1 - vp1);
    double b2 = (13.0 / 12.0) * sq(v0 - 2.0 * vp1 + vp2) + 0.25 * sq(3.0 * v0 - 4.0 * vp1 + vp2);
    return weno_mix(n0, n1, n2, b0, b1, b2);
}

// right-biased value at k+1/2 from (k-1..k+3), functions.py:289-318
__device__ __forceinline__ double weno_plus(double vm1, double v0, double vp1, double vp2, double vp3)
{
    double n0 = 2.0 * vp3 - 7.0 * vp2 + 11.0 * vp1;
    double n1 = -vp2 + 5.0 * vp1 + 2.0 * v0;
    double n2 = 2.0 * vp1 + 5.0 * v0 - vm1;
    double b0 = (13.0 / 12.0) * sq(vp3 - 2.0 * vp2 + vp1) + 0.25 * sq(3.0 * vp1 - 4.0 * vp2 + vp3);
    double b1 = (13.0 / 12.0) * sq(vp2 - 2.0 * vp1 + v0) + 0.25 * sq(vp2 - v0);
    double b2 = (13.0 / 12.0) * sq(vp1 - 2.0 * v0 + vm1) + 0.25 * sq(vp1 - 4.0 * v0 + 3.0 * vm1);
    return weno_mix(n0, n1, n2, b0, b1, b2);
}

// d q / d(line) at index k of n with the reference's face selection and rim
// fallbacks (functions.py:347-389).  `g(o)` returns q at line index k+o.
// For vel < 0 the reference feeds BOTH faces the same five points, so the
// difference is exactly zero unless the upper-rim fallback kicks in; that
// behaviour is reproduced, not repaired.
template <class G>
__device__ __forceinline__ double weno5_dq(G g, int k, int n, double vel, double h,
                                           double q_last /* q at line index n-1 */)
{
    double qp, qm;
    if (vel >= 0.0) {
        double m2 = g(-2), m1 = g(-1), c0 = g(0), p1 = g(1), p2 = g(2);
        qp = weno_minus(m2, m1, c0, p1, p2);
        qm = (k >= 3) ? weno_minus(g(-3), m2, m1, c0, p1) : qp;
    } else {
        double m1 = g(-1), c0 = g(0), p1 = g(1), p2 = g(2);
        if (k + 3 < n) {
            double p3 = g(3);
            // Both faces see the same five points: (qp - qp)/h is +0.0 exactly whenever
            // qp is finite, which moderate stencil values guarantee -- skip the work.
            const double big = 1e60;
            if (fabs(m1) < big && fabs(c0) < big && fabs(p1) < big && fabs(p2) < big && fabs(p3) < big)
                return 0.0;
            qp = weno_plus(m1, c0, p1, p2, p3);
            qm = qp;
        } else {
            qp = weno_minus(g(-2), m1, c0, p1, p2);
            qm = weno_plus(m1, c0, p1, p2, q_last);
        }
    }
    return (qp - qm) / h;
}

// SSP-RK3 stage:  out = c0*q0 + c1*(qs + dt*RHS(qs))   (functions.py:406-415)
//   stage 1: c0=0,   c1=1   (q0 unused)      q1 = q + dt R(q)
//   stage 2: c0=3/4, c1=1/4                  q2 = 3/4 q + 1/4 (q1 + dt R(q1))
//   stage 3: c0=1/3, c1=2/3                  q' = 1/3 q + 2/3 (q2 + dt R(q2))
// SCHEME: 0 central2, 1 weno5, 2 conservative.  RHS is zero where phi > w_cut
// or outside the scheme's interior range.
struct EulerFields {      // up to two transported fields that share (a, b, phi) -- xi1 and xi2
    const double *q0[2], *qs[2];
    double *out[2];
};

// One node of a stage, any scheme, any position (rim fallbacks included): the generic path.
template <int SCHEME, int NQ>
__device__ __forceinline__ void euler_cell(const EulerFields &F, const double *__restrict__ a,
                                           const double *__restrict__ b, const double *__restrict__ phi, int j,
                                           int i, int Ny, int Nx, double dx, double dy, double dt, double w_cut,
                                           double c0, double c1, int first, int mask_solid)
{
    size_t c = (size_t)j * Nx + i;
    const int halo = (SCHEME == 1) ? 2 : 1;
    const bool interior = (i >= halo && i < Nx - halo && j >= halo && j < Ny - halo);
    const double ph = phi[c];
    const bool active = interior && !(ph > w_cut);
    double u = 0.0, v = 0.0;
    if (active && SCHEME != 2) { u = a[c]; v = b[c]; }
    // final stage may apply the driver's "* (phi <= 0)" mask (soft_disc_in_lid_driven.py:88-91)
    const double m = (mask_solid && !(ph <= 0.0)) ? 0.0 : 1.0;
#pragma unroll
    for (int f = 0; f < NQ; ++f) {
        const double *qs = F.qs[f];
        const double qc = qs[c];
        double rhs = 0.0;
        if (active) {
            if (SCHEME == 0) {
                double dqdx = (__ldg(qs + c + 1) - __ldg(qs + c - 1)) * (0.5 / dx);
                double dqdy = (__ldg(qs + c + Nx) - __ldg(qs + c - Nx)) * (0.5 / dy);
                rhs = -(u * dqdx + v * dqdy);
            } else if (SCHEME == 2) {
                double fx = (__ldg(a + c + 1) * __ldg(qs + c + 1) - __ldg(a + c - 1) * __ldg(qs + c - 1)) * (0.5 / dx);
                double fy = (__ldg(b + c + Nx) * __ldg(qs + c + Nx) - __ldg(b + c - Nx) * __ldg(qs + c - Nx)) * (0.5 / dy);
                rhs = -(fx + fy);
            } else {
                const double *row = qs + (size_t)j * Nx;
                const double *col = qs + i;
                double dqdx = weno5_dq([&](int o) { return __ldg(row + i + o); }, i, Nx, u, dx,
                                       __ldg(row + Nx - 1));
                double dqdy = weno5_dq([&](int o) { return __ldg(col + (size_t)(j + o) * Nx); }, j, Ny, v,
                                       dy, __ldg(col + (size_t)(Ny - 1) * Nx));
                rhs = -(u * dqdx + v * dqdy);
            }
        }
        const double upd = qc + dt * rhs;
        const double r = first ? upd : (c0 * F.q0[f][c] + c1 * upd);
        F.out[f][c] = mask_solid ? r * m : r;
    }
}

template <int SCHEME, int NQ>
__global__ void __launch_bounds__(256, 4)
k_euler_stage(const EulerFields F, const double *__restrict__ a, const double *__restrict__ b,
              const double *__restrict__ phi, int Ny, int Nx, double dx, double dy, double dt,
              double w_cut, double c0, double c1, int first, int mask_solid)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    euler_cell<SCHEME, NQ>(F, a, b, phi, j, i, Ny, Nx, dx, dy, dt, w_cut, c0, c1, first, mask_solid);
}

// ------------------------------------------------------- WENO5, active chunks only
// The reference's RHS is zero wherever phi > w_cut (functions.py:341 `continue`), which is ~2/3 of
// the grid in the multi-body cases, so the WENO5 path works on CHUNKS of 31 columns x 16 rows:
//   * k_weno_classify / k_weno_lists sort the chunks into NEAR (an active node within 3 nodes of
//     the chunk: some active node's stencil can read it) and FAR, once per call;
//   * stages 1 and 2 touch near chunks only; nothing ever reads their output in a far chunk;
//   * stage 3 computes near chunks and writes far chunks from q directly (a far node's three
//     stage values are pointwise functions of q: RHS == 0 in every stage) -- the same expression
//     tree, so the result is what the node-by-node kernel writes, bit for bit.
// A near chunk is walked by ONE WARP, top to bottom, lane l on column x0 - 1 + l (lane 0 is a
// helper column).  Both left-biased face values of a node are the same function
//     Fm(k) = weno_minus(q[k-2 .. k+2])          (functions.py:256-286)
// at k and k-1 (functions.py:352-366), so each Fm is evaluated ONCE: along y it is carried
// from the previous row in a register (the five-row window of q lives in registers, one new
// row loaded per step), along x it comes from the left lane by a shuffle.  That halves the fp64
// work (the IEEE divisions of the nonlinear weights dominate this kernel) without changing a
// bit of the result.  Chunks that touch the rim of the grid take the generic node-by-node code.
constexpr int CW = 31, CH = 16;
constexpr int WSR = CH + 6, WSC = CW + 6;                 // staged rows / columns of a chunk (halo 3)
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
struct WenoLists {
    int *summary;        // per chunk: classification bits (k_weno_classify)
    int *near_list;      // chunk ids, bit 30 set: no node with phi <= 0 in the chunk
    int *far_list;
    int *counters;       // [0] near, [1] far, [2..4] work tickets of the three stages
};

__global__ void __launch_bounds__(256)
k_weno_classify(const double *__restrict__ phi, WenoLists W, int Ny, int Nx, int ncx, int ncy, double w_cut)
{
    if (blockIdx.x == 0 && threadIdx.x < 8) W.counters[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int ch = warp; ch < ncx * ncy; ch += nwarp) {
        const int ty = ch / ncx, tx = ch - ty * ncx;
        const int x0 = tx * CW, y0 = ty * CH;
        const int i = x0 + lane;
        const bool colok = lane < CW && i < Nx;
        const int h = min(CH, Ny - y0), wdt = min(CW, Nx - x0);
        unsigned any = 0, top = 0, bot = 0, neg = 0;
        if (colok) {
            const bool ci = i >= 2 && i < Nx - 2;
            for (int r = 0; r < h; ++r) {
                const int j = y0 + r;
                const double ph = __ldg(phi + (size_t)j * Nx + i);
                const bool act = ci && j >= 2 && j < Ny - 2 && !(ph > w_cut);
                any |= act;
                top |= act && r < 3;
                bot |= act && r >= h - 3;
                neg |= (ph <= 0.0);
            }
        }
        const unsigned many = __ballot_sync(0xffffffffu, any), mtop = __ballot_sync(0xffffffffu, top),
                       mbot = __ballot_sync(0xffffffffu, bot), mneg = __ballot_sync(0xffffffffu, neg);
        const unsigned L3 = 7u, R3 = 7u << (wdt >= 3 ? wdt - 3 : 0);
        int bits = 0;
        bits |= (many != 0) << 0;
        bits |= (mtop != 0) << 1;
        bits |= (mbot != 0) << 2;
        bits |= ((many & L3) != 0) << 3;
        bits |= ((many & R3) != 0) << 4;
        bits |= ((mtop & L3) != 0) << 5;
        bits |= ((mtop & R3) != 0) << 6;
        bits |= ((mbot & L3) != 0) << 7;
        bits |= ((mbot & R3) != 0) << 8;
        bits |= (mneg != 0) << 9;
        if (lane == 0) W.summary[ch] = bits;
    }
}

__global__ void __launch_bounds__(256)
k_weno_lists(WenoLists W, int ncx, int ncy)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool valid = ch < ncx * ncy, near = false;
    int own = 0;
    if (valid) {
        const int ty = ch / ncx, tx = ch - ty * ncx;
        own = W.summary[ch];
        auto S = [&](int yy, int xx) { return (yy >= 0 && yy < ncy && xx >= 0 && xx < ncx) ? W.summary[yy * ncx + xx] : 0; };
        near = (own & 1) | ((S(ty - 1, tx) >> 2) & 1) | ((S(ty + 1, tx) >> 1) & 1) | ((S(ty, tx - 1) >> 4) & 1) |
               ((S(ty, tx + 1) >> 3) & 1) | ((S(ty - 1, tx - 1) >> 8) & 1) | ((S(ty - 1, tx + 1) >> 7) & 1) |
               ((S(ty + 1, tx - 1) >> 6) & 1) | ((S(ty + 1, tx + 1) >> 5) & 1);
    }
    const int entry = ch | (((own >> 9) & 1) ? 0 : (1 << 30));
    const unsigned mn = __ballot_sync(0xffffffffu, valid && near), mf = __ballot_sync(0xffffffffu, valid && !near);
    int bn = 0, bf = 0;
    if (lane == 0) {
        if (mn) bn = atomicAdd(&W.counters[0], __popc(mn));
        if (mf) bf = atomicAdd(&W.counters[1], __popc(mf));
    }
    bn = __shfl_sync(0xffffffffu, bn, 0);
    bf = __shfl_sync(0xffffffffu, bf, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (valid && near) W.near_list[bn + __popc(mn & lt)] = entry;
    if (valid && !near) W.far_list[bf + __popc(mf & lt)] = entry;
}

__device__ __forceinline__ bool weno_tame(double v) { return fabs(v) < 1e60; }

template <int NQ>
__device__ __forceinline__ void weno_chunk_fast(const EulerFields &F, const double *__restrict__ a,
                                                const double *__restrict__ b, const double *__restrict__ phi,
                                                int x0, int y0, int Nx, double dx, double dy, double dt,
                                                double w_cut, double c0, double c1, int first, int mask_solid,
                                                int lane, double *__restrict__ sm)
{
    const int i = x0 - 1 + lane;
    const bool outl = lane >= 1;                          // lane 0: helper column x0 - 1
    // stage the chunk's q (rows y0-3 .. y0+CH+2, columns x0-3 .. x0+CW+2) in this warp's shared memory
    // with asynchronous copies: every load of the chunk is in flight at once, and the walk below
    // then runs at shared-memory latency
    __syncwarp();
#pragma unroll
    for (int f = 0; f < NQ; ++f) {
        const double *src = F.qs[f] + (size_t)(y0 - 3) * Nx + (x0 - 3) + lane;
        double *dst = sm + f * (WSR * WSC) + lane;
        for (int r = 0; r < WSR; ++r, src += Nx, dst += WSC) {
            cp_async8(dst, src);
            if (lane < WSC - 32) cp_async8(dst + 32, src + 32);
        }
    }
    cp_async_wait_all();
    __syncwarp();
    const double *sq[NQ];
#pragma unroll
    for (int f = 0; f < NQ; ++f) sq[f] = sm + f * (WSR * WSC) + lane + 2;   // column i of staged row 0 (= y0-3)
    double w[NQ][5];                                      // q on rows j-2 .. j+2 of this column
    double fprev[NQ];                                     // Fm_y(j-1)
    int j = y0 - 1;
#pragma unroll
    for (int f = 0; f < NQ; ++f) {
        fprev[f] = 0.0;
#pragma unroll
        for (int r = 0; r < 5; ++r) w[f][r] = sq[f][r * WSC];
    }
    // scalars of the row below the current one are fetched one step ahead
    size_t cn = (size_t)y0 * Nx + i;
    double ph_n = __ldg(phi + cn), u_n = __ldg(a + cn), v_n = __ldg(b + cn);
    double ph_c = 0.0, u_c = 0.0, v_c = 0.0;
    bool act_c = false;
    for (; j < y0 + CH; ++j) {
        const bool more = j + 1 < y0 + CH;
        const bool act_n = more && outl && !(ph_n > w_cut);
        const bool needy_c = act_c && v_c >= 0.0, needy_n = act_n && v_n >= 0.0;
        double wn[NQ];
#pragma unroll
        for (int f = 0; f < NQ; ++f) wn[f] = sq[f][(j + 3 - (y0 - 3)) * WSC];
        const double ph_k = ph_n, u_k = u_n, v_k = v_n;   // row j+1, becomes current next step
        if (j + 2 < y0 + CH) {
            cn += Nx;
            ph_n = __ldg(phi + cn); u_n = __ldg(a + cn); v_n = __ldg(b + cn);
        }
        double fy[NQ];
        if (needy_c || needy_n) {
#pragma unroll
            for (int f = 0; f < NQ; ++f) fy[f] = weno_minus(w[f][0], w[f][1], w[f][2], w[f][3], w[f][4]);
        } else {
#pragma unroll
            for (int f = 0; f < NQ; ++f) fy[f] = 0.0;
        }
        if (j >= y0) {                                    // warp-uniform
            const size_t c = (size_t)j * Nx + i;
            double q0c[NQ];
            if (!first && outl) {
#pragma unroll
                for (int f = 0; f < NQ; ++f) q0c[f] = __ldg(F.q0[f] + c);
            }
            const bool needx = act_c && u_c >= 0.0;
            const bool needx_r = __shfl_down_sync(0xffffffffu, (int)needx, 1) && lane < 31;
            double xl[NQ][5];                             // q[j][i-2], [i-1], [i+1], [i+2], [i+3]
            if (act_c || needx_r) {
#pragma unroll
                for (int f = 0; f < NQ; ++f) {
                    const double *r = sq[f] + (j - (y0 - 3)) * WSC;
                    xl[f][0] = r[-2]; xl[f][1] = r[-1]; xl[f][2] = r[1]; xl[f][3] = r[2]; xl[f][4] = r[3];
                }
            }
            double fx[NQ];
            if (needx || needx_r) {
#pragma unroll
                for (int f = 0; f < NQ; ++f) fx[f] = weno_minus(xl[f][0], xl[f][1], w[f][2], xl[f][2], xl[f][3]);
            } else {
#pragma unroll
                for (int f = 0; f < NQ; ++f) fx[f] = 0.0;
            }
            const double m = (mask_solid && !(ph_c <= 0.0)) ? 0.0 : 1.0;
#pragma unroll
            for (int f = 0; f < NQ; ++f) {
                const double fxm = __shfl_up_sync(0xffffffffu, fx[f], 1);
                double rhs = 0.0;
                if (act_c) {
                    double dqdx, dqdy;
                    if (u_c >= 0.0) {
                        dqdx = (fx[f] - fxm) / dx;
                    } else if (weno_tame(xl[f][1]) && weno_tame(w[f][2]) && weno_tame(xl[f][2]) &&
                               weno_tame(xl[f][3]) && weno_tame(xl[f][4])) {
                        dqdx = 0.0;                      // both faces see the same five points (functions.py:367-376)
                    } else {
                        const double qp = weno_plus(xl[f][1], w[f][2], xl[f][2], xl[f][3], xl[f][4]);
                        dqdx = (qp - qp) / dx;
                    }
                    if (v_c >= 0.0) {
                        dqdy = (fy[f] - fprev[f]) / dy;
                    } else if (weno_tame(w[f][1]) && weno_tame(w[f][2]) && weno_tame(w[f][3]) &&
                               weno_tame(w[f][4]) && weno_tame(wn[f])) {
                        dqdy = 0.0;
                    } else {
                        const double qp = weno_plus(w[f][1], w[f][2], w[f][3], w[f][4], wn[f]);
                        dqdy = (qp - qp) / dy;
                    }
                    rhs = -(u_c * dqdx + v_c * dqdy);
                }
                if (outl) {
                    const double upd = w[f][2] + dt * rhs;
                    const double r = first ? upd : (c0 * q0c[f] + c1 * upd);
                    F.out[f][c] = mask_solid ? r * m : r;
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NQ; ++f) {
            w[f][0] = w[f][1]; w[f][1] = w[f][2]; w[f][2] = w[f][3]; w[f][3] = w[f][4]; w[f][4] = wn[f];
            fprev[f] = fy[f];
        }
        ph_c = ph_k; u_c = u_k; v_c = v_k; act_c = act_n;
    }
}

// stage: 0, 1, 2.  Near chunks are computed; far chunks are written by the last stage only.
template <int NQ>
__global__ void __launch_bounds__(256, 2)
k_weno_stage(const EulerFields F, const double *__restrict__ a, const double *__restrict__ b,
             const double *__restrict__ phi, WenoLists W, int stage, int Ny, int Nx, int ncx, double dx,
             double dy, double dt, double w_cut, double c0, double c1, int mask_solid)
{
    extern __shared__ double s_weno[];
    const int lane = threadIdx.x & 31;
    double *sm = s_weno + (threadIdx.x >> 5) * (NQ * WSR * WSC);
    const int nnear = W.counters[0], nfar = (stage == 2) ? W.counters[1] : 0;
    const int first = stage == 0;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&W.counters[2 + stage], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nnear + nfar) break;
        const bool far = t >= nnear;
        const int e = far ? W.far_list[t - nnear] : W.near_list[t];
        const int ch = e & ((1 << 30) - 1);
        const int ty = ch / ncx, tx = ch - ty * ncx;
        const int x0 = tx * CW, y0 = ty * CH;
        const int i = x0 - 1 + lane;
        if (far) {
            // RHS == 0 in all three stages: q1 = q + dt*0, q2 = 3/4 q + 1/4 (q1 + dt*0), q' = 1/3 q + 2/3 (q2 + dt*0)
            const bool anyneg = !(e >> 30);
            if (lane >= 1 && i < Nx) {
                const int j1 = min(y0 + CH, Ny);
                for (int j = y0; j < j1; ++j) {
                    const size_t c = (size_t)j * Nx + i;
                    double m = 1.0;
                    if (mask_solid) m = (anyneg && __ldg(phi + c) <= 0.0) ? 1.0 : 0.0;
#pragma unroll
                    for (int f = 0; f < NQ; ++f) {
                        const double q = __ldg(F.q0[f] + c);
                        const double q1 = q + dt * 0.0;
                        const double q2 = 0.75 * q + 0.25 * (q1 + dt * 0.0);
                        const double r = c0 * q + c1 * (q2 + dt * 0.0);
                        F.out[f][c] = mask_solid ? r * m : r;
                    }
                }
            }
            continue;
        }
        const bool fast = x0 >= 3 && x0 + CW - 1 + 3 < Nx && y0 >= 3 && y0 + CH - 1 + 3 < Ny;
        if (fast) {
            weno_chunk_fast<NQ>(F, a, b, phi, x0, y0, Nx, dx, dy, dt, w_cut, c0, c1, first, mask_solid, lane, sm);
        } else if (lane >= 1 && i < Nx) {
            const int j1 = min(y0 + CH, Ny);
            for (int j = y0; j < j1; ++j)
                euler_cell<1, NQ>(F, a, b, phi, j, i, Ny, Nx, dx, dy, dt, w_cut, c0, c1, first, mask_solid);
        }
    }
}

// test probe: mode 0: a / 6 through div6; mode 1: a / b through div_by_recip with y = 1 / b
__global__ void k_weno_div_probe(const double *__restrict__ a, const double *__restrict__ b,
                                 double *__restrict__ out, long n, int mode)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        out[k] = mode == 0 ? div_by_recip(a[k], 6.0, 1.0 / 6.0) : div_by_recip(a[k], b[k], 1.0 / b[k]);
}

// standalone RHS (API parity for _weno5_rhs/_central2_rhs/_conservative_rhs)
template <int SCHEME>
__global__ void k_euler_rhs(const double *__restrict__ qs, const double *__restrict__ a,
                            const double *__restrict__ b, const double *__restrict__ phi,
                            double *__restrict__ out, int Ny, int Nx, double dx, double dy, double w_cut)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    size_t c = (size_t)j * Nx + i;
    const int halo = (SCHEME == 1) ? 2 : 1;
    double rhs = 0.0;
    bool interior = (i >= halo && i < Nx - halo && j >= halo && j < Ny - halo);
    if (interior && !(phi[c] > w_cut)) {
        if (SCHEME == 0) {
            double dqdx = (__ldg(qs + c + 1) - __ldg(qs + c - 1)) * (0.5 / dx);
            double dqdy = (__ldg(qs + c + Nx) - __ldg(qs + c - Nx)) * (0.5 / dy);
            rhs = -(a[c] * dqdx + b[c] * dqdy);
        } else if (SCHEME == 2) {
            double fx = (__ldg(a + c + 1) * __ldg(qs + c + 1) - __ldg(a + c - 1) * __ldg(qs + c - 1)) * (0.5 / dx);
            double fy = (__ldg(b + c + Nx) * __ldg(qs + c + Nx) - __ldg(b + c - Nx) * __ldg(qs + c - Nx)) * (0.5 / dy);
            rhs = -(fx + fy);
        } else {
            double u = a[c], v = b[c];
            const double *row = qs + (size_t)j * Nx;
            const double *col = qs + i;
            double dqdx = weno5_dq([&](int o) { return __ldg(row + i + o); }, i, Nx, u, dx,
                                   __ldg(row + Nx - 1));
            double dqdy = weno5_dq([&](int o) { return __ldg(col + (size_t)(j + o) * Nx); }, j, Ny, v,
                                   dy, __ldg(col + (size_t)(Ny - 1) * Nx));
            rhs = -(u * dqdx + v * dqdy);
        }
    }
    out[c] = rhs;
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

// per-device chunk lists of the WENO5 path (grown on demand; one stream per device at a time, like
// the other workspaces of this library)
struct WenoWs { int *buf; size_t cap; };
static WenoWs g_weno_ws[64];

static int weno_lists(int Ny, int Nx, WenoLists *W, int *ncx_out, int *ncy_out)
{
    int dev = 0;
    RMT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return RMT_EINVAL;
    const int ncx = rmt_cdiv(Nx, CW), ncy = rmt_cdiv(Ny, CH);
    const size_t nch = (size_t)ncx * ncy, need = 3 * nch + 64;
    WenoWs &ws = g_weno_ws[dev];
    if (ws.cap < need) {
        if (ws.buf) cudaFree(ws.buf);
        ws.buf = nullptr; ws.cap = 0;
        RMT_CUDA(cudaMalloc((void **)&ws.buf, need * sizeof(int)));
        ws.cap = need;
    }
    W->counters = ws.buf;
    W->summary = ws.buf + 64;
    W->near_list = W->summary + nch;
    W->far_list = W->near_list + nch;
    *ncx_out = ncx; *ncy_out = ncy;
    return RMT_OK;
}

template <int NQ>
static int weno_rk3(const double *const *q, const double *a, const double *b, const double *phi,
                    double *const *out, double *const *w1, double *const *w2, int Ny, int Nx, double dx,
                    double dy, double dt, double w_cut, int mask_solid, cudaStream_t s)
{
    WenoLists W;
    int ncx, ncy;
    int rc = weno_lists(Ny, Nx, &W, &ncx, &ncy);
    if (rc != RMT_OK) return rc;
    const int nch = ncx * ncy;
    EulerFields F1{}, F2{}, F3{};
    for (int f = 0; f < NQ; ++f) {
        F1.q0[f] = q[f]; F1.qs[f] = q[f];  F1.out[f] = w1[f];
        F2.q0[f] = q[f]; F2.qs[f] = w1[f]; F2.out[f] = w2[f];
        F3.q0[f] = q[f]; F3.qs[f] = w2[f]; F3.out[f] = out[f];
    }
    const int cblocks = min(rmt_cdiv(nch, 8), 148 * 8);
    k_weno_classify<<<cblocks, 256, 0, s>>>(phi, W, Ny, Nx, ncx, ncy, w_cut);
    RMT_LAUNCH_CHECK();
    k_weno_lists<<<rmt_cdiv(nch, 256), 256, 0, s>>>(W, ncx, ncy);
    RMT_LAUNCH_CHECK();
    const int sblocks = min(rmt_cdiv(nch, 8), 148 * 2);
    const size_t shm = (size_t)8 * NQ * WSR * WSC * sizeof(double);
    RMT_CUDA(cudaFuncSetAttribute(k_weno_stage<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    k_weno_stage<NQ><<<sblocks, 256, shm, s>>>(F1, a, b, phi, W, 0, Ny, Nx, ncx, dx, dy, dt, w_cut, 0.0, 1.0, 0);
    RMT_LAUNCH_CHECK();
    k_weno_stage<NQ><<<sblocks, 256, shm, s>>>(F2, a, b, phi, W, 1, Ny, Nx, ncx, dx, dy, dt, w_cut, 0.75, 0.25, 0);
    RMT_LAUNCH_CHECK();
    k_weno_stage<NQ><<<sblocks, 256, shm, s>>>(F3, a, b, phi, W, 2, Ny, Nx, ncx, dx, dy, dt, w_cut, 1.0 / 3.0,
                                             2.0 / 3.0, mask_solid);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

template <int SCHEME, int NQ>
static int euler_rk3(const double *const *q, const double *a, const double *b, const double *phi,
                     double *const *out, double *const *w1, double *const *w2, int Ny, int Nx, double dx,
                     double dy, double dt, double w_cut, int mask_solid, cudaStream_t s)
{
    if (SCHEME == 1 && Nx >= 2 * CW && Ny >= 2 * CH)
        return weno_rk3<NQ>(q, a, b, phi, out, w1, w2, Ny, Nx, dx, dy, dt, w_cut, mask_solid, s);
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    EulerFields F1{}, F2{}, F3{};
    for (int f = 0; f < NQ; ++f) {
        F1.q0[f] = q[f]; F1.qs[f] = q[f];  F1.out[f] = w1[f];
        F2.q0[f] = q[f]; F2.qs[f] = w1[f]; F2.out[f] = w2[f];
        F3.q0[f] = q[f]; F3.qs[f] = w2[f]; F3.out[f] = out[f];
    }
    k_euler_stage<SCHEME, NQ><<<grd, blk, 0, s>>>(F1, a, b, phi, Ny, Nx, dx, dy, dt, w_cut, 0.0, 1.0, 1, 0);
    RMT_LAUNCH_CHECK();
    k_euler_stage<SCHEME, NQ><<<grd, blk, 0, s>>>(F2, a, b, phi, Ny, Nx, dx, dy, dt, w_cut, 0.75, 0.25, 0, 0);
    RMT_LAUNCH_CHECK();
    k_euler_stage<SCHEME, NQ><<<grd, blk, 0, s>>>(F3, a, b, phi, Ny, Nx, dx, dy, dt, w_cut, 1.0 / 3.0, 2.0 / 3.0,
                                                 0, mask_solid);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // namespace

extern "C" {

int rmt_sample(const double *u, const double *xq, const double *yq, double *out, long nq, double dx,
               double dy, int Nx, int Ny, int cubic, void *stream)
{
    if (!u || !xq || !yq || !out || nq <= 0 || Nx < 2 || Ny < 2) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (cubic) k_sample<true><<<flat_blocks(nq), 256, 0, s>>>(u, xq, yq, out, nq, dx, dy, Nx, Ny);
    else k_sample<false><<<flat_blocks(nq), 256, 0, s>>>(u, xq, yq, out, nq, dx, dy, Nx, Ny);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_advect_sl_rk4(const double *q0, const double *q1, const double *a, const double *b,
                      const double *X, const double *Y, double *out0, double *out1, int Ny, int Nx,
                      double dt, double dx, double dy, int cubic, void *stream)
{
    return rmt_advect_sl_rk4_rows(q0, q1, a, b, X, Y, out0, out1, Ny, Nx, Ny, 0, dt, dx, dy, cubic, stream);
}

int rmt_advect_sl_rk4_rows(const double *q0, const double *q1, const double *a, const double *b,
                           const double *X, const double *Y, double *out0, double *out1, int Ny_local, int Nx,
                           int Ny, int row_offset, double dt, double dx, double dy, int cubic, void *stream)
{
    if (!q0 || !a || !b || !X || !Y || !out0 || Nx < 2 || Ny < 2 || Ny_local < 1 || row_offset < 0 ||
        row_offset + Ny_local > Ny)
        return RMT_EINVAL;
    if ((q1 == nullptr) != (out1 == nullptr)) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int Nyl = Ny_local, jo = row_offset;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Nyl, TY));
    if (q1) {
        if (cubic) k_advect_sl<true, 2><<<grd, blk, 0, s>>>(q0, q1, a, b, X, Y, out0, out1, Ny, Nx, dt, dx, dy, Nyl, jo);
        else k_advect_sl<false, 2><<<grd, blk, 0, s>>>(q0, q1, a, b, X, Y, out0, out1, Ny, Nx, dt, dx, dy, Nyl, jo);
    } else {
        if (cubic) k_advect_sl<true, 1><<<grd, blk, 0, s>>>(q0, q1, a, b, X, Y, out0, out1, Ny, Nx, dt, dx, dy, Nyl, jo);
        else k_advect_sl<false, 1><<<grd, blk, 0, s>>>(q0, q1, a, b, X, Y, out0, out1, Ny, Nx, dt, dx, dy, Nyl, jo);
    }
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_advect_euler_rk3(const double *q, const double *a, const double *b, const double *phi,
                         double *out, double *work1, double *work2, int Ny, int Nx, double dx,
                         double dy, double dt, double w_cut, int scheme, void *stream)
{
    if (!q || !a || !b || !phi || !out || !work1 || !work2 || Nx < 5 || Ny < 5) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const double *qq[1] = {q};
    double *oo[1] = {out}, *w1[1] = {work1}, *w2[1] = {work2};
    switch (scheme) {
    case 0: return euler_rk3<0, 1>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, 0, s);
    case 1: return euler_rk3<1, 1>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, 0, s);
    case 2: return euler_rk3<2, 1>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, 0, s);
    }
    return RMT_EINVAL;
}

int rmt_advect_euler_rk3_pair(const double *q0, const double *q1, const double *a, const double *b,
                              const double *phi, double *out0, double *out1, double *work /* 4 fields */,
                              int Ny, int Nx, double dx, double dy, double dt, double w_cut, int scheme,
                              int mask_solid, void *stream)
{
    if (!q0 || !q1 || !a || !b || !phi || !out0 || !out1 || !work || Nx < 5 || Ny < 5) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)Ny * Nx;
    const double *qq[2] = {q0, q1};
    double *oo[2] = {out0, out1}, *w1[2] = {work, work + n}, *w2[2] = {work + 2 * n, work + 3 * n};
    switch (scheme) {
    case 0: return euler_rk3<0, 2>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, mask_solid, s);
    case 1: return euler_rk3<1, 2>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, mask_solid, s);
    case 2: return euler_rk3<2, 2>(qq, a, b, phi, oo, w1, w2, Ny, Nx, dx, dy, dt, w_cut, mask_solid, s);
    }
    return RMT_EINVAL;
}

int rmt_weno_div_probe(const double *a, const double *b, double *out, long n, int mode, void *stream)
{
    if (!a || !out || n <= 0 || (mode != 0 && mode != 1) || (mode == 1 && !b)) return RMT_EINVAL;
    k_weno_div_probe<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, mode);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_euler_rhs(const double *q, const double *a, const double *b, const double *phi, double *out,
                  int Ny, int Nx, double dx, double dy, double w_cut, int scheme, void *stream)
{
    if (!q || !a || !b || !phi || !out || Nx < 5 || Ny < 5) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    switch (scheme) {
    case 0: k_euler_rhs<0><<<grd, blk, 0, s>>>(q, a, b, phi, out, Ny, Nx, dx, dy, w_cut); break;
    case 1: k_euler_rhs<1><<<grd, blk, 0, s>>>(q, a, b, phi, out, Ny, Nx, dx, dy, w_cut); break;
    case 2: k_euler_rhs<2><<<grd, blk, 0, s>>>(q, a, b, phi, out, Ny, Nx, dx, dy, w_cut); break;
    default: return RMT_EINVAL;
    }
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

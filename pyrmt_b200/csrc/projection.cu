// projection.cu -- the non-spectral half of the pressure projection.
//
// Replaces (reference file:line):
//   pyRMT/functions.py:1005-1014   _compute_divergence
//   pyRMT/functions.py:1016-1071   _compute_divergence_rc (constant-density branch)
//   pyRMT/functions.py:1073-1089   _compute_pressure_gradient
//   pyRMT/functions.py:1236-1252   periodic divergence / gradient on the reduced grid
//   pyRMT/functions.py:1284-1290,:1331,:1350-1362   rhs scaling, velocity correction,
//                                                   incremental pressure update
//
// Two fused kernels bracket the spectral solve: the front end turns (a*, b*,
// p_prev, rho) straight into the Poisson right-hand side (the Rhie-Chow face
// velocities, both pressure-gradient fields and the divergence never reach HBM),
// and the back end turns the Poisson solution into (a, b, p) in one pass.
#include "common.cuh"
#include "../../include/rmt_b200.h"

using namespace rmt;

namespace {

struct GField {
    const double *p;
    int Nx;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        return __ldg(p + (size_t)j * Nx + i);
    }
};

// plain central divergence at an interior node (functions.py:1010-1013)
__device__ __forceinline__ double div_plain(const GField &A, const GField &B, int j, int i,
                                            double i2dx, double i2dy)
{
    return (A(j, i + 1) - A(j, i - 1)) * i2dx + (B(j + 1, i) - B(j - 1, i)) * i2dy;
}

// Rhie-Chow divergence at an interior node (functions.py:1029-1069)
__device__ __forceinline__ double div_rc(const GField &A, const GField &B, const GField &P, int j,
                                         int i, int Ny, int Nx, double invdx, double invdy,
                                         double d_f)
{
    const double i2dx = 0.5 * invdx, i2dy = 0.5 * invdy;
    // x faces i-1/2 and i+1/2
    double gxm = ddx2(P, j, i - 1, Nx, i2dx), gx0 = ddx2(P, j, i, Nx, i2dx), gxp = ddx2(P, j, i + 1, Nx, i2dx);
    double pm = P(j, i - 1), p0 = P(j, i), pp = P(j, i + 1);
    double am = A(j, i - 1), a0 = A(j, i), ap = A(j, i + 1);
    double uf_e = 0.5 * (a0 + ap) - d_f * ((pp - p0) * invdx - 0.5 * (gx0 + gxp));
    double uf_w = 0.5 * (am + a0) - d_f * ((p0 - pm) * invdx - 0.5 * (gxm + gx0));
    // y faces j-1/2 and j+1/2
    double gym = ddy2(P, j - 1, i, Ny, i2dy), gy0 = ddy2(P, j, i, Ny, i2dy), gyp = ddy2(P, j + 1, i, Ny, i2dy);
    double qm = P(j - 1, i), qp = P(j + 1, i);
    double bm = B(j - 1, i), b0 = B(j, i), bp = B(j + 1, i);
    double vf_n = 0.5 * (b0 + bp) - d_f * ((qp - p0) * invdy - 0.5 * (gy0 + gyp));
    double vf_s = 0.5 * (bm + b0) - d_f * ((p0 - qm) * invdy - 0.5 * (gym + gy0));
    return (uf_e - uf_w) * invdx + (vf_n - vf_s) * invdy;
}

// Rhie-Chow divergence where every stencil point (i-2..i+2, j-2..j+2) is an interior node: the same
// expressions with the central-difference branch of ddx2 / ddy2 taken unconditionally
__device__ __forceinline__ double div_rc_interior(const GField &A, const GField &B, const GField &P, int j,
                                                  int i, double invdx, double invdy, double d_f)
{
    const double i2dx = 0.5 * invdx, i2dy = 0.5 * invdy;
    const double pm2 = P(j, i - 2), pm = P(j, i - 1), p0 = P(j, i), pp = P(j, i + 1), pp2 = P(j, i + 2);
    const double gxm = (p0 - pm2) * i2dx, gx0 = (pp - pm) * i2dx, gxp = (pp2 - p0) * i2dx;
    const double am = A(j, i - 1), a0 = A(j, i), ap = A(j, i + 1);
    const double uf_e = 0.5 * (a0 + ap) - d_f * ((pp - p0) * invdx - 0.5 * (gx0 + gxp));
    const double uf_w = 0.5 * (am + a0) - d_f * ((p0 - pm) * invdx - 0.5 * (gxm + gx0));
    const double qm2 = P(j - 2, i), qm = P(j - 1, i), qp = P(j + 1, i), qp2 = P(j + 2, i);
    const double gym = (p0 - qm2) * i2dy, gy0 = (qp - qm) * i2dy, gyp = (qp2 - p0) * i2dy;
    const double bm = B(j - 1, i), b0 = B(j, i), bp = B(j + 1, i);
    const double vf_n = 0.5 * (b0 + bp) - d_f * ((qp - p0) * invdy - 0.5 * (gy0 + gyp));
    const double vf_s = 0.5 * (bm + b0) - d_f * ((p0 - qm) * invdy - 0.5 * (gym + gy0));
    return (uf_e - uf_w) * invdx + (vf_n - vf_s) * invdy;
}

// periodic wide-central divergence at a reduced-grid node (functions.py:1236-1243)
__device__ __forceinline__ double div_periodic(const GField &A, const GField &B, int j, int i, int my,
                                               int mx, double i2dx, double i2dy)
{
    int ip = (i + 1 == mx) ? 0 : i + 1, im = (i == 0) ? mx - 1 : i - 1;
    int jp = (j + 1 == my) ? 0 : j + 1, jm = (j == 0) ? my - 1 : j - 1;
    return (A(j, ip) - A(j, im)) * i2dx + (B(jp, i) - B(jm, i)) * i2dy;
}

// mode 0 plain, 1 Rhie-Chow, 2 periodic, 3 periodic in x with the y neighbours taken from the rows
// above / below as stored (a row slab whose wrap rows were exchanged; rows 0 and Ny-1 are halo rows).
// scale: 0 none (API divergence),
// 1 rho[c]*div/dt, 2 rho_scalar*div/dt, 3 (rho_sum/n)*div/dt.
__global__ void __launch_bounds__(256)
k_divergence(const double *__restrict__ a, const double *__restrict__ b,
             const double *__restrict__ p_prev, const double *__restrict__ rho, double rho_scalar,
             const double *__restrict__ rho_sum, double *__restrict__ out, int Ny, int Nx, double dx,
             double dy, double dt, int mode, int scale)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const size_t c = (size_t)j * Nx + i;
    const GField A{a, Nx}, B{b, Nx};
    double d = 0.0;
    if (mode == 2) {
        int jr = (j == Ny - 1) ? 0 : j, ir = (i == Nx - 1) ? 0 : i;   // _tile_overlap :1205-1213
        d = div_periodic(A, B, jr, ir, Ny - 1, Nx - 1, 0.5 / dx, 0.5 / dy);
    } else if (mode == 3) {
        if (j >= 1 && j < Ny - 1) {
            const int mx = Nx - 1, ir = (i == mx) ? 0 : i;
            const int ip = (ir + 1 == mx) ? 0 : ir + 1, im = (ir == 0) ? mx - 1 : ir - 1;
            d = (A(j, ip) - A(j, im)) * (0.5 / dx) + (B(j + 1, ir) - B(j - 1, ir)) * (0.5 / dy);
        }
    } else if (i >= 1 && i < Nx - 1 && j >= 1 && j < Ny - 1) {
        if (mode == 1) {
            const GField P{p_prev, Nx};
            double d_f = dt / (rho_sum[0] / ((double)Ny * (double)Nx));
            if (i >= 2 && i < Nx - 2 && j >= 2 && j < Ny - 2) d = div_rc_interior(A, B, P, j, i, 1.0 / dx, 1.0 / dy, d_f);
            else d = div_rc(A, B, P, j, i, Ny, Nx, 1.0 / dx, 1.0 / dy, d_f);
        } else {
            d = div_plain(A, B, j, i, 0.5 / dx, 0.5 / dy);
        }
    }
    if (scale == 1) d = rho[c] * d / dt;
    else if (scale == 2) d = rho_scalar * d / dt;
    else if (scale == 3) d = (rho_sum[0] / ((double)Ny * (double)Nx)) * d / dt;
    out[c] = d;
}

struct Shifted {   // p_correction = sol - mean(sol), formed on the fly
    const double *p;
    int Nx;
    double shift;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        return __ldg(p + (size_t)j * Nx + i) - shift;
    }
};

template <class F>
__device__ __forceinline__ void pgrad_neumann(const F &P, int j, int i, int Ny, int Nx, double i2dx,
                                              double i2dy, double &gx, double &gy)
{
    const bool xin = (i >= 1 && i < Nx - 1), yin = (j >= 1 && j < Ny - 1);
    gx = (xin && !yin) ? 0.0 : ddx2(P, j, i, Nx, i2dx);   // :1078,:1082-1083
    gy = (yin && !xin) ? 0.0 : ddy2(P, j, i, Ny, i2dy);   // :1079,:1086-1087
}

template <class F>
__device__ __forceinline__ void pgrad_periodic(const F &P, int j, int i, int Ny, int Nx, double i2dx,
                                               double i2dy, double &gx, double &gy)
{
    const int my = Ny - 1, mx = Nx - 1;
    int jr = (j == my) ? 0 : j, ir = (i == mx) ? 0 : i;
    int ip = (ir + 1 == mx) ? 0 : ir + 1, im = (ir == 0) ? mx - 1 : ir - 1;
    int jp = (jr + 1 == my) ? 0 : jr + 1, jm = (jr == 0) ? my - 1 : jr - 1;
    gx = (P(jr, ip) - P(jr, im)) * i2dx;
    gy = (P(jp, ir) - P(jm, ir)) * i2dy;
}

// periodic in x, stored rows in y (slab with exchanged wrap rows; rows 0 and Ny-1 are halo rows)
template <class F>
__device__ __forceinline__ void pgrad_slab(const F &P, int j, int i, int Ny, int Nx, double i2dx,
                                           double i2dy, double &gx, double &gy)
{
    const int mx = Nx - 1, ir = (i == mx) ? 0 : i;
    const int ip = (ir + 1 == mx) ? 0 : ir + 1, im = (ir == 0) ? mx - 1 : ir - 1;
    const int jp = min(j + 1, Ny - 1), jm = max(j - 1, 0);
    gx = (P(j, ip) - P(j, im)) * i2dx;
    gy = (P(jp, ir) - P(jm, ir)) * i2dy;
}

__global__ void __launch_bounds__(256)
k_pressure_gradient(const double *__restrict__ p, double *__restrict__ gx, double *__restrict__ gy,
                    int Ny, int Nx, double dx, double dy, int periodic)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const GField P{p, Nx};
    double x, y;
    if (periodic) pgrad_periodic(P, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, x, y);
    else pgrad_neumann(P, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, x, y);
    const size_t c = (size_t)j * Nx + i;
    gx[c] = x;
    gy[c] = y;
}

__global__ void __launch_bounds__(256)
k_projection_correct(const double *__restrict__ sol, const double *__restrict__ sol_sum,
                     const double *__restrict__ a_star, const double *__restrict__ b_star,
                     const double *__restrict__ rho, double rho_scalar,
                     const double *__restrict__ p_prev, double *__restrict__ a,
                     double *__restrict__ b, double *__restrict__ p, int Ny, int Nx, double dx,
                     double dy, double dt, int periodic)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const size_t c = (size_t)j * Nx + i;
    const Shifted PC{sol, Nx, sol_sum ? sol_sum[0] / ((double)Ny * (double)Nx) : 0.0};
    double gx, gy;
    if (periodic == 2) pgrad_slab(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx, gy);
    else if (periodic) pgrad_periodic(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx, gy);
    else pgrad_neumann(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx, gy);
    const double r = rho ? rho[c] : rho_scalar;
    const double dtr = dt / r;
    a[c] = a_star[c] - dtr * gx;
    b[c] = b_star[c] - dtr * gy;
    const double pc = PC(j, i);
    p[c] = p_prev ? (p_prev[c] + pc) : pc;
}

// The same back end with mean(p) removed in the same pass: p = p_prev + pc - p_prev_sum/N (mean(pc) is zero
// to rounding, so mean(p_prev) IS mean(p_prev + pc) to rounding), and the block sums of the p written here
// go to `partial` so that the next step gets its sum(p_prev) without reading p again.
__global__ void __launch_bounds__(256)
k_projection_correct_centered(const double *__restrict__ sol, const double *__restrict__ sol_sum,
                              const double *__restrict__ a_star, const double *__restrict__ b_star,
                              const double *__restrict__ rho, double rho_scalar,
                              const double *__restrict__ p_prev, const double *__restrict__ p_prev_sum,
                              double *__restrict__ a, double *__restrict__ b, double *__restrict__ p,
                              double *__restrict__ partial, int Ny, int Nx, double dx, double dy, double dt,
                              int periodic)
{
    // 32 x 16 nodes per CTA, two rows per thread: the loads of both nodes are in flight together
    __shared__ double red[TY];
    const int i = blockIdx.x * TX + threadIdx.x;
    const double n = (double)Ny * (double)Nx;
    const Shifted PC{sol, Nx, sol_sum ? sol_sum[0] / n : 0.0};
    const double m = p_prev_sum ? p_prev_sum[0] / n : 0.0;
    double gx[2], gy[2], as[2], bs[2], rr[2], pp[2], pc[2];
    bool in[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = blockIdx.y * (2 * TY) + threadIdx.y + TY * r;
        in[r] = i < Nx && j < Ny;
        gx[r] = gy[r] = as[r] = bs[r] = pp[r] = pc[r] = 0.0;
        rr[r] = 1.0;
        if (in[r]) {
            const size_t c = (size_t)j * Nx + i;
            if (periodic == 2) pgrad_slab(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx[r], gy[r]);
            else if (periodic) pgrad_periodic(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx[r], gy[r]);
            else pgrad_neumann(PC, j, i, Ny, Nx, 0.5 / dx, 0.5 / dy, gx[r], gy[r]);
            rr[r] = rho ? __ldg(rho + c) : rho_scalar;
            as[r] = __ldg(a_star + c);
            bs[r] = __ldg(b_star + c);
            pc[r] = PC(j, i);
            pp[r] = p_prev ? __ldg(p_prev + c) : 0.0;
        }
    }
    double pv = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!in[r]) continue;
        const int j = blockIdx.y * (2 * TY) + threadIdx.y + TY * r;
        const size_t c = (size_t)j * Nx + i;
        const double dtr = dt / rr[r];
        a[c] = as[r] - dtr * gx[r];
        b[c] = bs[r] - dtr * gy[r];
        const double v = (p_prev ? (pp[r] + pc[r]) : pc[r]) - m;
        p[c] = v;
        pv += v;
    }
    // deterministic block sum (warp tree, then the TY warp sums in order)
    pv = warp_sum(pv);
    if (threadIdx.x == 0) red[threadIdx.y] = pv;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double s = 0.0;
        for (int w = 0; w < TY; ++w) s += red[w];
        partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void k_sum_partials(const double *__restrict__ part, long n, double *__restrict__ out)
{
    __shared__ double red[32];
    double s = 0.0;
    for (long k = threadIdx.x; k < n; k += blockDim.x) s += part[k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void k_subtract_mean(double *__restrict__ x, const double *__restrict__ sum, long n)
{
    const double m = sum[0] / (double)n;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        x[k] -= m;
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

int launch_div(const double *a, const double *b, const double *p_prev, const double *rho,
               double rho_scalar, const double *rho_sum, double *out, int Ny, int Nx, double dx,
               double dy, double dt, int mode, int scale, void *stream)
{
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_divergence<<<grd, blk, 0, (cudaStream_t)stream>>>(a, b, p_prev, rho, rho_scalar, rho_sum, out, Ny,
                                                       Nx, dx, dy, dt, mode, scale);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // namespace

extern "C" {

int rmt_divergence(const double *a, const double *b, double *div, int Ny, int Nx, double dx,
                   double dy, void *stream)
{
    if (!a || !b || !div || Ny < 3 || Nx < 3) return RMT_EINVAL;
    return launch_div(a, b, nullptr, nullptr, 0.0, nullptr, div, Ny, Nx, dx, dy, 1.0, 0, 0, stream);
}

int rmt_divergence_rc(const double *a, const double *b, const double *p_prev, const double *rho_sum,
                      double *div, int Ny, int Nx, double dx, double dy, double dt, void *stream)
{
    if (!a || !b || !p_prev || !rho_sum || !div || Ny < 4 || Nx < 4) return RMT_EINVAL;
    return launch_div(a, b, p_prev, nullptr, 0.0, rho_sum, div, Ny, Nx, dx, dy, dt, 1, 0, stream);
}

int rmt_divergence_periodic(const double *a, const double *b, double *div, int Ny, int Nx, double dx,
                            double dy, void *stream)
{
    if (!a || !b || !div || Ny < 4 || Nx < 4) return RMT_EINVAL;
    return launch_div(a, b, nullptr, nullptr, 0.0, nullptr, div, Ny, Nx, dx, dy, 1.0, 2, 0, stream);
}

int rmt_pressure_gradient(const double *p, double *gx, double *gy, int Ny, int Nx, double dx,
                          double dy, void *stream)
{
    if (!p || !gx || !gy || Ny < 3 || Nx < 3) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_pressure_gradient<<<grd, blk, 0, (cudaStream_t)stream>>>(p, gx, gy, Ny, Nx, dx, dy, 0);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_pressure_gradient_periodic(const double *p, double *gx, double *gy, int Ny, int Nx, double dx,
                                   double dy, void *stream)
{
    if (!p || !gx || !gy || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_pressure_gradient<<<grd, blk, 0, (cudaStream_t)stream>>>(p, gx, gy, Ny, Nx, dx, dy, 1);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_projection_rhs(const double *a, const double *b, const double *p_prev, const double *rho,
                       double rho_scalar, const double *rho_sum, double *rhs, int Ny, int Nx,
                       double dx, double dy, double dt, int periodic, void *stream)
{
    if (!a || !b || !rhs || Ny < 4 || Nx < 4) return RMT_EINVAL;
    if (periodic) {
        if (!rho_sum) return RMT_EINVAL;
        return launch_div(a, b, nullptr, nullptr, 0.0, rho_sum, rhs, Ny, Nx, dx, dy, dt, periodic == 2 ? 3 : 2, 3,
                          stream);
    }
    if (p_prev && !rho_sum) return RMT_EINVAL;
    return launch_div(a, b, p_prev, rho, rho_scalar, rho_sum, rhs, Ny, Nx, dx, dy, dt, p_prev ? 1 : 0,
                      rho ? 1 : 2, stream);
}

int rmt_projection_correct(const double *sol, const double *sol_sum, const double *a_star,
                           const double *b_star, const double *rho, double rho_scalar,
                           const double *p_prev, double *a, double *b, double *p, int Ny, int Nx,
                           double dx, double dy, double dt, int periodic, void *stream)
{
    if (!sol || !a_star || !b_star || !a || !b || !p || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_projection_correct<<<grd, blk, 0, (cudaStream_t)stream>>>(sol, sol_sum, a_star, b_star, rho,
                                                               rho_scalar, p_prev, a, b, p, Ny, Nx, dx,
                                                               dy, dt, periodic);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

long rmt_projection_partials(int Ny, int Nx) { return (long)rmt_cdiv(Nx, TX) * rmt_cdiv(Ny, 2 * TY); }

int rmt_projection_correct_centered(const double *sol, const double *sol_sum, const double *a_star,
                                    const double *b_star, const double *rho, double rho_scalar,
                                    const double *p_prev, const double *p_prev_sum, double *a, double *b,
                                    double *p, double *partial, double *p_sum_out, int Ny, int Nx, double dx,
                                    double dy, double dt, int periodic, void *stream)
{
    if (!sol || !a_star || !b_star || !a || !b || !p || !partial || !p_sum_out || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, 2 * TY));
    cudaStream_t s = (cudaStream_t)stream;
    k_projection_correct_centered<<<grd, blk, 0, s>>>(sol, sol_sum, a_star, b_star, rho, rho_scalar, p_prev,
                                                     p_prev_sum, a, b, p, partial, Ny, Nx, dx, dy, dt, periodic);
    RMT_LAUNCH_CHECK();
    k_sum_partials<<<1, 1024, 0, s>>>(partial, (long)grd.x * grd.y, p_sum_out);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_subtract_mean(double *x, const double *sum, long n, void *stream)
{
    if (!x || !sum || n <= 0) return RMT_EINVAL;
    k_subtract_mean<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, sum, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

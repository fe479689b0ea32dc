// stencil_ops.cu -- pointwise / small-stencil operators, reductions, the
// boundary-condition table kernel and the disc level-set kernel.
//
// Replaces (reference file:line):
//   pyRMT/utils.py:4-131            grad_central_{x,y}_{2nd,4th}, diff_upwind_3rd, lap_2nd
//   pyRMT/functions.py:660-671      smoothed_heaviside
//   pyRMT/functions.py:165-192      compute_timestep (the max|u| reduction)
//   pyRMT/functions.py:524          isfinite guard of advect_reference_map
//   benchmarks/common.py:27-57      BC callables (as a gather table) and the disc SDF
//   benchmarks/soft_disc_in_lid_driven.py:88-91,102-103   script-level glue
#include "common.cuh"
#include "../../include/rmt_b200.h"

using namespace rmt;

namespace {

struct Field {
    const double *p;
    int Nx;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        return __ldg(p + (size_t)j * Nx + i);
    }
};

// op: 0 grad_x_2nd, 1 grad_y_2nd, 2 grad_x_4th, 3 grad_y_4th, 4 lap_2nd
__global__ void k_stencil(const double *__restrict__ f, double *__restrict__ out, int Ny, int Nx,
                          double hx, double hy, int op)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    Field F{f, Nx};
    double r = 0.0;
    if (op == 0) {
        r = ddx2(F, j, i, Nx, 1.0 / (2.0 * hx));
    } else if (op == 1) {
        r = ddy2(F, j, i, Ny, 1.0 / (2.0 * hy));
    } else if (op == 2) {
        if (i >= 2 && i < Nx - 2)
            r = (-F(j, i + 2) + 8.0 * F(j, i + 1) - 8.0 * F(j, i - 1) + F(j, i - 2)) / (12.0 * hx);
        else
            r = ddx2(F, j, i, Nx, 1.0 / (2.0 * hx));
    } else if (op == 3) {
        if (j >= 2 && j < Ny - 2)
            r = (-F(j + 2, i) + 8.0 * F(j + 1, i) - 8.0 * F(j - 1, i) + F(j - 2, i)) / (12.0 * hy);
        else
            r = ddy2(F, j, i, Ny, 1.0 / (2.0 * hy));
    } else {
        double lx, ly;
        double ihx2 = 1.0 / (hx * hx), ihy2 = 1.0 / (hy * hy);
        if (i > 0 && i < Nx - 1)
            lx = (F(j, i + 1) - 2.0 * F(j, i) + F(j, i - 1)) * ihx2;
        else if (i == 0)
            lx = (2.0 * F(j, 0) - 5.0 * F(j, 1) + 4.0 * F(j, 2) - F(j, 3)) * ihx2;
        else
            lx = (2.0 * F(j, Nx - 1) - 5.0 * F(j, Nx - 2) + 4.0 * F(j, Nx - 3) - F(j, Nx - 4)) * ihx2;
        if (j > 0 && j < Ny - 1)
            ly = (F(j + 1, i) - 2.0 * F(j, i) + F(j - 1, i)) * ihy2;
        else if (j == 0)
            ly = (2.0 * F(0, i) - 5.0 * F(1, i) + 4.0 * F(2, i) - F(3, i)) * ihy2;
        else
            ly = (2.0 * F(Ny - 1, i) - 5.0 * F(Ny - 2, i) + 4.0 * F(Ny - 3, i) - F(Ny - 4, i)) * ihy2;
        r = lx + ly;
    }
    out[(size_t)j * Nx + i] = r;
}

__global__ void k_upwind3(const double *__restrict__ f, const double *__restrict__ u,
                          double *__restrict__ out, int Ny, int Nx, double h, int axis)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    double vel = u[(size_t)j * Nx + i];
    double r;
    if (axis == 1) {
        const double *row = f + (size_t)j * Nx;
        r = upwind3([&](int k) { return __ldg(row + k); }, i, Nx, vel, 1.0 / h);
    } else {
        const double *col = f + i;
        r = upwind3([&](int k) { return __ldg(col + (size_t)k * Nx); }, j, Ny, vel, 1.0 / h);
    }
    out[(size_t)j * Nx + i] = r;
}

__global__ void k_heaviside(const double *__restrict__ x, double *__restrict__ H, long n, double w_t)
{
    double inv_w = 1.0 / w_t;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        H[k] = heaviside_sin(x[k], w_t, inv_w);
}

// ---- reinitialize_phi_PDE (functions.py:1369-1411) ------------------------------------------
// S0 = phi0 / sqrt(phi0^2 + dx^2), fixed for all pseudo-time steps
__global__ void k_reinit_sign(const double *__restrict__ phi0, double *__restrict__ s0, long n, double dx)
{
    const double dx2 = dx * dx;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        const double p = phi0[k];
        s0[k] = p / sqrt(p * p + dx2);
    }
}

// one forward-Euler step: Godunov upwind |grad phi| from one-sided differences (edge value repeated
// outside the grid, np.pad mode='edge'), chosen by the sign of S0
__global__ void __launch_bounds__(256)
k_reinit_step(const double *__restrict__ phi, const double *__restrict__ s0, double *__restrict__ out, int Ny,
              int Nx, double dx, double dy, double dtau)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const size_t c = (size_t)j * Nx + i;
    const double p = phi[c], s = s0[c];
    const double pl = (i > 0) ? __ldg(phi + c - 1) : p, pr = (i < Nx - 1) ? __ldg(phi + c + 1) : p;
    const double pd = (j > 0) ? __ldg(phi + c - Nx) : p, pu = (j < Ny - 1) ? __ldg(phi + c + Nx) : p;
    const double bx = (p - pl) / dx, fx = (pr - p) / dx, by = (p - pd) / dy, fy = (pu - p) / dy;
    double gx2 = 0.0, gy2 = 0.0;
    if (s > 0.0) {
        const double a = fmax(bx, 0.0), b = fmin(fx, 0.0), e = fmax(by, 0.0), f = fmin(fy, 0.0);
        gx2 = fmax(a * a, b * b);
        gy2 = fmax(e * e, f * f);
    } else if (s < 0.0) {
        const double a = fmin(bx, 0.0), b = fmax(fx, 0.0), e = fmin(by, 0.0), f = fmax(fy, 0.0);
        gx2 = fmax(a * a, b * b);
        gy2 = fmax(e * e, f * f);
    }
    out[c] = p - dtau * (s * (sqrt(gx2 + gy2) - 1.0));
}

// H and rho_local = (1-H)*rho_s + H*rho_f in one pass (driver glue,
// soft_disc_in_lid_driven.py:102-103).
__global__ void k_heaviside_rho(const double *__restrict__ phi, double *__restrict__ H,
                                double *__restrict__ rho, long n, double w_t, double rho_s, double rho_f)
{
    double inv_w = 1.0 / w_t;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        double h = heaviside_sin(phi[k], w_t, inv_w);
        if (H) H[k] = h;
        rho[k] = (1.0 - h) * rho_s + h * rho_f;
    }
}

// out = q * (phi <= 0)   (the "* solid_mask" glue, soft_disc_in_lid_driven.py:88-91)
__global__ void k_mask_mul(const double *__restrict__ q, const double *__restrict__ phi,
                           double *__restrict__ out, long n)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        out[k] = q[k] * ((phi[k] <= 0.0) ? 1.0 : 0.0);
}

// ---------------------------------------------------------------- reductions
// Two-stage deterministic reductions: per-block partials, then one block.
// stats layout (doubles): [0]=max sqrt(a^2+b^2)  [1]=count of non-finite inputs
__global__ void k_speed_partial(const double *__restrict__ a, const double *__restrict__ b, long n,
                                double *__restrict__ part)
{
    __shared__ double sh[32];
    double m = 0.0, bad = 0.0;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        double x = a[k], y = b[k];
        if (!(isfinite(x) && isfinite(y))) bad += 1.0;
        m = fmax(m, sqrt(x * x + y * y));
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    m = warp_max(m);
    if (lane == 0) sh[wid] = m;
    __syncthreads();
    if (wid == 0) {
        m = (lane < nw) ? sh[lane] : 0.0;
        m = warp_max(m);
    }
    double s = block_sum(bad, sh);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = m;
        part[2 * blockIdx.x + 1] = s;
    }
}

__global__ void k_speed_final(const double *__restrict__ part, int nb, double *__restrict__ out)
{
    __shared__ double sh[32];
    double m = 0.0, bad = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        m = fmax(m, part[2 * k]);
        bad += part[2 * k + 1];
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    m = warp_max(m);
    if (lane == 0) sh[wid] = m;
    __syncthreads();
    if (wid == 0) {
        m = (lane < nw) ? sh[lane] : 0.0;
        m = warp_max(m);
    }
    double s = block_sum(bad, sh);
    if (threadIdx.x == 0) {
        out[0] = m;
        out[1] = s;
    }
}

// stats: [0]=sum [1]=min [2]=max [3]=count of non-finite
__global__ void k_stats_partial(const double *__restrict__ x, long n, double *__restrict__ part)
{
    __shared__ double sh[32];
    double s = 0.0, lo = INFINITY, hi = -INFINITY, bad = 0.0;
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        double v = x[k];
        s += v;
        lo = fmin(lo, v);
        hi = fmax(hi, v);
        if (!isfinite(v)) bad += 1.0;
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    lo = warp_min(lo);
    hi = warp_max(hi);
    __shared__ double shlo[32], shhi[32];
    if (lane == 0) { shlo[wid] = lo; shhi[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        lo = (lane < nw) ? shlo[lane] : INFINITY;
        hi = (lane < nw) ? shhi[lane] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
    }
    s = block_sum(s, sh);
    bad = block_sum(bad, sh);
    if (threadIdx.x == 0) {
        double *o = part + 4 * (size_t)blockIdx.x;
        o[0] = s; o[1] = lo; o[2] = hi; o[3] = bad;
    }
}

__global__ void k_stats_final(const double *__restrict__ part, int nb, double *__restrict__ out)
{
    __shared__ double sh[32];
    __shared__ double shlo[32], shhi[32];
    double s = 0.0, lo = INFINITY, hi = -INFINITY, bad = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        s += part[4 * k];
        lo = fmin(lo, part[4 * k + 1]);
        hi = fmax(hi, part[4 * k + 2]);
        bad += part[4 * k + 3];
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (lane == 0) { shlo[wid] = lo; shhi[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        lo = (lane < nw) ? shlo[lane] : INFINITY;
        hi = (lane < nw) ? shhi[lane] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
    }
    s = block_sum(s, sh);
    bad = block_sum(bad, sh);
    if (threadIdx.x == 0) { out[0] = s; out[1] = lo; out[2] = hi; out[3] = bad; }
}

// ------------------------------------------------------- boundary conditions
// A BC callable (benchmarks/common.py:27-50, tests/test_poisson.py:39-64) is
// classified once on the host into a gather table: for entry e,
//   field(dst[e])[cell(dst[e])] = a[e] * field(src[e])[cell(src[e])] + b[e]
// with src < 0 meaning "constant b".  Indices pack (field << 62 | cell); all
// sources are cells the BC does not itself overwrite, so in-place is safe.
__global__ void k_apply_bc(double *__restrict__ u, double *__restrict__ v,
                           const long long *__restrict__ dst, const long long *__restrict__ src,
                           const double *__restrict__ ca, const double *__restrict__ cb, int n)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    long long d = dst[e], s = src[e];
    double val = cb[e];
    if (s >= 0) {
        const double *sf = (s >> 62) ? v : u;
        val += ca[e] * sf[s & ((1LL << 62) - 1)];
    }
    double *df = (d >> 62) ? v : u;
    df[d & ((1LL << 62) - 1)] = val;
}

// -------------------------------------------------------------- disc level set
// phi0(xi) = min_k(|xi - c_k| - R_k).  Discs are binned on a uniform
// (gb x gb) grid over [0,Lx]x[0,Ly] in reference space; bin (by,bx) lists the
// candidates cand[start[bin] .. start[bin+1]) that can attain the minimum for
// some point of that bin (a conservative superset built on the host), so the
// result equals the full minimum bit for bit.
struct DiscSet {
    const double *cx, *cy, *R;
    int ndisc;
    const int *bin_start, *cand;
    int gb;
    double Lx, Ly, inv_bw_x, inv_bw_y;
};

__device__ __forceinline__ double disc_sdf_point(double x, double y, const DiscSet &D)
{
    double best = 0.0;
    bool have = false;
    // points outside the binned box (or non-finite) take the exhaustive loop
    if (D.gb > 0 && x >= 0.0 && x <= D.Lx && y >= 0.0 && y <= D.Ly) {
        int bx = min((int)floor(x * D.inv_bw_x), D.gb - 1);
        int by = min((int)floor(y * D.inv_bw_y), D.gb - 1);
        int b = by * D.gb + bx;
        for (int t = D.bin_start[b]; t < D.bin_start[b + 1]; ++t) {
            int k = D.cand[t];
            double ex = x - D.cx[k], ey = y - D.cy[k];
            double d = sqrt(ex * ex + ey * ey) - D.R[k];
            if (!have || d < best) { best = d; have = true; }
        }
    } else {
        for (int k = 0; k < D.ndisc; ++k) {
            double ex = x - D.cx[k], ey = y - D.cy[k];
            double d = sqrt(ex * ex + ey * ey) - D.R[k];
            if (!have || d < best) { best = d; have = true; }
        }
    }
    return best;
}

__global__ void k_disc_sdf(const double *__restrict__ X1, const double *__restrict__ X2,
                           double *__restrict__ phi, long n, const DiscSet D)
{
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x)
        phi[c] = disc_sdf_point(X1[c], X2[c], D);
}

// phi = phi0(xi) AND the solid stress of that xi, phi in one pass (the drivers rebuild the level set
// from the extrapolated map and hand both to the momentum predictor: soft_disc_in_lid_driven.py:93-97):
// the level set of the tile and a one-node halo is formed in shared memory, so xi is read once and
// phi never comes back from HBM for the stress.
struct SmemPhi {
    const double *s;
    int j0, i0;
    __device__ __forceinline__ double operator()(int j, int i) const { return s[(j - j0) * (TX + 2) + (i - i0)]; }
};
struct GlobalField {
    const double *p;
    int Nx;
    __device__ __forceinline__ double operator()(int j, int i) const { return __ldg(p + (size_t)j * Nx + i); }
};

// 32 x 16 output tile per 256-thread CTA: phi of the tile and its one-node halo (34 x 18 nodes, 2.4
// evaluations per thread for 2 outputs) goes to shared memory first
constexpr int STY = 2 * TY;
__global__ void __launch_bounds__(256, 4)
k_sdf_stress(const double *__restrict__ X1, const double *__restrict__ X2, double *__restrict__ phi,
             double *__restrict__ sxx, double *__restrict__ sxy, double *__restrict__ syy,
             double *__restrict__ J, int Ny, int Nx, double dx, double dy, double mu_s, double kappa,
             double w_cut, double detg_clamp, int isochoric, const DiscSet D)
{
    __shared__ double sphi[(STY + 2) * (TX + 2)];
    const int i0 = blockIdx.x * TX - 1, j0 = blockIdx.y * STY - 1;
    const int tid = threadIdx.y * TX + threadIdx.x;
    {   // 612 nodes for 256 threads: three per thread, their loads issued together (each evaluation is a
        // chain of dependent look-ups -- xi, bin, candidate, centre -- so the three chains overlap)
        constexpr int NE = (STY + 2) * (TX + 2);
        double xs[3], ys[3];
        bool ok[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + TX * TY * k;
            const int jj = j0 + e / (TX + 2), ii = i0 + e % (TX + 2);
            ok[k] = e < NE && jj >= 0 && jj < Ny && ii >= 0 && ii < Nx;
            xs[k] = ys[k] = 0.0;
            if (ok[k]) {
                const size_t c = (size_t)jj * Nx + ii;
                xs[k] = __ldg(X1 + c);
                ys[k] = __ldg(X2 + c);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + TX * TY * k;
            if (e < NE) sphi[e] = ok[k] ? disc_sdf_point(xs[k], ys[k], D) : 0.0;
        }
    }
    __syncthreads();
    const SmemPhi P{sphi, j0, i0};
    const GlobalField G1{X1, Nx}, G2{X2, Nx};
    const int i = blockIdx.x * TX + threadIdx.x;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = blockIdx.y * STY + threadIdx.y + TY * r;
        if (i >= Nx || j >= Ny) continue;
        const size_t c = (size_t)j * Nx + i;
        double oxx, oxy, oyy, oJ;
        solid_stress_cell(G1, G2, P, j, i, Ny, Nx, dx, dy, mu_s, kappa, w_cut, detg_clamp, isochoric, oxx, oxy, oyy, oJ);
        phi[c] = P(j, i);
        sxx[c] = oxx;
        sxy[c] = oxy;
        syy[c] = oyy;
        J[c] = oJ;
    }
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

}  // namespace

extern "C" {

int rmt_stencil_op(const double *f, double *out, int Ny, int Nx, double hx, double hy, int op,
                   void *stream)
{
    if (!f || !out || Ny < 4 || Nx < 4 || op < 0 || op > 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_stencil<<<grd, blk, 0, (cudaStream_t)stream>>>(f, out, Ny, Nx, hx, hy, op);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_diff_upwind_3rd(const double *f, const double *u, double *out, int Ny, int Nx, double h,
                        int axis, void *stream)
{
    if (!f || !u || !out || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_upwind3<<<grd, blk, 0, (cudaStream_t)stream>>>(f, u, out, Ny, Nx, h, axis);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_heaviside(const double *x, double *H, long n, double w_t, void *stream)
{
    if (!x || !H || n <= 0) return RMT_EINVAL;
    k_heaviside<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, H, n, w_t);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_reinit_sign(const double *phi0, double *s0, long n, double dx, void *stream)
{
    if (!phi0 || !s0 || n <= 0) return RMT_EINVAL;
    k_reinit_sign<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(phi0, s0, n, dx);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_reinit_step(const double *phi, const double *s0, double *out, int Ny, int Nx, double dx, double dy,
                    double dtau, void *stream)
{
    if (!phi || !s0 || !out || phi == out || Ny < 2 || Nx < 2) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_reinit_step<<<grd, blk, 0, (cudaStream_t)stream>>>(phi, s0, out, Ny, Nx, dx, dy, dtau);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_heaviside_rho(const double *phi, double *H, double *rho, long n, double w_t, double rho_s,
                      double rho_f, void *stream)
{
    if (!phi || !rho || n <= 0) return RMT_EINVAL;
    k_heaviside_rho<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(phi, H, rho, n, w_t, rho_s, rho_f);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_mask_mul(const double *q, const double *phi, double *out, long n, void *stream)
{
    if (!q || !phi || !out || n <= 0) return RMT_EINVAL;
    k_mask_mul<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(q, phi, out, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_reduce_workspace_doubles(void) { return 4 * 148 * 16 + 8; }

int rmt_max_speed(const double *a, const double *b, long n, double *work, double *out2, void *stream)
{
    if (!a || !b || !work || !out2 || n <= 0) return RMT_EINVAL;
    int nb = flat_blocks(n);
    k_speed_partial<<<nb, 256, 0, (cudaStream_t)stream>>>(a, b, n, work);
    RMT_LAUNCH_CHECK();
    k_speed_final<<<1, 256, 0, (cudaStream_t)stream>>>(work, nb, out2);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_field_stats(const double *x, long n, double *work, double *out4, void *stream)
{
    if (!x || !work || !out4 || n <= 0) return RMT_EINVAL;
    int nb = flat_blocks(n);
    k_stats_partial<<<nb, 256, 0, (cudaStream_t)stream>>>(x, n, work);
    RMT_LAUNCH_CHECK();
    k_stats_final<<<1, 256, 0, (cudaStream_t)stream>>>(work, nb, out4);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_apply_bc(double *u, double *v, const long long *dst, const long long *src, const double *ca,
                 const double *cb, int n, void *stream)
{
    if (!u || !v) return RMT_EINVAL;
    if (n <= 0) return RMT_OK;
    k_apply_bc<<<rmt_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(u, v, dst, src, ca, cb, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_disc_sdf(const double *X1, const double *X2, double *phi, long n, const double *cx,
                 const double *cy, const double *R, int ndisc, const int *bin_start, const int *cand,
                 int gb, double Lx, double Ly, void *stream)
{
    if (!X1 || !X2 || !phi || n <= 0 || ndisc <= 0) return RMT_EINVAL;
    double ibx = gb > 0 ? (double)gb / Lx : 0.0, iby = gb > 0 ? (double)gb / Ly : 0.0;
    const DiscSet D{cx, cy, R, ndisc, bin_start, cand, gb, Lx, Ly, ibx, iby};
    k_disc_sdf<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(X1, X2, phi, n, D);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_disc_sdf_stress(const double *X1, const double *X2, double *phi, double *sxx, double *sxy,
                        double *syy, double *J, int Ny, int Nx, double dx, double dy, double mu_s,
                        double kappa, double w_cut, double detg_clamp, int isochoric, const double *cx,
                        const double *cy, const double *R, int ndisc, const int *bin_start, const int *cand,
                        int gb, double Lx, double Ly, void *stream)
{
    if (!X1 || !X2 || !phi || !sxx || !sxy || !syy || !J || Ny < 3 || Nx < 3 || ndisc <= 0) return RMT_EINVAL;
    double ibx = gb > 0 ? (double)gb / Lx : 0.0, iby = gb > 0 ? (double)gb / Ly : 0.0;
    const DiscSet D{cx, cy, R, ndisc, bin_start, cand, gb, Lx, Ly, ibx, iby};
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, STY));
    k_sdf_stress<<<grd, blk, 0, (cudaStream_t)stream>>>(X1, X2, phi, sxx, sxy, syy, J, Ny, Nx, dx, dy, mu_s, kappa,
                                                       w_cut, detg_clamp, isochoric, D);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

// momentum.cu -- solid stress, blended-stress momentum RHS and the fused RK4
// stage kernel.
//
// Replaces (reference file:line):
//   pyRMT/functions.py:545-658   solid_cauchy_stress
//   pyRMT/functions.py:837-861   compute_curvature
//   pyRMT/functions.py:695-707   H, rho_local, surface-tension force
//   pyRMT/functions.py:897-944   velocity_rhs_blended_optimized
//   pyRMT/functions.py:711-758   rhs() closure + classical RK4 stage updates
//
// The RHS is a radius-2 stencil (gradient of a gradient, 3rd-order upwind) that
// turns radius-3 on the domain rim (one-sided 3-point formulas applied twice).
// One CTA owns a 32x16 tile of outputs: the stage velocities are staged in shared
// memory with a halo, the blended stress tensor is formed once per node in
// shared memory (halo 1 in the interior, 2 on rim tiles), and the divergence,
// upwind advection, pressure gradient, density division and the RK4 stage update
// are applied from there -- each input field is read once per stage and no RHS
// array ever reaches HBM.  The ~15 full-size NumPy temporaries per evaluation of
// the reference collapse into this one kernel.
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

// ------------------------------------------------------------ solid stress
__global__ void __launch_bounds__(256)
k_solid_stress(const double *__restrict__ X1, const double *__restrict__ X2,
               const double *__restrict__ phi, double *__restrict__ sxx, double *__restrict__ sxy,
               double *__restrict__ syy, double *__restrict__ J, int Ny, int Nx, double dx, double dy,
               double mu_s, double kappa, double w_cut, double detg_clamp, int isochoric)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    size_t c = (size_t)j * Nx + i;
    double oxx, oxy, oyy, oJ;
    const GField G1{X1, Nx}, G2{X2, Nx}, GP{phi, Nx};
    solid_stress_cell(G1, G2, GP, j, i, Ny, Nx, dx, dy, mu_s, kappa, w_cut, detg_clamp, isochoric, oxx, oxy, oyy, oJ);
    sxx[c] = oxx;
    sxy[c] = oxy;
    syy[c] = oyy;
    J[c] = oJ;
}

// -------------------------------------------------- curvature / surface tension
struct UnitNormal {   // n = grad(phi) / (|grad(phi)| + 1e-12), component C (0 = x, 1 = y)
    GField phi;
    int Ny, Nx, comp;
    double i2dx, i2dy;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        double gx = ddx2(phi, j, i, Nx, i2dx), gy = ddy2(phi, j, i, Ny, i2dy);
        double mag = sqrt(gx * gx + gy * gy) + 1e-12;
        return (comp ? gy : gx) / mag;
    }
};

struct HeavisideOf {
    GField phi;
    double w_t, inv_w;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        return heaviside_sin(phi(j, i), w_t, inv_w);
    }
};

// mode 0: curvature only (out0); mode 1: st_force (out0 = fsx, out1 = fsy)
__global__ void __launch_bounds__(256)
k_curvature(const double *__restrict__ phi, double *__restrict__ out0, double *__restrict__ out1,
            int Ny, int Nx, double dx, double dy, double w_t, double gamma, int mode)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const double i2dx = 1.0 / (2.0 * dx), i2dy = 1.0 / (2.0 * dy);
    GField P{phi, Nx};
    UnitNormal nx{P, Ny, Nx, 0, i2dx, i2dy}, ny{P, Ny, Nx, 1, i2dx, i2dy};
    double curv = ddx2(nx, j, i, Nx, i2dx) + ddy2(ny, j, i, Ny, i2dy);
    size_t c = (size_t)j * Nx + i;
    if (mode == 0) {
        out0[c] = curv;
    } else {
        HeavisideOf Hf{P, w_t, 1.0 / w_t};
        double hx = ddx2(Hf, j, i, Nx, i2dx), hy = ddy2(Hf, j, i, Ny, i2dy);
        out0[c] = -gamma * curv * hx;
        out1[c] = -gamma * curv * hy;
    }
}

// ------------------------------------------------------- contact force (functions.py:864-895)
struct MidSurface {     // phi12 = (phi1 - phi2) / 2
    const double *a, *b;
    int Nx;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        const size_t c = (size_t)j * Nx + i;
        return 0.5 * (__ldg(a + c) - __ldg(b + c));
    }
};

__global__ void __launch_bounds__(256)
k_contact_force(const double *__restrict__ phi1, const double *__restrict__ phi2, double *__restrict__ fx,
                double *__restrict__ fy, int Ny, int Nx, double dx, double dy, double k_rep, double w_c)
{
    int i = blockIdx.x * TX + threadIdx.x;
    int j = blockIdx.y * TY + threadIdx.y;
    if (i >= Nx || j >= Ny) return;
    const size_t c = (size_t)j * Nx + i;
    const MidSurface P{phi1, phi2, Nx};
    const double p12 = P(j, i);
    double ox = 0.0, oy = 0.0;
    const bool active = (phi1[c] < 0.0) || (phi2[c] < 0.0);
    if (fabs(p12) < w_c && active && p12 != 0.0) {
        const double pi = 3.141592653589793;
        const double delta = (1.0 + cos(pi * p12 / w_c)) / (2.0 * w_c);
        const double gx = ddx2(P, j, i, Nx, 1.0 / (2.0 * dx)), gy = ddy2(P, j, i, Ny, 1.0 / (2.0 * dy));
        const double gmag = sqrt(gx * gx + gy * gy) + 1e-12;
        const double sg = (p12 > 0.0) ? 1.0 : -1.0;
        ox = k_rep * delta * sg * (gx / gmag);
        oy = k_rep * delta * sg * (gy / gmag);
    }
    fx[c] = ox;
    fy[c] = oy;
}

__global__ void k_min2(const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out, long n)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        out[k] = fmin(a[k], b[k]);
}

// ------------------------------------------------------- diagnostics (output.py:6-193, common.py:110-115)
// One pass over the grid for the quantities the drivers log beside the step: kinetic energy density
// 0.5 rho (a^2 + b^2), neo-Hookean strain energy density on solid cells (central differences of the
// edge-padded reference map, output.py:70-83), viscous dissipation density 2 mu (Dxx^2 + Dyy^2 + 2 Dxy^2),
// and the solid cell count / coordinate sums of disc_centroid.  Any input group may be absent (NULL).
constexpr int kDiagBlocks = 148 * 8, kDiagVals = 6;

__global__ void __launch_bounds__(256)
k_diagnostics(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ X1,
              const double *__restrict__ X2, const double *__restrict__ phi, const double *__restrict__ Xc,
              const double *__restrict__ Yc, int Ny, int Nx, double dx, double dy, double rho_f, double rho_s,
              double mu_f, double mu_s, double kappa, double eta_s, double w_t, double *__restrict__ partial)
{
    __shared__ double red[32];
    const double inv_w = 1.0 / w_t, i2dx = 1.0 / (2.0 * dx), i2dy = 1.0 / (2.0 * dy);
    double acc[kDiagVals] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const long n = (long)Ny * Nx;
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
        const int j = (int)(c / Nx), i = (int)(c - (long)j * Nx);
        const double ph = phi[c];
        const double H = heaviside_sin(ph, w_t, inv_w);
        if (a && b) {
            const double u = a[c], v = b[c];
            acc[0] += 0.5 * ((1.0 - H) * rho_s + H * rho_f) * (u * u + v * v);
            const GField A{a, Nx}, B{b, Nx};
            const double dxx = ddx2(A, j, i, Nx, i2dx), dyy = ddy2(B, j, i, Ny, i2dy);
            const double dxy = 0.5 * (ddy2(A, j, i, Ny, i2dy) + ddx2(B, j, i, Nx, i2dx));
            acc[2] += 2.0 * (H * mu_f + (1.0 - H) * eta_s) * (dxx * dxx + dyy * dyy + 2.0 * dxy * dxy);
        }
        if (ph <= 0.0) {
            if (X1 && X2) {
                const int ip = min(i + 1, Nx - 1), im = max(i - 1, 0), jp = min(j + 1, Ny - 1), jm = max(j - 1, 0);
                const size_t r = (size_t)j * Nx, rp = (size_t)jp * Nx, rm = (size_t)jm * Nx;
                const double G11 = (X1[r + ip] - X1[r + im]) * i2dx, G12 = (X1[rp + i] - X1[rm + i]) * i2dy;
                const double G21 = (X2[r + ip] - X2[r + im]) * i2dx, G22 = (X2[rp + i] - X2[rm + i]) * i2dy;
                const double det = G11 * G22 - G12 * G21;
                if (fabs(det) > 1e-10) {
                    const double F11 = G22 / det, F12 = -G12 / det, F21 = -G21 / det, F22 = G11 / det;
                    const double I1 = (F11 * F11 + F21 * F21) + (F12 * F12 + F22 * F22), Jm1 = 1.0 / det - 1.0;
                    acc[1] += 0.5 * mu_s * (I1 - 2.0) + 0.5 * kappa * Jm1 * Jm1;
                }
            }
            acc[3] += 1.0;
            acc[4] += Xc ? Xc[c] : dx * i;
            acc[5] += Yc ? Yc[c] : dy * j;
        }
    }
#pragma unroll
    for (int k = 0; k < kDiagVals; ++k) {
        const double t = block_sum(acc[k], red);
        if (threadIdx.x == 0) partial[blockIdx.x * kDiagVals + k] = t;
    }
}

__global__ void k_diagnostics_final(const double *__restrict__ partial, int nblocks, double *__restrict__ out)
{
    __shared__ double red[32];
    for (int k = 0; k < kDiagVals; ++k) {
        double t = 0.0;
        for (int q = threadIdx.x; q < nblocks; q += blockDim.x) t += partial[q * kDiagVals + k];
        t = block_sum(t, red);
        if (threadIdx.x == 0) out[k] = t;
        __syncthreads();
    }
}

// ------------------------------------------------------- momentum RHS / stage
constexpr int MTX = 32, MTY = 16;            // output tile
constexpr int UHALO = 3, THALO = 2;          // shared-memory halos (rim tiles use all of it)
constexpr int UW = MTX + 2 * UHALO, UHT = MTY + 2 * UHALO;
constexpr int TW = MTX + 2 * THALO, THT = MTY + 2 * THALO;

struct STile {   // shared-memory tile addressed with global (j, i)
    const double *s;
    int j0, i0, W;
    __device__ __forceinline__ double operator()(int j, int i) const
    {
        return s[(j - j0) * W + (i - i0)];
    }
};

struct RhsArgs {
    const double *us, *vs, *p, *sxx, *sxy, *syy;
    const double *sbxx, *sbxy, *sbyy, *phi_b;   // MODE 2: the second solid (functions.py:765-835)
    const double *phi;          // FUSED: level set (H, rho, solid mask derived in-kernel)
    const double *H, *rho;      // !FUSED: arrays as passed to velocity_rhs_blended_optimized
    const double *fsx, *fsy;    // optional surface-tension force
    const double *u0, *v0;      // FUSED: base state of the RK4 step
    double *acc_u, *acc_v;      // FUSED: running k1 + 2 k2 + 2 k3
    double *out_u, *out_v;      // FUSED: next stage state / final;  !FUSED: ru, rv
    int Ny, Nx;
    double dx, dy, dt, mu_f, eta_s, w_t, rho_s, rho_f;
    int stage;
};

// MODE 0: velocity_rhs_blended_optimized on caller-supplied H / rho arrays; 1: one RK4 stage of
// momentum_step_rk4 (one solid); 2: one RK4 stage of momentum_step_rk4_2solids (n = 2 mixture
// sigma = (Ha+Hb-1) sigma_f + (1-Ha) sigma_A + (1-Hb) sigma_B, no Kelvin-Voigt term).
// Differences for tiles that cannot touch a rim formula (RIM = false): the interior branch of
// ddx2 / ddy2 / upwind3 without the position tests -- the kernel is instruction-issue bound.
template <bool RIM, class F>
__device__ __forceinline__ double tdx2(F f, int j, int i, int Nx, double inv2h)
{
    if (!RIM) return (f(j, i + 1) - f(j, i - 1)) * inv2h;
    return ddx2(f, j, i, Nx, inv2h);
}
template <bool RIM, class F>
__device__ __forceinline__ double tdy2(F f, int j, int i, int Ny, double inv2h)
{
    if (!RIM) return (f(j + 1, i) - f(j - 1, i)) * inv2h;
    return ddy2(f, j, i, Ny, inv2h);
}
template <bool RIM, class G>
__device__ __forceinline__ double tup3(G g, int k, int n, double vel, double invh)
{
    if (!RIM) {
        if (vel > 0.0)
            return (2.0 * g(k + 1) + 3.0 * g(k) - 6.0 * g(k - 1) + g(k - 2)) * (invh * (1.0 / 6.0));
        return (-g(k + 2) + 6.0 * g(k + 1) - 3.0 * g(k) - 2.0 * g(k - 1)) * (invh * (1.0 / 6.0));
    }
    return upwind3(g, k, n, vel, invh);
}

// One 32 x 16 output tile.  RIM = true: the tile touches (or comes within two nodes of) the domain
// edge -- clamped halos of run-time extent, one-sided formulas.  RIM = false: full tile, halo
// extents known at compile time (no run-time divisions in the index maps), interior formulas only.
template <int MODE, bool RIM>
__device__ __forceinline__ void momentum_tile(const RhsArgs &A, double *sU, double *sV, double *sTxx,
                                              double *sTxy, double *sTyy, double *sH, double *sHb)
{
    constexpr bool FUSED = (MODE != 0), TWO = (MODE == 2);
    const int Ny = A.Ny, Nx = A.Nx;
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int i1 = RIM ? min(i0 + MTX, Nx) : i0 + MTX, j1 = RIM ? min(j0 + MTY, Ny) : j0 + MTY;   // tile end
    // rim tiles need the blended stress two nodes away (one-sided divergence)
    constexpr int th = RIM ? 2 : 1, uh = th + 1;

    // Every global load whose address is known up front is issued HERE, before the first barrier:
    // the velocity tile (<= 4 points per thread), the level set / Heaviside at this thread's stress
    // nodes (<= 3), and the output phase's streaming fields -- ~19 loads in flight per thread, so the
    // kernel pays the HBM latency once instead of once per loop iteration and per phase.
    const int uja = RIM ? max(j0 - uh, 0) : j0 - uh, ujb = RIM ? min(j1 + uh, Ny) : j1 + uh;
    const int uia = RIM ? max(i0 - uh, 0) : i0 - uh, uib = RIM ? min(i1 + uh, Nx) : i1 + uh;
    const int uw = RIM ? uib - uia : MTX + 2 * uh, un = RIM ? (ujb - uja) * uw : (MTY + 2 * uh) * (MTX + 2 * uh);
    double ru[4], rv[4];
    int us_idx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = tid + 256 * k;
        us_idx[k] = -1;
        ru[k] = rv[k] = 0.0;
        if (e < un) {
            const int jj = uja + e / uw, ii = uia + e % uw;
            const size_t g = (size_t)jj * Nx + ii;
            us_idx[k] = (jj - (j0 - UHALO)) * UW + (ii - (i0 - UHALO));
            ru[k] = __ldg(A.us + g);
            rv[k] = __ldg(A.vs + g);
        }
    }
    const int tja = RIM ? max(j0 - th, 0) : j0 - th, tjb = RIM ? min(j1 + th, Ny) : j1 + th;
    const int tia = RIM ? max(i0 - th, 0) : i0 - th, tib = RIM ? min(i1 + th, Nx) : i1 + th;
    const int tw_ = RIM ? tib - tia : MTX + 2 * th, tn = RIM ? (tjb - tja) * tw_ : (MTY + 2 * th) * (MTX + 2 * th);
    double rh[3];                                  // phi (FUSED) or H at this thread's stress nodes
    double rhb[TWO ? 3 : 1];                       // phi of the second solid
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = tid + 256 * k;
        rh[k] = 1.0;
        if (TWO) rhb[k] = 1.0;
        if (e < tn) {
            const size_t g = (size_t)(tja + e / tw_) * Nx + (tia + e % tw_);
            rh[k] = FUSED ? __ldg(A.phi + g) : __ldg(A.H + g);
            if (TWO) rhb[k] = __ldg(A.phi_b + g);
        }
    }
    double pf_u0[MTY / 8], pf_v0[MTY / 8], pf_au[MTY / 8], pf_av[MTY / 8];
#pragma unroll
    for (int r = 0; r < MTY / 8; ++r) {
        const int i = i0 + threadIdx.x, j = j0 + threadIdx.y + 8 * r;
        pf_u0[r] = pf_v0[r] = pf_au[r] = pf_av[r] = 0.0;
        if (FUSED && (!RIM || (i < Nx && j < Ny))) {
            const size_t c = (size_t)j * Nx + i;
            pf_u0[r] = __ldg(A.u0 + c);
            pf_v0[r] = __ldg(A.v0 + c);
            if (A.stage > 1) {
                pf_au[r] = A.acc_u[c];
                pf_av[r] = A.acc_v[c];
            }
        }
    }
    // ---- stage velocities with halo -> shared memory ------------------------
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (us_idx[k] >= 0) {
            sU[us_idx[k]] = ru[k];
            sV[us_idx[k]] = rv[k];
        }
    __syncthreads();
    const STile U{sU, j0 - UHALO, i0 - UHALO, UW}, V{sV, j0 - UHALO, i0 - UHALO, UW};
    const double i2dx = 1.0 / (2.0 * A.dx), i2dy = 1.0 / (2.0 * A.dy);
    const double inv_w = FUSED ? 1.0 / A.w_t : 0.0;

    // ---- blended stress  T = H sigma_f + (1-H) sigma_s  --------------------
    {
        // solid stress only where (1-H) != 0: issue those loads for all of this thread's nodes first
        double hh[3], sx[3], sy[3], sq[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * k;
            hh[k] = FUSED ? heaviside_sin(rh[k], A.w_t, inv_w) : rh[k];
            sx[k] = sy[k] = sq[k] = 0.0;
            if (!TWO && e < tn && hh[k] != 1.0) {
                const size_t g = (size_t)(tja + e / tw_) * Nx + (tia + e % tw_);
                sx[k] = __ldg(A.sxx + g);
                sy[k] = __ldg(A.syy + g);
                sq[k] = __ldg(A.sxy + g);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * k;
            if (e >= tn) continue;
            const int jj = tja + e / tw_, ii = tia + e % tw_;
            const double ux = tdx2<RIM>(U, jj, ii, Nx, i2dx), vy = tdy2<RIM>(V, jj, ii, Ny, i2dy);
            const double uy = tdy2<RIM>(U, jj, ii, Ny, i2dy), vx = tdx2<RIM>(V, jj, ii, Nx, i2dx);
            const double h = hh[k];
            if (TWO) {      // functions.py:806-810
                const size_t g = (size_t)jj * Nx + ii;
                const double hb = heaviside_sin(rhb[k], A.w_t, inv_w), hf = h + hb - 1.0;
                const double oma = 1.0 - h, omb = 1.0 - hb;
                const int s = (jj - (j0 - THALO)) * TW + (ii - (i0 - THALO));
                sTxx[s] = hf * (2.0 * A.mu_f * ux) + oma * __ldg(A.sxx + g) + omb * __ldg(A.sbxx + g);
                sTyy[s] = hf * (2.0 * A.mu_f * vy) + oma * __ldg(A.syy + g) + omb * __ldg(A.sbyy + g);
                sTxy[s] = hf * (A.mu_f * (uy + vx)) + oma * __ldg(A.sxy + g) + omb * __ldg(A.sbxy + g);
                sH[s] = h;
                sHb[s] = hb;
                continue;
            }
            double txx = h * (2.0 * A.mu_f * ux);
            double tyy = h * (2.0 * A.mu_f * vy);
            double txy = h * (A.mu_f * (uy + vx));
            if (h != 1.0) {   // (1-H)*sigma_s vanishes identically in the pure fluid
                double ex = sx[k], ey = sy[k], exy = sq[k];
                if (MODE == 1 && A.eta_s > 0.0 && rh[k] <= 0.0) {   // Kelvin-Voigt, :717-730 (rh = phi)
                    ex += A.eta_s * ux;
                    ey += A.eta_s * vy;
                    exy += A.eta_s * 0.5 * (uy + vx);
                }
                const double omh = 1.0 - h;
                txx += omh * ex;
                tyy += omh * ey;
                txy += omh * exy;
            }
            const int s = (jj - (j0 - THALO)) * TW + (ii - (i0 - THALO));
            sTxx[s] = txx;
            sTxy[s] = txy;
            sTyy[s] = tyy;
            sH[s] = h;
        }
    }
    __syncthreads();
    const STile Txx{sTxx, j0 - THALO, i0 - THALO, TW}, Txy{sTxy, j0 - THALO, i0 - THALO, TW};
    const STile Tyy{sTyy, j0 - THALO, i0 - THALO, TW}, Hs{sH, j0 - THALO, i0 - THALO, TW};
    const GField P{A.p, Nx};
    const double invdx = 1.0 / A.dx, invdy = 1.0 / A.dy;

    // ---- outputs: two rows per thread -------------------------------------
#pragma unroll
    for (int r = 0; r < MTY / 8; ++r) {
        const int i = i0 + threadIdx.x, j = j0 + threadIdx.y + 8 * r;
        if (RIM && (i >= Nx || j >= Ny)) continue;
        const size_t c = (size_t)j * Nx + i;
        const double u = U(j, i), v = V(j, i);
        double div_x = tdx2<RIM>(Txx, j, i, Nx, i2dx) + tdy2<RIM>(Txy, j, i, Ny, i2dy);
        double div_y = tdx2<RIM>(Txy, j, i, Nx, i2dx) + tdy2<RIM>(Tyy, j, i, Ny, i2dy);
        double dux = tup3<RIM>([&](int k) { return U(j, k); }, i, Nx, u, invdx);
        double duy = tup3<RIM>([&](int k) { return U(k, i); }, j, Ny, v, invdy);
        double dvx = tup3<RIM>([&](int k) { return V(j, k); }, i, Nx, u, invdx);
        double dvy = tup3<RIM>([&](int k) { return V(k, i); }, j, Ny, v, invdy);
        double adv_u = -u * dux - v * duy;
        double adv_v = -u * dvx - v * dvy;
        double px = tdx2<RIM>(P, j, i, Nx, i2dx), py = tdy2<RIM>(P, j, i, Ny, i2dy);
        double rho;
        if (TWO) {          // functions.py:793
            const int s = (j - (j0 - THALO)) * TW + (i - (i0 - THALO));
            const double ha = sH[s], hb = sHb[s];
            rho = (ha + hb - 1.0) * A.rho_f + (1.0 - ha) * A.rho_s + (1.0 - hb) * A.rho_s;
        } else if (FUSED) {
            double h = Hs(j, i);
            rho = (1.0 - h) * A.rho_s + h * A.rho_f;
        } else {
            rho = __ldg(A.rho + c);
        }
        double fx = A.fsx ? __ldg(A.fsx + c) : 0.0, fy = A.fsy ? __ldg(A.fsy + c) : 0.0;
        double ku = adv_u + (div_x + fx - px) / (rho + 1e-12);
        double kv = adv_v + (div_y + fy - py) / (rho + 1e-12);
        if (!FUSED) {
            A.out_u[c] = ku;
            A.out_v[c] = kv;
        } else {
            const double ub = pf_u0[r], vb = pf_v0[r];
            if (A.stage == 1) {
                A.acc_u[c] = ku;
                A.acc_v[c] = kv;
                A.out_u[c] = ub + (0.5 * A.dt) * ku;
                A.out_v[c] = vb + (0.5 * A.dt) * kv;
            } else if (A.stage == 2 || A.stage == 3) {
                A.acc_u[c] = pf_au[r] + 2.0 * ku;
                A.acc_v[c] = pf_av[r] + 2.0 * kv;
                double cdt = (A.stage == 2) ? 0.5 * A.dt : A.dt;
                A.out_u[c] = ub + cdt * ku;
                A.out_v[c] = vb + cdt * kv;
            } else {
                A.out_u[c] = ub + (A.dt / 6.0) * (pf_au[r] + ku);
                A.out_v[c] = vb + (A.dt / 6.0) * (pf_av[r] + kv);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 4)
k_momentum_rhs(const RhsArgs A)
{
    __shared__ double sU[UHT * UW], sV[UHT * UW];
    __shared__ double sTxx[THT * TW], sTxy[THT * TW], sTyy[THT * TW], sH[THT * TW];
    __shared__ double sHb[MODE == 2 ? THT * TW : 1];
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY;
    // the one-sided / first-order formulas reach two nodes in from the edge (upwind3: k < 2, k >= n - 2)
    const bool rim = (i0 == 0) || (j0 == 0) || (i0 + MTX >= A.Nx - 1) || (j0 + MTY >= A.Ny - 1);
    if (rim) momentum_tile<MODE, true>(A, sU, sV, sTxx, sTxy, sTyy, sH, sHb);
    else momentum_tile<MODE, false>(A, sU, sV, sTxx, sTxy, sTyy, sH, sHb);
}

}  // namespace

extern "C" {

int rmt_abi_version(void) { return 1; }
unsigned long long g_rmt_launches = 0;
unsigned long long rmt_launch_count(void) { return g_rmt_launches; }

int rmt_solid_stress(const double *X1, const double *X2, const double *phi, double *sxx, double *sxy,
                     double *syy, double *J, int Ny, int Nx, double dx, double dy, double mu_s,
                     double kappa, double w_cut, double detg_clamp, int isochoric, void *stream)
{
    if (!X1 || !X2 || !phi || !sxx || !sxy || !syy || !J || Ny < 3 || Nx < 3) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_solid_stress<<<grd, blk, 0, (cudaStream_t)stream>>>(X1, X2, phi, sxx, sxy, syy, J, Ny, Nx, dx, dy,
                                                         mu_s, kappa, w_cut, detg_clamp, isochoric);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_curvature(const double *phi, double *curv, int Ny, int Nx, double dx, double dy, void *stream)
{
    if (!phi || !curv || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_curvature<<<grd, blk, 0, (cudaStream_t)stream>>>(phi, curv, nullptr, Ny, Nx, dx, dy, 1.0, 0.0, 0);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_surface_tension(const double *phi, double *fsx, double *fsy, int Ny, int Nx, double dx,
                        double dy, double w_t, double gamma, void *stream)
{
    if (!phi || !fsx || !fsy || Ny < 4 || Nx < 4) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_curvature<<<grd, blk, 0, (cudaStream_t)stream>>>(phi, fsx, fsy, Ny, Nx, dx, dy, w_t, gamma, 1);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_velocity_rhs(const double *u, const double *v, const double *p, const double *sxx,
                     const double *sxy, const double *syy, const double *H, const double *rho,
                     const double *fsx, const double *fsy, double *ru, double *rv, int Ny, int Nx,
                     double dx, double dy, double mu_f, void *stream)
{
    if (!u || !v || !p || !sxx || !sxy || !syy || !H || !rho || !ru || !rv || Ny < 4 || Nx < 4)
        return RMT_EINVAL;
    RhsArgs A{};
    A.us = u; A.vs = v; A.p = p; A.sxx = sxx; A.sxy = sxy; A.syy = syy;
    A.H = H; A.rho = rho; A.fsx = fsx; A.fsy = fsy; A.out_u = ru; A.out_v = rv;
    A.Ny = Ny; A.Nx = Nx; A.dx = dx; A.dy = dy; A.mu_f = mu_f;
    dim3 blk(32, 8), grd(rmt_cdiv(Nx, MTX), rmt_cdiv(Ny, MTY));
    k_momentum_rhs<0><<<grd, blk, 0, (cudaStream_t)stream>>>(A);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_momentum_stage(const double *us, const double *vs, const double *p, const double *sxx,
                       const double *sxy, const double *syy, const double *phi, const double *fsx,
                       const double *fsy, const double *u0, const double *v0, double *acc_u,
                       double *acc_v, double *out_u, double *out_v, int Ny, int Nx, double dx,
                       double dy, double dt, double mu_f, double eta_s, double w_t, double rho_s,
                       double rho_f, int stage, void *stream)
{
    if (!us || !vs || !p || !sxx || !sxy || !syy || !phi || !u0 || !v0 || !acc_u || !acc_v || !out_u ||
        !out_v || Ny < 4 || Nx < 4 || stage < 1 || stage > 4)
        return RMT_EINVAL;
    RhsArgs A{};
    A.us = us; A.vs = vs; A.p = p; A.sxx = sxx; A.sxy = sxy; A.syy = syy; A.phi = phi;
    A.fsx = fsx; A.fsy = fsy; A.u0 = u0; A.v0 = v0; A.acc_u = acc_u; A.acc_v = acc_v;
    A.out_u = out_u; A.out_v = out_v; A.Ny = Ny; A.Nx = Nx; A.dx = dx; A.dy = dy; A.dt = dt;
    A.mu_f = mu_f; A.eta_s = eta_s; A.w_t = w_t; A.rho_s = rho_s; A.rho_f = rho_f; A.stage = stage;
    dim3 blk(32, 8), grd(rmt_cdiv(Nx, MTX), rmt_cdiv(Ny, MTY));
    k_momentum_rhs<1><<<grd, blk, 0, (cudaStream_t)stream>>>(A);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_momentum_stage_2solids(const double *us, const double *vs, const double *p, const double *sAxx,
                               const double *sAxy, const double *sAyy, const double *sBxx, const double *sBxy,
                               const double *sByy, const double *phi_a, const double *phi_b, const double *fcx,
                               const double *fcy, const double *u0, const double *v0, double *acc_u,
                               double *acc_v, double *out_u, double *out_v, int Ny, int Nx, double dx, double dy,
                               double dt, double mu_f, double w_t, double rho_s, double rho_f, int stage,
                               void *stream)
{
    if (!us || !vs || !p || !sAxx || !sAxy || !sAyy || !sBxx || !sBxy || !sByy || !phi_a || !phi_b || !u0 ||
        !v0 || !acc_u || !acc_v || !out_u || !out_v || Ny < 4 || Nx < 4 || stage < 1 || stage > 4)
        return RMT_EINVAL;
    RhsArgs A{};
    A.us = us; A.vs = vs; A.p = p; A.sxx = sAxx; A.sxy = sAxy; A.syy = sAyy;
    A.sbxx = sBxx; A.sbxy = sBxy; A.sbyy = sByy; A.phi = phi_a; A.phi_b = phi_b;
    A.fsx = fcx; A.fsy = fcy; A.u0 = u0; A.v0 = v0; A.acc_u = acc_u; A.acc_v = acc_v;
    A.out_u = out_u; A.out_v = out_v; A.Ny = Ny; A.Nx = Nx; A.dx = dx; A.dy = dy; A.dt = dt;
    A.mu_f = mu_f; A.eta_s = 0.0; A.w_t = w_t; A.rho_s = rho_s; A.rho_f = rho_f; A.stage = stage;
    dim3 blk(32, 8), grd(rmt_cdiv(Nx, MTX), rmt_cdiv(Ny, MTY));
    k_momentum_rhs<2><<<grd, blk, 0, (cudaStream_t)stream>>>(A);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_contact_force(const double *phi1, const double *phi2, double *fx, double *fy, int Ny, int Nx, double dx,
                      double dy, double k_rep, double w_c, void *stream)
{
    if (!phi1 || !phi2 || !fx || !fy || Ny < 3 || Nx < 3 || !(w_c > 0.0)) return RMT_EINVAL;
    dim3 blk(TX, TY), grd(rmt_cdiv(Nx, TX), rmt_cdiv(Ny, TY));
    k_contact_force<<<grd, blk, 0, (cudaStream_t)stream>>>(phi1, phi2, fx, fy, Ny, Nx, dx, dy, k_rep, w_c);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_diagnostics_workspace_doubles(void) { return kDiagBlocks * kDiagVals; }

int rmt_diagnostics(const double *a, const double *b, const double *X1, const double *X2, const double *phi,
                    const double *Xc, const double *Yc, int Ny, int Nx, double dx, double dy, double rho_f,
                    double rho_s, double mu_f, double mu_s, double kappa, double eta_s, double w_t, double *work,
                    double *out6, void *stream)
{
    if (!phi || !work || !out6 || Ny < 3 || Nx < 3 || !(w_t > 0.0) || (a == nullptr) != (b == nullptr) ||
        (X1 == nullptr) != (X2 == nullptr))
        return RMT_EINVAL;
    long n = (long)Ny * Nx, blocks = (n + 255) / 256;
    if (blocks > kDiagBlocks) blocks = kDiagBlocks;
    k_diagnostics<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, X1, X2, phi, Xc, Yc, Ny, Nx, dx, dy, rho_f,
                                                                  rho_s, mu_f, mu_s, kappa, eta_s, w_t, work);
    RMT_LAUNCH_CHECK();
    k_diagnostics_final<<<1, 256, 0, (cudaStream_t)stream>>>(work, (int)blocks, out6);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_min2(const double *a, const double *b, double *out, long n, void *stream)
{
    if (!a || !b || !out || n <= 0) return RMT_EINVAL;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_min2<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

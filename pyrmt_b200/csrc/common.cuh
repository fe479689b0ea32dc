// common.cuh -- shared device helpers for the sm_100a RMT kernels.
//
// All fields are fp64, C-contiguous (Ny, Nx), index [j*Nx + i] (x fastest),
// exactly the layout of the reference's NumPy arrays ("Sizes").
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define RMT_OK 0
#define RMT_EINVAL (-1)

// Kernels launched by this library since load (rmt_launch_count; single host thread, no atomics).
extern "C" unsigned long long g_rmt_launches;

// Launch epilogue: count the launch and report launch-time errors through the C ABI as an int.
#define RMT_LAUNCH_CHECK()                         \
    do {                                           \
        ++g_rmt_launches;                          \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define RMT_CUDA(call)                             \
    do {                                           \
        cudaError_t e__ = (call);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

static inline int rmt_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

namespace rmt {

// 2-D tile shape used by the pointwise/stencil kernels: 32 lanes along x so
// each warp touches 256 contiguous bytes per row.
constexpr int TX = 32;
constexpr int TY = 8;

struct Grid {
    int Ny, Nx;
    __host__ __device__ size_t at(int j, int i) const { return (size_t)j * (size_t)Nx + (size_t)i; }
};

// ---- edge-aware 2nd-order derivatives of an arbitrary cell functor --------
// Mirrors pyRMT/utils.py:4-25: central in the interior, 3-point one-sided on
// the domain rim.  `inv2h` = 1/(2h).
template <class F>
__device__ __forceinline__ double ddx2(F f, int j, int i, int Nx, double inv2h)
{
    if (i > 0 && i < Nx - 1) return (f(j, i + 1) - f(j, i - 1)) * inv2h;
    if (i == 0) return (-3.0 * f(j, 0) + 4.0 * f(j, 1) - f(j, 2)) * inv2h;
    return (3.0 * f(j, Nx - 1) - 4.0 * f(j, Nx - 2) + f(j, Nx - 3)) * inv2h;
}

template <class F>
__device__ __forceinline__ double ddy2(F f, int j, int i, int Ny, double inv2h)
{
    if (j > 0 && j < Ny - 1) return (f(j + 1, i) - f(j - 1, i)) * inv2h;
    if (j == 0) return (-3.0 * f(0, i) + 4.0 * f(1, i) - f(2, i)) * inv2h;
    return (3.0 * f(Ny - 1, i) - 4.0 * f(Ny - 2, i) + f(Ny - 3, i)) * inv2h;
}

// ---- 3rd-order upwind-biased difference along a line ----------------------
// pyRMT/utils.py:61-114.  `g(k)` returns the differenced field at line index
// k (0..n-1); `vel` is the advecting velocity at k; `invh` = 1/h.
template <class G>
__device__ __forceinline__ double upwind3(G g, int k, int n, double vel, double invh)
{
    if (k >= 2 && k < n - 2) {
        if (vel > 0.0)
            return (2.0 * g(k + 1) + 3.0 * g(k) - 6.0 * g(k - 1) + g(k - 2)) * (invh * (1.0 / 6.0));
        return (-g(k + 2) + 6.0 * g(k + 1) - 3.0 * g(k) - 2.0 * g(k - 1)) * (invh * (1.0 / 6.0));
    }
    if (vel > 0.0 && k > 0) return (g(k) - g(k - 1)) * invh;
    if (vel <= 0.0 && k < n - 1) return (g(k + 1) - g(k)) * invh;
    if (k > 0) return (g(k) - g(k - 1)) * invh;
    if (k < n - 1) return (g(k + 1) - g(k)) * invh;
    return 0.0;
}

// ---- smoothed Heaviside, pyRMT/functions.py:660-671 (the sin form) --------
__device__ __forceinline__ double heaviside_sin(double x, double w_t, double inv_w)
{
    if (x > w_t) return 1.0;
    if (x < -w_t) return 0.0;
    const double inv_pi = 0.31830988618379067154;   // 1/pi
    const double pi = 3.14159265358979323846;
    return 0.5 * (1.0 + x * inv_w + inv_pi * sin(pi * x * inv_w));
}

// ---- neo-Hookean Cauchy stress of one node, pyRMT/functions.py:545-658 ------
// X1, X2, PHI: functors (j, i) -> value.  w_cut > 0: band mode (central differences where
// phi < w_cut); w_cut == 0: legacy mode (phi <= 0, one-sided differences next to the interface).
template <class FX, class FP>
__device__ __forceinline__ void solid_stress_cell(const FX &X1, const FX &X2, const FP &PHI, int j, int i, int Ny,
                                                  int Nx, double dx, double dy, double mu_s, double kappa,
                                                  double w_cut, double detg_clamp, int isochoric, double &oxx,
                                                  double &oxy, double &oyy, double &oJ)
{
    oxx = 0.0; oxy = 0.0; oyy = 0.0; oJ = 1.0;
    if (!(i >= 1 && i < Nx - 1 && j >= 1 && j < Ny - 1)) return;
    const double ph = PHI(j, i);
    const bool band = (w_cut > 0.0) ? (ph < w_cut) : (ph <= 0.0);
    if (!band) return;
    const double inv_2dx = 1.0 / (2.0 * dx), inv_2dy = 1.0 / (2.0 * dy);
    const double x1c = X1(j, i), x2c = X2(j, i);
    const double x1l = X1(j, i - 1), x1r = X1(j, i + 1), x2l = X2(j, i - 1), x2r = X2(j, i + 1);
    const double x1b = X1(j - 1, i), x1t = X1(j + 1, i), x2b = X2(j - 1, i), x2t = X2(j + 1, i);
    double g11, g21, g12, g22;
    if (w_cut > 0.0) {
        g11 = (x1r - x1l) * inv_2dx;
        g21 = (x2r - x2l) * inv_2dx;
        g12 = (x1t - x1b) * inv_2dy;
        g22 = (x2t - x2b) * inv_2dy;
    } else {
        const bool lf = PHI(j, i - 1) > 0.0, rf = PHI(j, i + 1) > 0.0;
        if (lf && !rf) {
            g11 = (x1r - x1c) / dx;
            g21 = (x2r - x2c) / dx;
        } else if (rf && !lf) {
            g11 = (x1c - x1l) / dx;
            g21 = (x2c - x2l) / dx;
        } else {
            g11 = (x1r - x1l) * inv_2dx;
            g21 = (x2r - x2l) * inv_2dx;
        }
        const bool bf = PHI(j - 1, i) > 0.0, tf = PHI(j + 1, i) > 0.0;
        if (bf && !tf) {
            g12 = (x1t - x1c) / dy;
            g22 = (x2t - x2c) / dy;
        } else if (tf && !bf) {
            g12 = (x1c - x1b) / dy;
            g22 = (x2c - x2b) / dy;
        } else {
            g12 = (x1t - x1b) * inv_2dy;
            g22 = (x2t - x2b) * inv_2dy;
        }
    }
    double detG = g11 * g22 - g12 * g21;
    if (fabs(detG) < 1e-10) return;
    if (detg_clamp > 0.0) {
        const double lo = 1.0 / detg_clamp;
        if (detG < lo) detG = lo;
        else if (detG > detg_clamp) detG = detg_clamp;
    }
    const double f11 = g22 / detG, f12 = -g12 / detG, f21 = -g21 / detG, f22 = g11 / detG;
    const double b11 = f11 * f11 + f12 * f12;
    const double b12 = f11 * f21 + f12 * f22;
    const double b22 = f21 * f21 + f22 * f22;
    const double jv = 1.0 / detG;
    oJ = jv;
    const double vol = kappa * (jv - 1.0);
    if (isochoric) {
        const double trh = 0.5 * (b11 + b22), jm2 = 1.0 / (jv * jv);
        oxx = mu_s * jm2 * (b11 - trh) + vol;
        oxy = mu_s * jm2 * b12;
        oyy = mu_s * jm2 * (b22 - trh) + vol;
    } else {
        oxx = mu_s * b11 + vol;
        oxy = mu_s * b12;
        oyy = mu_s * b22 + vol;
    }
}

// ---- block-wide reductions (deterministic tree) ----------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Sum over a block of up to 1024 threads; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *sh /* >=32 doubles */)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

}  // namespace rmt

// extrap.cu -- narrow-band least-squares extrapolation of the reference map.
//
// Replaces pyRMT/functions.py:48-163 (extrapolate_reference_map) and
// pyRMT/utils.py:134-167 (fast_solve_3x3).
//
// The reference is a SERIAL raster sweep: a fitted cell is marked known at once
// and later cells of the same layer gather it.  The fit
// is ill-conditioned (absolute coordinates + Cramer's rule) and amplifies 1-ulp
// differences to 1e-6 at 4096^2 (H2), so this kernel reproduces the reference
// BIT FOR BIT:
//   * same gather order (jj-major, ii-minor), same sequential accumulation,
//     same expression trees; this file is compiled with -fmad=false;
//   * exp() is glibc's table-driven algorithm with the FMA placement of its
//     x86-64 FMA build (exp_table.inc), bit-identical to host libm on [-1,0];
//   * IEEE division / sqrt (CUDA double defaults).
// Parallelism comes from a ROW-PIPELINED dataflow sweep: one warp owns one grid
// row at a time and walks its targets left to right; before fitting (j,i) it
// waits until rows j-4..j-1 have finished every target with column <= i+4
// (per-row progress counters, release/acquire through L2).  That is exactly the
// dependency set of the serial sweep -- a target depends on same-layer targets
// in rows j-4..j-1 (|di|<=4) and in row j to its left -- so discs and rows
// overlap while every fit sees the same inputs as the serial reference.
// Inside a warp the 81 window cells are evaluated in parallel (weights,
// products); only the 12 running sums are accumulated in order.
#include "common.cuh"
#include "../../include/rmt_b200.h"
#include <cooperative_groups.h>
#include <limits.h>

namespace {

// global (L1-cached) rather than __constant__: lanes index it divergently
__device__ const unsigned long long c_exp_tab[256] = {
#include "exp_table.inc"
};

// glibc exp(), x in the range the weights use ([-1, 0]); no special cases are
// needed there (|x| < 2^-54 -> 1+x is kept for completeness).
__device__ __forceinline__ double exp_glibc(double x)
{
    if (fabs(x) < 0x1p-54) return 1.0 + x;
    const double InvLn2N = 0x1.71547652b82fep0 * 128, Shift = 0x1.8p52;
    const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3;
    const double C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    double z = InvLn2N * x;
    double kd = z + Shift;
    unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd -= Shift;
    double r = fma(kd, NegLn2loN, fma(kd, NegLn2hiN, x));
    unsigned idx = 2u * (unsigned)(ki % 128);
    unsigned long long top = ki << 45;
    double tail = __longlong_as_double((long long)__ldg(c_exp_tab + idx));
    unsigned long long sbits = __ldg(c_exp_tab + idx + 1) + top;
    double r2 = r * r;
    double tmp = fma(r2 * r2, fma(r, C5, C4), fma(r2, fma(r, C3, C2), tail + r));
    double scale = __longlong_as_double((long long)sbits);
    return fma(scale, tmp, scale);
}

__global__ void k_exp_probe(const double *__restrict__ x, double *__restrict__ y, long n)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
        y[k] = exp_glibc(x[k]);
}

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// state byte per cell: 0 unknown, 1 known, 2 target of the current layer
constexpr unsigned char ST_UNKNOWN = 0, ST_KNOWN = 1, ST_TARGET = 2;

// seed: copies of X1/X2, known = (phi < 0)          functions.py:69-74
__global__ void k_ext_seed(const double *__restrict__ X1, const double *__restrict__ X2,
                           const double *__restrict__ phi, double *__restrict__ X1e,
                           double *__restrict__ X2e, unsigned char *__restrict__ st, long n)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        X1e[k] = X1[k];
        X2e[k] = X2[k];
        st[k] = (phi[k] < 0.0) ? ST_KNOWN : ST_UNKNOWN;
    }
}

// frontier of the current layer: unknown interior cells with a known 3x3
// neighbour (functions.py:79-90); one warp per row, also counts them.
__global__ void k_ext_flag(unsigned char *__restrict__ st, int *__restrict__ row_cnt, int Ny, int Nx)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < Ny; j += nwarp) {
        int cnt = 0;
        if (j >= 1 && j < Ny - 1) {
            const unsigned char *r0 = st + (size_t)(j - 1) * Nx, *r1 = r0 + Nx, *r2 = r1 + Nx;
            for (int base = 0; base < Nx; base += 32) {
                int i = base + lane;
                bool tgt = false;
                if (i >= 1 && i < Nx - 1 && r1[i] != ST_KNOWN) {
                    // neighbours in rows j-1 / j+1 can only be KNOWN or not; a
                    // concurrent UNKNOWN<->TARGET rewrite there is harmless.
                    tgt = (r0[i - 1] == ST_KNOWN) | (r0[i] == ST_KNOWN) | (r0[i + 1] == ST_KNOWN) |
                          (r1[i - 1] == ST_KNOWN) | (r1[i + 1] == ST_KNOWN) |
                          (r2[i - 1] == ST_KNOWN) | (r2[i] == ST_KNOWN) | (r2[i + 1] == ST_KNOWN);
                }
                if (i < Nx && r1[i] != ST_KNOWN) st[(size_t)j * Nx + i] = tgt ? ST_TARGET : ST_UNKNOWN;
                cnt += __popc(__ballot_sync(0xffffffffu, tgt));
            }
        }
        if (lane == 0) row_cnt[j] = cnt;
    }
}

// exclusive scan of the per-row counts -> row_off[0..Ny]; resets the per-row
// progress counters (rows without targets are "finished").
__global__ void k_ext_scan(const int *__restrict__ row_cnt, int *__restrict__ row_off,
                           int *__restrict__ prog, int Ny)
{
    __shared__ int sh[1024];
    int per = (Ny + blockDim.x - 1) / blockDim.x;
    int lo = threadIdx.x * per, hi = min(lo + per, Ny);
    int s = 0;
    for (int j = lo; j < hi; ++j) s += row_cnt[j];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
        int v = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - s;
    for (int j = lo; j < hi; ++j) {
        row_off[j] = run;
        int c = row_cnt[j];
        prog[j] = c ? 0 : INT_MAX;
        run += c;
    }
    if (threadIdx.x == blockDim.x - 1) row_off[Ny] = sh[threadIdx.x];
}

// column indices of each row's targets, in increasing column order
__global__ void k_ext_fill(const unsigned char *__restrict__ st, const int *__restrict__ row_off,
                           int *__restrict__ tcol, int Ny, int Nx)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp + 1; j < Ny - 1; j += nwarp) {
        int off = row_off[j];
        if (row_off[j + 1] == off) continue;
        const unsigned char *r1 = st + (size_t)j * Nx;
        for (int base = 0; base < Nx; base += 32) {
            int i = base + lane;
            bool tgt = (i < Nx) && (r1[i] == ST_TARGET);
            unsigned m = __ballot_sync(0xffffffffu, tgt);
            if (tgt) tcol[off + __popc(m & ((1u << lane) - 1u))] = i;
            off += __popc(m);
        }
    }
}

constexpr int FIT_WARPS = 4;           // warps per CTA in the sweep kernel (31 KB of products)
constexpr int WIN = 81;                // 9x9 window
constexpr int NACC = 12;               // running sums: B1[3], B2[3], A00 A01 A02 A11 A12 A22

// The row-pipelined sweep (functions.py:95-161).
__global__ void __launch_bounds__(FIT_WARPS * 32)
k_ext_sweep(double *__restrict__ X1e, double *__restrict__ X2e, unsigned char *__restrict__ st,
            const int *__restrict__ row_off, const int *__restrict__ tcol, int *__restrict__ prog,
            int Ny, int Nx, double dx, double dy, double r2)
{
    __shared__ double prod_s[FIT_WARPS][WIN * NACC];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * FIT_WARPS + wib, nw = gridDim.x * FIT_WARPS;
    double *P = prod_s[wib];

    for (int j = 1 + gw; j < Ny - 1; j += nw) {
        const int t0 = row_off[j], t1 = row_off[j + 1];
        for (int t = t0; t < t1; ++t) {
            const int i = tcol[t];
            // ---- wait for the rows above to pass column i+4 ---------------
            if (lane < 4) {
                int jr = j - 1 - lane;
                if (jr >= 1)
                    while (ld_acquire(prog + jr) <= i + 4) { /* spin */ }
            }
            __syncwarp();
            __threadfence();

            const double x0 = dx * i, y0 = dy * j;
            unsigned vmask[3];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                int n = lane + 32 * s;
                bool valid = false;
                if (n < WIN) {
                    int dj = n / 9 - 4, di = n % 9 - 4;
                    int jj = j + dj, ii = i + di;
                    if (jj >= 0 && jj < Ny && ii >= 0 && ii < Nx) {
                        size_t cc = (size_t)jj * Nx + ii;
                        if (__ldcg(st + cc) == ST_KNOWN) {
                            double xi = dx * ii, yi = dy * jj;
                            double ex = xi - x0, ey = yi - y0;
                            double dist_sq = ex * ex + ey * ey;
                            if (dist_sq <= r2) {
                                valid = true;
                                double w = exp_glibc(-dist_sq / r2);
                                double v1 = __ldcg(X1e + cc), v2 = __ldcg(X2e + cc);
                                double wx = w * xi, wy = w * yi;
                                double *p = P + n * NACC;
                                p[0] = w * v1;  p[1] = wx * v1;  p[2] = wy * v1;
                                p[3] = w * v2;  p[4] = wx * v2;  p[5] = wy * v2;
                                p[6] = w;       p[7] = wx;       p[8] = wy;
                                p[9] = wx * xi; p[10] = wx * yi; p[11] = wy * yi;
                            }
                        }
                    }
                }
                vmask[s] = __ballot_sync(0xffffffffu, valid);
            }
            __syncwarp();
            const int count = __popc(vmask[0]) + __popc(vmask[1]) + __popc(vmask[2]);
            bool fitted = false;
            if (count >= 3) {
                // ---- ordered accumulation: lane a owns running sum a -------
                double acc = 0.0;
                if (lane < NACC) {
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        unsigned m = vmask[s];
                        while (m) {
                            int bit = __ffs(m) - 1;
                            m &= m - 1;
                            acc += P[(32 * s + bit) * NACC + lane];
                        }
                    }
                }
                double B10 = __shfl_sync(0xffffffffu, acc, 0), B11 = __shfl_sync(0xffffffffu, acc, 1);
                double B12 = __shfl_sync(0xffffffffu, acc, 2), B20 = __shfl_sync(0xffffffffu, acc, 3);
                double B21 = __shfl_sync(0xffffffffu, acc, 4), B22 = __shfl_sync(0xffffffffu, acc, 5);
                double A00 = __shfl_sync(0xffffffffu, acc, 6), A01 = __shfl_sync(0xffffffffu, acc, 7);
                double A02 = __shfl_sync(0xffffffffu, acc, 8), A11 = __shfl_sync(0xffffffffu, acc, 9);
                double A12 = __shfl_sync(0xffffffffu, acc, 10), A22 = __shfl_sync(0xffffffffu, acc, 11);
                const double A10 = A01, A20 = A02, A21 = A12;
                double det = (A00 * (A11 * A22 - A12 * A21) - A01 * (A10 * A22 - A12 * A20) +
                              A02 * (A10 * A21 - A11 * A20));
                if (fabs(det) > 1e-10) {
                    fitted = true;
                    if (lane < 2) {
                        double b0 = lane ? B20 : B10, b1 = lane ? B21 : B11, b2 = lane ? B22 : B12;
                        double inv_det = 1.0 / det;
                        double cx = (b0 * (A11 * A22 - A12 * A21) - A01 * (b1 * A22 - A12 * b2) +
                                     A02 * (b1 * A21 - A11 * b2)) * inv_det;
                        double cy = (A00 * (b1 * A22 - A12 * b2) - b0 * (A10 * A22 - A12 * A20) +
                                     A02 * (A10 * b2 - b1 * A20)) * inv_det;
                        double cz = (A00 * (A11 * b2 - b1 * A21) - A01 * (A10 * b2 - b1 * A20) +
                                     b0 * (A10 * A21 - A11 * A20)) * inv_det;
                        double val = cx + cy * x0 + cz * y0;
                        size_t c = (size_t)j * Nx + i;
                        if (lane) __stcg(X2e + c, val);
                        else __stcg(X1e + c, val);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (fitted) __stcg(st + (size_t)j * Nx + i, ST_KNOWN);
                __threadfence();
                st_release(prog + j, (t + 1 < t1) ? tcol[t + 1] : INT_MAX);
            }
            __syncwarp();
        }
    }
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

}  // namespace

extern "C" {

long rmt_extrapolate_workspace_bytes(int Ny, int Nx)
{
    size_t ncell = (size_t)Ny * (size_t)Nx;
    size_t st = (ncell + 255) & ~(size_t)255;
    size_t rows = ((size_t)(3 * (Ny + 1)) * sizeof(int) + 255) & ~(size_t)255;
    size_t cols = ncell * sizeof(int);
    return (long)(st + rows + cols);
}

int rmt_extrapolate(const double *X1, const double *X2, const double *phi, double *X1e, double *X2e,
                    int Ny, int Nx, double dx, double dy, int max_layers, void *workspace,
                    void *stream)
{
    if (!X1 || !X2 || !phi || !X1e || !X2e || !workspace || Ny < 3 || Nx < 3) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    size_t ncell = (size_t)Ny * (size_t)Nx;
    unsigned char *st = (unsigned char *)workspace;
    size_t st_bytes = (ncell + 255) & ~(size_t)255;
    int *row_cnt = (int *)((char *)workspace + st_bytes);
    int *row_off = row_cnt + (Ny + 1);
    int *prog = row_off + (Ny + 1);
    size_t rows_bytes = ((size_t)(3 * (Ny + 1)) * sizeof(int) + 255) & ~(size_t)255;
    int *tcol = (int *)((char *)workspace + st_bytes + rows_bytes);

    // stencil_radius_sq = (4*sqrt(dx**2+dy**2))**2, functions.py:76 (no contraction)
    volatile double dx2 = dx * dx, dy2 = dy * dy;
    volatile double ssum = dx2 + dy2;
    volatile double rr = 4.0 * sqrt(ssum);
    double r2 = rr * rr;

    k_ext_seed<<<flat_blocks((long)ncell), 256, 0, s>>>(X1, X2, phi, X1e, X2e, st, (long)ncell);
    RMT_LAUNCH_CHECK();

    static int sweep_blocks = 0;
    if (!sweep_blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        RMT_CUDA(cudaGetDevice(&dev));
        RMT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        RMT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ext_sweep, FIT_WARPS * 32, 0));
        if (per_sm < 1) return RMT_EINVAL;
        sweep_blocks = sms * per_sm;
    }
    int row_warps_blocks = rmt_cdiv((long)Ny * 32, 256);
    if (row_warps_blocks > 148 * 8) row_warps_blocks = 148 * 8;

    for (int layer = 0; layer < max_layers; ++layer) {
        k_ext_flag<<<row_warps_blocks, 256, 0, s>>>(st, row_cnt, Ny, Nx);
        RMT_LAUNCH_CHECK();
        k_ext_scan<<<1, 1024, 0, s>>>(row_cnt, row_off, prog, Ny);
        RMT_LAUNCH_CHECK();
        k_ext_fill<<<row_warps_blocks, 256, 0, s>>>(st, row_off, tcol, Ny, Nx);
        RMT_LAUNCH_CHECK();
        // all CTAs must be co-resident (warps wait on each other): cooperative launch
        int blocks = sweep_blocks;
        int need = rmt_cdiv(Ny, FIT_WARPS);
        if (blocks > need) blocks = need;
        void *args[] = {&X1e, &X2e, &st, &row_off, &tcol, &prog, &Ny, &Nx, &dx, &dy, &r2};
        RMT_CUDA(cudaLaunchCooperativeKernel((void *)k_ext_sweep, dim3(blocks), dim3(FIT_WARPS * 32),
                                             args, 0, s));
    }
    return RMT_OK;
}

// device exp() on an array -- lets the tests prove bit-equality with host libm
int rmt_exp_probe(const double *x, double *y, long n, void *stream)
{
    if (!x || !y || n <= 0) return RMT_EINVAL;
    k_exp_probe<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

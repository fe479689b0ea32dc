// fft.cu -- spectral Poisson solves: hand-written DCT-I (Neumann) and FFT
// (periodic) transforms.  No cuFFT.
//
// Replaces (reference file:line):
//   pyRMT/functions.py:1107-1119  _solve_poisson_dct   (scipy.fft.dctn/idctn type 1 -> pocketfft)
//   pyRMT/functions.py:1216-1233  _solve_poisson_fft   (numpy.fft.fft2/ifft2       -> pocketfft)
//   pyRMT/functions.py:1205-1213  _tile_overlap
//
// DCT-I of a line x[0..M] equals the DFT of its even extension of length 2M
// (that is also how pocketfft evaluates it).  The DFT of a real even sequence is
// real, so TWO real lines a, b are transformed at once as the complex even
// sequence a + i b: Re = DCT-I(a), Im = DCT-I(b), no split post-processing.
//
// Fast path (M a power of two): the 2M-point complex FFT of a line pair lives in
// shared memory (split re/im planes, padded against bank conflicts) and runs as
// in-place radix-4 (+ one radix-2) butterflies:
//   rows:    load 2 rows -> DIF FFT -> gather out of digit-reversed order -> store
//   columns: load 2 columns -> DIF FFT -> scale by 1/(4 Mx My eig) in scrambled
//            order -> DIT FFT (consumes the scrambled order) -> store in place
// so the 2-D solve is three passes over HBM (rows, columns fwd+inv, rows) and
// the digit reversal is never materialised.  The inverse DCT-I is the forward
// one scaled by 1/(2M) (functions.py:1117 idctn).
// Other sizes (N <= RMT_DENSE_MAX) use dense cosine/sine matrices and a small
// fp64 GEMM kernel -- O(N^3) but exact to rounding and only used on the small
// awkward grids of the reference's benchmarks (N = 128 -> 2*127 and 127).
#include "common.cuh"
#include "../../include/rmt_b200.h"

#include <cmath>
#include <vector>

using namespace rmt;

namespace {

constexpr int kMaxSmemL = 8192;   // longest complex FFT held in one CTA's shared memory

__host__ __device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

// Storage position of frequency k after the in-place DIF (radices 4,...,4[,2]).
__device__ __forceinline__ int rev_pos(int k, int L)
{
    int p = 0, n = L;
    while (n >= 4) {
        n >>= 2;
        p += (k & 3) * n;
        k >>= 2;
    }
    if (n == 2) p += (k & 1);
    return p;
}

struct cplx {
    double x, y;
};
__device__ __forceinline__ cplx cmul(cplx a, double2 w)
{
    return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x};
}

// radix-2 stage on adjacent pairs (the n = 2 level; no twiddles)
__device__ __forceinline__ void stage_radix2(double *re, double *im, int L)
{
    for (int b = threadIdx.x; b < (L >> 1); b += blockDim.x) {
        int p0 = padi(2 * b), p1 = padi(2 * b + 1);
        double ar = re[p0], ai = im[p0], br = re[p1], bi = im[p1];
        re[p0] = ar + br; im[p0] = ai + bi;
        re[p1] = ar - br; im[p1] = ai - bi;
    }
    __syncthreads();
}

// Forward DFT, natural order in -> digit-reversed order out.  tw[m] = exp(-2 pi i m / L).
__device__ void fft_dif(double *re, double *im, int L, const double2 *__restrict__ tw)
{
    int n = L;
    for (; n >= 4; n >>= 2) {
        const int q = n >> 2, tstep = L / n;
        for (int b = threadIdx.x; b < (L >> 2); b += blockDim.x) {
            const int k = b & (q - 1), g = (b - k) << 2;
            const int p0 = padi(g + k), p1 = padi(g + k + q), p2 = padi(g + k + 2 * q), p3 = padi(g + k + 3 * q);
            cplx a0{re[p0], im[p0]}, a1{re[p1], im[p1]}, a2{re[p2], im[p2]}, a3{re[p3], im[p3]};
            cplx t0{a0.x + a2.x, a0.y + a2.y}, t1{a0.x - a2.x, a0.y - a2.y};
            cplx t2{a1.x + a3.x, a1.y + a3.y}, t3{a1.y - a3.y, -(a1.x - a3.x)};   // -i (a1 - a3)
            cplx y0{t0.x + t2.x, t0.y + t2.y}, y2{t0.x - t2.x, t0.y - t2.y};
            cplx y1{t1.x + t3.x, t1.y + t3.y}, y3{t1.x - t3.x, t1.y - t3.y};
            if (k) {
                y1 = cmul(y1, __ldg(tw + k * tstep));
                y2 = cmul(y2, __ldg(tw + 2 * k * tstep));
                y3 = cmul(y3, __ldg(tw + 3 * k * tstep));
            }
            re[p0] = y0.x; im[p0] = y0.y;
            re[p1] = y1.x; im[p1] = y1.y;
            re[p2] = y2.x; im[p2] = y2.y;
            re[p3] = y3.x; im[p3] = y3.y;
        }
        __syncthreads();
    }
    if (n == 2) stage_radix2(re, im, L);
}

// Forward DFT, digit-reversed order in (as left by fft_dif) -> natural order out.
__device__ void fft_dit(double *re, double *im, int L, const double2 *__restrict__ tw)
{
    int n = L;
    while (n >= 4) n >>= 2;          // n = 2 iff log2(L) is odd
    int first = 4;
    if (n == 2) {
        stage_radix2(re, im, L);
        first = 8;
    }
    for (n = first; n <= L; n <<= 2) {
        const int q = n >> 2, tstep = L / n;
        for (int b = threadIdx.x; b < (L >> 2); b += blockDim.x) {
            const int k = b & (q - 1), g = (b - k) << 2;
            const int p0 = padi(g + k), p1 = padi(g + k + q), p2 = padi(g + k + 2 * q), p3 = padi(g + k + 3 * q);
            cplx b0{re[p0], im[p0]}, b1{re[p1], im[p1]}, b2{re[p2], im[p2]}, b3{re[p3], im[p3]};
            if (k) {
                b1 = cmul(b1, __ldg(tw + k * tstep));
                b2 = cmul(b2, __ldg(tw + 2 * k * tstep));
                b3 = cmul(b3, __ldg(tw + 3 * k * tstep));
            }
            cplx t0{b0.x + b2.x, b0.y + b2.y}, t1{b0.x - b2.x, b0.y - b2.y};
            cplx t2{b1.x + b3.x, b1.y + b3.y}, t3{b1.y - b3.y, -(b1.x - b3.x)};
            re[p0] = t0.x + t2.x; im[p0] = t0.y + t2.y;
            re[p1] = t1.x + t3.x; im[p1] = t1.y + t3.y;
            re[p2] = t0.x - t2.x; im[p2] = t0.y - t2.y;
            re[p3] = t1.x - t3.x; im[p3] = t1.y - t3.y;
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------- DCT-I, rows
// One CTA transforms rows 2*blockIdx.x and 2*blockIdx.x+1 (unnormalised DCT-I
// times `scale`).  `partial` (optional) receives the CTA's sum of outputs.
__global__ void k_dct_rows(const double *__restrict__ in, double *__restrict__ out, int Ny, int Nx,
                           const double2 *__restrict__ tw, double scale, double *__restrict__ partial)
{
    extern __shared__ double sm[];
    const int M = Nx - 1, L = 2 * M;
    double *re = sm, *im = sm + padi(L);
    const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
    const bool has1 = r1 < Ny;
    const double *a = in + (size_t)r0 * Nx, *b = in + (size_t)(has1 ? r1 : r0) * Nx;
    for (int m = threadIdx.x; m <= M; m += blockDim.x) {
        double va = __ldg(a + m), vb = has1 ? __ldg(b + m) : 0.0;
        re[padi(m)] = va;
        im[padi(m)] = vb;
        if (m > 0 && m < M) {
            re[padi(L - m)] = va;
            im[padi(L - m)] = vb;
        }
    }
    __syncthreads();
    fft_dif(re, im, L, tw);
    double s = 0.0;
    double *oa = out + (size_t)r0 * Nx, *ob = out + (size_t)r1 * Nx;
    for (int k = threadIdx.x; k <= M; k += blockDim.x) {
        int p = padi(rev_pos(k, L));
        double va = re[p] * scale;
        oa[k] = va;
        s += va;
        if (has1) {
            double vb = im[p] * scale;
            ob[k] = vb;
            s += vb;
        }
    }
    if (partial) {
        __shared__ double red[32];
        s = block_sum(s, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

// -------------------------------------------- DCT-I, columns: fwd, 1/eig, inverse
// In place on T (Ny, Nx): columns 2*blockIdx.x and +1.
__global__ void k_dct_cols_solve(double *__restrict__ T, const double *__restrict__ eig, int Ny, int Nx,
                                 const double2 *__restrict__ tw, double scale)
{
    extern __shared__ double sm[];
    const int M = Ny - 1, L = 2 * M;
    double *re = sm, *im = sm + padi(L);
    const int c0 = 2 * blockIdx.x;
    const bool has1 = (c0 + 1) < Nx;
    for (int m = threadIdx.x; m <= M; m += blockDim.x) {
        const double *row = T + (size_t)m * Nx + c0;
        double va = row[0], vb = has1 ? row[1] : 0.0;
        re[padi(m)] = va;
        im[padi(m)] = vb;
        if (m > 0 && m < M) {
            re[padi(L - m)] = va;
            im[padi(L - m)] = vb;
        }
    }
    __syncthreads();
    fft_dif(re, im, L, tw);
    for (int k = threadIdx.x; k <= M; k += blockDim.x) {
        const double *er = eig + (size_t)k * Nx + c0;
        double fa = scale / __ldg(er), fb = has1 ? scale / __ldg(er + 1) : 0.0;
        int p = padi(rev_pos(k, L));
        re[p] *= fa;
        im[p] *= fb;
        if (k > 0 && k < M) {
            int p2 = padi(rev_pos(L - k, L));
            re[p2] *= fa;
            im[p2] *= fb;
        }
    }
    __syncthreads();
    fft_dit(re, im, L, tw);
    for (int m = threadIdx.x; m <= M; m += blockDim.x) {
        double *row = T + (size_t)m * Nx + c0;
        row[0] = re[padi(m)];
        if (has1) row[1] = im[padi(m)];
    }
}

__global__ void k_sum_final(const double *__restrict__ part, int n, double *__restrict__ out)
{
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) s += part[k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// ------------------------------------------------------------ dense fallback
// C (M x N, ldc) = alpha * A (M x K, lda) * op(B) + beta * C ;  op(B) = B (K x N, ldb)
// or B^T with B stored (N x K, ldb).  16x16 tiles, one output per thread.
__global__ void __launch_bounds__(256)
k_gemm(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int M, int N,
       int K, int lda, int ldb, int ldc, int transB, double alpha, double beta)
{
    __shared__ double sA[16][17], sB[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        int ka = k0 + tx;
        sA[ty][tx] = (row < M && ka < K) ? A[(size_t)row * lda + ka] : 0.0;
        if (transB) {
            int n = blockIdx.x * 16 + ty, kb = k0 + tx;      // B[n][k]
            sB[tx][ty] = (n < N && kb < K) ? B[(size_t)n * ldb + kb] : 0.0;
        } else {
            int kb = k0 + ty;                                // B[k][col]
            sB[ty][tx] = (kb < K && col < N) ? B[(size_t)kb * ldb + col] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += sA[ty][k] * sB[k][tx];
        __syncthreads();
    }
    if (row < M && col < N) {
        size_t o = (size_t)row * ldc + col;
        C[o] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * C[o];
    }
}

// x[j][i] *= scale / eig[j][i]  (and zero where null) on an (R x Cn) block
__global__ void k_spectral_divide(double *__restrict__ xr, double *__restrict__ xi,
                                  const double *__restrict__ eig, const unsigned char *__restrict__ null_mask,
                                  long n, double scale)
{
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        double f = scale / eig[k];
        if (null_mask && null_mask[k]) f = 0.0;
        xr[k] *= f;
        if (xi) xi[k] *= f;
    }
}

// _tile_overlap (functions.py:1205-1213) in place on a full (Ny, Nx) array whose
// [:-1,:-1] block holds the reduced solution.
__global__ void k_tile_overlap(double *__restrict__ s, int Ny, int Nx)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Ny - 1) s[(size_t)t * Nx + (Nx - 1)] = s[(size_t)t * Nx];
    if (t < Nx) {
        int src = (t == Nx - 1) ? 0 : t;
        s[(size_t)(Ny - 1) * Nx + t] = s[src];
    }
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }
inline bool is_pow2(int v) { return v >= 2 && (v & (v - 1)) == 0; }

int gemm(const double *A, const double *B, double *C, int M, int N, int K, int lda, int ldb, int ldc,
         int transB, double alpha, double beta, cudaStream_t s)
{
    dim3 grd(rmt_cdiv(N, 16), rmt_cdiv(M, 16));
    k_gemm<<<grd, 256, 0, s>>>(A, B, C, M, N, K, lda, ldb, ldc, transB, alpha, beta);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // namespace

struct rmt_poisson_plan {
    int Ny, Nx, kind;
    bool fast;
    // fast path
    double2 *tw_x, *tw_y;      // exp(-2 pi i m / L), L = 2(N-1) (DCT) per direction
    int Lx, Ly;
    // dense path
    double *Cx, *Cy, *Sx, *Sy; // cosine (and sine, periodic) matrices
    int nx, ny;                // transform sizes (DCT: N ; periodic: N-1)
    double *w[4];              // work arrays
    double *red;               // reduction partials / stats
};

namespace {

template <class T>
int upload(T **dst, const std::vector<T> &h)
{
    RMT_CUDA(cudaMalloc((void **)dst, h.size() * sizeof(T)));
    RMT_CUDA(cudaMemcpy(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return RMT_OK;
}

int make_twiddles(double2 **dst, int L)
{
    std::vector<double2> h((size_t)L);
    const double pi = 3.14159265358979323846;
    for (int m = 0; m < L; ++m) {
        // reduce to the first octant for uniformly small error
        double ang = 2.0 * pi * (double)m / (double)L;
        h[m].x = std::cos(ang);
        h[m].y = -std::sin(ang);
    }
    // exact values on the axes
    h[0] = {1.0, 0.0};
    if (L % 4 == 0) {
        h[L / 4] = {0.0, -1.0};
        h[L / 2] = {-1.0, 0.0};
        h[3 * L / 4] = {0.0, 1.0};
    } else if (L % 2 == 0) {
        h[L / 2] = {-1.0, 0.0};
    }
    return upload(dst, h);
}

// DCT-I matrix: C[k][n] = w_n cos(pi k n / M), w = 1 at n in {0, M}, else 2
int make_dct_matrix(double **dst, int N)
{
    const int M = N - 1;
    const double pi = 3.14159265358979323846;
    std::vector<double> h((size_t)N * N);
    for (int k = 0; k < N; ++k)
        for (int n = 0; n < N; ++n) {
            long r = ((long)k * n) % (2L * M);
            double w = (n == 0 || n == M) ? 1.0 : 2.0;
            h[(size_t)k * N + n] = w * std::cos(pi * (double)r / (double)M);
        }
    return upload(dst, h);
}

int make_dft_matrices(double **c, double **s, int m)
{
    const double pi = 3.14159265358979323846;
    std::vector<double> hc((size_t)m * m), hs((size_t)m * m);
    for (int k = 0; k < m; ++k)
        for (int n = 0; n < m; ++n) {
            long r = ((long)k * n) % m;
            double ang = 2.0 * pi * (double)r / (double)m;
            hc[(size_t)k * m + n] = std::cos(ang);
            hs[(size_t)k * m + n] = std::sin(ang);
        }
    int e = upload(c, hc);
    if (e) return e;
    return upload(s, hs);
}

int fft_threads(int L) { return L >= 4096 ? 512 : (L >= 1024 ? 256 : (L >= 256 ? 64 : 32)); }

}  // namespace

extern "C" {

int rmt_poisson_plan_create(int Ny, int Nx, int kind, rmt_poisson_plan **out)
{
    if (!out || Ny < 4 || Nx < 4 || kind < 0 || kind > 1) return RMT_EINVAL;
    rmt_poisson_plan *P = new rmt_poisson_plan();
    *P = rmt_poisson_plan{};
    P->Ny = Ny; P->Nx = Nx; P->kind = kind;
    int e = RMT_OK;
    const size_t ncell = (size_t)Ny * Nx;
    if (kind == 0) {
        P->Lx = 2 * (Nx - 1); P->Ly = 2 * (Ny - 1);
        P->fast = is_pow2(Nx - 1) && is_pow2(Ny - 1) && P->Lx <= kMaxSmemL && P->Ly <= kMaxSmemL &&
                  P->Lx >= 8 && P->Ly >= 8;
        P->nx = Nx; P->ny = Ny;
    } else {
        P->fast = false;   // periodic shared-memory path: see rmt_poisson_solve_fft
        P->nx = Nx - 1; P->ny = Ny - 1;
    }
    if (!P->fast && (P->nx > RMT_DENSE_MAX || P->ny > RMT_DENSE_MAX)) {
        delete P;
        return -2;
    }
    if (P->fast) {
        e = make_twiddles(&P->tw_x, P->Lx);
        if (!e) e = make_twiddles(&P->tw_y, P->Ly);
        if (!e) e = (int)cudaMalloc((void **)&P->w[0], ncell * sizeof(double));
        int smem = 2 * padi(P->Lx > P->Ly ? P->Lx : P->Ly) * (int)sizeof(double);
        if (!e) e = (int)cudaFuncSetAttribute(k_dct_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (!e) e = (int)cudaFuncSetAttribute(k_dct_cols_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    } else if (kind == 0) {
        e = make_dct_matrix(&P->Cx, Nx);
        if (!e) e = make_dct_matrix(&P->Cy, Ny);
        for (int k = 0; k < 2 && !e; ++k) e = (int)cudaMalloc((void **)&P->w[k], ncell * sizeof(double));
    } else {
        e = make_dft_matrices(&P->Cx, &P->Sx, P->nx);
        if (!e) e = make_dft_matrices(&P->Cy, &P->Sy, P->ny);
        for (int k = 0; k < 4 && !e; ++k)
            e = (int)cudaMalloc((void **)&P->w[k], (size_t)P->nx * P->ny * sizeof(double));
    }
    if (!e) e = (int)cudaMalloc((void **)&P->red, (size_t)(rmt_reduce_workspace_doubles() + Ny + 16) * sizeof(double));
    if (e) {
        rmt_poisson_plan_destroy(P);
        return e;
    }
    *out = P;
    return RMT_OK;
}

void rmt_poisson_plan_destroy(rmt_poisson_plan *P)
{
    if (!P) return;
    cudaFree(P->tw_x); cudaFree(P->tw_y);
    cudaFree(P->Cx); cudaFree(P->Cy); cudaFree(P->Sx); cudaFree(P->Sy);
    for (int k = 0; k < 4; ++k) cudaFree(P->w[k]);
    cudaFree(P->red);
    delete P;
}

int rmt_poisson_plan_is_fast(const rmt_poisson_plan *P) { return (P && P->fast) ? 1 : 0; }

int rmt_poisson_solve_dct(rmt_poisson_plan *P, const double *rhs, const double *eig, double *sol,
                          double *sum_out, void *stream)
{
    if (!P || P->kind != 0 || !rhs || !eig || !sol) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int Ny = P->Ny, Nx = P->Nx;
    const long ncell = (long)Ny * Nx;
    const double scale = 1.0 / (4.0 * (double)(Nx - 1) * (double)(Ny - 1));
    double *sum_dst = sum_out ? sum_out : P->red + rmt_reduce_workspace_doubles() + Ny + 8;
    if (P->fast) {
        double *T = P->w[0];
        double *partial = P->red;                       // ceil(Ny/2) partial sums
        const int nrow_cta = (Ny + 1) / 2, ncol_cta = (Nx + 1) / 2;
        const int smx = 2 * padi(P->Lx) * (int)sizeof(double), smy = 2 * padi(P->Ly) * (int)sizeof(double);
        k_dct_rows<<<nrow_cta, fft_threads(P->Lx), smx, s>>>(rhs, T, Ny, Nx, P->tw_x, 1.0, nullptr);
        RMT_LAUNCH_CHECK();
        k_dct_cols_solve<<<ncol_cta, fft_threads(P->Ly), smy, s>>>(T, eig, Ny, Nx, P->tw_y, scale);
        RMT_LAUNCH_CHECK();
        k_dct_rows<<<nrow_cta, fft_threads(P->Lx), smx, s>>>(T, sol, Ny, Nx, P->tw_x, 1.0, partial);
        RMT_LAUNCH_CHECK();
        k_sum_final<<<1, 256, 0, s>>>(partial, nrow_cta, sum_dst);
        RMT_LAUNCH_CHECK();
    } else {
        double *W1 = P->w[0], *W2 = P->w[1];
        int e;
        if ((e = gemm(rhs, P->Cx, W1, Ny, Nx, Nx, Nx, Nx, Nx, 1, 1.0, 0.0, s))) return e;   // rows
        if ((e = gemm(P->Cy, W1, W2, Ny, Nx, Ny, Ny, Nx, Nx, 0, 1.0, 0.0, s))) return e;    // columns
        k_spectral_divide<<<flat_blocks(ncell), 256, 0, s>>>(W2, nullptr, eig, nullptr, ncell, scale);
        RMT_LAUNCH_CHECK();
        if ((e = gemm(W2, P->Cx, W1, Ny, Nx, Nx, Nx, Nx, Nx, 1, 1.0, 0.0, s))) return e;
        if ((e = gemm(P->Cy, W1, sol, Ny, Nx, Ny, Ny, Nx, Nx, 0, 1.0, 0.0, s))) return e;
        double *stats = P->red + rmt_reduce_workspace_doubles();
        if ((e = rmt_field_stats(sol, ncell, P->red, stats, stream))) return e;
        if (sum_out) RMT_CUDA(cudaMemcpyAsync(sum_out, stats, sizeof(double), cudaMemcpyDeviceToDevice, s));
        else sum_dst = stats;
    }
    if (!sum_out) return rmt_subtract_mean(sol, sum_dst, ncell, stream);
    return RMT_OK;
}

int rmt_poisson_solve_fft(rmt_poisson_plan *P, const double *rhs, const double *eig,
                          const unsigned char *null_mask, double *sol, double *sum_out, void *stream)
{
    if (!P || P->kind != 1 || !rhs || !eig || !null_mask || !sol) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int Ny = P->Ny, Nx = P->Nx, my = P->ny, mx = P->nx;
    const long nred = (long)my * mx, ncell = (long)Ny * Nx;
    const double scale = 1.0 / ((double)mx * (double)my);
    // The mean removal before the transform (functions.py:1226) only changes the
    // DC mode, which the null mask zeroes (eig[0,0] = 0 exactly) -- skipped.
    double *Zr = P->w[0], *Zi = P->w[1], *Yr = P->w[2], *Yi = P->w[3];
    int e;
    // rows:  Z = r (Cx - i Sx)        (the DFT matrices are symmetric)
    if ((e = gemm(rhs, P->Cx, Zr, my, mx, mx, Nx, mx, mx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(rhs, P->Sx, Zi, my, mx, mx, Nx, mx, mx, 0, -1.0, 0.0, s))) return e;
    // columns:  Y = (Cy - i Sy) Z
    if ((e = gemm(P->Cy, Zr, Yr, my, mx, my, my, mx, mx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(P->Sy, Zi, Yr, my, mx, my, my, mx, mx, 0, 1.0, 1.0, s))) return e;
    if ((e = gemm(P->Cy, Zi, Yi, my, mx, my, my, mx, mx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(P->Sy, Zr, Yi, my, mx, my, my, mx, mx, 0, -1.0, 1.0, s))) return e;
    k_spectral_divide<<<flat_blocks(nred), 256, 0, s>>>(Yr, Yi, eig, null_mask, nred, scale);
    RMT_LAUNCH_CHECK();
    // inverse rows:  W = Y (Cx + i Sx)
    if ((e = gemm(Yr, P->Cx, Zr, my, mx, mx, mx, mx, mx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(Yi, P->Sx, Zr, my, mx, mx, mx, mx, mx, 0, -1.0, 1.0, s))) return e;
    if ((e = gemm(Yr, P->Sx, Zi, my, mx, mx, mx, mx, mx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(Yi, P->Cx, Zi, my, mx, mx, mx, mx, mx, 0, 1.0, 1.0, s))) return e;
    // inverse columns, real part:  out = Cy Wr - Sy Wi   (into sol[:-1,:-1], ld = Nx)
    if ((e = gemm(P->Cy, Zr, sol, my, mx, my, my, mx, Nx, 0, 1.0, 0.0, s))) return e;
    if ((e = gemm(P->Sy, Zi, sol, my, mx, my, my, mx, Nx, 0, -1.0, 1.0, s))) return e;
    int nt = Ny > Nx ? Ny : Nx;
    k_tile_overlap<<<rmt_cdiv(nt, 256), 256, 0, s>>>(sol, Ny, Nx);
    RMT_LAUNCH_CHECK();
    double *stats = P->red + rmt_reduce_workspace_doubles();
    if ((e = rmt_field_stats(sol, ncell, P->red, stats, stream))) return e;
    if (sum_out) {
        RMT_CUDA(cudaMemcpyAsync(sum_out, stats, sizeof(double), cudaMemcpyDeviceToDevice, s));
        return RMT_OK;
    }
    return rmt_subtract_mean(sol, stats, ncell, stream);
}

}  // extern "C"

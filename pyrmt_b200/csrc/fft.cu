// fft.cu -- spectral Poisson solves: hand-written DCT-I (Neumann) and FFT
// (periodic) transforms.  No cuFFT.
//
// Replaces (reference file:line):
//   pyRMT/functions.py:1107-1119  _solve_poisson_dct   (scipy.fft.dctn/idctn type 1 -> pocketfft)
//   pyRMT/functions.py:1216-1233  _solve_poisson_fft   (numpy.fft.fft2/ifft2       -> pocketfft)
//   pyRMT/functions.py:1205-1213  _tile_overlap
//
// DCT-I of a line x[0..M] equals the DFT of its even extension e of length 2M
// (that is also how pocketfft evaluates it).  Fast path (M a power of two): the
// real sequence e is packed into M complex points z[n] = e[2n] + i e[2n+1], one
// M-point complex FFT is run in shared memory, and the DCT coefficients are
// unpacked as  X_k = (Re Z_k + Re Z_{M-k} + c_k (Im Z_k + Im Z_{M-k}) - s_k (Re Z_k - Re Z_{M-k})) / 2
// with (c_k, s_k) = (cos, sin)(pi k / M)  -- half the flops and half the shared
// memory of transforming the extension directly.  The FFT is decimation in
// frequency, in place on split re/im planes (padded: every pass is bank-conflict
// free), and runs as fused radix-16 passes: a thread pulls 16 points into
// registers, does two radix-4 levels, and writes them back -- three passes and
// three barriers for M = 4096 instead of twelve radix-2 levels.  The output is
// left in digit-reversed order and the unpack step reads it through rev_pos(),
// so the permutation is never materialised.
//   rows:    one CTA per row:    load+pack -> FFT -> unpack -> store (x scale)
//   columns: one CTA per column: load+pack -> FFT -> unpack x 1/(4 Mx My eig) ->
//            pack -> FFT -> unpack -> store in place   (forward and inverse DCT-I
//            are the same transform; functions.py:1117 idctn = dctn / (2M))
// so the 2-D solve is three passes over HBM.  4096^2 is fp64-pipe and
// shared-memory bound at about the same level as HBM.
// Periodic solve (fast path, reduced grid lengths m a power of two): the symbol of
// functions.py:1177-1202 is even in each wavenumber separately, so fft2 -> /eig -> ifft2 of a
// REAL field equals  DHT2 -> /eig -> DHT2 / (mx my)  with the separable discrete Hartley
// transform  H_k = sum_n x_n (cos + sin)(2 pi k n / m) = Re X_k - Im X_k  (cas(k.) and
// cas(-k.) span the same eigenspace as exp(+-ik.)).  A Hartley line is real -> real, so the
// 2-D solve has the same row / transpose / column structure as the DCT one and never stores
// a complex plane: the m-point real line is packed into m/2 complex points, transformed with
// the same radix-16 kernel and unpacked to (H_k, H_{m-k}).
// Other sizes (N <= RMT_DENSE_MAX) use dense cosine/sine matrices and a small
// fp64 GEMM kernel -- O(N^3) but exact to rounding and only used on the small
// awkward grids of the reference's benchmarks (N = 128 -> 2*127 and 127).
#include "common.cuh"
#include "../../include/rmt_b200.h"

#include <cmath>
#include <vector>

using namespace rmt;

namespace {

constexpr int kMaxSmemL = 8192;   // longest complex FFT held in one CTA's shared memory (N <= 8193)

// Shared-memory index padding.  With 8-byte words a half-warp is conflict-free iff its
// 16 indices differ mod 16.  i + (i>>4) does that for every FFT pass (strides 1, 16, 256,
// ...); the extra (i>>8) term also spreads the digit-reversed addresses the unpack step
// reads (16 consecutive frequencies sit 256 and 1024 words apart) and is constant within a
// half-warp for all the passes.
__host__ __device__ __forceinline__ int padi(int i) { return i + (i >> 4) + (i >> 8); }

// Storage position of frequency k after the in-place DIF (radices 4,...,4[,2]): the base-4
// digits of k (least significant first) become the most significant digits of the
// position; a trailing radix-2 level contributes the lowest bit.  lg = log2(L).
__device__ __forceinline__ int rev_pos(int k, int lg)
{
    const int s2 = lg & ~1;                                   // bits covered by radix-4 digits
    const unsigned lowk = (unsigned)k & ((1u << s2) - 1u);
    unsigned r = s2 ? (__brev(lowk) >> (32 - s2)) : 0u;       // bit reversal ...
    r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);  // ... with the bit pairs kept in order
    return (lg & 1) ? (int)((r << 1) | ((unsigned)k >> s2)) : (int)r;
}

struct cplx {
    double x, y;
};
__device__ __forceinline__ cplx cmul(cplx a, double2 w)
{
    return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x};
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }

// w_16^m = exp(-2 pi i m / 16)
#define RMT_C8 0.92387953251128673848   /* cos(pi/8) */
#define RMT_S8 0.38268343236508978178   /* sin(pi/8) */
#define RMT_R2 0.70710678118654752440   /* sqrt(1/2)  */
__device__ constexpr double2 kW16[16] = {
    {1.0, 0.0},      {RMT_C8, -RMT_S8},  {RMT_R2, -RMT_R2},  {RMT_S8, -RMT_C8},
    {0.0, -1.0},     {-RMT_S8, -RMT_C8}, {-RMT_R2, -RMT_R2}, {-RMT_C8, -RMT_S8},
    {-1.0, 0.0},     {-RMT_C8, RMT_S8},  {-RMT_R2, RMT_R2},  {-RMT_S8, RMT_C8},
    {0.0, 1.0},      {RMT_S8, RMT_C8},   {RMT_R2, RMT_R2},   {RMT_C8, RMT_S8}};

// radix-4 forward DFT kernel:  y_q = sum_m a_m (-i)^(q m)
__device__ __forceinline__ void bfly4(cplx &a0, cplx &a1, cplx &a2, cplx &a3)
{
    cplx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
    cplx d = csub(a1, a3), t3{d.y, -d.x};                 // -i (a1 - a3)
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// A CTA hosts G independent thread groups (one line each); a group synchronises on its
// own named barrier so the groups drift apart and overlap each other's latencies.
struct Grp {
    int tid, nthr, bar;
    __device__ __forceinline__ void sync() const
    {
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
    }
};

// radix-2 level on adjacent pairs (the n = 2 level; no twiddles)
__device__ __forceinline__ void stage_radix2(double *re, double *im, int L, const Grp &g)
{
    for (int b = g.tid; b < (L >> 1); b += g.nthr) {
        int p0 = padi(2 * b), p1 = padi(2 * b + 1);
        double ar = re[p0], ai = im[p0], br = re[p1], bi = im[p1];
        re[p0] = ar + br; im[p0] = ai + bi;
        re[p1] = ar - br; im[p1] = ai - bi;
    }
    g.sync();
}

// one radix-4 DIF level of sub-length n
__device__ __forceinline__ void stage_radix4(double *re, double *im, int L, int n,
                                             const double2 *__restrict__ tw, const Grp &g)
{
    const int q = n >> 2, tstep = L / n;
    for (int b = g.tid; b < (L >> 2); b += g.nthr) {
        const int k = b & (q - 1), gb = (b - k) << 2;
        const int p0 = padi(gb + k), p1 = padi(gb + k + q), p2 = padi(gb + k + 2 * q), p3 = padi(gb + k + 3 * q);
        cplx a0{re[p0], im[p0]}, a1{re[p1], im[p1]}, a2{re[p2], im[p2]}, a3{re[p3], im[p3]};
        bfly4(a0, a1, a2, a3);
        if (k) {
            a1 = cmul(a1, __ldg(tw + k * tstep));
            a2 = cmul(a2, __ldg(tw + 2 * k * tstep));
            a3 = cmul(a3, __ldg(tw + 3 * k * tstep));
        }
        re[p0] = a0.x; im[p0] = a0.y;
        re[p1] = a1.x; im[p1] = a1.y;
        re[p2] = a2.x; im[p2] = a2.y;
        re[p3] = a3.x; im[p3] = a3.y;
    }
    g.sync();
}

// two fused radix-4 DIF levels (sub-lengths n and n/4) on 16 register-resident points.
// xline != nullptr (first pass of a DCT line, n == L): the inputs are taken straight from the real line
// x[0..L] in shared memory -- point p of the packed even extension is (e[2p], e[2p+1]), e[m] = x[m] for
// m <= L and x[2L - m] beyond -- so no separate packing pass (and its barrier) is needed.
__device__ __forceinline__ void stage_radix16(double *re, double *im, int L, int n,
                                              const double2 *__restrict__ tw, const Grp &g,
                                              const double *xline = nullptr)
{
    const int st = n >> 4;                       // spacing of the 16 points
    const int ts1 = L / n;                       // twiddle stride: w_n^k = tw[k * ts1]
    for (int b = g.tid; b < (L >> 4); b += g.nthr) {
        const int k0 = b & (st - 1), base = ((b - k0) << 4) + k0;
        cplx a[16];
        if (xline) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int i0 = 2 * (base + m * st), i1 = i0 + 1;
                if (i1 > L) { i0 = 2 * L - i0; i1 = 2 * L - i1; }
                a[m] = {xline[i0], xline[i1]};
            }
        } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int p = padi(base + m * st);
            a[m] = {re[p], im[p]};
        }
        }
        if (st == 1) {
            // the n = 16 pass (the last one of a 4096-point transform): k0 = 0, so every w_n^k0 is exactly 1 --
            // only the constant w_16^(q r) remain in the first level and the second level has no twiddles
            // (multiplying by (1, 0) returns the operand bit for bit, so this is the same arithmetic with the
            // no-ops left out; n is a compile-time constant on the specialised paths and the branch folds)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                bfly4(a[r], a[r + 4], a[r + 8], a[r + 12]);
                if (r) {
#pragma unroll
                    for (int q = 1; q < 4; ++q) a[r + 4 * q] = cmul(a[r + 4 * q], kW16[(q * r) & 15]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) bfly4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int p = padi(base + m * st);
                re[p] = a[m].x;
                im[p] = a[m].y;
            }
            continue;
        }
        // ONE table lookup per pass: b1 = w_n^k0.  Everything else follows by powers:
        // w_n^(q (k0 + r n/16)) = b1^q * w_16^(q r)  and  w_(n/4)^(q' k0) = (b1^4)^q'.
        const double2 b1 = __ldg(tw + k0 * ts1);
        const cplx B1{b1.x, b1.y};
        const cplx B2 = cmul(B1, b1), B3 = cmul(B2, b1);
        const cplx C1 = cmul(B2, double2{B2.x, B2.y});
        const cplx C2 = cmul(C1, double2{C1.x, C1.y}), C3 = cmul(C2, double2{C1.x, C1.y});
        const cplx Bq[4] = {{1.0, 0.0}, B1, B2, B3};
        // level n: butterflies over m = r, r+4, r+8, r+12; output q lands at slot 4q + r
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            bfly4(a[r], a[r + 4], a[r + 8], a[r + 12]);
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                cplx t = Bq[q];
                if (r) t = cmul(t, kW16[(q * r) & 15]);
                a[r + 4 * q] = cmul(a[r + 4 * q], double2{t.x, t.y});
            }
        }
        // level n/4: inside quarter q, butterflies over slots 4q .. 4q+3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bfly4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
            a[4 * q + 1] = cmul(a[4 * q + 1], double2{C1.x, C1.y});
            a[4 * q + 2] = cmul(a[4 * q + 2], double2{C2.x, C2.y});
            a[4 * q + 3] = cmul(a[4 * q + 3], double2{C3.x, C3.y});
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int p = padi(base + m * st);
            re[p] = a[m].x;
            im[p] = a[m].y;
        }
    }
    g.sync();
}

// Forward DFT of length L (power of two >= 2), natural order in -> digit-reversed
// order out (see rev_pos).  tw[m] = exp(-2 pi i m / L).  Caller syncs before.
__device__ void fft_dif(double *re, double *im, int L, const double2 *__restrict__ tw, const Grp &g)
{
    int n = L;
    for (; n >= 16; n >>= 4) stage_radix16(re, im, L, n, tw, g);
    for (; n >= 4; n >>= 2) stage_radix4(re, im, L, n, tw, g);
    if (n == 2) stage_radix2(re, im, L, g);
}

// DCT lines of compile-time length: first pass straight from the real line (see stage_radix16), ...
template <int L>
__device__ __forceinline__ void fft_first_from_line(const double *xline, double *re, double *im,
                                                    const double2 *__restrict__ tw, const Grp &g)
{
    static_assert(L >= 16, "radix-16 first pass");
    stage_radix16(re, im, L, L, tw, g, xline);
}
// ... then the remaining passes
template <int L>
__device__ __forceinline__ void fft_rest_ct(double *re, double *im, const double2 *__restrict__ tw, const Grp &g)
{
    int n = L >> 4;
#pragma unroll
    for (; n >= 16; n >>= 4) stage_radix16(re, im, L, n, tw, g);
#pragma unroll
    for (; n >= 4; n >>= 2) stage_radix4(re, im, L, n, tw, g);
    if (n == 2) stage_radix2(re, im, L, g);
}

// the same with the length known at compile time: inlined, passes unrolled, strides folded
template <int L>
__device__ __forceinline__ void fft_dif_ct(double *re, double *im, const double2 *__restrict__ tw, const Grp &g)
{
    int n = L;
#pragma unroll
    for (; n >= 16; n >>= 4) stage_radix16(re, im, L, n, tw, g);
#pragma unroll
    for (; n >= 4; n >>= 2) stage_radix4(re, im, L, n, tw, g);
    if (n == 2) stage_radix2(re, im, L, g);
}

// pack the even extension of the real line x[0..M] (shared memory) into M complex points
__device__ __forceinline__ void pack_even(const double *x, double *re, double *im, int M, const Grp &g)
{
    for (int n = g.tid; n < M; n += g.nthr) {
        int i0 = 2 * n, i1 = 2 * n + 1;                    // e[2n], e[2n+1]
        if (i1 > M) { i0 = 2 * M - i0; i1 = 2 * M - i1; }  // mirrored half: e[m] = x[2M - m]
        re[padi(n)] = x[i0];
        im[padi(n)] = x[i1];
    }
}

// DCT-I coefficient k from the digit-reversed packed spectrum; tw2[k] = (cos, sin)(pi k / M)
__device__ __forceinline__ double unpack_dct(const double *re, const double *im, int k, int M, int lg,
                                             const double2 *__restrict__ tw2)
{
    if (k == 0) return re[0] + im[0];
    if (k == M) return re[0] - im[0];
    const int pa = padi(rev_pos(k, lg)), pb = padi(rev_pos(M - k, lg));
    const double ar = re[pa], ai = im[pa], br = re[pb], bi = im[pb];
    const double2 w = __ldg(tw2 + k);
    return 0.5 * ((ar + br) + w.x * (ai + bi) - w.y * (ar - br));
}

// The pair (X_k, X_{M-k}), 0 < k <= M/2, from ONE set of shared-memory reads: both coefficients are built from
// Z_k and Z_{M-k}, and (cos, sin)(pi (M-k) / M) = (-cos, sin)(pi k / M).
__device__ __forceinline__ void unpack_dct_pair(const double *re, const double *im, int k, int M, int lg,
                                                const double2 *__restrict__ tw2, double &xk, double &xmk)
{
    const int pa = padi(rev_pos(k, lg)), pb = padi(rev_pos(M - k, lg));
    const double ar = re[pa], ai = im[pa], br = re[pb], bi = im[pb];
    const double2 w = __ldg(tw2 + k);
    const double sr = ar + br, si = ai + bi, dr = ar - br;
    const double t = w.x * si - w.y * dr;
    xk = 0.5 * (sr + t);
    xmk = 0.5 * (sr - t);
}

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------- DCT-I along rows, persistent
// Lines are the rows of a (nrows, N) array, N = M + 1.  A CTA hosts `G` groups of
// `tpg` threads; each group owns one line at a time (planes re/im + a staging copy of
// the NEXT line, fetched with cp.async while the current one is transformed) and walks
// lines grp, grp + ngroups, ...  MODE 0: out = scale * DCT-I(in).  MODE 1 (Poisson
// solve along this axis): out = DCT-I( DCT-I(in) * scale / eig ), eig laid out like in.
// in == out is allowed (a group holds its whole line on chip before it writes).
// LGT > 0: the line length is 2^LGT + 1 at compile time -- pass strides, twiddle strides, the padded
// index maps and the digit reversal then fold to constants and the passes unroll (the index arithmetic
// is a third of this kernel's instructions otherwise); LGT = 0: any power of two at run time.
template <int MODE, int LGT = 0>
__global__ void __launch_bounds__(512, 1)
k_dct_lines(const double *in, double *out, const double *__restrict__ eig, int nrows, int N_rt,
            const double2 *__restrict__ tw, const double2 *__restrict__ tw2, double scale, int tpg_rt,
            double *__restrict__ partial)
{
    extern __shared__ double sm[];
    const int N = LGT ? (1 << LGT) + 1 : N_rt;
    const int M = N - 1, lg = LGT ? LGT : 31 - __clz(M);
    const int tpg = LGT ? ((1 << LGT) / 16 > 512 ? 512 : ((1 << LGT) / 16 < 32 ? 32 : (1 << LGT) / 16)) : tpg_rt;
    const int G = blockDim.x / tpg, gi = threadIdx.x / tpg;
    const Grp g{(int)threadIdx.x % tpg, tpg, 1 + gi};
    const int plane = padi(M) + 1, per_group = 2 * plane + (N + 1);
    double *re = sm + (size_t)gi * per_group, *im = re + plane, *stage = im + plane;
    const int ngroups = gridDim.x * G;
    int r = blockIdx.x * G + gi;
    double s = 0.0;

    if (r < nrows)
        for (int k = g.tid; k < N; k += g.nthr) cp_async8(stage + k, in + (size_t)r * N + k);
    for (; r < nrows; r += ngroups) {
        cp_async_wait_all();
        g.sync();                                   // staged line complete; previous unpack done
        const int rn = r + ngroups;                 // prefetch the next line behind the (last) FFT
        if (LGT) {                                  // first pass reads the line itself: no packing pass
            fft_first_from_line<(LGT ? (1 << LGT) : 16)>(stage, re, im, tw, g);
            if (MODE == 0 && rn < nrows)            // (the staging buffer is free once that pass is through)
                for (int k = g.tid; k < N; k += g.nthr) cp_async8(stage + k, in + (size_t)rn * N + k);
            fft_rest_ct<(LGT ? (1 << LGT) : 16)>(re, im, tw, g);
        } else {
            pack_even(stage, re, im, M, g);
            g.sync();
            if (MODE == 0 && rn < nrows)
                for (int k = g.tid; k < N; k += g.nthr) cp_async8(stage + k, in + (size_t)rn * N + k);
            fft_dif(re, im, M, tw, g);
        }
        if (MODE == 1) {
            // spectrum x scale/eig -> staging buffer (natural order) -> packed again -> second FFT
            const double *er = eig + (size_t)r * N;
            for (int k = g.tid; k <= (M >> 1); k += g.nthr) {       // coefficients k and M - k together
                double xk, xmk;
                if (k == 0) { xk = re[0] + im[0]; xmk = re[0] - im[0]; }
                else unpack_dct_pair(re, im, k, M, lg, tw2, xk, xmk);
                stage[k] = xk * (scale / __ldg(er + k));
                if (k != M - k) stage[M - k] = xmk * (scale / __ldg(er + M - k));
            }
            g.sync();
            if (LGT) {
                fft_first_from_line<(LGT ? (1 << LGT) : 16)>(stage, re, im, tw, g);
                if (rn < nrows)
                    for (int k = g.tid; k < N; k += g.nthr) cp_async8(stage + k, in + (size_t)rn * N + k);
                fft_rest_ct<(LGT ? (1 << LGT) : 16)>(re, im, tw, g);
            } else {
                pack_even(stage, re, im, M, g);
                g.sync();
                if (rn < nrows)
                    for (int k = g.tid; k < N; k += g.nthr) cp_async8(stage + k, in + (size_t)rn * N + k);
                fft_dif(re, im, M, tw, g);
            }
        }
        double *o = out + (size_t)r * N;
        const double sc = (MODE == 1) ? 1.0 : scale;
        for (int k = g.tid; k <= (M >> 1); k += g.nthr) {           // coefficients k and M - k together
            double xk, xmk;
            if (k == 0) { xk = re[0] + im[0]; xmk = re[0] - im[0]; }
            else unpack_dct_pair(re, im, k, M, lg, tw2, xk, xmk);
            xk *= sc;
            o[k] = xk;
            s += xk;
            if (k != M - k) {
                xmk *= sc;
                o[M - k] = xmk;
                s += xmk;
            }
        }
    }
    if (partial) {
        __shared__ double red[32];
        __syncthreads();
        s = block_sum(s, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}


// ----------------------------------------------------- DHT along rows, persistent
// Lines are the rows (length m = 2M, stride ldi) of a real array; out row stride ldo.
//   out[k] = DHT(in)[k] * (mul ? mul[r * m + k] : scale)
// The packing z[n] = x[2n] + i x[2n+1] is the load itself (straight into the planes), so a
// group needs no staging copy: more groups per CTA hide each other's load latency instead.
// in == out is allowed (a group holds its whole line on chip before it writes).
// LGT > 0: m = 2^(LGT+1) at compile time (see k_dct_lines).
template <int LGT = 0>
__global__ void __launch_bounds__(512, 1)
k_dht_lines(const double *in, double *out, const double *__restrict__ mul, int nrows, int m_rt, long ldi,
            long ldo, const double2 *__restrict__ tw, const double2 *__restrict__ tw2, double scale, int tpg_rt)
{
    extern __shared__ double sm[];
    const int m = LGT ? (2 << LGT) : m_rt;
    const int M = m >> 1, lg = LGT ? LGT : 31 - __clz(M);
    const int tpg = LGT ? ((1 << LGT) / 16 > 512 ? 512 : ((1 << LGT) / 16 < 32 ? 32 : (1 << LGT) / 16)) : tpg_rt;
    const int G = blockDim.x / tpg, gi = threadIdx.x / tpg;
    const Grp g{(int)threadIdx.x % tpg, tpg, 1 + gi};
    const int plane = padi(M) + 1;
    double *re = sm + (size_t)gi * 2 * plane, *im = re + plane;
    const int ngroups = gridDim.x * G;
    for (int r = blockIdx.x * G + gi; r < nrows; r += ngroups) {
        const double *x = in + (size_t)r * ldi;
        g.sync();                                   // previous line's unpack has finished with the planes
        for (int k = g.tid; k < m; k += g.nthr) {
            const double v = __ldg(x + k);
            ((k & 1) ? im : re)[padi(k >> 1)] = v;
        }
        g.sync();
        if (LGT) fft_dif_ct<(LGT ? (1 << LGT) : 2)>(re, im, tw, g); else fft_dif(re, im, M, tw, g);
        double *o = out + (size_t)r * ldo;
        const double *f = mul ? mul + (size_t)r * m : nullptr;
        for (int k = g.tid; k <= M; k += g.nthr) {
            if (k == 0 || k == M) {
                const double v = (k == 0) ? re[0] + im[0] : re[0] - im[0];     // X_0, X_M are real
                o[k] = v * (f ? __ldg(f + k) : scale);
                continue;
            }
            const int pa = padi(rev_pos(k, lg)), pb = padi(rev_pos(M - k, lg));
            const double ar = re[pa], ai = im[pa], br = re[pb], bi = im[pb];
            const double2 w = __ldg(tw2 + k);                                   // (cos, sin)(2 pi k / m)
            const double er = 0.5 * (ar + br), ei = 0.5 * (ai - bi);            // even-sample spectrum
            const double pr = 0.5 * (ai + bi), pi_ = -0.5 * (ar - br);          // odd-sample spectrum
            const double xr = er + w.x * pr + w.y * pi_, xi = ei + w.x * pi_ - w.y * pr;
            o[k] = (xr - xi) * (f ? __ldg(f + k) : scale);
            o[m - k] = (xr + xi) * (f ? __ldg(f + m - k) : scale);
        }
    }
}

// out (C, R; row stride ldo) = in (R, C; row stride ldi)^T
__global__ void __launch_bounds__(256)
k_transpose_ld(const double *__restrict__ in, double *__restrict__ out, int R, int C, long ldi, long ldo)
{
    __shared__ double t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < R && c < C) t[ty + 8 * k][tx] = __ldg(in + (size_t)r * ldi + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < R && c < C) out[(size_t)c * ldo + r] = t[tx][ty + 8 * k];
    }
}

// transposed multiplier table of the periodic solve: fT[i][j] = null[j][i] ? 0 : scale / eig[j][i]
__global__ void __launch_bounds__(256)
k_inv_symbol_T(const double *__restrict__ eig, const unsigned char *__restrict__ null_mask,
               double *__restrict__ fT, int R, int C, double scale)
{
    __shared__ double t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < R && c < C) {
            const size_t q = (size_t)r * C + c;
            t[ty + 8 * k][tx] = null_mask[q] ? 0.0 : scale / eig[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < R && c < C) fT[(size_t)c * R + r] = t[tx][ty + 8 * k];
    }
}

// out (C, R) = in (R, C)^T, 32x32 tiles through padded shared memory (both sides coalesced)
__global__ void __launch_bounds__(256)
k_transpose(const double *__restrict__ in, double *__restrict__ out, int R, int C)
{
    __shared__ double t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < R && c < C) t[ty + 8 * k][tx] = __ldg(in + (size_t)r * C + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < R && c < C) out[(size_t)c * R + r] = t[tx][ty + 8 * k];
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

__global__ void k_copy2d(const double *__restrict__ src, double *__restrict__ dst, int rows, int cols,
                         long src_ld, long dst_ld)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) dst[(size_t)r * dst_ld + c] = src[(size_t)r * src_ld + c];
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
    double2 *tw_x, *tw_y;      // exp(-2 pi i m / M), M = N-1, per direction
    double2 *tw2_x, *tw2_y;    // (cos, sin)(pi k / M), k = 0..M: the real-FFT unpack twiddles
    int Lx, Ly;                // complex FFT lengths M
    // dense path
    double *Cx, *Cy, *Sx, *Sy; // cosine (and sine, periodic) matrices
    int nx, ny;                // transform sizes (DCT: N ; periodic: N-1)
    double *w[4];              // work arrays (fast DCT: w[0] = transposed field, w[1] = transposed eig)
    const double *eig_src;     // eigenvalue table w[1] was transposed from
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

int make_half_twiddles(double2 **dst, int M)
{
    std::vector<double2> h((size_t)M + 1);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k <= M; ++k) {
        h[k].x = std::cos(pi * (double)k / (double)M);
        h[k].y = std::sin(pi * (double)k / (double)M);
    }
    h[0] = {1.0, 0.0};
    h[M] = {-1.0, 0.0};
    if (M % 2 == 0) h[M / 2] = {0.0, 1.0};
    return upload(dst, h);
}

// one thread per 16 points (a radix-16 pass is one butterfly per thread)
int fft_threads(int L) { int t = L / 16; return t < 32 ? 32 : (t > 512 ? 512 : t); }
int group_doubles(int M) { return 2 * (padi(M) + 1) + (M + 2); }
constexpr int kMaxDyn = 227 * 1024 - 1024;

// launch geometry of k_dct_lines for line length M+1: groups per CTA and CTAs
struct LineLaunch { int tpg, G, threads, smem, ctas; };
LineLaunch line_launch(int M, int nrows)
{
    LineLaunch L;
    L.tpg = fft_threads(M);
    int by_smem = kMaxDyn / (group_doubles(M) * (int)sizeof(double));
    int by_thr = 512 / L.tpg;
    L.G = by_smem < by_thr ? by_smem : by_thr;
    if (L.G < 1) L.G = 1;
    if (L.G > 8) L.G = 8;                                  // named barriers 1..8
    L.threads = L.G * L.tpg;
    L.smem = L.G * group_doubles(M) * (int)sizeof(double);
    int want = (nrows + L.G - 1) / L.G;
    L.ctas = want < 148 ? want : 148;                     // one persistent CTA per SM
    return L;
}

// launch k_dct_lines<MODE>, with the compile-time specialisation for the common line lengths
template <int MODE>
void launch_dct_lines(const LineLaunch &L, cudaStream_t s, const double *in, double *out, const double *eig,
                      int nrows, int N, const double2 *tw, const double2 *tw2, double scale, double *partial)
{
    static bool attr_done = false;
    if (!attr_done) {       // every instantiation that can be launched needs the shared-memory opt-in
        cudaFuncSetAttribute(k_dct_lines<MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        cudaFuncSetAttribute(k_dct_lines<MODE, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        cudaFuncSetAttribute(k_dct_lines<MODE, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        cudaFuncSetAttribute(k_dct_lines<MODE, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        cudaFuncSetAttribute(k_dct_lines<MODE, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        attr_done = true;
    }
    switch (N - 1) {
    case 1024: k_dct_lines<MODE, 10><<<L.ctas, L.threads, L.smem, s>>>(in, out, eig, nrows, N, tw, tw2, scale, L.tpg, partial); break;
    case 2048: k_dct_lines<MODE, 11><<<L.ctas, L.threads, L.smem, s>>>(in, out, eig, nrows, N, tw, tw2, scale, L.tpg, partial); break;
    case 4096: k_dct_lines<MODE, 12><<<L.ctas, L.threads, L.smem, s>>>(in, out, eig, nrows, N, tw, tw2, scale, L.tpg, partial); break;
    case 8192: k_dct_lines<MODE, 13><<<L.ctas, L.threads, L.smem, s>>>(in, out, eig, nrows, N, tw, tw2, scale, L.tpg, partial); break;
    default:   k_dct_lines<MODE, 0><<<L.ctas, L.threads, L.smem, s>>>(in, out, eig, nrows, N, tw, tw2, scale, L.tpg, partial);
    }
}

// k_dht_lines for line length m = 2M: no staging copy, so more groups fit
LineLaunch dht_launch(int M, int nrows)
{
    LineLaunch L;
    L.tpg = fft_threads(M);
    const int per_group = 2 * (padi(M) + 1) * (int)sizeof(double);
    int by_smem = kMaxDyn / per_group, by_thr = 512 / L.tpg;
    L.G = by_smem < by_thr ? by_smem : by_thr;
    if (L.G < 1) L.G = 1;
    if (L.G > 8) L.G = 8;
    L.threads = L.G * L.tpg;
    L.smem = L.G * per_group;
    int want = (nrows + L.G - 1) / L.G;
    L.ctas = want < 148 ? want : 148;
    return L;
}

// twiddle tables per complex FFT length (per device), shared by the line entry points
struct LineTab { int M, dev; double2 *tw, *tw2; };
int line_tables(int M, const LineTab **out)
{
    static std::vector<LineTab> cache;
    int dev = 0;
    RMT_CUDA(cudaGetDevice(&dev));
    for (const LineTab &t : cache)
        if (t.M == M && t.dev == dev) { *out = &t; return RMT_OK; }
    LineTab t{M, dev, nullptr, nullptr};
    int e = make_twiddles(&t.tw, M);
    if (!e) e = make_half_twiddles(&t.tw2, M);
    if (!e) e = (int)cudaFuncSetAttribute(k_dct_lines<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (!e) e = (int)cudaFuncSetAttribute(k_dct_lines<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (!e) e = (int)cudaFuncSetAttribute(k_dht_lines<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (!e) e = (int)cudaFuncSetAttribute(k_dht_lines<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (!e) e = (int)cudaFuncSetAttribute(k_dht_lines<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (!e) e = (int)cudaFuncSetAttribute(k_dht_lines<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (e) return e;
    cache.reserve(64);
    cache.push_back(t);
    *out = &cache.back();
    return RMT_OK;
}

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
        P->Lx = Nx - 1; P->Ly = Ny - 1;
        P->fast = is_pow2(Nx - 1) && is_pow2(Ny - 1) && P->Lx <= kMaxSmemL && P->Ly <= kMaxSmemL &&
                  P->Lx >= 8 && P->Ly >= 8;
        P->nx = Nx; P->ny = Ny;
    } else {
        P->nx = Nx - 1; P->ny = Ny - 1;
        P->Lx = P->nx / 2; P->Ly = P->ny / 2;         // a real line of m points is m/2 packed complex points
        P->fast = is_pow2(P->nx) && is_pow2(P->ny) && P->Lx <= kMaxSmemL && P->Ly <= kMaxSmemL &&
                  P->Lx >= 8 && P->Ly >= 8;
    }
    if (!P->fast && (P->nx > RMT_DENSE_MAX || P->ny > RMT_DENSE_MAX)) {
        delete P;
        return -2;
    }
    if (P->fast && kind == 1) {
        // Hartley path: transposed work array + transposed multiplier table
        for (int k = 0; k < 2 && !e; ++k)
            e = (int)cudaMalloc((void **)&P->w[k], (size_t)P->nx * P->ny * sizeof(double));
    } else if (P->fast) {
        e = make_twiddles(&P->tw_x, P->Lx);
        if (!e) e = make_twiddles(&P->tw_y, P->Ly);
        if (!e) e = make_half_twiddles(&P->tw2_x, P->Lx);
        if (!e) e = make_half_twiddles(&P->tw2_y, P->Ly);
        // opt in to the full 227 KB once (a later, smaller plan must not lower it)
        if (!e) e = (int)cudaFuncSetAttribute(k_dct_lines<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        if (!e) e = (int)cudaFuncSetAttribute(k_dct_lines<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        const int mxy = P->Lx > P->Ly ? P->Lx : P->Ly;
        if (!e && group_doubles(mxy) * (int)sizeof(double) > kMaxDyn) e = -2;
        // transposed work arrays: T2 (Nx, Ny) and the transposed eigenvalue table
        for (int k = 0; k < 2 && !e; ++k) e = (int)cudaMalloc((void **)&P->w[k], ncell * sizeof(double));
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
    cudaFree(P->tw_x); cudaFree(P->tw_y); cudaFree(P->tw2_x); cudaFree(P->tw2_y);
    cudaFree(P->Cx); cudaFree(P->Cy); cudaFree(P->Sx); cudaFree(P->Sy);
    for (int k = 0; k < 4; ++k) cudaFree(P->w[k]);
    cudaFree(P->red);
    delete P;
}

int rmt_poisson_plan_is_fast(const rmt_poisson_plan *P) { return (P && P->fast) ? 1 : 0; }

// The plan keeps tables DERIVED from the caller's eigenvalue table (its transpose / inverse symbol).  A raw
// device pointer is not an identity (allocators recycle addresses), so the owner of the table says when it
// changed: the derived tables are rebuilt on the next solve.
int rmt_poisson_plan_invalidate(rmt_poisson_plan *P)
{
    if (!P) return RMT_EINVAL;
    P->eig_src = nullptr;
    return RMT_OK;
}

// ---- building blocks of the slab-decomposed (multi-GPU) DCT solve ------------------------
// DCT-I along the rows of a (nrows, N) array, N - 1 a power of two in [8, 8192].
//   eig == NULL: out = scale * DCT-I(in);   eig != NULL: out = DCT-I(DCT-I(in) * scale / eig).
// Twiddle tables are cached per line length (per device); in == out is allowed.
int rmt_dct_lines(const double *in, double *out, const double *eig, int nrows, int N, double scale,
                  void *stream)
{
    if (!in || !out || nrows < 1 || N < 9) return RMT_EINVAL;
    const int M = N - 1;
    if (!is_pow2(M) || M > kMaxSmemL) return -2;
    const LineTab *T = nullptr;
    if (int e = line_tables(M, &T)) return e;
    const LineLaunch L = line_launch(M, nrows);
    cudaStream_t s = (cudaStream_t)stream;
    if (eig)
        launch_dct_lines<1>(L, s, in, out, eig, nrows, N, T->tw, T->tw2, scale, nullptr);
    else
        launch_dct_lines<0>(L, s, in, out, nullptr, nrows, N, T->tw, T->tw2, scale, nullptr);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

// Discrete Hartley transform along the rows (length m, a power of two in [16, 16384]) of a real array:
//   out[r][k] = DHT(in[r])[k] * (mul ? mul[r*m + k] : scale);  row strides ldi / ldo doubles.
int rmt_dht_lines(const double *in, double *out, const double *mul, int nrows, int m, long ldi, long ldo,
                  double scale, void *stream)
{
    if (!in || !out || nrows < 1 || m < 16 || ldi < m || ldo < m) return RMT_EINVAL;
    const int M = m / 2;
    if (!is_pow2(m) || M > kMaxSmemL) return -2;
    const LineTab *T = nullptr;
    if (int e = line_tables(M, &T)) return e;
    const LineLaunch L = dht_launch(M, nrows);
    cudaStream_t s = (cudaStream_t)stream;
    switch (M) {
    case 2048: k_dht_lines<11><<<L.ctas, L.threads, L.smem, s>>>(in, out, mul, nrows, m, ldi, ldo, T->tw, T->tw2, scale, L.tpg); break;
    case 4096: k_dht_lines<12><<<L.ctas, L.threads, L.smem, s>>>(in, out, mul, nrows, m, ldi, ldo, T->tw, T->tw2, scale, L.tpg); break;
    case 8192: k_dht_lines<13><<<L.ctas, L.threads, L.smem, s>>>(in, out, mul, nrows, m, ldi, ldo, T->tw, T->tw2, scale, L.tpg); break;
    default:   k_dht_lines<0><<<L.ctas, L.threads, L.smem, s>>>(in, out, mul, nrows, m, ldi, ldo, T->tw, T->tw2, scale, L.tpg);
    }
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

// out (C, R) = in (R, C)^T
int rmt_transpose(const double *in, double *out, int R, int C, void *stream)
{
    if (!in || !out || R < 1 || C < 1) return RMT_EINVAL;
    dim3 grd(rmt_cdiv(C, 32), rmt_cdiv(R, 32));
    k_transpose<<<grd, 256, 0, (cudaStream_t)stream>>>(in, out, R, C);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

// dst[r*dst_ld + c] = src[r*src_ld + c] for a rows x cols block (pack / unpack of the all-to-all)
int rmt_copy2d(const double *src, double *dst, int rows, int cols, long src_ld, long dst_ld, void *stream)
{
    if (!src || !dst || rows < 1 || cols < 1) return RMT_EINVAL;
    dim3 grd(rmt_cdiv(cols, 128), rows > 65535 ? 65535 : rows);
    k_copy2d<<<grd, 128, 0, (cudaStream_t)stream>>>(src, dst, rows, cols, src_ld, dst_ld);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

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
        // rows: rhs -> sol | transpose -> T2 | rows of T2 = columns: fwd, x scale/eig, inverse |
        // transpose back -> sol | rows in place on sol (+ partial sums).  The only FFT kernel is the
        // row kernel; the two transposes are plain tiled copies.
        double *T2 = P->w[0], *eigT = P->w[1];
        dim3 tg_f(rmt_cdiv(Nx, 32), rmt_cdiv(Ny, 32)), tg_b(rmt_cdiv(Ny, 32), rmt_cdiv(Nx, 32));
        if (P->eig_src != eig) {                         // transposed eigenvalues, once per table
            k_transpose<<<tg_f, 256, 0, s>>>(eig, eigT, Ny, Nx);
            RMT_LAUNCH_CHECK();
            P->eig_src = eig;
        }
        const LineLaunch lx = line_launch(P->Lx, Ny), ly = line_launch(P->Ly, Nx);
        double *partial = P->red;                        // one partial sum per CTA
        const int nrow_cta = lx.ctas;
        launch_dct_lines<0>(lx, s, rhs, sol, nullptr, Ny, Nx, P->tw_x, P->tw2_x, 1.0, nullptr);
        RMT_LAUNCH_CHECK();
        k_transpose<<<tg_f, 256, 0, s>>>(sol, T2, Ny, Nx);
        RMT_LAUNCH_CHECK();
        launch_dct_lines<1>(ly, s, T2, T2, eigT, Nx, Ny, P->tw_y, P->tw2_y, scale, nullptr);
        RMT_LAUNCH_CHECK();
        k_transpose<<<tg_b, 256, 0, s>>>(T2, sol, Nx, Ny);
        RMT_LAUNCH_CHECK();
        launch_dct_lines<0>(lx, s, sol, sol, nullptr, Ny, Nx, P->tw_x, P->tw2_x, 1.0, partial);
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
    if (P->fast) {
        // rows DHT: rhs[:-1,:-1] -> sol[:-1,:-1] | transpose -> T2 (mx, my) | columns: DHT x fT, DHT |
        // transpose back | rows DHT in place | overlap row / column
        double *T2 = P->w[0], *fT = P->w[1];
        dim3 tg_f(rmt_cdiv(mx, 32), rmt_cdiv(my, 32)), tg_b(rmt_cdiv(my, 32), rmt_cdiv(mx, 32));
        if (P->eig_src != eig) {
            k_inv_symbol_T<<<tg_f, 256, 0, s>>>(eig, null_mask, fT, my, mx, scale);
            RMT_LAUNCH_CHECK();
            P->eig_src = eig;
        }
        int e;
        if ((e = rmt_dht_lines(rhs, sol, nullptr, my, mx, Nx, Nx, 1.0, stream))) return e;
        k_transpose_ld<<<tg_f, 256, 0, s>>>(sol, T2, my, mx, Nx, my);
        RMT_LAUNCH_CHECK();
        if ((e = rmt_dht_lines(T2, T2, fT, mx, my, my, my, 1.0, stream))) return e;
        if ((e = rmt_dht_lines(T2, T2, nullptr, mx, my, my, my, 1.0, stream))) return e;
        k_transpose_ld<<<tg_b, 256, 0, s>>>(T2, sol, mx, my, my, Nx);
        RMT_LAUNCH_CHECK();
        if ((e = rmt_dht_lines(sol, sol, nullptr, my, mx, Nx, Nx, 1.0, stream))) return e;
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

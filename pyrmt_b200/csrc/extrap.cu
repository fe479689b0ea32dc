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
#include <stdlib.h>

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

// block-scope (L1-coherent within the SM) relaxed loads for rows this CTA owns
__device__ __forceinline__ unsigned char ld_cta_u8(const unsigned char *p)
{
    unsigned v;
    asm volatile("ld.relaxed.cta.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return (unsigned char)v;
}
__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ double ld_cta_f64(const double *p)
{
    double v;
    asm volatile("ld.relaxed.cta.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// state byte per cell.  Odd = known: 1 = known before this layer's sweep, 3 = fitted by
// this layer's sweep (turned into 1 by the next layer's flag pass).  Even = unknown:
// 0 = not a target, 2 = target of the current layer (stays 2 if its fit is rejected).
constexpr unsigned char ST_UNKNOWN = 0, ST_KNOWN = 1, ST_TARGET = 2, ST_FRESH = 3;

// seed: copies of X1/X2, known = (phi < 0)          functions.py:69-74
__global__ void k_ext_seed(const double *X1, const double *X2, const double *__restrict__ phi, double *X1e,
                           double *X2e, unsigned char *__restrict__ st, long n)
{
    const bool copy = (X1e != X1) || (X2e != X2);      // in place (X1e == X1, X2e == X2): only the state bytes
    for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        if (copy) {
            X1e[k] = X1[k];
            X2e[k] = X2[k];
        }
        st[k] = (phi[k] < 0.0) ? ST_KNOWN : ST_UNKNOWN;
    }
}

// Sweep tiling (tunable at start-up through RMT_EXT_XT / RMT_EXT_MRB / RMT_EXT_SLEEP):
//   XT  columns per x-tile (multiple of 32)
//   MRB row blocks a CTA sweeps in sequence (macro-tile = MRB*RB rows x XT columns)
struct ExtTune { int XT, MRB, sleep_ns; };
static ExtTune ext_tune()
{
    static ExtTune t = {0, 0, 0};
    if (!t.XT) {
        const char *e;
        t.XT = 256; t.MRB = 32; t.sleep_ns = 20;
        if ((e = getenv("RMT_EXT_XT")) && atoi(e) >= 32) t.XT = (atoi(e) / 32) * 32;
        if ((e = getenv("RMT_EXT_MRB")) && atoi(e) >= 1) t.MRB = atoi(e);
        if ((e = getenv("RMT_EXT_SLEEP"))) t.sleep_ns = atoi(e);
    }
    return t;
}

// frontier of the current layer: unknown interior cells with a known 3x3
// neighbour (functions.py:79-90); one warp per row, counts them per x-tile.
__global__ void k_ext_flag(unsigned char *__restrict__ st, int *__restrict__ seg_cnt, int Ny, int Nx,
                           int nxt, int XT, const int *__restrict__ mode)
{
    if (mode && *mode != 0) return;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < Ny; j += nwarp) {
        const unsigned char *r1 = st + (size_t)j * Nx;
        const bool inner = (j >= 1 && j < Ny - 1);
        const unsigned char *r0 = inner ? r1 - Nx : r1, *r2 = inner ? r1 + Nx : r1;
        for (int xt = 0; xt < nxt; ++xt) {
            int cnt = 0;
            const int cend = min((xt + 1) * XT, Nx);
            for (int base = xt * XT; base < cend; base += 32) {
                int i = base + lane;
                bool tgt = false;
                unsigned char me = (i < Nx) ? r1[i] : ST_KNOWN;
                if (inner && i >= 1 && i < Nx - 1 && !(me & 1)) {
                    // neighbour rows are rewritten concurrently only between values of the
                    // same parity (0<->2, 3->1), so the "known" bit read here is stable.
                    tgt = ((r0[i - 1] | r0[i] | r0[i + 1] | r1[i - 1] | r1[i + 1] | r2[i - 1] | r2[i] |
                            r2[i + 1]) & 1) != 0;
                }
                if (i < Nx) {
                    unsigned char nv = (me & 1) ? ST_KNOWN : (tgt ? ST_TARGET : ST_UNKNOWN);
                    if (nv != me) st[(size_t)j * Nx + i] = nv;
                }
                cnt += __popc(__ballot_sync(0xffffffffu, tgt));
            }
            if (lane == 0) seg_cnt[j * nxt + xt] = cnt;
        }
    }
}

// exclusive scan of the per-(row, x-tile) counts -> seg_off[0..Ny*nxt]; resets the tile counter.
// One thread per row sums its nxt segments, the row totals are scanned in shared memory, each
// thread then writes its row's segment offsets.
__global__ void k_ext_scan(const int *__restrict__ seg_cnt, int *__restrict__ seg_off, int Ny, int nxt,
                           int *__restrict__ tile_counter, const int *__restrict__ mode)
{
    if (mode && *mode != 0) return;
    __shared__ int sh[1024];
    const int per = (Ny + blockDim.x - 1) / blockDim.x;       // rows per thread (contiguous)
    const int lo = min((int)threadIdx.x * per, Ny), hi = min(lo + per, Ny);
    int s = 0;
    for (int j = lo; j < hi; ++j)
        for (int x = 0; x < nxt; ++x) s += seg_cnt[j * nxt + x];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
        int v = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - s;
    for (int j = lo; j < hi; ++j)
        for (int x = 0; x < nxt; ++x) {
            seg_off[j * nxt + x] = run;
            run += seg_cnt[j * nxt + x];
        }
    if (threadIdx.x == blockDim.x - 1) seg_off[Ny * nxt] = sh[threadIdx.x];
    if (threadIdx.x == 0) *tile_counter = 0;
}

// column indices of each row's targets in increasing column order (so the list of
// segment (j, xt) is tcol[seg_off[j*nxt+xt] .. seg_off[j*nxt+xt+1])), and each
// segment's initial progress marker = column of its first target (INT_MAX if
// none): "segment (j, xt) has dealt with every column < prog[j*nxt+xt]".
__global__ void k_ext_fill(const unsigned char *__restrict__ st, const int *__restrict__ seg_off,
                           int *__restrict__ tcol, int *__restrict__ trow, int *__restrict__ prog, int Ny,
                           int Nx, int nxt, int XT, const int *__restrict__ mode)
{
    if (mode && *mode != 0) return;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < Ny; j += nwarp) {
        const unsigned char *r1 = st + (size_t)j * Nx;
        for (int xt = 0; xt < nxt; ++xt) {
            int off = seg_off[j * nxt + xt];
            if (seg_off[j * nxt + xt + 1] == off) {
                if (lane == 0) prog[j * nxt + xt] = INT_MAX;
                continue;
            }
            bool first = true;
            const int cend = min((xt + 1) * XT, Nx);
            for (int base = xt * XT; base < cend; base += 32) {
                int i = base + lane;
                bool tgt = (i < Nx) && (r1[i] == ST_TARGET);
                unsigned m = __ballot_sync(0xffffffffu, tgt);
                if (tgt) {
                    const int o = off + __popc(m & ((1u << lane) - 1u));
                    tcol[o] = i;
                    trow[o] = j;
                }
                if (first && m) {
                    if (lane == 0) prog[j * nxt + xt] = base + __ffs(m) - 1;
                    first = false;
                }
                off += __popc(m);
            }
        }
    }
}

#ifdef RMT_EXT_TIMING
__device__ unsigned long long g_ext_dbg[8];
#define EXT_T(k) do { if (lane == 0) { long long now__ = clock64(); atomicAdd(&g_ext_dbg[k], (unsigned long long)(now__ - tmark)); tmark = now__; } } while (0)
__device__ unsigned long long g_body_dbg[16 * 8];   // [layer][counter]
#define BODY_T(k) do { if (lane == 0) { long long now__ = clock64(); atomicAdd(&g_body_dbg[(k) + 16 * layer], (unsigned long long)(now__ - bmark)); bmark = now__; } } while (0)
#define BODY_T0() long long bmark = clock64(); long long dmark = bmark
#define DISC_T(k) do { if (lane == 0) { long long now__ = clock64(); atomicAdd(&g_body_dbg[(k) + 16 * (layer + 3)], (unsigned long long)(now__ - dmark)); dmark = now__; } } while (0)
#define BODY_CNT(k) do { if (lane == 0) atomicAdd(&g_body_dbg[(k) + 16 * layer], 1ull); } while (0)
#else
#define EXT_T(k) do { } while (0)
#define BODY_T(k) do { } while (0)
#define BODY_T0() do { } while (0)
#define DISC_T(k) do { } while (0)
#define BODY_CNT(k) do { } while (0)
#endif

constexpr int RB = 16;                 // rows per CTA block = warps per CTA in the sweep
constexpr int WIN = 81;                // 9x9 window
constexpr int NACC = 12;               // running sums: B1[3], B2[3], A00 A01 A02 A11 A12 A22
constexpr int RING = 16;               // direct-mapped cache of a row's freshly fitted cells
constexpr int MAXPEND = 40;            // window cells that precede the target in raster order

// Everything about a target that does not depend on this layer's earlier fits ("phase A").
// Same layout in shared memory (the warp's working buffer) and in global memory (the
// records k_ext_prepare writes for the whole layer up front).
struct alignas(16) ExtRec {
    double prod[WIN][NACC];            // per contributing cell (in gather order): the 12 products
    double pw[MAXPEND], px[MAXPEND], py[MAXPEND];   // weight / coordinates of the undecided cells
    unsigned char pn[MAXPEND], pslot[MAXPEND];      // window index and prod[] slot of the undecided cells
    int ptag[MAXPEND];                              // body variant: jj << 15 | ii of the undecided cells
};
static_assert(sizeof(ExtRec) % 16 == 0, "ExtRec must be copyable in 16-byte chunks");
constexpr int REC_PEND_OFF = WIN * NACC * 8;                   // byte offset of pw
constexpr int REC_PEND_BYTES = (int)sizeof(ExtRec) - REC_PEND_OFF;

struct SweepWarp {
    ExtRec rec;
    double sums[NACC];                 // the twelve running sums, handed from 12 lanes to all
    double ring_v[RING][2];            // freshly fitted (xi1, xi2) of this row, slot = column & (RING-1)
    int ring_tag[RING];                // column held by the slot (-1: none / being rewritten)
};
struct SweepSmem {
    SweepWarp w[RB];
    double prev_v[4][RING][2];         // rings of the last four rows of the previous row block
    int prev_tag[4][RING];
    int prev_row0;                     // grid row of prev_*[0] (-100: none)
    int prog[RB];                      // per-row progress marker of the tile being swept
    int tile;                          // tile index handed out by the global counter
};

__device__ __forceinline__ void st_relaxed_gpu(int *p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The cross-CTA progress counters are POLLED with relaxed loads (an acquire load per poll would invalidate
// L1 every time); the acquire that orders the dependent reads of the other CTA's results behind the poll
// that saw the counter is one fence after the spin loop (PTX memory model: relaxed load + fence.acq_rel
// synchronises with the writer's fence + relaxed store).
#define ACQUIRE_GPU() asm volatile("fence.acq_rel.gpu;" ::: "memory")
__device__ __forceinline__ int ld_relaxed_gpu(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#define SMEM_ORDER() asm volatile("" ::: "memory")   // compiler barrier; shared memory itself is in-order per warp
// 16-byte shared-memory accesses done as ONE instruction (header + ready flag travel together)
__device__ __forceinline__ void sts_v4(void *p, int4 v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_generic_v4(void *p, int4 v)      // generic address (distributed shared memory)
{
    asm volatile("st.volatile.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int4 lds_v4(const void *p)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// the 12 products of one contributing cell, in the order of the running sums:
//   B1[k] += (w*a_k)*v1, B2[k] += (w*a_k)*v2, A[r][c] += (w*a_r)*a_c   with a = (1, x, y)
// (w*1.0 == w exactly, so the reference's wg*a0 factors drop out bit for bit)
__device__ __forceinline__ void store_products(double *q, double w, double x, double y, double v1, double v2)
{
    const double wx = w * x, wy = w * y;
    q[0] = w * v1;  q[1] = wx * v1; q[2] = wy * v1;
    q[3] = w * v2;  q[4] = wx * v2; q[5] = wy * v2;
    q[6] = w;       q[7] = wx;      q[8] = wy;
    q[9] = wx * x;  q[10] = wx * y; q[11] = wy * y;
}

// Phase A of one target (warp-collective): classify the 81 window cells, weight them,
// turn the cells known before the layer into their 12 products (compacted in gather
// order) and list the undecided ones.  Returns nslots | npend << 8.
// FUSED = false: per-layer launches -- known == state 1, undecided == marked 2/3 and raster-earlier.
// FUSED = true : all layers in one launch -- state 2l+3 = fitted by layer l; "known" is any odd state
// below this layer's code `fresh`; every raster-earlier cell that is not known is a candidate (there
// are no target marks); all reads go to L2 (other SMs wrote the previous layers).
// MODE 0: per-layer launches (plain loads); 1: all-layers kernel across SMs (L2 loads); 2: all-layers
// kernel with a whole tile on ONE SM (block-scope coherent loads, served by L1); 3: as 2 for the helper CTA
// on the neighbouring SM (L2 loads: the tile's own SM wrote the data).
struct ExtNoWait { __device__ __forceinline__ bool operator()() const { return true; } };
// `before_store()` runs after the window has been loaded, classified and weighted (registers only) and
// before the first store into R: a caller whose record slot may still be in use waits there, so the
// expensive part of phase A overlaps the wait.  It returns false to abandon the record (result -1).
template <int MODE, class Hook = ExtNoWait>
__device__ __forceinline__ int ext_phase_a(ExtRec *R, const double *__restrict__ X1e,
                                           const double *__restrict__ X2e,
                                           const unsigned char *__restrict__ st, int j, int i, int Ny, int Nx,
                                           int joff, double dx, double dy, double r2, int lane, int fresh = ST_FRESH,
                                           Hook before_store = Hook())
{
    const unsigned lt = (1u << lane) - 1u;
    const double x0 = dx * i, y0 = dy * (j + joff);     // absolute coordinates of the GLOBAL grid (functions.py:105)
    constexpr bool FUSED = MODE != 0;
    int cls[3];                                 // 0 no contribution, 1 known, 2 undecided
    double cw[3], cx_[3], cy_[3], c1[3], c2[3];
    unsigned mk[3], mp[3];
    // all loads of the window first (three cells per lane), then the arithmetic
    unsigned sv[3];
    bool in[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int n = lane + 32 * s;
        const int jj = j + n / 9 - 4, ii = i + n % 9 - 4;
        in[s] = n < WIN && jj >= 0 && jj < Ny && ii >= 0 && ii < Nx;
        sv[s] = 0;
        c1[s] = c2[s] = 0.0;
        if (in[s]) {
            const size_t cc = (size_t)jj * Nx + ii;
            sv[s] = (MODE == 1 || MODE == 3) ? __ldcg(st + cc) : MODE == 2 ? ld_cta_u8(st + cc) : st[cc];   // known cells are final
            c1[s] = (MODE == 1 || MODE == 3) ? __ldcg(X1e + cc) : MODE == 2 ? ld_cta_f64(X1e + cc) : X1e[cc];   // only meaningful
            c2[s] = (MODE == 1 || MODE == 3) ? __ldcg(X2e + cc) : MODE == 2 ? ld_cta_f64(X2e + cc) : X2e[cc];   // if known
        }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int n = lane + 32 * s;
        cls[s] = 0;
        cw[s] = cx_[s] = cy_[s] = 0.0;
        if (in[s]) {
            const int dj = n / 9 - 4, di = n % 9 - 4;
            const int jj = j + dj, ii = i + di;
            const double xi = dx * ii, yi = dy * (jj + joff);
            const double ex = xi - x0, ey = yi - y0;
            const double dist_sq = ex * ex + ey * ey;
            const bool earlier = (dj < 0) || (dj == 0 && di < 0);
            const bool known = FUSED ? ((sv[s] & 1) && (int)sv[s] < fresh) : (sv[s] == ST_KNOWN);
            // undecided = a raster-earlier target of this layer.  MODE 0: marked 2 / 3 by the flag pass and the
            // sweep; MODE 1: no marks, every raster-earlier cell that is not known is a candidate; MODE 2: the
            // discovery warp marked this layer's targets with fresh - 1 (fresh once fitted).
            const bool cand = MODE == 1 ? (!known && earlier)
                            : MODE >= 2 ? (earlier && ((int)sv[s] == fresh || (int)sv[s] == fresh - 1))
                                        : (sv[s] != ST_UNKNOWN && sv[s] != ST_KNOWN && earlier);
            if (dist_sq <= r2 && (known || cand)) {
                cls[s] = known ? 1 : 2;
                cw[s] = exp_glibc(-dist_sq / r2);
                cx_[s] = xi; cy_[s] = yi;
            }
        }
        mk[s] = __ballot_sync(0xffffffffu, cls[s] == 1);
        mp[s] = __ballot_sync(0xffffffffu, cls[s] == 2);
    }
    if (!before_store()) return -1;
    int slot_base = 0, pend_base = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const unsigned mv = mk[s] | mp[s];
        const int slot = slot_base + __popc(mv & lt);
        if (cls[s] == 1) {
            store_products(R->prod[slot], cw[s], cx_[s], cy_[s], c1[s], c2[s]);
        } else if (cls[s] == 2) {
            const int r = pend_base + __popc(mp[s] & lt);
            R->pw[r] = cw[s]; R->px[r] = cx_[s]; R->py[r] = cy_[s];
            R->pn[r] = (unsigned char)(lane + 32 * s);
            R->pslot[r] = (unsigned char)slot;
            if (MODE >= 2) {
                const int n = lane + 32 * s;
                R->ptag[r] = ((j + n / 9 - 4) << 15) | (i + n % 9 - 4);
            }
        }
        slot_base += __popc(mv);
        pend_base += __popc(mp[s]);
    }
    // pad the slot list with zero products to a multiple of 8 (exact no-ops in the sums):
    // the accumulation loop then runs in fixed, fully unrolled chunks of 8
    const int npad = min((slot_base + 7) & ~7, WIN);
    for (int e = slot_base * NACC + lane; e < npad * NACC; e += 32) (&R->prod[0][0])[e] = 0.0;
    return slot_base | (pend_base << 8);
}

// Phase A for every target of the layer, fully parallel (one warp per target): nothing in
// it depends on this layer's fits, so it is taken off the serial chain of the sweep.
__global__ void __launch_bounds__(256)
k_ext_prepare(const double *__restrict__ X1e, const double *__restrict__ X2e,
              const unsigned char *__restrict__ st, const int *__restrict__ seg_off, int nseg,
              const int *__restrict__ tcol, const int *__restrict__ trow, ExtRec *__restrict__ recs,
              int *__restrict__ tinfo, int cap, int Ny, int Nx, int joff, double dx, double dy, double r2,
              const int *__restrict__ mode)
{
    if (mode && *mode != 0) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int ntot = min(seg_off[nseg], cap);
    for (int t = warp; t < ntot; t += nwarp) {
        const int info = ext_phase_a<0>(recs + t, X1e, X2e, st, trow[t], tcol[t], Ny, Nx, joff, dx, dy, r2, lane);
        if (lane == 0) tinfo[t] = info;
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// start copying target t's record into the warp's buffer (only the used part)
__device__ __forceinline__ void ext_fetch(ExtRec *dst, const ExtRec *__restrict__ src, int info, int lane)
{
    const int nslots = info & 255, npend = info >> 8;
    const char *s8 = (const char *)src;
    char *d8 = (char *)dst;
    const int npad = min((nslots + 7) & ~7, WIN);
    for (int c = lane; c < npad * (NACC * 8 / 16); c += 32) cp_async16(d8 + c * 16, s8 + c * 16);
    if (npend)
        for (int c = lane; c < REC_PEND_BYTES / 16; c += 32)
            cp_async16(d8 + REC_PEND_OFF + c * 16, s8 + REC_PEND_OFF + c * 16);
}

// The row-pipelined sweep (functions.py:95-161).
//
// The grid is cut into tiles of RB rows x XT columns, handed out to the
// (co-resident) CTAs in raster order by a global counter; one warp sweeps one row
// of the tile, left to right.  Before fitting (j, i) a warp waits until rows
// j-4..j-1 have dealt with every column <= i+4 and its own row with every column
// < i -- that is exactly the dependency set of the serial raster sweep, so every
// fit sees the inputs the reference's fit saw.  The serial dependency chain that
// runs along the flank of a body is hundreds of targets long and a lone warp
// retires only one instruction every few cycles, so the work that sits between
// "my predecessor published" and "I publish" is cut to the bone and kept inside
// one SM:
//   * rows of the same tile publish progress AND their freshly fitted values
//     through shared memory (a small direct-mapped ring per row, seqlock-style);
//     a row publishes through global memory only where another tile can depend
//     on it (last four rows, columns next to the tile's x-edges, end of its list),
//     and other tiles poll those markers with relaxed L2 loads;
//   * BEFORE the wait: cells known before the sweep (state 1, final) are gathered,
//     weighted and turned into their 12 products, compacted in gather order; the
//     cells that are this layer's earlier targets get a slot, their weight and a
//     place in a to-do list;
//   * AFTER the wait: one lane per to-do cell looks the value up (ring, else L2)
//     and fills its slot -- with zeros if that target was rejected: adding +-0.0
//     never changes a running sum that started at +0.0, so the sums still equal
//     the reference's sums over the known cells bit for bit;
//   * twelve lanes then run the twelve dependent-add chains over the compacted
//     slots (1 shared-memory load + 1 DADD per term), the sums are exchanged
//     through shared memory (no shuffles), lanes 0/1 solve for xi1/xi2.
// Different bodies fall into different tiles and are swept concurrently.
__global__ void __launch_bounds__(RB * 32, 1)
k_ext_sweep(double *__restrict__ X1e, double *__restrict__ X2e, unsigned char *__restrict__ st,
            const int *__restrict__ seg_off, const int *__restrict__ tcol, int *__restrict__ prog,
            int *__restrict__ tile_counter, const ExtRec *__restrict__ recs, const int *__restrict__ tinfo,
            int cap, int Ny, int Nx, int joff, int nxt, int XT, int MRB, int sleep_ns, double dx, double dy,
            double r2, const int *__restrict__ mode)
{
    if (mode && *mode != 0) return;
    extern __shared__ unsigned char s_raw[];
    SweepSmem &S = *reinterpret_cast<SweepSmem *>(s_raw);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SweepWarp &W = S.w[wib];
    volatile int *sp = S.prog;
    const int nrb = (Ny - 2 + RB - 1) / RB;            // row blocks over rows 1 .. Ny-2
    const int nmrb = (nrb + MRB - 1) / MRB;            // macro row blocks
    const int ntiles = nmrb * nxt;
    const int la = (lane < NACC) ? lane : 0;

    // A CTA takes a macro-tile (MRB row blocks x one x-tile) and sweeps its row blocks
    // top to bottom itself, so the chain that runs along a flank stays on this SM for
    // MRB*RB rows; macro-tiles are handed out in raster order (their dependencies point
    // to raster-earlier macro-tiles, which are resident or finished: no deadlock).
    for (;;) {
        if (threadIdx.x == 0) S.tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = S.tile;
        if (tile >= ntiles) break;
        const int mrb = tile / nxt, xt = tile - mrb * nxt;
        const int xc0 = xt * XT, xc1 = min(xc0 + XT, Nx); // its column range
        const int rb_end = min((mrb + 1) * MRB, nrb);
        if (threadIdx.x == 0) S.prev_row0 = -100;
      for (int rb = mrb * MRB; rb < rb_end; ++rb) {
        const int row0 = 1 + rb * RB;                  // first row of this row block
        const int j = row0 + wib;
        const bool live = j < Ny - 1;
        const int t0 = live ? seg_off[j * nxt + xt] : 0, t1 = live ? seg_off[j * nxt + xt + 1] : 0;
        if (__syncthreads_or(t1 > t0) == 0) continue;  // nothing to fit in this row block
        if (lane == 0) sp[wib] = (t1 > t0) ? tcol[t0] : INT_MAX;
        if (lane < RING) W.ring_tag[lane] = -1;
        __syncthreads();
        const int prow0 = S.prev_row0;                 // rows prow0..prow0+3 were swept by this CTA just before
        const bool prev_ok = (prow0 == row0 - 4);
        const bool last_rb = (rb == rb_end - 1);
        // does another tile possibly read this row's results?  (x-edges, or the macro-tile's last rows)
        const bool edgy = (t1 > t0) && (tcol[t0] < xc0 + 8 || tcol[t1 - 1] >= xc1 - 8 || (last_rb && wib >= RB - 4));

        // the first target's record is fetched while this warp waits for its dependencies;
        // every later one right after the previous accumulation has drained the buffer
        int info_next = (t0 < t1 && t0 < cap) ? tinfo[t0] : 0;
        if (t0 < t1 && t0 < cap) ext_fetch(&W.rec, recs + t0, info_next, lane);

        for (int t = t0; t < t1; ++t) {
            const int i = tcol[t];
            const double x0 = dx * i, y0 = dy * (j + joff);
#ifdef RMT_EXT_TIMING
            long long tmark = clock64();
#endif
            // ---- phase A: prepared record (or inline beyond the record capacity) -------
            int info;
            if (t < cap) {
                info = info_next;
                cp_async_wait_all();
            } else {
                info = ext_phase_a<0>(&W.rec, X1e, X2e, st, j, i, Ny, Nx, joff, dx, dy, r2, lane);
            }
            info_next = (t + 1 < t1 && t + 1 < cap) ? tinfo[t + 1] : 0;
            const int next = (t + 1 < t1) ? tcol[t + 1] : INT_MAX;   // loaded early: it gates the publish
            const int nslots = info & 255, npend = info >> 8;
            const int nknown = nslots - npend;
            __syncwarp();
            // this lane's undecided cell (the first 32; a second round is rare): everything that does
            // not depend on the wait is read now
            int p_n = 0, p_slot = 0;
            double p_w = 0.0, p_x = 0.0, p_y = 0.0;
            if (lane < npend) {
                p_n = W.rec.pn[lane]; p_slot = W.rec.pslot[lane];
                p_w = W.rec.pw[lane]; p_x = W.rec.px[lane]; p_y = W.rec.py[lane];
            }
            EXT_T(0);
            // ---- wait: rows j-4..j-1 past column i+4, own row past column i-1 ---------
            //   lanes 0-3: the tile holding column i+4 (normally this one), rows j-1-lane
            //   lanes 4-7: the tile holding column i-4 when that is the left neighbour
            //   lane 8   : the left neighbour's part of the own row
            if (lane < 9) {
                const int xr = min(i + 4, Nx - 1) / XT, xl = max(i - 4, 0) / XT;
                int jr = j, xq = -1, need = i + 4;
                if (lane < 4) { jr = j - 1 - lane; xq = xr; }
                else if (lane < 8) { jr = j - 1 - (lane - 4); xq = (xl != xr) ? xl : -1; }
                else { xq = (xl != xt) ? xl : -1; need = i - 1; }
                if (jr >= 1 && xq >= 0 && !(xq == xt && prev_ok && jr < row0)) {   // (own previous block: done)
                    if (xq == xt && jr >= row0) {
                        if (jr != j) {
                            if (sleep_ns > 0) while (sp[jr - row0] <= need) __nanosleep(sleep_ns);
                            else while (sp[jr - row0] <= need) { }
                        }
                    } else {
                        const int *g = prog + jr * nxt + xq;
                        { while (ld_relaxed_gpu(g) <= need) __nanosleep(100); ACQUIRE_GPU(); }
                    }
                }
            }
            __syncwarp();
            SMEM_ORDER();
            EXT_T(1);
            // ---- phase C: one lane per undecided cell ----------------------------------
            int nfail = 0;
            for (int r0 = 0; r0 < npend; r0 += 32) {
                const int r = r0 + lane;
                bool fail = false;
                if (r < npend) {
                    const bool pre = (r0 == 0);
                    const int n = pre ? p_n : W.rec.pn[r];
                    const int dj = n / 9 - 4, di = n % 9 - 4;
                    const int jj = j + dj, ii = i + di;
                    const int slot = ii & (RING - 1);
                    const bool mine = (ii >= xc0 && ii < xc1);
                    double v1 = 0.0, v2 = 0.0;
                    bool got = false, smem_row = false;
                    if (mine && jj >= row0) {                 // a row of this block: its live ring (seqlock)
                        smem_row = true;
                        SweepWarp &R = S.w[jj - row0];
                        volatile int *tag = &R.ring_tag[slot];
                        volatile double *rv = R.ring_v[slot];
                        if (*tag == ii) {
                            v1 = rv[0]; v2 = rv[1];
                            SMEM_ORDER();
                            got = (*tag == ii);
                        }
                    } else if (mine && prev_ok && jj >= prow0) {   // last rows of the previous block: frozen rings
                        smem_row = true;
                        if (S.prev_tag[jj - prow0][slot] == ii) {
                            v1 = S.prev_v[jj - prow0][slot][0]; v2 = S.prev_v[jj - prow0][slot][1];
                            got = true;
                        }
                    }
                    if (!got) {
                        const size_t cc = (size_t)jj * Nx + ii;
                        if (smem_row) {                       // slot recycled long ago / target rejected
                            if (ld_cta_u8(st + cc) == ST_FRESH) {
                                v1 = ld_cta_f64(X1e + cc); v2 = ld_cta_f64(X2e + cc);
                                got = true;
                            }
                        } else if (__ldcg(st + cc) == ST_FRESH) {   // fitted by another tile (L2)
                            v1 = __ldcg(X1e + cc); v2 = __ldcg(X2e + cc);
                            got = true;
                        }
                    }
                    // a rejected target contributes exact zeros (weight 0)
                    const double cw = pre ? p_w : W.rec.pw[r], cx = pre ? p_x : W.rec.px[r],
                                 cy = pre ? p_y : W.rec.py[r];
                    store_products(W.rec.prod[pre ? p_slot : W.rec.pslot[r]], got ? cw : 0.0, got ? cx : 0.0,
                                   got ? cy : 0.0, v1, v2);
                    fail = !got;
                }
                nfail += __popc(__ballot_sync(0xffffffffu, fail));
            }
            const int count = nknown + npend - nfail;    // known cells in the window (functions.py:147)
            __syncwarp();
            EXT_T(2);
            // ---- ordered accumulation: lane a owns running sum a ----------------------
            double acc = 0.0;
            {
                const double *q = &W.rec.prod[0][la];
                int k = 0;
                for (; k + 8 <= nslots + 7 && k + 8 <= WIN; k += 8) {      // zero-padded chunks of 8
                    const double q0 = q[(k + 0) * NACC], q1 = q[(k + 1) * NACC], q2 = q[(k + 2) * NACC],
                                 q3 = q[(k + 3) * NACC], q4 = q[(k + 4) * NACC], q5 = q[(k + 5) * NACC],
                                 q6 = q[(k + 6) * NACC], q7 = q[(k + 7) * NACC];
                    acc += q0; acc += q1; acc += q2; acc += q3; acc += q4; acc += q5; acc += q6; acc += q7;
                }
                for (; k < nslots; ++k) acc += q[k * NACC];                // (window of 81: last odd slot)
            }
            if (lane < NACC) W.sums[lane] = acc;
            __syncwarp();
            EXT_T(3);
            const volatile double *sm = W.sums;
            const int h = (lane & 1) * 3;                 // lane 0 solves for xi1, lane 1 for xi2
            const double b0 = sm[h], b1 = sm[h + 1], b2 = sm[h + 2];
            const double A00 = sm[6], A01 = sm[7], A02 = sm[8], A11 = sm[9], A12 = sm[10], A22 = sm[11];
            const double A10 = A01, A20 = A02, A21 = A12;
            const double m00 = A11 * A22 - A12 * A21, m01 = A10 * A22 - A12 * A20, m02 = A10 * A21 - A11 * A20;
            const double det = (A00 * m00 - A01 * m01 + A02 * m02);
            const bool fitted = (count >= 3) && (fabs(det) > 1e-10);
            // Cramer's rule (fast_solve_3x3; its own |det| >= 1e-15 gate holds whenever fitted)
            const double inv_det = 1.0 / det;
            const double cx = (b0 * m00 - A01 * (b1 * A22 - A12 * b2) + A02 * (b1 * A21 - A11 * b2)) * inv_det;
            const double cy = (A00 * (b1 * A22 - A12 * b2) - b0 * m01 + A02 * (A10 * b2 - b1 * A20)) * inv_det;
            const double cz = (A00 * (A11 * b2 - b1 * A21) - A01 * (A10 * b2 - b1 * A20) + b0 * m02) * inv_det;
            const double val = cx + cy * x0 + cz * y0;
            EXT_T(4);
            // ---- publish: shared memory first (that is what the next row waits for) ----
            {
                const size_t c = (size_t)j * Nx + i;
                const int slot = i & (RING - 1);
                if (fitted) {
                    if (lane == 0) {
                        // the entry being recycled must be readable from global memory first
                        if (W.ring_tag[slot] >= 0) __threadfence_block();
                        *((volatile int *)&W.ring_tag[slot]) = -1;        // seqlock: invalidate, write, validate
                    }
                    SMEM_ORDER();
                    if (lane < 2) *((volatile double *)&W.ring_v[slot][lane]) = val;
                    SMEM_ORDER();
                    if (lane == 0) *((volatile int *)&W.ring_tag[slot]) = i;
                    SMEM_ORDER();
                }
                if (lane == 0) sp[wib] = next;                // shared-memory stores retire in order
                SMEM_ORDER();
                // the record buffer was drained by the accumulation: fetch the next target's record
                if (t + 1 < t1 && t + 1 < cap) ext_fetch(&W.rec, recs + t + 1, info_next, lane);
                if (fitted) {
                    if (lane == 0) X1e[c] = val;
                    if (lane == 1) X2e[c] = val;
                    __syncwarp();
                    if (lane == 0) st[c] = ST_FRESH;
                }
                // global marker: only where another tile can be waiting on this row; the fence is
                // needed only if that tile can also read this row's values
                if (lane == 0) {
                    const bool data = (last_rb && wib >= RB - 4) || i < xc0 + 8 || i >= xc1 - 8;
                    if (data || (next == INT_MAX && edgy)) {
                        __threadfence();
                        st_release(prog + j * nxt + xt, next);
                    } else if (next == INT_MAX) {
                        st_relaxed_gpu(prog + j * nxt + xt, next);
                    }
                }
            }
            __syncwarp();
            EXT_T(5);
#ifdef RMT_EXT_TIMING
            if (lane == 0) atomicAdd(&g_ext_dbg[7], 1ull);
#endif
        }
        __syncthreads();
        if (wib >= RB - 4) {                                   // freeze the last four rows' rings
            if (lane < RING) {
                S.prev_tag[wib - (RB - 4)][lane] = W.ring_tag[lane];
                S.prev_v[wib - (RB - 4)][lane][0] = W.ring_v[lane][0];
                S.prev_v[wib - (RB - 4)][lane][1] = W.ring_v[lane][1];
            }
            if (wib == RB - 4 && lane == 0) S.prev_row0 = row0 + RB - 4;
        }
        __syncthreads();                                       // shared state is reused by the next row block
      }
    }
}

// ===========================================================================================
// All layers in ONE launch.
//
// Layer l+1 may trail layer l by a few rows: a layer-(l+1) fit at row j needs layer l complete
// on rows j-4..j+4 (its window) and rows j-1..j+1 (to know whether it is a target at all).  So
// the three sweeps run as three wavefronts a few rows apart, on different SMs, and the serial
// chain along a body costs ~one layer instead of the sum of all layers.  Because a rejected fit
// (|det| <= 1e-10, count < 3) changes which cells the NEXT layer fits, targets cannot be listed
// up front: every (layer, row, x-tile) warp discovers its own targets from the state bytes once
// the previous layer has passed, prepares their phase-A records (to its own global scratch
// slots), and then sweeps them exactly like k_ext_sweep.  State byte: 0 unknown, 1 known at
// entry, 2l+3 fitted by layer l.
//
// Tasks = (macro-row m, layer l, x-tile xt), handed out in that lexicographic order by a global
// counter to co-resident CTAs.  A task waits only on tasks handed out before it, or at most
// (2L-1) groups of nxt tasks after it (layer l-1 of the next macro-row, for its last four rows);
// the launcher keeps (2L-1)*nxt below the number of resident CTAs, so every awaited task is
// resident or done: no deadlock.
constexpr int CAPW = 16;               // phase-A records prepared per row warp (beyond: inline)
constexpr int LMAX = 512;              // targets listed per (row, x-tile)

template <int RBT>
struct FusedSmemT {
    SweepWarp w[RBT];
    double prev_v[4][RING][2];
    int prev_tag[4][RING];
    int prev_row0;
    int prog[RBT];
    int tile;
    int tinf[2][RBT][CAPW];            // double-buffered: a warp prepares its next row block's list and
    unsigned short tl[2][RBT][LMAX];   // records while the warps below it are still sweeping the current one
};

// RBT = 16: one CTA per SM; RBT = 8: two per SM (half the rows per block, half the shared memory) --
// twice as many chains in flight for grids with more bodies than SMs, at a higher per-block overhead.
template <int RBT>
__global__ void __launch_bounds__(RBT * 32, 16 / RBT)
k_ext_fused(double *__restrict__ X1e, double *__restrict__ X2e, unsigned char *__restrict__ st,
            const int *__restrict__ cnt0 /* layer-0 target counts per (row, x-tile), then their first and last
                                            columns and their 32-column chunk masks ([Ny*nxt] each) */,
            int *__restrict__ prog /* [L][Ny*nxt] */, int *__restrict__ tile_counter,
            ExtRec *__restrict__ scratch /* [gridDim][RBT][CAPW] */, const int *__restrict__ mode, int L, int Ny,
            int Nx, int joff, int nxt, int XT, double dx, double dy, double r2)
{
    if (mode[0] != RBT) return;                 // k_ext_decide picked another variant
    const int MRB = mode[1] / RBT;              // macro-tile height (rows, chosen by k_ext_decide) in row blocks
    extern __shared__ unsigned char s_raw[];
    FusedSmemT<RBT> &S = *reinterpret_cast<FusedSmemT<RBT> *>(s_raw);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SweepWarp &W = S.w[wib];
    volatile int *sp = S.prog;
    const int nseg = Ny * nxt;
    const int nrb = (Ny - 2 + RBT - 1) / RBT, nmrb = (nrb + MRB - 1) / MRB;
    const int ntasks = nmrb * L * nxt;
    const int la = (lane < NACC) ? lane : 0;
    ExtRec *myrecs2 = scratch + ((size_t)blockIdx.x * RBT + wib) * 2 * CAPW;   // two buffers of CAPW records

    for (;;) {
        if (threadIdx.x == 0) S.tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int task = S.tile;
        if (task >= ntasks) break;
        const int xt = task % nxt, grp = task / nxt, layer = grp % L, mrb = grp / L;
        const int fresh = 2 * layer + 3;
        int *progL = prog + (size_t)layer * nseg;
        const int *progP = prog + (size_t)(layer - 1) * nseg;       // previous layer (layer > 0)
        const int xc0 = xt * XT, xc1 = min(xc0 + XT, Nx);
        const int rb_end = min((mrb + 1) * MRB, nrb);
        if (threadIdx.x == 0) S.prev_row0 = -100;
        int cur = 0, pre_rb = -1, pre_nt = 0;                       // list / record buffer in use; block prepared ahead
        const int pre_warps = mode[2] ? mode[2] : RBT / 2;
      for (int rb = mrb * MRB; rb < rb_end; ++rb) {
        const int row0 = 1 + rb * RBT;
        const int j = row0 + wib;
        const bool live = j < Ny - 1;
        // ---- can this block hold targets of this layer at all?  (layer-l targets lie within l cells of
        //      layer-0 targets; cnt0 is exact for layer 0 and conservative beyond)
        int any = 0;
        for (int e = threadIdx.x; e < (RBT + 2 * layer + 2) * 3; e += blockDim.x) {
            const int jr = row0 - layer - 1 + e / 3, xq = xt - 1 + e % 3;
            if (jr >= 0 && jr < Ny && xq >= 0 && xq < nxt) any |= cnt0[jr * nxt + xq];
        }
        if (__syncthreads_or(any) == 0) {
            if (live && lane == 0) st_relaxed_gpu(progL + j * nxt + xt, INT_MAX);   // nothing here, ever
            continue;
        }
        if (lane == 0) sp[wib] = 0;                                 // 0 = list not built yet
        if (lane < RING) W.ring_tag[lane] = -1;
        __syncthreads();
        const int prow0 = S.prev_row0;
        const bool prev_ok = (prow0 == row0 - 4);
        const bool last_rb = (rb == rb_end - 1);
        // ---- per-row preparation (steps 1, 2, 4 below) for row jr into list / record buffer bf; returns
        //      the number of targets.  Depends only on EARLIER layers, so a warp also runs it for its row
        //      of the next block as soon as its own sweep is done (the chain is then further down).
        auto prepare_row = [&](int jrow, int bf) -> int {
            int cnt = 0;
            unsigned short *tl = S.tl[bf][wib];
            // (1) previous layer complete on rows jrow-4..jrow+4, tiles xt-1..xt+1.
            //     A neighbouring tile matters only through the few columns next to the shared boundary
            //     (window: 4, discovery: 1).  Targets of layer l-1 lie within l-1 cells of layer-0 targets,
            //     so the static first / last layer-0 target columns tell when it cannot have any there
            //     -- then its progress (its body may sit dozens of rows lower) is not waited for.
            if (layer > 0 && lane < 27) {
                const int jr = jrow - 4 + lane / 3, xq = xt - 1 + lane % 3;
                if (jr >= 1 && jr < Ny - 1 && xq >= 0 && xq < nxt) {
                    bool need = true;
                    if (xq != xt) {
                        const int *cmin0 = cnt0 + nseg, *cmax0 = cnt0 + 2 * nseg;
                        const int reach = (layer - 1) + 6, bnd = (xq < xt) ? xc0 : xc1;
                        need = false;
                        for (int r = max(jr - (layer - 1), 0); r <= min(jr + (layer - 1), Ny - 1) && !need; ++r) {
                            const int far_lo = cmin0[r * nxt + xq], far_hi = cmax0[r * nxt + xq];
                            const int own_lo = cmin0[r * nxt + xt], own_hi = cmax0[r * nxt + xt];
                            need = (xq < xt) ? (far_hi + reach >= bnd || own_lo - reach < bnd)
                                             : (far_lo - reach < bnd || own_hi + reach >= bnd);
                        }
                    }
                    if (need)
                    {
                        while (ld_relaxed_gpu(progP + jr * nxt + xq) != INT_MAX) __nanosleep(200);
                        ACQUIRE_GPU();
                    }
                }
            }
            __syncwarp();
            // (2) discover the row's targets: not known before this layer, with a known neighbour.  They lie
            //     within `layer` cells of layer-0 targets, so only the 32-column chunks that the layer-0 chunk
            //     masks of rows jrow-layer..jrow+layer (dilated by one chunk, plus the neighbouring tiles'
            //     edge columns) mark are scanned.
            unsigned chunks;
            {
                unsigned m = 0;
                const int r = jrow - layer + lane;
                if (lane <= 2 * layer && r >= 0 && r < Ny) {
                    const int *cmin0 = cnt0 + nseg, *cmax0 = cnt0 + 2 * nseg, *chunk0 = cnt0 + 3 * nseg;
                    m = (unsigned)chunk0[r * nxt + xt];
                    if (xt > 0 && cmax0[r * nxt + xt - 1] + layer + 1 >= xc0) m |= 1u;
                    if (xt + 1 < nxt && cmin0[r * nxt + xt + 1] - layer - 1 < xc1) m |= 1u << ((xc1 - 1 - xc0) >> 5);
                }
                m = __reduce_or_sync(0xffffffffu, m);
                chunks = m | (m << 1) | (m >> 1);
            }
            const unsigned char *rowc = st + (size_t)jrow * Nx, *rowa = rowc - Nx, *rowb = rowc + Nx;
            for (int base = xc0; base < xc1; base += 32) {
                if (!((chunks >> ((base - xc0) >> 5)) & 1u)) continue;
                const int i = base + lane;
                bool tgt = false;
                if (i < xc1 && i >= 1 && i < Nx - 1) {
                    const unsigned me = __ldcg(rowc + i);
                    if (!((me & 1) && me < fresh)) {
                        unsigned kb = 0;
#pragma unroll
                        for (int d = -1; d <= 1; ++d) {
                            const unsigned a = __ldcg(rowa + i + d), c = __ldcg(rowb + i + d);
                            kb |= ((a & 1) && a < fresh) | ((c & 1) && c < fresh);
                            if (d) { const unsigned b = __ldcg(rowc + i + d); kb |= ((b & 1) && b < fresh); }
                        }
                        tgt = kb != 0;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, tgt);
                if (tgt) {
                    const int o = cnt + __popc(m & ((1u << lane) - 1u));
                    if (o < LMAX) tl[o] = (unsigned short)(i - xc0);
                }
                cnt += __popc(m);
            }
            // (more than LMAX targets in one row segment cannot be listed: the launcher keeps XT <= LMAX)
            __syncwarp();
            // (4) phase A of the first CAPW targets into this warp's scratch records
            const int nprep = min(cnt, CAPW);
            for (int k = 0; k < nprep; ++k) {
                const int info = ext_phase_a<1>(myrecs2 + bf * CAPW + k, X1e, X2e, st, jrow, xc0 + tl[k], Ny, Nx,
                                                   joff, dx, dy, r2, lane, fresh);
                if (lane == 0) S.tinf[bf][wib][k] = info;
            }
            if (nprep) __threadfence();                            // records readable by this warp's cp.async
            __syncwarp();
            return cnt;
        };
        int nt = 0;
        if (live) {
            if (pre_rb == rb) {                                    // prepared while the previous block was swept
                nt = pre_nt;
            } else {
                nt = prepare_row(j, cur);
            }
            // ---- (3) announce the list: the row is now "at" its first target -----------------
            const int first = nt ? xc0 + S.tl[cur][wib][0] : INT_MAX;
            if (lane == 0) {
                sp[wib] = first;
                st_relaxed_gpu(progL + j * nxt + xt, first);
            }
        } else if (lane == 0) {
            sp[wib] = INT_MAX;
        }
        const unsigned short *tlw = S.tl[cur][wib];
        const int *tinfw = S.tinf[cur][wib];
        ExtRec *myrecs = myrecs2 + cur * CAPW;

        // ---- (5) the sweep of this row, as in k_ext_sweep -----------------------------------------
        int info_next = nt ? (0 < CAPW ? tinfw[0] : 0) : 0;
        if (nt) ext_fetch(&W.rec, myrecs, info_next, lane);
        for (int t = 0; t < nt; ++t) {
            const int i = xc0 + tlw[t];
            const double x0 = dx * i, y0 = dy * (j + joff);
            int info;
            if (t < CAPW) {
                info = info_next;
                cp_async_wait_all();
            } else {
                info = ext_phase_a<1>(&W.rec, X1e, X2e, st, j, i, Ny, Nx, joff, dx, dy, r2, lane, fresh);
            }
            info_next = (t + 1 < nt && t + 1 < CAPW) ? tinfw[t + 1] : 0;
            const int next = (t + 1 < nt) ? xc0 + tlw[t + 1] : INT_MAX;
            const int nslots = info & 255, npend = info >> 8;
            const int nknown = nslots - npend;
            __syncwarp();
            // ---- wait: rows j-4..j-1 past column i+4, own row past column i-1 (same layer) ---------
            if (lane < 9) {
                const int xr = min(i + 4, Nx - 1) / XT, xl = max(i - 4, 0) / XT;
                int jr = j, xq = -1, need = i + 4;
                if (lane < 4) { jr = j - 1 - lane; xq = xr; }
                else if (lane < 8) { jr = j - 1 - (lane - 4); xq = (xl != xr) ? xl : -1; }
                else { xq = (xl != xt) ? xl : -1; need = i - 1; }
                if (jr >= 1 && xq >= 0 && !(xq == xt && prev_ok && jr < row0)) {
                    if (xq == xt && jr >= row0) {
                        if (jr != j) while (sp[jr - row0] <= need) __nanosleep(20);
                    } else {
                        const int *g = progL + jr * nxt + xq;
                        { while (ld_relaxed_gpu(g) <= need) __nanosleep(100); ACQUIRE_GPU(); }
                    }
                }
            }
            __syncwarp();
            SMEM_ORDER();
            // ---- phase C: one lane per undecided cell ----------------------------------
            int nfail = 0;
            for (int q0 = 0; q0 < npend; q0 += 32) {
                const int r = q0 + lane;
                bool fail = false;
                if (r < npend) {
                    const int n = W.rec.pn[r];
                    const int dj = n / 9 - 4, di = n % 9 - 4;
                    const int jj = j + dj, ii = i + di;
                    const int slot = ii & (RING - 1);
                    const bool mine = (ii >= xc0 && ii < xc1);
                    double v1 = 0.0, v2 = 0.0;
                    bool got = false, smem_row = false;
                    if (mine && jj >= row0) {
                        smem_row = true;
                        SweepWarp &R = S.w[jj - row0];
                        volatile int *tag = &R.ring_tag[slot];
                        volatile double *rv = R.ring_v[slot];
                        if (*tag == ii) {
                            v1 = rv[0]; v2 = rv[1];
                            SMEM_ORDER();
                            got = (*tag == ii);
                        }
                    } else if (mine && prev_ok && jj >= prow0) {
                        smem_row = true;
                        if (S.prev_tag[jj - prow0][slot] == ii) {
                            v1 = S.prev_v[jj - prow0][slot][0]; v2 = S.prev_v[jj - prow0][slot][1];
                            got = true;
                        }
                    }
                    if (!got) {
                        const size_t cc = (size_t)jj * Nx + ii;
                        if (smem_row) {
                            if (ld_cta_u8(st + cc) == fresh) {
                                v1 = ld_cta_f64(X1e + cc); v2 = ld_cta_f64(X2e + cc);
                                got = true;
                            }
                        } else if (__ldcg(st + cc) == fresh) {
                            v1 = __ldcg(X1e + cc); v2 = __ldcg(X2e + cc);
                            got = true;
                        }
                    }
                    store_products(W.rec.prod[W.rec.pslot[r]], got ? W.rec.pw[r] : 0.0, got ? W.rec.px[r] : 0.0,
                                   got ? W.rec.py[r] : 0.0, v1, v2);
                    fail = !got;
                }
                nfail += __popc(__ballot_sync(0xffffffffu, fail));
            }
            const int count = nknown + npend - nfail;
            __syncwarp();
            // ---- ordered accumulation ----------------------------------------------------
            double acc = 0.0;
            {
                const double *q = &W.rec.prod[0][la];
                int k = 0;
                for (; k + 8 <= nslots + 7 && k + 8 <= WIN; k += 8) {
                    const double q0 = q[(k + 0) * NACC], q1 = q[(k + 1) * NACC], q2 = q[(k + 2) * NACC],
                                 q3 = q[(k + 3) * NACC], q4 = q[(k + 4) * NACC], q5 = q[(k + 5) * NACC],
                                 q6 = q[(k + 6) * NACC], q7 = q[(k + 7) * NACC];
                    acc += q0; acc += q1; acc += q2; acc += q3; acc += q4; acc += q5; acc += q6; acc += q7;
                }
                for (; k < nslots; ++k) acc += q[k * NACC];
            }
            if (lane < NACC) W.sums[lane] = acc;
            __syncwarp();
            const volatile double *sm = W.sums;
            const int h = (lane & 1) * 3;
            const double b0 = sm[h], b1 = sm[h + 1], b2 = sm[h + 2];
            const double A00 = sm[6], A01 = sm[7], A02 = sm[8], A11 = sm[9], A12 = sm[10], A22 = sm[11];
            const double A10 = A01, A20 = A02, A21 = A12;
            const double m00 = A11 * A22 - A12 * A21, m01 = A10 * A22 - A12 * A20, m02 = A10 * A21 - A11 * A20;
            const double det = (A00 * m00 - A01 * m01 + A02 * m02);
            const bool fitted = (count >= 3) && (fabs(det) > 1e-10);
            const double inv_det = 1.0 / det;
            const double cx = (b0 * m00 - A01 * (b1 * A22 - A12 * b2) + A02 * (b1 * A21 - A11 * b2)) * inv_det;
            const double cy = (A00 * (b1 * A22 - A12 * b2) - b0 * m01 + A02 * (A10 * b2 - b1 * A20)) * inv_det;
            const double cz = (A00 * (A11 * b2 - b1 * A21) - A01 * (A10 * b2 - b1 * A20) + b0 * m02) * inv_det;
            const double val = cx + cy * x0 + cz * y0;
            // ---- publish ---------------------------------------------------------------------
            {
                const size_t c = (size_t)j * Nx + i;
                const int slot = i & (RING - 1);
                if (fitted) {
                    if (lane == 0) {
                        if (W.ring_tag[slot] >= 0) __threadfence_block();
                        *((volatile int *)&W.ring_tag[slot]) = -1;
                    }
                    SMEM_ORDER();
                    if (lane < 2) *((volatile double *)&W.ring_v[slot][lane]) = val;
                    SMEM_ORDER();
                    if (lane == 0) *((volatile int *)&W.ring_tag[slot]) = i;
                    SMEM_ORDER();
                }
                if (lane == 0) sp[wib] = next;
                SMEM_ORDER();
                if (t + 1 < nt && t + 1 < CAPW) ext_fetch(&W.rec, myrecs + t + 1, info_next, lane);
                if (fitted) {
                    if (lane == 0) X1e[c] = val;
                    if (lane == 1) X2e[c] = val;
                    __syncwarp();
                    if (lane == 0) st[c] = (unsigned char)fresh;
                }
                // global marker: the next layer reads every row (fence at the row's end); this layer's
                // other tiles read the rows/columns next to them
                if (lane == 0) {
                    const bool data = (last_rb && wib >= RBT - 4) || i < xc0 + 8 || i >= xc1 - 8;
                    if (data || next == INT_MAX) {
                        __threadfence();
                        st_release(progL + j * nxt + xt, next);
                    }
                }
            }
            __syncwarp();
        }
        // ---- (6) this warp's row is done: prepare its row of the NEXT block while the warps below sweep.
        //      The emptiness test is the block-wide one of the loop head, evaluated per warp (static data).
        //      Only the upper half of the warps: they finish first (slack until the chain has run down the
        //      block) and their rows are needed first in the next block; the lower half prepares at the next
        //      block's start, behind the chain's passage through the upper rows.
        pre_rb = -1;
        if (rb + 1 < rb_end && wib < pre_warps) {
            const int jn = j + RBT, rown = row0 + RBT;
            int anyn = 0;
            for (int e = lane; e < (RBT + 2 * layer + 2) * 3; e += 32) {
                const int jr = rown - layer - 1 + e / 3, xq = xt - 1 + e % 3;
                if (jr >= 0 && jr < Ny && xq >= 0 && xq < nxt) anyn |= cnt0[jr * nxt + xq];
            }
            if (__any_sync(0xffffffffu, anyn != 0)) {
                pre_rb = rb + 1;
                pre_nt = 0;
                if (jn < Ny - 1) {
                    pre_nt = prepare_row(jn, cur ^ 1);
                    // the marker may run ahead: "every column before the first target is dealt with"
                    if (lane == 0)
                        st_relaxed_gpu(progL + jn * nxt + xt, pre_nt ? xc0 + S.tl[cur ^ 1][wib][0] : INT_MAX);
                }
            }
        }
        __syncthreads();
        if (wib >= RBT - 4) {
            if (lane < RING) {
                S.prev_tag[wib - (RBT - 4)][lane] = W.ring_tag[lane];
                S.prev_v[wib - (RBT - 4)][lane][0] = W.ring_v[lane][0];
                S.prev_v[wib - (RBT - 4)][lane][1] = W.ring_v[lane][1];
            }
            if (wib == RBT - 4 && lane == 0) S.prev_row0 = row0 + RBT - 4;
        }
        __syncthreads();
        if (pre_rb == rb + 1) cur ^= 1;
      }
    }
}

// ===========================================================================================
// All layers in one launch, ONE CTA PER ISOLATED TILE ("body" variant).
//
// When every band of targets keeps a margin of (layers + 5) cells to the borders of its tile, no fit
// of any layer reads a cell that another tile writes: tiles are independent, and a tile's whole
// dependency chain can stay inside one CTA.  The chain is then run by ONE warp per layer, which fits
// the layer's targets strictly in raster order -- exactly the reference's serial sweep restricted to
// the tile -- so nothing on the chain waits for another warp: no progress polling, no seqlocks, no
// cross-SM traffic.  Everything that does not depend on the layer's own fits is done ahead of the
// chain warp by helper warps of the same CTA:
//   discovery warp (one per layer): scans the rows top to bottom (only the 32-column chunks the
//       layer-0 chunk masks mark) once the previous layer has finished rows <= j+4, and appends the
//       targets to a queue in raster order;
//   prepare warps (P per layer): take queue entries in turn and run phase A (81 window
//       classifications, weights, the 12 products of every cell known before the layer, the to-do list
//       of raster-earlier undecided cells) straight into a ring of records in shared memory;
//   chain warp: for each record in order: one lane per undecided cell reads its state and value
//       (written by this very warp a few fits ago: L1-coherent block-scope loads), fills its slot,
//       twelve lanes run the ordered sums, lanes 0/1 solve and publish.
// Layer l+1 trails layer l by five rows inside the same CTA (row_done hand-off through shared
// memory).  Cost: ~0.5 us per target of the fullest tile and layer, against ~1.8 us per DEPENDENT fit
// plus per-row-block overheads in the row-pipelined variants -- it wins whenever the bodies are
// isolated and a tile holds at most a few thousand targets (k_ext_decide compares the models).
// Spin-loop watchdog of the body kernel: a wait that lasts ~2^23 polls (seconds) is a protocol failure;
// the warp records where it was and leaves, so the kernel ends (and the result is wrong) instead of
// hanging the device.  [layer][role 0 chain / 1 discovery / 2 prepare][8]
__device__ int g_body_watch[8 * 3 * 8];
#define BODY_WATCH_NOW(code, tval)                                                                     \
    do {                                                                                               \
        if (lane == 0) {                                                                               \
            int *w__ = g_body_watch + (layer * 3 + (role == 0 ? 0 : role <= W ? 1 : 2)) * 8;                         \
            w__[0] = (code); w__[1] = (tval); w__[2] = C.q_count; w__[3] = C.consumed;                 \
            w__[4] = C.row_done; w__[5] = C.disc_row; w__[6] = C.prep_next; w__[7] = C.q_final;        \
        }                                                                                              \
    } while (0)
#define BODY_WATCH(code, tval)                                                                         \
    do {                                                                                               \
        if (++spins > (1 << 23)) {                                                                     \
            if (lane == 0) {                                                                           \
                int *w__ = g_body_watch + (layer * 3 + (role == 0 ? 0 : role <= W ? 1 : 2)) * 8;                     \
                w__[0] = (code); w__[1] = (tval); w__[2] = C.q_count; w__[3] = C.consumed;             \
                w__[4] = C.row_done; w__[5] = C.disc_row; w__[6] = C.prep_next; w__[7] = C.q_final;    \
            }                                                                                          \
            return;                                                                                    \
        }                                                                                              \
    } while (0)
// discovery-side fence: block scope, or cluster scope when a helper CTA on another SM reads what this CTA wrote
#define BODY_FENCE() do { if (clustered) asm volatile("fence.acq_rel.cluster;" ::: "memory"); else __threadfence_block(); } while (0)
constexpr int BQ = 256;                // target queue entries per layer (ring)
constexpr int BSLOT_MAX = 6;           // records per layer in shared memory (ring)
constexpr int BODY_KMAX = 6;           // candidate tile sizes: (512 << k) rows x (XT << k) columns
constexpr int BODY_LMAX = 5;           // layers the 640-thread CTA has warps for

constexpr int BC_ROWS = 8, BC_COLS = 128;  // chain-private cache of the layer's latest fits: [row & 7][column & 127]
__device__ __forceinline__ int body_cache_slot(int jj, int ii)
{
    // bodies on a lattice sit a multiple of 128 columns apart: fold the high column bits in
    return (jj & (BC_ROWS - 1)) * BC_COLS + ((ii + (ii >> 7) * 41) & (BC_COLS - 1));
}
constexpr int BODY_DEP = 1 << 16;          // record header: the target's window holds the PREVIOUS queue entry
struct BodyCtl {
    int q_count;                       // targets discovered so far
    int q_final;                       // discovery complete
    int disc_row;                      // discovery has scanned every row <= disc_row
    int row_done;                      // the chain warp has fitted every target of rows <= row_done
    int prep_next;                     // next queue entry a prepare warp takes
    int consumed;                      // the chain warp has consumed entries < consumed
    int disc_turn;                     // the row whose discovery warp may append now
    int q_tail;                        // queue length including the rows appended so far
    int4 rhdr[BSLOT_MAX];              // {info, j, i, t + 1} once the record of queue entry t sits in the slot
    int qrow[BQ], qcol[BQ];
    double sums[2][NACC];              // one set per half warp (two independent targets are fitted at once)
    // what the chain warp itself fitted lately (single writer, single reader: no synchronisation):
    // tag = j << 15 | i, bit 30 set for a rejected target; values (xi1, xi2)
    int ctag[BC_ROWS][BC_COLS];
    double cval[BC_ROWS][BC_COLS][2];
};
static_assert(sizeof(BodyCtl) % 16 == 0, "BodyCtl must keep 16-byte alignment in an array");

__device__ __forceinline__ int body_tile_offset(int k, int Ny, int nxt)
{
    int off = 0;
    for (int q = 0; q < k; ++q) off += ((Ny + (512 << q) - 1) / (512 << q)) * ((nxt + (1 << q) - 1) >> q);
    return off;
}

// The work of one warp of the body kernel.  `ctl` / `recs` are generic pointers to the shared memory of the
// cluster's first CTA (the warp's own CTA, or -- for the prepare warps of the helper CTA -- the neighbouring
// SM's, reached through distributed shared memory).
template <bool REMOTE>
__device__ __forceinline__ void
body_work(double *__restrict__ X1e, double *__restrict__ X2e, unsigned char *__restrict__ st,
          const int *__restrict__ cnt0, const int *__restrict__ cmin0, const int *__restrict__ cmax0,
          const int *__restrict__ chunk0, BodyCtl *ctl, ExtRec *recs, const int layer, const int role,
          const bool clustered, const int W, const int nslot, const int Ny, const int Nx, const int joff, const int nxt,
          const int XT, const int seg0, const int seg1, const int R0, const int R1, const double dx,
          const double dy, const double r2, const int flags)
{
    const int lane = threadIdx.x & 31;
    const int fresh = 2 * layer + 3;
    volatile BodyCtl &C = ctl[layer];
    int spins = 0;
    ExtRec *myrecs = recs + (size_t)layer * nslot;

    if (role >= 1 && role <= W) {
        // ---------------- discovery: the layer's targets, in raster order, into the queue ----------
        // W warps take the rows in turn (row r of the tile belongs to warp r % W).  SCANNING a row for
        // targets reads only cells that are final for this layer (known / not known never depends on
        // this layer's own marks or fits), so a warp scans its row while the other one is still busy
        // with the row above; APPENDING to the queue is done strictly in row order (disc_turn).
        volatile BodyCtl *CP = layer > 0 ? &ctl[layer - 1] : nullptr;
        for (int j = R0 + (role - 1); j < R1; j += W) {
            // layer-l targets lie within l cells of layer-0 targets: rows j-l .. j+l of this tile
            int any = 0;
            spins = 0;
            if (lane <= 2 * layer) {
                const int r = j - layer + lane;
                if (r >= 0 && r < Ny)
                    for (int sg = seg0; sg < seg1; ++sg) any |= cnt0[r * nxt + sg];
            }
            const bool has = __any_sync(0xffffffffu, any != 0);
            // ---- scan: lane c keeps (first column, target mask) of the c-th flagged 32-column chunk
            int fbase = 0, nfc = 0;
            unsigned fmask = 0;
            bool overflow = false;                      // > 32 flagged chunks: the rest is scanned in turn
            const unsigned char *rowc = st + (size_t)j * Nx, *rowa = rowc - Nx, *rowb = rowc + Nx;
            if (has) {
                if (CP) {                               // previous layer complete on rows <= j+4 (window)
                    const int need = min(j + 4, R1 - 1);
                    while (CP->row_done < need) { __nanosleep(60); BODY_WATCH(1, j); }
                    BODY_FENCE();                        // cumulative: the helper CTA reads the previous layer's fits
                }
                for (int sg = seg0; sg < seg1 && !overflow; ++sg) {
                    const int xc0 = sg * XT, xc1 = min(xc0 + XT, Nx);
                    unsigned m = 0;
                    const int r = j - layer + lane;
                    if (lane <= 2 * layer && r >= 0 && r < Ny) {
                        m = (unsigned)chunk0[r * nxt + sg];
                        if (sg > seg0 && cmax0[r * nxt + sg - 1] + layer + 1 >= xc0) m |= 1u;
                        if (sg + 1 < seg1 && cmin0[r * nxt + sg + 1] - layer - 1 < xc1)
                            m |= 1u << ((xc1 - 1 - xc0) >> 5);
                    }
                    m = __reduce_or_sync(0xffffffffu, m);
                    // layer 0: the chunk masks of row j are exact; deeper layers reach `layer` cells further
                    const unsigned chunks = layer ? (m | (m << 1) | (m >> 1)) : m;
                    // Block-scope coherent loads cost an L2 round trip each, and this scan has to stay ahead of
                    // the chain warp: the nine bytes of THREE flagged chunks are requested at once, then judged.
                    const int nchunk = (xc1 - xc0 + 31) >> 5;
                    unsigned rem = chunks & (nchunk >= 32 ? 0xffffffffu : ((1u << nchunk) - 1u));
                    while (rem) {
                        if (nfc + 3 > 32) { overflow = true; break; }
                        int bs[3];
                        unsigned v[3][9];
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            bs[u] = -1;
                            if (rem) { bs[u] = xc0 + 32 * (__ffs(rem) - 1); rem &= rem - 1; }
                            const int i = bs[u] + lane;
                            const bool ok = bs[u] >= 0 && i < xc1 && i >= 1 && i < Nx - 1;
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                v[u][d] = ok ? ld_cta_u8(rowa + i + d - 1) : 0u;
                                v[u][3 + d] = ok ? ld_cta_u8(rowc + i + d - 1) : 0u;
                                v[u][6 + d] = ok ? ld_cta_u8(rowb + i + d - 1) : 0u;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            if (bs[u] < 0) continue;     // warp-uniform
                            const int i = bs[u] + lane;
                            const unsigned me = v[u][4];
                            bool tgt = false;
                            if (i < xc1 && i >= 1 && i < Nx - 1 && !((me & 1) && me < fresh)) {
                                unsigned kb = 0;
#pragma unroll
                                for (int e = 0; e < 9; ++e)
                                    if (e != 4) kb |= ((v[u][e] & 1) && v[u][e] < fresh);
                                tgt = kb != 0;
                            }
                            const unsigned tm = __ballot_sync(0xffffffffu, tgt);
                            if (tm) {
                                if (lane == nfc) { fbase = bs[u]; fmask = tm; }
                                ++nfc;
                            }
                        }
                    }
                }
            }
            // ---- my turn: append the row's targets, mark them, publish
            while (C.disc_turn != j) { __nanosleep(20); BODY_WATCH(6, j); }
            if (nfc || overflow) {
                int qn = C.q_tail;
                const int row_start = qn;               // entries [row_start, qn) of this row are not published yet
                bool streaming = overflow;              // a long row is published as it is found (no reordering)
                auto append = [&](int base, unsigned tm) -> bool {
                    if (!streaming && qn - row_start + __popc(tm) > 32) {   // too long to reorder: publish
                        streaming = true;
                        BODY_FENCE();
                        __syncwarp();
                        if (lane == 0) C.q_count = qn;
                    }
                    // room in the ring (one spare entry: the prepare warps look at entry t - 1)
                    while (qn + 33 - C.consumed > BQ) {
                        __nanosleep(200);
                        if (++spins > (1 << 23)) return false;
                    }
                    if ((tm >> lane) & 1u) {
                        const int o = (qn + __popc(tm & ((1u << lane) - 1u))) & (BQ - 1);
                        ((int *)C.qrow)[o] = j;
                        ((int *)C.qcol)[o] = base + lane;
                        // mark: "target of this layer, undecided" (even = still unknown to every
                        // known-test); phase A of later targets lists exactly the marked cells
                        ((unsigned char *)rowc)[base + lane] = (unsigned char)(fresh - 1);
                    }
                    qn += __popc(tm);
                    if (streaming) {
                        BODY_FENCE();
                        __syncwarp();
                        if (lane == 0) C.q_count = qn;
                    }
                    return true;
                };
                for (int c = 0; c < nfc; ++c) {
                    const int base = __shfl_sync(0xffffffffu, fbase, c);
                    const unsigned tm = __shfl_sync(0xffffffffu, fmask, c);
                    if (!append(base, tm)) { BODY_WATCH_NOW(2, qn); return; }
                }
                if (overflow) {
                    // the rare wide row: every chunk the scan above did not reach, one at a time
                    int seen = 0;
                    for (int sg = seg0; sg < seg1; ++sg) {
                        const int xc0 = sg * XT, xc1 = min(xc0 + XT, Nx);
                        for (int base = xc0; base < xc1; base += 32) {
                            const int i = base + lane;
                            bool tgt = false;
                            if (i < xc1 && i >= 1 && i < Nx - 1) {
                                const unsigned me = ld_cta_u8(rowc + i);
                                if (!((me & 1) && me < fresh)) {
                                    unsigned kb = 0;
#pragma unroll
                                    for (int d = -1; d <= 1; ++d) {
                                        const unsigned a = ld_cta_u8(rowa + i + d), c = ld_cta_u8(rowb + i + d);
                                        kb |= ((a & 1) && a < fresh) | ((c & 1) && c < fresh);
                                        if (d) { const unsigned b = ld_cta_u8(rowc + i + d); kb |= ((b & 1) && b < fresh); }
                                    }
                                    tgt = kb != 0;
                                }
                            }
                            const unsigned tm = __ballot_sync(0xffffffffu, tgt);
                            if (!tm) continue;
                            if (seen++ < nfc) continue;  // already appended from the scan (same order)
                            if (!append(base, tm)) { BODY_WATCH_NOW(2, qn); return; }
                        }
                    }
                }
                if (!streaming && qn > row_start) {
                    // Queue order within the row.  Runs of targets separated by >= 5 columns are independent
                    // inside the row (a fit reads its own row only 4 columns to the left), so the runs left
                    // and right of a gap near the middle are INTERLEAVED: L0 R0 L1 R1 ...  Neighbouring
                    // queue entries are then usually independent and the chain warp fits them two at a time.
                    // Every entry still follows all the raster-earlier targets of its window.
                    const int nrow = qn - row_start;
                    __syncwarp();
                    int col = 0;
                    if (lane < nrow) col = ((int *)C.qcol)[(row_start + lane) & (BQ - 1)];
                    const int prevc = __shfl_up_sync(0xffffffffu, col, 1);
                    const unsigned bnd = (flags & 1) ? __ballot_sync(0xffffffffu, lane > 0 && lane < nrow && col - prevc >= 5) : 0u;
                    if (bnd) {
                        // the boundary nearest the middle of the list
                        const int mid = nrow >> 1;
                        const unsigned lo = bnd & ((2u << mid) - 1u), hi = bnd & ~((2u << mid) - 1u);
                        int sp = -1;
                        if (lo) sp = 31 - __clz(lo);
                        if (hi) { const int h2 = __ffs(hi) - 1; if (sp < 0 || h2 - mid < mid - sp) sp = h2; }
                        const int na = sp, nb = nrow - sp;
                        __syncwarp();
                        if (lane < nrow) {
                            int pos;
                            if (lane < sp) pos = (lane < nb) ? 2 * lane : lane + nb;
                            else { const int k2 = lane - sp; pos = (k2 < na) ? 2 * k2 + 1 : k2 + na; }
                            ((int *)C.qcol)[(row_start + pos) & (BQ - 1)] = col;
                        }
                    }
                    BODY_FENCE();
                    __syncwarp();
                    if (lane == 0) C.q_count = qn;
                }
                if (lane == 0) C.q_tail = qn;
            }
            __syncwarp();
            if (lane == 0) {
                C.disc_row = j;
                if (j == R1 - 1) C.q_final = 1;
                __threadfence_block();
                C.disc_turn = j + 1;
            }
            __syncwarp();
        }
        return;
    }

    if (role > W) {
        // ---------------- prepare: phase A of queue entry t into record slot t % nslot --------------
        for (;;) {
            BODY_T0();
            int t = 0;
            spins = 0;
            if (lane == 0) t = atomicAdd((int *)&C.prep_next, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            for (;;) {
                const int fin = C.q_final;
                const int cnt = C.q_count;
                if (t < cnt) break;
                if (fin) return;
                __nanosleep(250);                       // idle helpers must not take issue slots from the chain warps
                BODY_WATCH(3, t);
            }
            BODY_T(8);
            if (REMOTE) asm volatile("fence.acq_rel.cluster;" ::: "memory"); else __threadfence_block();
            const int j = C.qrow[t & (BQ - 1)], i = C.qcol[t & (BQ - 1)];
            const int slot = t % nslot;
            // the record slot may still hold entry t - nslot: wait for it only when phase A is ready to store
            bool timed_out = false;
            auto slot_free = [&]() -> bool {
                BODY_T(10);
                while (t - C.consumed >= nslot) {
                    __nanosleep(150);
                    if (++spins > (1 << 23)) { timed_out = true; return false; }
                }
                if (REMOTE) asm volatile("fence.acq_rel.cluster;" ::: "memory"); else __threadfence_block();
                BODY_T(9);
                return true;
            };
            int info = ext_phase_a<REMOTE ? 3 : 2>(myrecs + slot, X1e, X2e, st, j, i, Ny, Nx, joff, dx, dy, r2, lane, fresh,
                                                   slot_free);
            if (timed_out) { BODY_WATCH_NOW(4, t); return; }
            if (t > 0) {                                // is the previous queue entry inside this target's window?
                const int jp = C.qrow[(t - 1) & (BQ - 1)], ip = C.qcol[(t - 1) & (BQ - 1)];
                if (j - jp >= 0 && j - jp <= 4 && abs(i - ip) <= 4 && (jp < j || ip < i)) info |= BODY_DEP;
            }
            if (REMOTE) asm volatile("fence.acq_rel.cluster;" ::: "memory"); else __threadfence_block();
            __syncwarp();
            if (lane == 0) {                            // one 16-byte store: header and ready flag travel together
                if (REMOTE) st_generic_v4((void *)&C.rhdr[slot], make_int4(info, j, i, t + 1));
                else sts_v4((void *)&C.rhdr[slot], make_int4(info, j, i, t + 1));
            }
            BODY_T(10);
            BODY_CNT(11);
        }
    }

    // ---------------- the chain: fit the layer's targets in raster order ---------------------------
    // Nothing this warp reads on the chain is written by another warp except the prepared record (shared
    // memory, in order per warp: a compiler barrier after the ready flag is enough), so the loop carries no
    // hardware fence; a release fence precedes each row_done update (once per row) for the next layer.
    // TWO targets per pass: lanes 0-15 fit the head of the queue, lanes 16-31 the entry behind it whenever
    // that one does not read the head's result (its record says so) -- on the flanks of a body the
    // interleaved queue order makes that the common case, and a pass costs the same for one target or two.
    const int half = lane >> 4, hl = lane & 15;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int la = (hl < NACC) ? hl : 0;
    int rdone = R0 - 1, cur_row = -1;
    int *ctag = (int *)&C.ctag[0][0];
    double *cval = (double *)&C.cval[0][0][0];
    int slot = 0;
    for (int t = 0;;) {
        BODY_T0();
        int4 hdr = lds_v4((const void *)&C.rhdr[slot]);
        if (hdr.w != t + 1) {
            bool finished = false;
            for (;;) {
                hdr = lds_v4((const void *)&C.rhdr[slot]);
                if (hdr.w == t + 1) break;
                const int dr = C.disc_row;              // read BEFORE the count: rows <= dr are all queued
                const int fin = C.q_final;
                const int cnt = C.q_count;
                if (cnt == t) {                         // everything discovered so far is fitted
                    if (fin) { finished = true; break; }
                    if (dr > rdone) {
                        rdone = dr;
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) C.row_done = dr;
                    }
                }
                __nanosleep(20);
                BODY_WATCH(5, t);
            }
            if (finished) break;
        }
        spins = 0;
        const int slot2 = (slot + 1 == nslot) ? 0 : slot + 1;
        // the decision is a snapshot of a flag another warp sets: take lane 0's view for the whole warp
        // (lanes that left the wait loop at different times could otherwise disagree)
        int4 hdr2 = lds_v4((const void *)&C.rhdr[slot2]);
        hdr2.x = __shfl_sync(0xffffffffu, hdr2.x, 0); hdr2.y = __shfl_sync(0xffffffffu, hdr2.y, 0);
        hdr2.z = __shfl_sync(0xffffffffu, hdr2.z, 0); hdr2.w = __shfl_sync(0xffffffffu, hdr2.w, 0);
        const bool paired = (flags & 2) && (hdr2.w == t + 2) && !(hdr2.x & BODY_DEP);
        SMEM_ORDER();
        BODY_T(0);
        if (hdr.y != cur_row) {                         // rows above the head's row are complete
            cur_row = hdr.y;
            if (cur_row - 1 > rdone) {
                rdone = cur_row - 1;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) C.row_done = rdone;
            }
        }
        const bool mine = half == 0 || paired;           // this half warp has a target
        const int4 hm = half ? hdr2 : hdr;
        ExtRec &Rc = myrecs[half ? slot2 : slot];
        const int info = hm.x, j = hm.y, i = hm.z;
        const double x0 = dx * i, y0 = dy * (j + joff);
        const int nslots = mine ? (info & 255) : 0, npend = mine ? ((info >> 8) & 255) : 0;
        const int nknown = nslots - npend;
        const int npend_max = max(__shfl_sync(0xffffffffu, npend, 0), __shfl_sync(0xffffffffu, npend, 16));
        BODY_T(1);
        // ---- one lane per undecided cell (a raster-earlier target of this layer, fitted or rejected by
        //      this warp): its value from the warp's own cache, else (entry recycled) from global memory
        int nfail = 0;
        for (int q0 = 0; q0 < npend_max; q0 += 16) {
            const int r = q0 + hl;
            bool fail = false;
            if (r < npend) {
                const int want = Rc.ptag[r];              // jj << 15 | ii
                const int ce = body_cache_slot(want >> 15, want & 32767);
                const int tag = ctag[ce];
                double v1 = cval[2 * ce], v2 = cval[2 * ce + 1];
                bool got;
                if ((tag & ~(1 << 30)) == want) {
                    got = !(tag & (1 << 30));
                } else {
                    const size_t cc = (size_t)(want >> 15) * Nx + (want & 32767);
                    if (flags & 4) {
                        got = (__ldcg(st + cc) == (unsigned)fresh);
                        v1 = __ldcg(X1e + cc); v2 = __ldcg(X2e + cc);
                    } else {
                        got = (ld_cta_u8(st + cc) == (unsigned)fresh);
                        v1 = ld_cta_f64(X1e + cc); v2 = ld_cta_f64(X2e + cc);
                    }
#ifdef RMT_EXT_TIMING
                    atomicAdd(&g_body_dbg[12 + 16 * layer], 1ull);
#endif
                }
                // a rejected target contributes exact zeros (weight 0)
                store_products(Rc.prod[Rc.pslot[r]], got ? Rc.pw[r] : 0.0, got ? Rc.px[r] : 0.0,
                               got ? Rc.py[r] : 0.0, got ? v1 : 0.0, got ? v2 : 0.0);
                fail = !got;
            }
            nfail += __popc(__ballot_sync(0xffffffffu, fail) & hmask);
        }
        const int count = nknown + npend - nfail;        // known cells in the window (functions.py:147)
        __syncwarp();
        BODY_T(2);
        // ---- ordered accumulation: lane a of each half owns running sum a; the next chunk of 8 is loaded
        //      while the current one is added (the list is zero-padded to a multiple of 8: exact no-ops)
        double acc = 0.0;
        {
            const double *q = &Rc.prod[0][la];
            const int nch = (nslots + 7) >> 3;            // nslots <= 80: the target itself is never a slot
            const int nch_max = max(__shfl_sync(0xffffffffu, nch, 0), __shfl_sync(0xffffffffu, nch, 16));
            double c0 = q[0 * NACC], c1 = q[1 * NACC], c2 = q[2 * NACC], c3 = q[3 * NACC], c4 = q[4 * NACC],
                   c5 = q[5 * NACC], c6 = q[6 * NACC], c7 = q[7 * NACC];
            for (int ch = 0; ch < nch_max; ++ch) {
                const double *qn = q + (ch + 1 < nch ? ch + 1 : (nch ? nch - 1 : 0)) * 8 * NACC;
                const double n0 = qn[0 * NACC], n1 = qn[1 * NACC], n2 = qn[2 * NACC], n3 = qn[3 * NACC],
                             n4 = qn[4 * NACC], n5 = qn[5 * NACC], n6 = qn[6 * NACC], n7 = qn[7 * NACC];
                if (ch < nch) { acc += c0; acc += c1; acc += c2; acc += c3; acc += c4; acc += c5; acc += c6; acc += c7; }
                c0 = n0; c1 = n1; c2 = n2; c3 = n3; c4 = n4; c5 = n5; c6 = n6; c7 = n7;
            }
        }
        if (hl < NACC) ((double *)C.sums[half])[hl] = acc;
        __syncwarp();
        BODY_T(3);
        const volatile double *sm = C.sums[half];
        const int h = (lane & 1) * 3;                     // even lanes solve for xi1, odd lanes for xi2
        const double b0 = sm[h], b1 = sm[h + 1], b2 = sm[h + 2];
        const double A00 = sm[6], A01 = sm[7], A02 = sm[8], A11 = sm[9], A12 = sm[10], A22 = sm[11];
        const double A10 = A01, A20 = A02, A21 = A12;
        const double m00 = A11 * A22 - A12 * A21, m01 = A10 * A22 - A12 * A20, m02 = A10 * A21 - A11 * A20;
        const double det = (A00 * m00 - A01 * m01 + A02 * m02);
        const bool fitted = (count >= 3) && (fabs(det) > 1e-10);
        // Cramer's rule (fast_solve_3x3; its own |det| >= 1e-15 gate holds whenever fitted)
        const double inv_det = 1.0 / det;
        const double cx = (b0 * m00 - A01 * (b1 * A22 - A12 * b2) + A02 * (b1 * A21 - A11 * b2)) * inv_det;
        const double cy = (A00 * (b1 * A22 - A12 * b2) - b0 * m01 + A02 * (A10 * b2 - b1 * A20)) * inv_det;
        const double cz = (A00 * (A11 * b2 - b1 * A21) - A01 * (A10 * b2 - b1 * A20) + b0 * m02) * inv_det;
        const double val = cx + cy * x0 + cz * y0;
        BODY_T(4);
        // ---- publish: the warp's cache first (the next fit reads it), then global memory.  Every lane
        //      of a half stores (even lanes hold xi1, odd lanes xi2; same address, same value)
        const int ce = body_cache_slot(j, i);
        // two targets that map to the same cache entry: only the second keeps it (no torn entry)
        const bool clash = paired && __shfl_sync(0xffffffffu, ce, 0) == __shfl_sync(0xffffffffu, ce, 16);
        if (mine && !(clash && half == 0)) {
            cval[2 * ce + (lane & 1)] = val;
            ctag[ce] = ((j << 15) | i) | (fitted ? 0 : (1 << 30));
        }
        if (mine && fitted) {
            const size_t c = (size_t)j * Nx + i;
            ((lane & 1) ? X2e : X1e)[c] = val;
            st[c] = (unsigned char)fresh;                 // readers in other warps order through row_done
        }
        const int adv = paired ? 2 : 1;
        t += adv;
        slot = paired ? ((slot2 + 1 == nslot) ? 0 : slot2 + 1) : slot2;
        __syncwarp();
        C.consumed = t;                                   // frees the record slots and the queue entries
        BODY_T(5);
        BODY_CNT(6);
        if (paired) BODY_CNT(7);
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) C.row_done = INT_MAX;
}

// grid = tiles x cluster size.  With a cluster of two CTAs the second one is a helper on the neighbouring SM:
// all of its warps are prepare warps that write their records straight into the first CTA's shared memory
// (the chain consumes records faster than the 4 prepare warps per layer of one CTA produce them).
__global__ void __launch_bounds__(768, 1)
k_ext_body(double *__restrict__ X1e, double *__restrict__ X2e, unsigned char *__restrict__ st,
           const int *__restrict__ cnt0 /* + cmin0, cmax0, chunk0, [Ny*nxt] each */,
           const int *__restrict__ tilecnt /* layer-0 targets per tile, all candidate sizes */,
           const int *__restrict__ mode, int L, int P, int W, int nslot, int Ny, int Nx, int joff, int nxt, int XT,
           double dx, double dy, double r2, int flags /* 1: interleave the queue, 2: fit two targets per pass */)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int tile = (int)blockIdx.x / csize;
    // every exit below is taken by all CTAs of a cluster alike (no cluster barrier is left waiting)
    if (mode[0] != 1) return;                           // k_ext_decide picked another variant
    const int k = mode[1];
    const int Tr = 512 << k, nsegt = 1 << k;
    const int ntx = (nxt + nsegt - 1) >> k, nty = (Ny + Tr - 1) / Tr;
    if (tile >= ntx * nty) return;
    if (tilecnt[body_tile_offset(k, Ny, nxt) + tile] == 0) return;          // no band in this tile
    const int ty = tile / ntx, tx = tile - ty * ntx;
    const int seg0 = tx * nsegt, seg1 = min(seg0 + nsegt, nxt);
    const int R0 = max(ty * Tr, 1), R1 = min((ty + 1) * Tr, Ny - 1);       // target rows [R0, R1)
    const int nseg = Ny * nxt;
    const int *cmin0 = cnt0 + nseg, *cmax0 = cnt0 + 2 * nseg, *chunk0 = cnt0 + 3 * nseg;

    extern __shared__ unsigned char s_raw[];
    // the tile's own CTA addresses its shared memory directly; the helper goes through the cluster window
    BodyCtl *ctl = reinterpret_cast<BodyCtl *>(s_raw);
    ExtRec *recs = reinterpret_cast<ExtRec *>(s_raw + (((size_t)L * sizeof(BodyCtl) + 15) & ~(size_t)15));
    const int wib = threadIdx.x >> 5;
    if (crank == 0) {
        for (int e = threadIdx.x; e < L * (int)(sizeof(BodyCtl) / 4); e += blockDim.x) ((int *)ctl)[e] = 0;
        __syncthreads();
        if (threadIdx.x < L) { ctl[threadIdx.x].disc_row = R0 - 1; ctl[threadIdx.x].row_done = R0 - 1; ctl[threadIdx.x].disc_turn = R0; }
        for (int e = threadIdx.x; e < L * BC_ROWS * BC_COLS; e += blockDim.x)
            (&ctl[e / (BC_ROWS * BC_COLS)].ctag[0][0])[e % (BC_ROWS * BC_COLS)] = -1;
        __syncthreads();
    }
    if (csize > 1) cluster.sync();                      // the helper CTA sees initialised control blocks
    int layer, role;
    if (crank == 0) {
        const int per = 1 + W + P;                      // warps per layer: chain, W x discovery, P x prepare
        layer = wib / per;
        role = wib - layer * per;
    } else {
        layer = wib % L;                                // helper CTA: prepare warps only
        role = W + 1;
    }
    if (layer < L) {
        if (crank == 0) {
            body_work<false>(X1e, X2e, st, cnt0, cmin0, cmax0, chunk0, ctl, recs, layer, role, csize > 1, W, nslot, Ny, Nx,
                             joff, nxt, XT, seg0, seg1, R0, R1, dx, dy, r2, flags);
        } else {
            unsigned char *base = (unsigned char *)cluster.map_shared_rank((void *)s_raw, 0);
            BodyCtl *rctl = reinterpret_cast<BodyCtl *>(base);
            ExtRec *rrecs = reinterpret_cast<ExtRec *>(base + (((size_t)L * sizeof(BodyCtl) + 15) & ~(size_t)15));
            body_work<true>(X1e, X2e, st, cnt0, cmin0, cmax0, chunk0, rctl, rrecs, layer, role, true, W, nslot, Ny, Nx,
                            joff, nxt, XT, seg0, seg1, R0, R1, dx, dy, r2, flags);
        }
    }
    if (csize > 1) cluster.sync();                      // the first CTA's shared memory outlives the helper's accesses
}

// Which variant sweeps this call?  The all-layers kernel keeps one CTA (SM) per (layer, macro-tile) for
// the whole length of a chain, so it wins while there are fewer such tasks than SMs and loses when bodies
// outnumber the SMs (then the per-layer launches pack the SMs better).  Decided on the device from the
// layer-0 counts -- no host round trip; the kernels of the variant not chosen return at once.
// Cost model (microseconds, measured at 4097^2 on B200): a chain advances one target row in ~3.4 us within
// one layer; the all-layers kernel runs the L layers of a tile side by side and needs ~7.1 us per row of its
// longest macro-tile, the per-layer launches L * (0.36 ms + 3.4 us per row of the longest 512-row band).
// longest chain of target rows through vertically adjacent tiles of x-tile xt (a body crossing a tile
// boundary chains the two tiles); also counts the non-empty tiles
__device__ inline int ext_longest_chain(const int *__restrict__ cnt0, const int *__restrict__ rows, int ntile_rows,
                                        int height, int nxt, int xt, int Ny, int *nonempty, int *links)
{
    int chain = 0, longest = 0;
    for (int m = 0; m < ntile_rows; ++m) {
        const int r = rows[m * nxt + xt], jb = m * height;               // jb: last row of the tile above
        const bool linked = m > 0 && jb + 1 < Ny && cnt0[jb * nxt + xt] && cnt0[(jb + 1) * nxt + xt];
        chain = r + (linked ? chain : 0);
        *links += linked;
        *nonempty += r != 0;
        longest = max(longest, chain);
    }
    return longest;
}

__global__ void k_ext_decide(const int *__restrict__ cnt0, const int *__restrict__ rows_macro, int nmrb,
                             int macro, const int *__restrict__ rows_band, int nbands, int band, int nxt,
                             int Ny, int L, int limit16, int limit8, int force_variant, int force_rows,
                             int pre_warps, const int *__restrict__ tilecnt, const int *__restrict__ bviol,
                             int body_ok, int fused_ok, int *__restrict__ mode)
{
    __shared__ int body_max[BODY_KMAX];
    if (threadIdx.x < BODY_KMAX) body_max[threadIdx.x] = 0;
    __shared__ int busy, busy_band, longest_macro, longest_band, links_macro, links_band;
    if (threadIdx.x == 0) busy = busy_band = longest_macro = longest_band = links_macro = links_band = 0;
    __syncthreads();
    for (int xt = threadIdx.x; xt < nxt; xt += blockDim.x) {
        int nb = 0, nbb = 0, km = 0, kb = 0;
        const int lm = ext_longest_chain(cnt0, rows_macro, nmrb, macro, nxt, xt, Ny, &nb, &km);
        const int lb = ext_longest_chain(cnt0, rows_band, nbands, band, nxt, xt, Ny, &nbb, &kb);
        atomicAdd(&busy, nb);
        atomicAdd(&busy_band, nbb);
        atomicAdd(&links_macro, km);
        atomicAdd(&links_band, kb);
        atomicMax(&longest_macro, lm);
        atomicMax(&longest_band, lb);
    }
    {   // fullest tile of every candidate body-tile size
        int off = 0;
        for (int k = 0; k < BODY_KMAX; ++k) {
            const int Tr = 512 << k, ns = 1 << k;
            const int nt = ((nxt + ns - 1) >> k) * ((Ny + Tr - 1) / Tr);
            int m = 0;
            for (int e = threadIdx.x; e < nt; e += blockDim.x) m = max(m, tilecnt[off + e]);
            if (m) atomicMax(&body_max[k], m);
            off += nt;
            if (nt == 1) break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // all-layers kernel (16 or 8 rows per block; macro-tiles of `macro` or of `band` rows: shorter
        // tasks let bodies stacked in one macro-tile run side by side) or the per-layer launches.
        // us per chained row: 6.6 with 16-row blocks, 7.4 with 8-row blocks (twice the CTAs per SM).  A body
        // cut by a task boundary costs the 8-row variant far more than its row count says (measured: 36
        // discs on a 6 x 6 lattice, 6.3 ms against 4.9 ms per-layer), so it only runs when nothing is cut.
        const float kRow16 = 6.6f, kRow8 = 7.4f, big = 3.0e38f;
        const int cut[2] = {links_macro, links_band};
        const float cost_layers = L * (360.f + 3.4f * longest_band);
        const int nb[2] = {busy, busy_band}, ln[2] = {longest_macro, longest_band}, ht[2] = {macro, band};
        float cost_f[2] = {big, big};                       // [0]: 16-row blocks, [1]: 8-row blocks
        int rows_f[2] = {macro, macro};
        for (int g = 0; g < 2 && fused_ok; ++g) {
            const float c16 = (nb[g] * L <= limit16) ? kRow16 * ln[g] : big;
            const float c8 = (nb[g] * L <= limit8 && cut[g] == 0) ? kRow8 * ln[g] : big;
            if (c16 < cost_f[0]) { cost_f[0] = c16; rows_f[0] = ht[g]; }
            if (c8 < cost_f[1]) { cost_f[1] = c8; rows_f[1] = ht[g]; }
        }
        // body variant: the smallest tile size whose tiles are all isolated; one warp fits a tile's layer
        // serially (kBody us per target), the layers run side by side a few rows apart
        int body_k = -1;
        for (int k = 0; k < BODY_KMAX && body_ok; ++k) {
            const int Tr = 512 << k, ns = 1 << k;
            const int nt = ((nxt + ns - 1) >> k) * ((Ny + Tr - 1) / Tr);
            if (!bviol[k]) { body_k = k; break; }
            if (nt == 1) break;
        }
        const float kBody = 0.55f;
        const float cost_b = body_k >= 0 ? 20.f + kBody * body_max[body_k] * (1.f + 0.05f * (L - 1)) : big;
        int variant = 0, rows = macro;
        float best = cost_layers;
        if (cost_f[0] < best) { best = cost_f[0]; variant = 16; rows = rows_f[0]; }
        if (cost_f[1] < best) { best = cost_f[1]; variant = 8; rows = rows_f[1]; }
        if (cost_b < best) { best = cost_b; variant = 1; rows = body_k; }
        // rmt_extrapolate_set_mode / RMT_EXT_FORCE: a forced variant that cannot run here (bands too close
        // to the tile borders for the body variant, workspace too small for the all-layers kernels) falls
        // back to the model's choice
        if (force_variant == 0) { variant = 0; rows = macro; }
        if (force_variant == 1 && body_k >= 0) { variant = 1; rows = body_k; }
        if ((force_variant == 16 || force_variant == 8) && fused_ok) {
            variant = force_variant;
            rows = force_rows >= 16 ? force_rows : rows_f[force_variant == 8];
        }
        mode[0] = variant;
        mode[1] = rows;
        mode[2] = pre_warps;
    }
}

// layer-0 target counts per (row, x-tile), without touching the state bytes.
// SEED: the state bytes do not exist yet -- "known" is phi < 0 (functions.py:72), read for the three rows as
// coalesced doubles and turned into neighbour masks with ballots; the task also WRITES the state bytes of its
// own row segment, so the seed pass is not needed (in-place extrapolation: nothing to copy either).
template <bool SEED>
__global__ void k_ext_count0(const unsigned char *__restrict__ st, const double *__restrict__ phi,
                             unsigned char *__restrict__ st_out, int *__restrict__ cnt0,
                             int *__restrict__ rows_macro /* [macro-row][x-tile], zeroed */, int macro,
                             int *__restrict__ rows_band /* [row band][x-tile], zeroed */, int band,
                             int *__restrict__ cmin0, int *__restrict__ cmax0 /* first / last target column */,
                             int *__restrict__ chunk0 /* bit k: targets in columns [32k, 32k+32) of the tile */,
                             int *__restrict__ tilecnt /* zeroed: layer-0 targets per body tile, sizes k = 0.. */,
                             int *__restrict__ bviol /* zeroed [BODY_KMAX]: a band comes within `margin` of a tile border */,
                             int margin, int Ny, int Nx, int nxt, int XT)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarp = (gridDim.x * blockDim.x) >> 5;
    // one warp per (row, x-tile): the rows are short dependent chains of byte loads, so the more warps
    // in flight the better
    for (int task = warp; task < Ny * nxt; task += nwarp) {
        const int j = task / nxt;
        const bool inner = (j >= 1 && j < Ny - 1);
        const size_t o1 = (size_t)j * Nx, o0 = inner ? o1 - Nx : o1, o2 = inner ? o1 + Nx : o1;
        const unsigned char *r1 = st + o1, *r0 = st + o0, *r2 = st + o2;
        {
            const int xt = task - j * nxt;
            int cnt = 0, cmin = INT_MAX, cmax = -1, chunks = 0;
            const int cend = min((xt + 1) * XT, Nx);
            // four 32-column chunks per round, every byte requested before any is looked at: the scan is a chain
            // of L2 round trips otherwise (96 us at 4097^2 when the eight neighbours waited for the cell itself)
            for (int base0 = xt * XT; base0 < cend; base0 += 128) {
                unsigned me[4], nb[4];
                unsigned tm[4];                      // SEED: target mask of each chunk
                if (SEED) {
                    double p0[4], p1[4], p2[4], q0[4], q1[4], q2[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int base = base0 + 32 * u, i = base + lane;
                        p0[u] = p1[u] = p2[u] = q0[u] = q1[u] = q2[u] = 1.0;     // "not known"
                        if (i < Nx && base < cend) {
                            p0[u] = __ldg(phi + o0 + i);
                            p1[u] = __ldg(phi + o1 + i);
                            p2[u] = __ldg(phi + o2 + i);
                        }
                        // the columns just outside the chunk: lane 0 fetches base - 1, lane 31 base + 32
                        const int e = (lane == 0) ? base - 1 : ((lane == 31) ? base + 32 : -1);
                        if (e >= 0 && e < Nx && base < cend) {
                            q0[u] = __ldg(phi + o0 + e);
                            q1[u] = __ldg(phi + o1 + e);
                            q2[u] = __ldg(phi + o2 + e);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int base = base0 + 32 * u, i = base + lane;
                        const unsigned k0 = __ballot_sync(0xffffffffu, p0[u] < 0.0);
                        const unsigned k1 = __ballot_sync(0xffffffffu, p1[u] < 0.0);
                        const unsigned k2 = __ballot_sync(0xffffffffu, p2[u] < 0.0);
                        const unsigned e0 = __ballot_sync(0xffffffffu, q0[u] < 0.0);   // bit 0: left column, bit 31: right
                        const unsigned e1 = __ballot_sync(0xffffffffu, q1[u] < 0.0);
                        const unsigned e2 = __ballot_sync(0xffffffffu, q2[u] < 0.0);
                        if (i < cend) st_out[o1 + i] = ((k1 >> lane) & 1u) ? ST_KNOWN : ST_UNKNOWN;
                        // known neighbours: left / right shifts of the three rows (edge columns spliced in) + above, below
                        const unsigned l0 = (k0 << 1) | (e0 & 1u), rr0 = (k0 >> 1) | (e0 & 0x80000000u);
                        const unsigned l1 = (k1 << 1) | (e1 & 1u), rr1 = (k1 >> 1) | (e1 & 0x80000000u);
                        const unsigned l2 = (k2 << 1) | (e2 & 1u), rr2 = (k2 >> 1) | (e2 & 0x80000000u);
                        const unsigned nbm = l0 | k0 | rr0 | l1 | rr1 | l2 | k2 | rr2;
                        const bool valid = inner && i >= 1 && i < Nx - 1 && i < cend;
                        tm[u] = ~k1 & nbm & __ballot_sync(0xffffffffu, valid);
                    }
                } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base0 + 32 * u + lane;
                    me[u] = 1u;
                    nb[u] = 0u;
                    if (inner && i >= 1 && i < Nx - 1 && i < cend) {
                        me[u] = r1[i];
                        nb[u] = r0[i - 1] | r0[i] | r0[i + 1] | r1[i - 1] | r1[i + 1] | r2[i - 1] | r2[i] | r2[i + 1];
                    }
                }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int base = base0 + 32 * u;
                    const bool tgt = !(me[u] & 1) && (nb[u] & 1);
                    const unsigned m = SEED ? tm[u] : __ballot_sync(0xffffffffu, tgt);
                    cnt += __popc(m);
                    if (m) {
                        cmin = min(cmin, base + __ffs(m) - 1);
                        cmax = base + 31 - __clz(m);
                        chunks |= 1 << ((base - xt * XT) >> 5);
                    }
                }
            }
            if (lane == 0) {
                cnt0[j * nxt + xt] = cnt;
                cmin0[j * nxt + xt] = cmin;
                cmax0[j * nxt + xt] = cmax;
                chunk0[j * nxt + xt] = chunks;
                if (cnt) {       // rows with targets per tile: the length of the dependency chain through it
                    atomicAdd(&rows_macro[((j - 1) / macro) * nxt + xt], 1);
                    atomicAdd(&rows_band[((j - 1) / band) * nxt + xt], 1);
                    // body variant: tiles of (512 << k) rows x (XT << k) columns; a border only counts where
                    // targets can exist on its far side (rows < Ny-1, columns < Nx-1)
                    int off = 0;
                    for (int k = 0; k < BODY_KMAX; ++k) {
                        const int Tr = 512 << k, ns = 1 << k;
                        const int ntx = (nxt + ns - 1) >> k, nty = (Ny + Tr - 1) / Tr;
                        const int ty = j / Tr, tx = xt >> k, rj = j - ty * Tr;
                        atomicAdd(&tilecnt[off + ty * ntx + tx], cnt);
                        off += ntx * nty;
                        bool v = (ty > 0 && rj < margin) || ((ty + 1) * Tr < Ny - 1 && rj >= Tr - margin);
                        if ((xt & (ns - 1)) == 0 && xt > 0 && cmin < xt * XT + margin) v = true;
                        if (((xt + 1) & (ns - 1)) == 0 && (xt + 1) * XT < Nx - 1 && cmax >= (xt + 1) * XT - margin) v = true;
                        if (v) atomicOr(&bviol[k], 1);
                        if (ntx * nty == 1) break;        // larger tiles are the same single tile
                    }
                }
            }
        }
    }
}

inline int flat_blocks(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

}  // namespace

extern "C" {

static inline int ext_nxt(int Nx) { int XT = ext_tune().XT; return (Nx + XT - 1) / XT; }

// Variant selection override (rmt_extrapolate_set_mode; the RMT_EXT_* environment variables only seed it):
//   variant -1 = chosen on the device by k_ext_decide, 0 = per-layer launches, 16 / 8 = all-layers kernel
//   rows    macro-tile height of the all-layers kernel (0 = chosen on the device)
//   cap     prepared-record capacity of the per-layer sweep (0 = default; a small value exercises inline phase A)
struct ExtForce { int init, variant, rows, pre_warps; long cap; };
static ExtForce g_force = {0, -1, 0, 0, 0};
static ExtForce &ext_force()
{
    if (!g_force.init) {
        const char *e;
        g_force.init = 1;
        if ((e = getenv("RMT_EXT_FUSED")) && atoi(e) == 0) g_force.variant = 0;
        if ((e = getenv("RMT_EXT_FORCE"))) { g_force.variant = atoi(e) / 10000; g_force.rows = atoi(e) % 10000; }
        if ((e = getenv("RMT_EXT_PREWARPS"))) g_force.pre_warps = atoi(e);
        if ((e = getenv("RMT_EXT_CAP"))) g_force.cap = atol(e);
    }
    return g_force;
}
// occupancy-derived grid sizes and the shared-memory opt-ins are per DEVICE (CUDA keeps function
// attributes per device): one slot per device ordinal
struct ExtDev { int init, resident16, resident8, sweep_blocks; };
static ExtDev g_ext_dev[64];

// prepared-record capacity (targets per layer); beyond it the sweep computes phase A inline
static inline long ext_cap_default(long ncell)
{
    long c = ncell / 48;
    if (c < 4096) c = 4096;
    if (c > 300000) c = 300000;
    return c;
}
static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

struct ExtLayout {
    size_t st, segs, tcol, trow, tinfo, recs, total;
    long cap;
};
static ExtLayout ext_layout(int Ny, int Nx)
{
    ExtLayout L;
    size_t ncell = (size_t)Ny * (size_t)Nx;
    size_t nseg = (size_t)Ny * ext_nxt(Nx) + 2;
    L.cap = ext_cap_default((long)ncell);
    {   // the all-layers kernel keeps CAPW records per row warp of every resident CTA
        long nmrb = ((Ny - 2 + RB - 1) / RB + 31) / 32, tasks = nmrb * 8 * ext_nxt(Nx);
        long fused = 2 * (tasks < 148 ? tasks : 148) * 16 * CAPW;   // two buffers; = 296 CTAs x 8 warps of the 8-row variant
        if (L.cap < fused) L.cap = fused;
    }
    size_t off = 0;
    L.st = off;    off += al256(ncell);
    L.segs = off;  off += al256((3 * nseg + 16) * sizeof(int));
    L.tcol = off;  off += al256(ncell * sizeof(int));
    L.trow = off;  off += al256(ncell * sizeof(int));
    L.tinfo = off; off += al256((size_t)L.cap * sizeof(int));
    L.recs = off;  off += al256((size_t)L.cap * sizeof(ExtRec));
    L.total = off;
    return L;
}

long rmt_extrapolate_workspace_bytes(int Ny, int Nx) { return (long)ext_layout(Ny, Nx).total; }

int rmt_extrapolate(const double *X1, const double *X2, const double *phi, double *X1e, double *X2e,
                    int Ny, int Nx, double dx, double dy, int max_layers, void *workspace,
                    void *stream)
{
    return rmt_extrapolate_rows(X1, X2, phi, X1e, X2e, Ny, Nx, 0, dx, dy, max_layers, workspace, stream);
}

// where the device-side decision lives inside the workspace (mode[0] variant, mode[1] rows / tile size, mode[2])
static int *ext_mode_ptr(void *workspace, int Ny, int Nx)
{
    const ExtLayout L = ext_layout(Ny, Nx);
    const int nseg = Ny * ext_nxt(Nx);
    int *seg_cnt = (int *)((char *)workspace + L.segs);
    return seg_cnt + 3 * (nseg + 2) + 2;
}

int rmt_extrapolate_set_mode(int variant, int rows, long cap)
{
    if (!(variant == -1 || variant == 0 || variant == 1 || variant == 8 || variant == 16) || rows < 0 || cap < 0)
        return RMT_EINVAL;
    ExtForce &f = ext_force();
    f.variant = variant;
    f.rows = rows;
    f.cap = cap;
    return RMT_OK;
}

int rmt_extrapolate_last_mode(const void *workspace, int Ny, int Nx, int *out3, void *stream)
{
    if (!workspace || !out3 || Ny < 3 || Nx < 3) return RMT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    RMT_CUDA(cudaMemcpyAsync(out3, ext_mode_ptr((void *)workspace, Ny, Nx), 3 * sizeof(int), cudaMemcpyDeviceToHost, s));
    RMT_CUDA(cudaStreamSynchronize(s));
    return RMT_OK;
}

int rmt_extrapolate_rows(const double *X1, const double *X2, const double *phi, double *X1e, double *X2e,
                         int Ny, int Nx, int row_offset, double dx, double dy, int max_layers,
                         void *workspace, void *stream)
{
    if (!X1 || !X2 || !phi || !X1e || !X2e || !workspace || Ny < 3 || Nx < 3) return RMT_EINVAL;
    int joff = row_offset;
    cudaStream_t s = (cudaStream_t)stream;
    size_t ncell = (size_t)Ny * (size_t)Nx;
    int nxt = ext_nxt(Nx);
    int nseg = Ny * nxt;
    ExtTune tune = ext_tune();
    int XT = tune.XT, MRB = tune.MRB, sleep_ns = tune.sleep_ns;
    const ExtLayout L = ext_layout(Ny, Nx);
    const ExtForce force = ext_force();
    char *ws = (char *)workspace;
    unsigned char *st = (unsigned char *)(ws + L.st);
    int *seg_cnt = (int *)(ws + L.segs);
    int *seg_off = seg_cnt + (nseg + 2);
    int *prog = seg_off + (nseg + 2);
    int *tile_counter = prog + (nseg + 2);
    int *mode = tile_counter + 2;       // device: mode[0] = 0 per-layer launches, 16 / 8 all-layers kernel with that
                                        // many rows per block, 1 body variant; mode[1] = macro-tile rows / body tile size
    int *tcol = (int *)(ws + L.tcol), *trow = (int *)(ws + L.trow), *tinfo = (int *)(ws + L.tinfo);
    ExtRec *recs = (ExtRec *)(ws + L.recs);
    int cap = (int)L.cap;
    if (force.cap > 0 && force.cap < cap) cap = (int)force.cap;     // test hook: inline phase A beyond `cap` targets

    // per-device launch geometry (occupancy, shared-memory opt-ins)
    int dev = 0;
    RMT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return RMT_EINVAL;
    ExtDev &D = g_ext_dev[dev];
    if (!D.init) {
        int sms = 0, per_sm = 0;
        RMT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        RMT_CUDA(cudaFuncSetAttribute(k_ext_fused<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(FusedSmemT<16>)));
        RMT_CUDA(cudaFuncSetAttribute(k_ext_fused<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(FusedSmemT<8>)));
        RMT_CUDA(cudaFuncSetAttribute(k_ext_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(SweepSmem)));
        RMT_CUDA(cudaFuncSetAttribute(k_ext_body, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        RMT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ext_fused<16>, 16 * 32,
                                                               sizeof(FusedSmemT<16>)));
        if (per_sm < 1) return RMT_EINVAL;
        D.resident16 = sms * per_sm;
        RMT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ext_fused<8>, 8 * 32,
                                                               sizeof(FusedSmemT<8>)));
        if (per_sm < 1) return RMT_EINVAL;
        D.resident8 = sms * per_sm;
        RMT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ext_sweep, RB * 32, sizeof(SweepSmem)));
        if (per_sm < 1) return RMT_EINVAL;
        D.sweep_blocks = sms * per_sm;
        D.init = 1;
    }

    // stencil_radius_sq = (4*sqrt(dx**2+dy**2))**2, functions.py:76 (no contraction)
    volatile double dx2 = dx * dx, dy2 = dy * dy;
    volatile double ssum = dx2 + dy2;
    volatile double rr = 4.0 * sqrt(ssum);
    double r2 = rr * rr;

    // The seed pass (copies + state bytes) is skipped when there is nothing to copy (in place) AND the layer-0
    // count below runs: that kernel then derives "known" from phi itself and writes the state bytes.
    const bool in_place = (X1e == X1) && (X2e == X2);
    bool seeded = false;
    if (!in_place) {
        k_ext_seed<<<flat_blocks((long)ncell), 256, 0, s>>>(X1, X2, phi, X1e, X2e, st, (long)ncell);
        RMT_LAUNCH_CHECK();
        seeded = true;
    }
    RMT_CUDA(cudaMemsetAsync(tile_counter, 0, 6 * sizeof(int), s));      // tile counter, mode[0..3] = per-layer

    // ---- all layers in one launch when that is faster; which variant is decided on the device ----
    const int Lyr = max_layers;
    if (force.variant != 0 && Lyr >= 1 && Lyr <= 8) {
        // x-tiles wide enough that a task never waits on a task more than ~3/4 of the resident CTAs ahead
        int max_tiles = (D.resident16 * 3 / 4) / (2 * Lyr - 1);
        if (max_tiles < 1) max_tiles = 1;
        int XTf = (Nx > 1024) ? 512 : XT;             // both flanks of a body in one tile: fewer tasks
        const int need_xt = ((Nx + max_tiles - 1) / max_tiles + 31) / 32 * 32;
        if (need_xt > XTf) XTf = need_xt;
        const int nxtf = (Nx + XTf - 1) / XTf, nsegf = Ny * nxtf;
        const long prog_ints = (long)Lyr * nsegf;
        const int macro = 1024, band = 512;            // candidate macro-tile heights (rows)
        const int nmrb = rmt_cdiv(Ny - 2, macro);
        const int nmac = nmrb * nxtf, nbnd = rmt_cdiv(Ny - 2, band) * nxtf, nbusy = nmac + nbnd;
        long ntile_total = 0;                          // body tiles over all candidate sizes
        for (int k = 0; k < BODY_KMAX; ++k) {
            const long nt = (long)((nxtf + (1 << k) - 1) >> k) * rmt_cdiv(Ny, 512 << k);
            ntile_total += nt;
            if (nt == 1) break;
        }
        const int ntile0 = nxtf * rmt_cdiv(Ny, 512);
        if (XTf <= 1024 && prog_ints + 4L * nsegf + nbusy + ntile_total + BODY_KMAX + 8 <= (long)ncell) {
            int *progF = trow;                         // [L][nsegf] ints, then chain lengths, body tile counts and
            int *busy = trow + prog_ints;              //  flags, then cnt0 (trow holds ncell ints; the per-layer path
            int *tilecnt = busy + nbusy;               //  rewrites it afterwards if it is the one that runs)
            int *bviol = tilecnt + ntile_total;
            int *cnt0 = bviol + BODY_KMAX;
            const long ntasks = (long)rmt_cdiv(Ny - 2, band) * Lyr * nxtf;   // with the smaller macro-tile
            int blocks16 = D.resident16, blocks8 = D.resident8;
            if ((long)blocks16 > ntasks) blocks16 = (int)ntasks;
            if ((long)blocks8 > ntasks) blocks8 = (int)ntasks;
            const long need_recs = 2L * (blocks16 * 16 > blocks8 * 8 ? blocks16 * 16 : blocks8 * 8) * CAPW;
            const int fused_ok = (XTf <= LMAX && need_recs <= L.cap) ? 1 : 0;
            // body variant: warps per layer = chain + discovery + P prepare warps in a CTA of <= 640 threads
            int P = 0, Wd = 1, nslot = 0;
            size_t body_smem = 0;
            if (Lyr <= BODY_LMAX) {
                int per = 24 / Lyr;                // warps per layer in a CTA of 768 threads
                if (per > 8) per = 8;
                Wd = per >= 6 ? 2 : 1;             // discovery warps (rows in turn)
                P = per - 1 - Wd;                  // prepare warps
                const size_t ctl_bytes = ((size_t)Lyr * sizeof(BodyCtl) + 15) & ~(size_t)15;
                nslot = (int)((220 * 1024 - ctl_bytes) / ((size_t)Lyr * sizeof(ExtRec)));
                if (nslot > BSLOT_MAX) nslot = BSLOT_MAX;
                body_smem = ctl_bytes + (size_t)Lyr * nslot * sizeof(ExtRec);
            }
            const int body_ok = (P >= 1 && nslot >= 2) ? 1 : 0;
            if (fused_ok || body_ok) {
                RMT_CUDA(cudaMemsetAsync(progF, 0, (size_t)(prog_ints + nbusy + ntile_total + BODY_KMAX) * sizeof(int), s));
                int rwb = rmt_cdiv((long)Ny * nxtf * 32, 256);
                if (rwb > 148 * 8) rwb = 148 * 8;
                if (seeded)
                    k_ext_count0<false><<<rwb, 256, 0, s>>>(st, phi, st, cnt0, busy, macro, busy + nmac, band, cnt0 + nsegf,
                                                            cnt0 + 2 * nsegf, cnt0 + 3 * nsegf, tilecnt, bviol, Lyr + 5,
                                                            Ny, Nx, nxtf, XTf);
                else
                    k_ext_count0<true><<<rwb, 256, 0, s>>>(st, phi, st, cnt0, busy, macro, busy + nmac, band, cnt0 + nsegf,
                                                           cnt0 + 2 * nsegf, cnt0 + 3 * nsegf, tilecnt, bviol, Lyr + 5,
                                                           Ny, Nx, nxtf, XTf);
                RMT_LAUNCH_CHECK();
                seeded = true;
                k_ext_decide<<<1, 64, 0, s>>>(cnt0, busy, nmrb, macro, busy + nmac, nbnd / nxtf, band, nxtf, Ny,
                                              Lyr, D.resident16 * 9 / 10, D.resident8 * 9 / 10, force.variant,
                                              force.rows, force.pre_warps, tilecnt, bviol, body_ok, fused_ok, mode);
                RMT_LAUNCH_CHECK();
                if (body_ok) {
                    static int body_flags = -1;      // debugging hook: RMT_BODY_FLAGS (default 3 = everything on)
                    if (body_flags < 0) { const char *e = getenv("RMT_BODY_FLAGS"); body_flags = e ? atoi(e) : 3; }
                    // a cluster of two CTAs per tile when the tiles leave SMs free for the helpers
                    static int body_cluster = -1;    // RMT_BODY_CLUSTER=2: a helper CTA per tile (measured slower: see DESIGN.md)
                    if (body_cluster < 0) { const char *e = getenv("RMT_BODY_CLUSTER"); body_cluster = e ? atoi(e) : 1; }
                    const int csz = body_cluster >= 2 ? 2 : 1;
                    if (csz == 1) {
                        k_ext_body<<<ntile0, 768, body_smem, s>>>(X1e, X2e, st, cnt0, tilecnt, mode, Lyr, P, Wd, nslot, Ny, Nx,
                                                                  joff, nxtf, XTf, dx, dy, r2, body_flags);
                    } else {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(ntile0 * csz);
                    cfg.blockDim = dim3(768);
                    cfg.dynamicSmemBytes = body_smem;
                    cfg.stream = s;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    RMT_CUDA(cudaLaunchKernelEx(&cfg, k_ext_body, X1e, X2e, st, (const int *)cnt0, (const int *)tilecnt,
                                                (const int *)mode, Lyr, P, Wd, nslot, Ny, Nx, joff, nxtf, XTf, dx, dy, r2,
                                                body_flags));
                    }
                    RMT_LAUNCH_CHECK();
                }
                if (fused_ok) {
                    int Lv = Lyr, nxv = nxtf, xtv = XTf;
                    void *args[] = {&X1e, &X2e, &st, &cnt0, &progF, &tile_counter, &recs, &mode, &Lv, &Ny, &Nx,
                                    &joff, &nxv, &xtv, &dx, &dy, &r2};
                    ++g_rmt_launches;
                    RMT_CUDA(cudaLaunchCooperativeKernel((void *)k_ext_fused<16>, dim3(blocks16), dim3(16 * 32),
                                                         args, sizeof(FusedSmemT<16>), s));
                    ++g_rmt_launches;
                    RMT_CUDA(cudaLaunchCooperativeKernel((void *)k_ext_fused<8>, dim3(blocks8), dim3(8 * 32), args,
                                                         sizeof(FusedSmemT<8>), s));
                }
            }
        }
    }

    if (!seeded) {        // in place and no all-layers candidate: the per-layer kernels need the state bytes
        k_ext_seed<<<flat_blocks((long)ncell), 256, 0, s>>>(X1, X2, phi, X1e, X2e, st, (long)ncell);
        RMT_LAUNCH_CHECK();
    }
    const size_t sweep_smem = sizeof(SweepSmem);
    int row_warps_blocks = rmt_cdiv((long)Ny * 32, 256);
    if (row_warps_blocks > 148 * 8) row_warps_blocks = 148 * 8;

    // the per-layer launches: they return at once when the device picked an all-layers variant
    for (int layer = 0; layer < max_layers; ++layer) {
        k_ext_flag<<<row_warps_blocks, 256, 0, s>>>(st, seg_cnt, Ny, Nx, nxt, XT, mode);
        RMT_LAUNCH_CHECK();
        k_ext_scan<<<1, 1024, 0, s>>>(seg_cnt, seg_off, Ny, nxt, tile_counter, mode);
        RMT_LAUNCH_CHECK();
        k_ext_fill<<<row_warps_blocks, 256, 0, s>>>(st, seg_off, tcol, trow, prog, Ny, Nx, nxt, XT, mode);
        RMT_LAUNCH_CHECK();
        k_ext_prepare<<<148 * 12, 256, 0, s>>>(X1e, X2e, st, seg_off, nseg, tcol, trow, recs, tinfo, cap, Ny, Nx,
                                               joff, dx, dy, r2, mode);
        RMT_LAUNCH_CHECK();
        // all CTAs must be co-resident (warps wait on each other): cooperative launch
        int blocks = D.sweep_blocks;
        int need = rmt_cdiv(rmt_cdiv(Ny - 2, RB), MRB) * nxt;
        if (blocks > need) blocks = need;
        void *args[] = {&X1e, &X2e, &st, &seg_off, &tcol, &prog, &tile_counter, &recs, &tinfo, &cap,
                        &Ny, &Nx, &joff, &nxt, &XT, &MRB, &sleep_ns, &dx, &dy, &r2, &mode};
        ++g_rmt_launches;
        RMT_CUDA(cudaLaunchCooperativeKernel((void *)k_ext_sweep, dim3(blocks), dim3(RB * 32), args,
                                             sweep_smem, s));
    }
    return RMT_OK;
}

#ifdef RMT_EXT_TIMING
int rmt_body_watch_read(int *out192, int reset)
{
    RMT_CUDA(cudaMemcpyFromSymbol(out192, g_body_watch, sizeof(int) * 192));
    if (reset) {
        int z[192] = {0};
        RMT_CUDA(cudaMemcpyToSymbol(g_body_watch, z, sizeof(z)));
    }
    return RMT_OK;
}
int rmt_body_debug_read(unsigned long long *out128, int reset)
{
    RMT_CUDA(cudaMemcpyFromSymbol(out128, g_body_dbg, sizeof(unsigned long long) * 128));
    if (reset) {
        unsigned long long z[128] = {0};
        RMT_CUDA(cudaMemcpyToSymbol(g_body_dbg, z, sizeof(z)));
    }
    return RMT_OK;
}
int rmt_ext_debug_read(unsigned long long *out8, int reset)
{
    RMT_CUDA(cudaMemcpyFromSymbol(out8, g_ext_dbg, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0};
        RMT_CUDA(cudaMemcpyToSymbol(g_ext_dbg, z, sizeof(z)));
    }
    return RMT_OK;
}
#endif

// device exp() on an array -- lets the tests prove bit-equality with host libm
int rmt_exp_probe(const double *x, double *y, long n, void *stream)
{
    if (!x || !y || n <= 0) return RMT_EINVAL;
    k_exp_probe<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

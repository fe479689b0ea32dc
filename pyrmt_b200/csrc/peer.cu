// peer.cu -- the slab exchanges over NVLink peer memory (one process per GPU on one node).
//
// The reference is a single-process NumPy/Numba program; its whole-grid stencils
// (pyRMT/utils.py:4-114), its whole-grid transforms (pyRMT/functions.py:1107-1119 dctn/idctn,
// :1216-1233 fft2/ifft2) and its whole-grid means become, once the grid is cut into row slabs,
// a halo exchange, an all-to-all transpose and an all-reduce.  Here they are plain kernels that
// store straight into the neighbour's memory (CUDA IPC mapping, NVLink / NVSwitch underneath):
//
//   * rmt_peer_alloc / export / import      an arena per rank, mapped into every other rank
//   * rmt_peer_put2d                        batched 2-D block copies (halo rows, overlap rows) -- one launch
//   * rmt_transpose_scatter                 the local transpose AND the all-to-all of the distributed
//                                           DCT / Hartley solve in one kernel: every 32x32 tile is turned
//                                           in shared memory and written to the rank that owns its columns
//   * rmt_peer_barrier                      one release-store per involved peer + acquire polls (system scope)
//   * rmt_peer_reduce                       the local half of the all-reduce: every rank has stored its vector
//                                           into slot [rank] of every rank (rmt_peer_put2d); after the barrier
//                                           each rank reduces the slots in rank order (same bits on all ranks)
//
// Ordering contract (pyrmt_b200/slab.py PeerComm): a barrier names the peers that trade data in this exchange;
// each PAIR of ranks counts the barriers it shares, and both sides of a pair issue the same sequence (SPMD).
// A staging region holds two copies used alternately: a copy written for use i is read after that use's
// barrier and written again for use i + 2 at the earliest -- by then the writer has passed the barrier of use
// i + 1, which its reader signalled after (in stream order) its reads of use i.
#include <string.h>

#include "common.cuh"
#include "../../include/rmt_b200.h"

namespace {

constexpr int kMaxPeers = RMT_PEER_MAX;
constexpr int kMaxPut = RMT_PUT_MAX;

struct PutBatch {
    const double *src[kMaxPut];
    double *dst[kMaxPut];
    int rows[kMaxPut], cols[kMaxPut];
    long sld[kMaxPut], dld[kMaxPut];
};

// blockIdx.z = descriptor, blockIdx.y strides over rows, blockIdx.x over 256-column chunks
__global__ void __launch_bounds__(256) k_put2d(const PutBatch B)
{
    const int d = blockIdx.z;
    const int rows = B.rows[d], cols = B.cols[d];
    const double *__restrict__ s = B.src[d];
    double *__restrict__ o = B.dst[d];
    const long sld = B.sld[d], dld = B.dld[d];
    for (int c = blockIdx.x * 256 + threadIdx.x; c < cols; c += gridDim.x * 256)
        for (int r = blockIdx.y; r < rows; r += gridDim.y) o[(size_t)r * dld + c] = s[(size_t)r * sld + c];
}

struct ScatterParts {
    double *dst[kMaxPeers];     // part q: element (c - start[q], r) goes to dst[q][(c - start[q]) * ld[q] + r]
    long ld[kMaxPeers];
    int start[kMaxPeers + 1];
    int n;
};

// in (R, C; row stride ldi) -> transposed, columns [start[q], start[q+1]) to part q
__global__ void __launch_bounds__(256)
k_transpose_scatter(const double *__restrict__ in, int R, int C, long ldi, const ScatterParts P)
{
    __shared__ double t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < R && c < C) t[ty + 8 * k][tx] = __ldg(in + (size_t)r * ldi + c);
    }
    __syncthreads();
    int q0 = 0;
    while (q0 + 1 < P.n && P.start[q0 + 1] <= c0) ++q0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < R && c < C) {
            int q = q0;
            while (q + 1 < P.n && P.start[q + 1] <= c) ++q;
            P.dst[q][(size_t)(c - P.start[q]) * P.ld[q] + r] = t[tx][ty + 8 * k];
        }
    }
}

struct PeerFlags {
    unsigned long long *p[kMaxPeers];   // flag array of every rank (>= world entries), as mapped HERE
    unsigned long long epoch[kMaxPeers];   // per peer: the value to publish / to wait for; 0 = peer not involved
};

__device__ __forceinline__ void st_release_sys(unsigned long long *a, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *a)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
    return v;
}

// Thread q: tell rank q that this rank has reached `epoch` (everything this stream did before is visible
// first: kernel boundary + system fence), then wait for rank q's own arrival.  A watchdog ends the wait
// with *err = 1 instead of hanging the device when a peer never arrives.
__global__ void k_peer_barrier(const PeerFlags F, int rank, int world, long long limit_cycles, int *err)
{
    const int q = threadIdx.x;
    if (q >= world || q == rank || F.epoch[q] == 0) return;
    const unsigned long long epoch = F.epoch[q];
    __threadfence_system();
    st_release_sys(F.p[q] + rank, epoch);
    const unsigned long long *mine = F.p[rank] + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < epoch) {
        // after one time-out every later barrier falls through at once: the run ends quickly and wrong
        // (PeerComm.check() reports it) instead of spinning for the limit again and again
        if (*(volatile int *)err || clock64() - t0 > limit_cycles) {
            *err = 1;
            break;
        }
        __nanosleep(64);
    }
}

// out[k] = op over ranks q = 0 .. world-1 (in that order) of slots[q * stride + k]
__global__ void k_peer_reduce(const double *__restrict__ slots, int world, long stride, int n, int op,
                              double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double acc = slots[k];
    for (int q = 1; q < world; ++q) {
        const double v = slots[(size_t)q * stride + k];
        acc = op == 0 ? acc + v : (op == 1 ? fmax(acc, v) : fmin(acc, v));
    }
    out[k] = acc;
}

}  // namespace

extern "C" {

int rmt_peer_alloc(size_t bytes, void **ptr)
{
    if (!ptr || bytes == 0) return RMT_EINVAL;
    RMT_CUDA(cudaMalloc(ptr, bytes));
    RMT_CUDA(cudaMemset(*ptr, 0, bytes));
    RMT_CUDA(cudaDeviceSynchronize());
    return RMT_OK;
}

int rmt_peer_free(void *ptr)
{
    if (!ptr) return RMT_EINVAL;
    RMT_CUDA(cudaFree(ptr));
    return RMT_OK;
}

int rmt_peer_export(void *ptr, unsigned char *handle64)
{
    if (!ptr || !handle64) return RMT_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    RMT_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return RMT_OK;
}

int rmt_peer_import(const unsigned char *handle64, void **ptr)
{
    if (!ptr || !handle64) return RMT_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RMT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RMT_OK;
}

int rmt_peer_release(void *ptr)
{
    if (!ptr) return RMT_EINVAL;
    RMT_CUDA(cudaIpcCloseMemHandle(ptr));
    return RMT_OK;
}

int rmt_peer_put2d(const rmt_put2d *desc, int n, void *stream)
{
    if (!desc || n < 1 || n > kMaxPut) return RMT_EINVAL;
    PutBatch B;
    int max_rows = 1, max_cols = 1;
    for (int k = 0; k < n; ++k) {
        if (!desc[k].src || !desc[k].dst || desc[k].rows < 1 || desc[k].cols < 1) return RMT_EINVAL;
        B.src[k] = desc[k].src;
        B.dst[k] = desc[k].dst;
        B.rows[k] = desc[k].rows;
        B.cols[k] = desc[k].cols;
        B.sld[k] = desc[k].src_ld;
        B.dld[k] = desc[k].dst_ld;
        if (desc[k].rows > max_rows) max_rows = desc[k].rows;
        if (desc[k].cols > max_cols) max_cols = desc[k].cols;
    }
    dim3 grd(rmt_cdiv(max_cols, 256) > 32 ? 32 : rmt_cdiv(max_cols, 256), max_rows > 2048 ? 2048 : max_rows, n);
    k_put2d<<<grd, 256, 0, (cudaStream_t)stream>>>(B);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_transpose_scatter(const double *in, int R, int C, long ldi, int nparts, const int *start,
                          double *const *dst, const long *dst_ld, void *stream)
{
    if (!in || R < 1 || C < 1 || nparts < 1 || nparts > kMaxPeers || !start || !dst || !dst_ld) return RMT_EINVAL;
    ScatterParts P;
    P.n = nparts;
    for (int q = 0; q < nparts; ++q) {
        if (start[q + 1] < start[q] || (start[q + 1] > start[q] && !dst[q])) return RMT_EINVAL;
        P.dst[q] = dst[q];
        P.ld[q] = dst_ld[q];
        P.start[q] = start[q];
    }
    P.start[nparts] = start[nparts];
    if (start[0] != 0 || start[nparts] != C) return RMT_EINVAL;
    dim3 grd(rmt_cdiv(C, 32), rmt_cdiv(R, 32));
    k_transpose_scatter<<<grd, 256, 0, (cudaStream_t)stream>>>(in, R, C, ldi, P);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_peer_barrier(void *const *flags, int rank, int world, const unsigned long long *epochs, double timeout_s,
                     int *err, void *stream)
{
    if (!flags || !epochs || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !err) return RMT_EINVAL;
    if (world == 1) return RMT_OK;
    PeerFlags F;
    bool any = false;
    for (int q = 0; q < world; ++q) {
        if (!flags[q]) return RMT_EINVAL;
        F.p[q] = (unsigned long long *)flags[q];
        F.epoch[q] = q == rank ? 0 : epochs[q];
        any = any || F.epoch[q] != 0;
    }
    if (!any) return RMT_OK;
    const long long limit = (long long)((timeout_s > 0 ? timeout_s : 5.0) * 1.9e9);
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(F, rank, world, limit, err);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

int rmt_peer_reduce(const double *slots, int world, long stride, int n, int op, double *out, void *stream)
{
    if (!slots || !out || world < 1 || n < 1 || op < 0 || op > 2) return RMT_EINVAL;
    k_peer_reduce<<<rmt_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(slots, world, stride, n, op, out);
    RMT_LAUNCH_CHECK();
    return RMT_OK;
}

}  // extern "C"

// fp64 dependent-issue latency / throughput micro-benchmark (B200 design input)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, double a, double b, int n, long long *cyc)
{
    double x = a + threadIdx.x;
    double y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == 0) x = x + y;
            if (OP == 1) x = x * y;
            if (OP == 2) x = fma(x, y, y);
            if (OP == 3) x = 1.0 / x;
            if (OP == 4) x = __drcp_rn(x);
            if (OP == 5) { float f = (float)x; f = f + 1.0f; x = f; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
__global__ void thr(double *out, double a, double b, int n, long long *cyc)
{
    double x[8];
    for (int k = 0; k < 8; ++k) x[k] = a + threadIdx.x + k;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) x[k] = x[k] + b;
            if (OP == 2) x[k] = fma(x[k], b, b);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "1.0/x", "drcp_rn", "F2F+FADD+F2F"};
    int n = 256;
#define RUN(OP, T) chain<OP><<<1, T>>>(out, 1.0, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-14s dependent chain, %4d thr/SM: %.1f cycles/op\n", names[OP], T, (double)h / (n * 16));
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32)
    RUN(0, 1024) RUN(2, 1024)
    for (int T : {32, 128, 256, 512, 1024}) {
        thr<2><<<1, T>>>(out, 1.0, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA 8 independent chains, %4d thr/SM: %.2f warp-instr/cycle/SM (%.1f DFMA lanes/clk)\n", T,
               (double)n * 8 * (T / 32) / h, (double)n * 8 * T / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// Micro-benchmark: FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) latency / issue rate, and whether it runs
// concurrently with the FP64 FMA pipe (DFMA).  Run on the B200 box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC, int NFMA>
__global__ void k(double *out, int iters, double a, double b, long long *cyc) {
    double c[NACC > 0 ? NACC : 1][2], x[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
#pragma unroll
    for (int i = 0; i < NFMA; ++i) x[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < NFMA; ++i) x[i] = fma(x[i], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NFMA; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC, int NFMA>
void run(int threads) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<NACC, NFMA><<<1, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    k<NACC, NFMA><<<1, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_it = (double)h / (iters * 4.0);
    printf("threads=%4d  DMMA accumulators=%d  DFMA chains=%2d : %.1f cycles per round (%d DMMA + %d DFMA per warp)", threads, NACC, NFMA, per_it, NACC, NFMA);
    if (NACC) printf("  -> %.2f cycles per DMMA per sub-partition", per_it / NACC / (threads / 128.0 > 1 ? threads / 128.0 : 1));
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>(32); run<2, 0>(32); run<4, 0>(32); run<8, 0>(32);
    run<4, 0>(128); run<8, 0>(128); run<8, 0>(256); run<8, 0>(512);
    run<0, 8>(128); run<4, 8>(128); run<4, 16>(128); run<4, 32>(128); run<8, 32>(256);
    return 0;
}

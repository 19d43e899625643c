// Micro-benchmark: FP64 DFMA latency / throughput on one SM sub-partition (run on the B200 box).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dfma_chain(double *out, int iters, double a, double b, long long *cyc) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double *out, int iters, long long *cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rsq_chain(double *out, int iters, long long *cyc) {
    double x = threadIdx.x + 2.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y + 2.0; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double *out, int iters, long long *cyc) {
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double)((i * 7 + 3) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) idx = (int)s[idx];
    }
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int threads, const char *name) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    dfma_chain<CH><<<1, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    dfma_chain<CH><<<1, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0 * CH);
    printf("%s CH=%d threads=%d: %.2f cycles per DFMA warp-instr per warp; SM-wide %.3f cyc/warp-instr\n", name, CH, threads, per, per / (threads / 32));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(32, "dfma"); run<2>(32, "dfma"); run<4>(32, "dfma"); run<8>(32, "dfma"); run<16>(32, "dfma");
    run<1>(128, "dfma"); run<2>(128, "dfma"); run<4>(128, "dfma"); run<8>(128, "dfma"); run<16>(128, "dfma");
    run<1>(256, "dfma"); run<4>(256, "dfma"); run<8>(256, "dfma");
    run<1>(512, "dfma"); run<4>(512, "dfma"); run<8>(512, "dfma");
    run<1>(1024, "dfma"); run<4>(1024, "dfma");
    double *out; long long *cyc, h; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
    shfl_chain<<<1, 32>>>(out, 1000, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("shfl64+dadd chain: %.1f cycles per link\n", h / 8000.0);
    rsq_chain<<<1, 32>>>(out, 1000, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rsqrt.approx.f64+dadd chain: %.1f cycles per link\n", h / 8000.0);
    lds_chain<<<1, 32>>>(out, 1000, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("lds64+cvt chain: %.1f cycles per link\n", h / 8000.0);
    return 0;
}

// Micro-benchmark: straight-line (non-looping) code vs looped code -- instruction-fetch limits on B200.
// W warps per SM sub-partition, each running 8 independent DFMA chains; the loop body holds UNROLL*8 DFMAs.
#include <cstdio>
#include <cuda_runtime.h>
template <int UNROLL, bool SKEW>
__global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-3 + c;
    if (SKEW) {      // de-phase the warps of an SM so that they do not walk the code in lockstep
        int w = (blockIdx.x * 4 + (threadIdx.x >> 5)) % 16;
        for (int i = 0; i < w * 37; ++i) x[0] = fma(x[0], a, b);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int UNROLL, bool SKEW>
void run(int blocks_per_sm) {
    double *out;
    cudaMalloc(&out, 148 * 8 * 128 * 8);
    int total = 1 << 17;                 // DFMA-groups per thread
    int iters = total / UNROLL;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<UNROLL, SKEW><<<148 * blocks_per_sm, 128>>>(out, 16, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<UNROLL, SKEW><<<148 * blocks_per_sm, 128>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma_per_smsp = (double)iters * UNROLL * 8 * blocks_per_sm;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("body=%6d DFMA (%4d KB) skew=%d warps/SMSP=%d: %.3f cycles per DFMA per SMSP (%.2f ms)\n", UNROLL * 8, UNROLL * 8 * 16 / 1024,
           (int)SKEW, blocks_per_sm, cyc / dfma_per_smsp, ms);
    cudaFree(out);
}
int main() {
    run<4, false>(2); run<32, false>(2); run<128, false>(2); run<256, false>(2); run<512, false>(2); run<1024, false>(2); run<4096, false>(2);
    run<4, true>(2); run<32, true>(2); run<128, true>(2); run<256, true>(2); run<512, true>(2); run<1024, true>(2); run<4096, true>(2);
    run<1024, true>(1); run<1024, true>(3); run<1024, true>(4);
    return 0;
}

// Probe: CUDA green contexts on B200 -- can a 16-SM partition host a 16-CTA cluster kernel while a 132-SM partition runs a
// long kernel, launched through the runtime API?  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o greenctx greenctx.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s -> %s (line %d)\n", #x, s_, __LINE__); exit(1); } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(r_), __LINE__); exit(1); } } while (0)

__global__ void spin_kernel(long long cycles, unsigned *smids) {
    const long long t0 = clock64();
    if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = s; }
    while (clock64() - t0 < cycles) {}
}
__global__ void __launch_bounds__(256, 1) cluster_kernel(long long cycles, unsigned *smids) {
    __shared__ double pad[1024];
    const long long t0 = clock64();
    pad[threadIdx.x] = 1.0;
    if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = s; }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    while (clock64() - t0 < cycles) {}
}

int main(int argc, char **argv) {
    const unsigned want = argc > 1 ? atoi(argv[1]) : 16;
    const unsigned flags = argc > 2 ? atoi(argv[2]) : CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE;
    RK(cudaSetDevice(0));
    RK(cudaFree(0));
    CUdevice dev;
    CK(cuDeviceGet(&dev, 0));
    CUdevResource all, grp[4], rem;
    CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", all.sm.smCount);
    unsigned ng = 1;
    CK(cuDevSmResourceSplitByCount(grp, &ng, &all, &rem, flags, want));
    printf("split: groups=%u group0=%u SMs, remaining=%u SMs\n", ng, grp[0].sm.smCount, rem.sm.smCount);
    CUdevResourceDesc d0, d1;
    CK(cuDevResourceGenerateDesc(&d0, &grp[0], 1));
    CK(cuDevResourceGenerateDesc(&d1, &rem, 1));
    CUgreenCtx g0, g1;
    CK(cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CK(cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s0, s1;
    CK(cuGreenCtxStreamCreate(&s0, g0, CU_STREAM_NON_BLOCKING, 0));
    CK(cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0));
    unsigned *sm0, *sm1;
    RK(cudaMalloc(&sm0, 4096 * 4));
    RK(cudaMalloc(&sm1, 4096 * 4));
    RK(cudaMemset(sm0, 0xff, 4096 * 4));
    RK(cudaMemset(sm1, 0xff, 4096 * 4));
    RK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1, e2, e3;
    RK(cudaEventCreate(&e0)); RK(cudaEventCreate(&e1)); RK(cudaEventCreate(&e2)); RK(cudaEventCreate(&e3));
    // long kernel on the big partition: 4 waves of 2 CTAs/SM worth of blocks, 200 us each
    const long long cyc = 400000;
    RK(cudaEventRecord(e0, (cudaStream_t)s1));
    spin_kernel<<<148 * 8, 256, 0, (cudaStream_t)s1>>>(cyc, sm1);
    RK(cudaGetLastError());
    RK(cudaEventRecord(e1, (cudaStream_t)s1));
    // meanwhile: 10 cluster-16 kernels of 40 us on the small partition
    for (int csz = 16; csz >= 8; csz /= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz, 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.stream = (cudaStream_t)s0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = -1;
        cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, cluster_kernel, &cfg);
        printf("cluster size %d on the small partition: maxActiveClusters=%d (%s)\n", csz, nclusters, cudaGetErrorString(oe));
        RK(cudaEventRecord(e2, (cudaStream_t)s0));
        cudaError_t le = cudaSuccess;
        for (int i = 0; i < 10 && le == cudaSuccess; ++i) le = cudaLaunchKernelEx(&cfg, cluster_kernel, (long long)80000, sm0);
        printf("  launch: %s\n", cudaGetErrorString(le));
        (void)cudaGetLastError();
        RK(cudaEventRecord(e3, (cudaStream_t)s0));
        cudaError_t se = cudaStreamSynchronize((cudaStream_t)s0);
        float ms = 0;
        if (se == cudaSuccess) { cudaEventElapsedTime(&ms, e2, e3); printf("  10 cluster kernels: %.3f ms (expected ~0.4 ms if concurrent with the long kernel)\n", ms); }
        else { printf("  sync: %s\n", cudaGetErrorString(se)); break; }
        if (le == cudaSuccess) break;
    }
    RK(cudaStreamSynchronize((cudaStream_t)s1));
    float msl = 0;
    RK(cudaEventElapsedTime(&msl, e0, e1));
    printf("long kernel on the big partition: %.3f ms\n", msl);
    unsigned h0[64], h1[148 * 8];
    RK(cudaMemcpy(h0, sm0, sizeof(h0), cudaMemcpyDeviceToHost));
    RK(cudaMemcpy(h1, sm1, sizeof(h1), cudaMemcpyDeviceToHost));
    bool used[256] = {false}, usedb[256] = {false};
    printf("small-partition smids:");
    for (int i = 0; i < 16; ++i) { printf(" %u", h0[i]); if (h0[i] < 256) used[h0[i]] = true; }
    printf("\n");
    int nb = 0, overlap = 0;
    for (int i = 0; i < 148 * 8; ++i) if (h1[i] < 256 && !usedb[h1[i]]) { usedb[h1[i]] = true; ++nb; if (used[h1[i]]) ++overlap; }
    printf("big partition used %d distinct SMs, %d of them also used by the small partition\n", nb, overlap);
    // events across partitions + a kernel on the primary context stream using all SMs
    cudaStream_t sp;
    RK(cudaStreamCreateWithFlags(&sp, cudaStreamNonBlocking));
    RK(cudaEventRecord(e0, (cudaStream_t)s0));
    RK(cudaStreamWaitEvent(sp, e0, 0));
    spin_kernel<<<148 * 2, 256, 0, sp>>>(1000, sm1);
    RK(cudaGetLastError());
    RK(cudaEventRecord(e1, sp));
    RK(cudaStreamWaitEvent((cudaStream_t)s1, e1, 0));
    spin_kernel<<<8, 256, 0, (cudaStream_t)s1>>>(1000, sm1);
    RK(cudaDeviceSynchronize());
    printf("cross-partition events + primary-stream launch: ok\n");
    return 0;
}

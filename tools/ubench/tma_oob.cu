// Probe: TMA tiled copies whose box starts at NEGATIVE coordinates / is larger than the tensor (zero fill on load, skipped on store).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_oob tma_oob.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, double *out, int c0, int c1, int N, int PITCH, int do_store) {
    extern __shared__ __align__(128) unsigned char sm[];
    double *img = (double *)sm;
    uint64_t *bar = (uint64_t *)(sm + N * PITCH * 8);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(N * PITCH * 8));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(img)),
                     "l"(&tm), "r"(c0), "r"(c1), "r"(0), "r"(s32(bar))
                     : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(s32(bar)));
    for (int i = threadIdx.x; i < N * PITCH; i += blockDim.x) out[i] = img[i];
    __syncthreads();
    if (do_store) {
        for (int i = threadIdx.x; i < N * PITCH; i += blockDim.x) img[i] += 1000.0;
        asm volatile("fence.proxy.async.shared::cta;");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm), "r"(c0), "r"(c1), "r"(0), "r"(s32(img)) : "memory");
            asm volatile("cp.async.bulk.commit_group;");
            asm volatile("cp.async.bulk.wait_group 0;");
        }
    }
}
int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 24, N = argc > 2 ? atoi(argv[2]) : 32, start = argc > 3 ? atoi(argv[3]) : -(N - n);
    int PITCH = N + 2, batch = 4;
    double *A, *out;
    cudaMalloc(&A, sizeof(double) * n * n * batch);
    cudaMalloc(&out, sizeof(double) * N * PITCH);
    double *hA = (double *)malloc(sizeof(double) * n * n * batch);
    for (int i = 0; i < n * n * batch; ++i) hA[i] = 1 + i;
    cudaMemcpy(A, hA, sizeof(double) * n * n * batch, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8};
    cuuint32_t box[3] = {(cuuint32_t)PITCH, (cuuint32_t)N, 1}, es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode n=%d N=%d start=%d -> %d\n", n, N, start, (int)r);
    size_t smem = N * PITCH * 8 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int st = 0; st < 2; ++st) {
        probe<<<1, 128, smem>>>(tm, out, start, start, N, PITCH, st);
        cudaError_t e = cudaDeviceSynchronize();
        printf("do_store=%d: %s\n", st, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        double *ho = (double *)malloc(sizeof(double) * N * PITCH);
        cudaMemcpy(ho, out, sizeof(double) * N * PITCH, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < N; ++c)
            for (int rr = 0; rr < PITCH; ++rr) {
                int gc = c + start, gr = rr + start;
                double want = (gc >= 0 && gc < n && gr >= 0 && gr < n) ? hA[gc * n + gr] : 0.0;
                if (ho[c * PITCH + rr] != want) ++bad;
            }
        printf("  load mismatches: %d\n", bad);
        if (st) {
            cudaMemcpy(hA + 0, A, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
            int bs = 0;
            for (int i = 0; i < n * n; ++i)
                if (hA[i] != 1 + i + 1000.0) ++bs;
            printf("  store mismatches: %d\n", bs);
        }
    }
    return 0;
}

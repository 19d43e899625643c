// capi.cu -- extern "C" boundary of libqlb200.so (see include/qlb200.h): handle management and the
// host-pointer / device-pointer entry points.  No torch types, no exceptions, no CPU fallback.
#include "common.cuh"
#include "internal.h"
#include <new>
#include <thread>
#include <vector>

#define QLB_REQUIRE(h, cond, argno, msg)                      \
    do {                                                      \
        if (!(cond)) return qlb_fail(h, -(argno), msg "%s");  \
    } while (0)

#include <cuda.h>      // driver-API types; the entry points are fetched through the runtime (no link-time libcuda dependency)

// SM partitions for the look-ahead ql! (see common.cuh).  Every failure simply leaves gc_ok = false.
template <typename F>
static bool drv(const char *name, F *fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
        (void)cudaGetLastError();
        return false;
    }
    *fn = (F)p;
    return true;
}
static void green_contexts_create(qlb200_ctx *h, unsigned panel_sms) {
    CUresult (*pDeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*pGetRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*pSplit)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned) = nullptr;
    CUresult (*pDesc)(CUdevResourceDesc *, CUdevResource *, unsigned) = nullptr;
    CUresult (*pCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned) = nullptr;
    CUresult (*pStream)(CUstream *, CUgreenCtx, unsigned, int) = nullptr;
    if (!drv("cuDeviceGet", &pDeviceGet) || !drv("cuDeviceGetDevResource", &pGetRes) || !drv("cuDevSmResourceSplitByCount", &pSplit) ||
        !drv("cuDevResourceGenerateDesc", &pDesc) || !drv("cuGreenCtxCreate", &pCreate) || !drv("cuGreenCtxStreamCreate", &pStream))
        return;
    CUdevice dev;
    CUdevResource all, grp, rem;
    if (pDeviceGet(&dev, h->device) != CUDA_SUCCESS || pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
    unsigned ng = 1;
    // MAX_POTENTIAL_CLUSTER_SIZE: the group must be able to host a 16-CTA cluster (SMs of one GPC)
    if (pSplit(&grp, &ng, &all, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE, panel_sms) != CUDA_SUCCESS || ng != 1) return;
    if (grp.sm.smCount < 16 || rem.sm.smCount < 64) return;
    CUdevResourceDesc d0, d1;
    if (pDesc(&d0, &grp, 1) != CUDA_SUCCESS || pDesc(&d1, &rem, 1) != CUDA_SUCCESS) return;
    CUgreenCtx g0 = nullptr, g1 = nullptr;
    if (pCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    if (pCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    int lo = 0, hi = 0;
    (void)cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CUstream s0 = nullptr, s1 = nullptr;
    if (pStream(&s0, g0, CU_STREAM_NON_BLOCKING, hi) != CUDA_SUCCESS || pStream(&s1, g1, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return;
    for (int i = 0; i < 4; ++i)
        if (cudaEventCreateWithFlags(&h->ev_gc[i], cudaEventDisableTiming) != cudaSuccess) return;
    h->gc_ctx[0] = g0;
    h->gc_ctx[1] = g1;
    h->gc_panel = (cudaStream_t)s0;
    h->gc_gemm = (cudaStream_t)s1;
    h->gc_panel_sms = (int)grp.sm.smCount;
    h->gc_gemm_sms = (int)rem.sm.smCount;
    h->gc_ok = true;
}
static void green_contexts_destroy(qlb200_ctx *h) {
    if (h->gc_panel) cudaStreamDestroy(h->gc_panel);
    if (h->gc_gemm) cudaStreamDestroy(h->gc_gemm);
    CUresult (*pDestroy)(CUgreenCtx) = nullptr;
    if (drv("cuGreenCtxDestroy", &pDestroy))
        for (int i = 0; i < 2; ++i)
            if (h->gc_ctx[i]) pDestroy((CUgreenCtx)h->gc_ctx[i]);
    for (int i = 0; i < 4; ++i)
        if (h->ev_gc[i]) cudaEventDestroy(h->ev_gc[i]);
    h->gc_ok = false;
}

extern "C" {

int qlb200_version(void) { return QLB200_VERSION; }

int qlb200_create(qlb200_handle *out, int device) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return 1000 + (int)e;
    if (device < 0 || device >= ndev) return -2;
    qlb200_ctx *h = new (std::nothrow) qlb200_ctx();
    if (!h) return 1000 + (int)cudaErrorMemoryAllocation;
    h->device = device;
    QLB_CUDA(h, cudaSetDevice(device));
    cudaDeviceProp prop;
    QLB_CUDA(h, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete h;
        return 1000 + (int)cudaErrorInvalidDevice;      // sm_100a only -- no fallback path
    }
    h->sm_count = prop.multiProcessorCount;
    QLB_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;      // side stream (look-ahead panel) gets the highest priority
        QLB_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        QLB_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, hi));
    }
    QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    QLB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    QLB_CUDA(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 32; ++i) QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) QLB_CUDA(h, cudaEventCreate(&h->ev_up[i]));
    QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_home, cudaEventDisableTiming));
    QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_rest[i], cudaEventDisableTiming));
    for (int i = 0; i < 8; ++i) QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_panel[i], cudaEventDisableTiming));
    QLB_CUDA(h, cudaMalloc(&h->dinfo, 64 + 256 * sizeof(unsigned)));      // status word + per-SM tickets
    QLB_CUDA(h, cudaMemset(h->dinfo, 0, 64 + 256 * sizeof(unsigned)));
    QLB_CUDA(h, cudaMalloc(&h->tc_sum_dev, 2 * sizeof(unsigned long long)));
    QLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_tc, cudaEventDisableTiming));
    h->stream = h->own_stream;
    green_contexts_create(h, 16);
    (void)cudaGetLastError();
    *out = h;
    return 0;
}

int qlb200_destroy(qlb200_handle h) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws) cudaFree(h->ws);
    if (h->ws2) cudaFree(h->ws2);
    if (h->stage) cudaFree(h->stage);
    if (h->stage2) cudaFree(h->stage2);
    if (h->dinfo) cudaFree(h->dinfo);
    if (h->tc_sum_dev) cudaFree(h->tc_sum_dev);
    if (h->ev_tc) cudaEventDestroy(h->ev_tc);
    if (h->redo) cudaFree(h->redo);
    for (int i = 0; i < 3; ++i)
        if (h->cvt[i]) cudaFree(h->cvt[i]);
    if (h->zws) cudaFree(h->zws);
    if (h->tall) cudaFree(h->tall);
    if (h->tcache) cudaFree(h->tcache);
    if (h->joinV) cudaFree(h->joinV);
    if (h->hgraph_exec) cudaGraphExecDestroy(h->hgraph_exec);
    if (h->ev_home) cudaEventDestroy(h->ev_home);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < 2; ++i)
        if (h->ev_rest[i]) cudaEventDestroy(h->ev_rest[i]);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->prof_ev) {
        for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
        delete[] h->prof_ev;
        delete[] h->prof_flops;
    }
    green_contexts_destroy(h);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    for (int i = 0; i < 8; ++i)
        if (h->ev_panel[i]) cudaEventDestroy(h->ev_panel[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 32; ++i)
        if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
    for (int i = 0; i < 2; ++i)
        if (h->ev_up[i]) cudaEventDestroy(h->ev_up[i]);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    delete h;
    return 0;
}

int qlb200_set_stream(qlb200_handle h, void *cuda_stream) {
    if (!h) return -1;
    h->stream = (cudaStream_t)cuda_stream;        // NULL = the CUDA default stream
    return 0;
}

int qlb200_reset_stream(qlb200_handle h) {
    if (!h) return -1;
    h->stream = h->own_stream;
    return 0;
}

int qlb200_sync(qlb200_handle h) {
    if (!h) return -1;
    QLB_CUDA(h, cudaSetDevice(h->device));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

const char *qlb200_last_error(qlb200_handle h) { return h ? h->err : "null handle"; }

int qlb200_profile_collect(qlb200_handle h, double *ms_total, double *flops_total, int64_t *launches) {
    if (!h) return -1;
    QLB_CUDA(h, cudaSetDevice(h->device));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    double ms = 0.0, fl = 0.0;
    int64_t cnt = 0;
    for (int i = 0; i < h->prof_n; ++i) {
        if (h->prof_flops[i] <= 0.0) continue;
        float t = 0.f;
        QLB_CUDA(h, cudaEventElapsedTime(&t, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        ms += t;
        fl += h->prof_flops[i];
        ++cnt;
    }
    h->prof_n = 0;
    if (ms_total) *ms_total = ms;
    if (flops_total) *flops_total = fl;
    if (launches) *launches = cnt;
    return 0;
}

int64_t qlb200_launch_count(qlb200_handle h) { return h ? h->launches : -1; }
double qlb200_upload_rate(qlb200_handle h) { return h ? h->h2d_gbps : -1.0; }
// Page-lock / release a caller's host array (cudaHostRegister, portable): host-pointer calls then copy at PCIe speed
// (pageable: ~10 GB/s here, pinned: ~53 GB/s) and the large ql! may replay its whole schedule as a CUDA graph.
int qlb200_host_register(qlb200_handle h, void *ptr, int64_t bytes) {
    if (!h) return -1;
    QLB_REQUIRE(h, ptr != nullptr, 2, "ptr is null");
    QLB_REQUIRE(h, bytes > 0, 3, "bytes must be positive");
    QLB_CUDA(h, cudaSetDevice(h->device));
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return 0;
    }
    QLB_CUDA(h, e);
    return 0;
}
int qlb200_host_unregister(qlb200_handle h, void *ptr) {
    if (!h) return -1;
    QLB_REQUIRE(h, ptr != nullptr, 2, "ptr is null");
    QLB_CUDA(h, cudaSetDevice(h->device));
    QLB_CUDA(h, cudaDeviceSynchronize());          // nothing of this handle may still be copying from / to the array
    if (h->hgraph_exec) {                          // the cached host schedule holds this array's address
        cudaGraphExecDestroy(h->hgraph_exec);
        h->hgraph_exec = nullptr;
    }
    memset(h->hgraph_key, 0, sizeof(h->hgraph_key));
    memset(h->hgraph_seen_key, 0, sizeof(h->hgraph_seen_key));
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        (void)cudaGetLastError();
        return 0;
    }
    QLB_CUDA(h, e);
    return 0;
}

int qlb200_set_option(qlb200_handle h, const char *name, int64_t value) {
    if (!h) return -1;
    if (!name) return -2;
    if (!strcmp(name, "nb")) {
        if (value != 0 && (value < 16 || value > 256 || value % 16)) return qlb_fail(h, -3, "nb must be 0 (auto) or a multiple of 16 in [16,256]%s");
        h->opt_nb = value;
    } else if (!strcmp(name, "gemm_cfg")) {
        if (value < -1 || value > 3) return qlb_fail(h, -3, "gemm_cfg must be in [-1,3]%s");
        h->opt_gemm_cfg = value;
    } else if (!strcmp(name, "profile")) {
        if (value && !h->prof_ev) {
            h->prof_cap = 8192;
            h->prof_ev = new (std::nothrow) cudaEvent_t[2 * h->prof_cap];
            h->prof_flops = new (std::nothrow) double[h->prof_cap];
            if (!h->prof_ev || !h->prof_flops) return qlb_fail(h, 1000 + (int)cudaErrorMemoryAllocation, "profile buffers%s");
            for (int i = 0; i < 2 * h->prof_cap; ++i) QLB_CUDA(h, cudaEventCreate(&h->prof_ev[i]));
        }
        h->opt_profile = value ? 1 : 0;
        h->prof_n = 0;
    } else if (!strcmp(name, "strip_rpt")) {
        if (value != 1 && value != 2 && value != 4) return qlb_fail(h, -3, "strip_rpt must be 1, 2 or 4%s");
        h->opt_strip_rpt = value;
    } else if (!strcmp(name, "trail_splits")) {
        if (value < 0 || value > 8) return qlb_fail(h, -3, "trail_splits must be in [0,8]%s");
        h->opt_trail_splits = value;
    } else if (!strcmp(name, "tcache")) {
        if (value < 0 || value > 2) return qlb_fail(h, -3, "tcache must be 0 (off), 1 (validated by checksum) or 2 (trusted: keyed on pointers)%s");
        h->opt_tcache = value;
        h->tc_valid = false;
    } else if (!strcmp(name, "overlap_t")) {
        h->opt_overlap_t = value ? 1 : 0;
    } else if (!strcmp(name, "ul_cluster")) {
        h->opt_ul_cluster = value ? 1 : 0;
    } else if (!strcmp(name, "la_pnarrow_min")) {
        if (value < 0) return qlb_fail(h, -3, "la_pnarrow_min must be >= 0%s");
        h->opt_la_pnarrow_min = value;
        h->graph_key[0] = 0; h->graph_seen_key[0] = 0; h->hgraph_key[0] = 0; h->hgraph_seen_key[0] = 0;
    } else if (!strcmp(name, "la_join_ramp")) {
        if (value < 0 || value > 64) return qlb_fail(h, -3, "la_join_ramp must be in [0,64]%s");
        h->opt_la_join_ramp = value;
    } else if (!strcmp(name, "la_join_chunks")) {
        if (value < 1 || value > 32) return qlb_fail(h, -3, "la_join_chunks must be in [1,32]%s");
        h->opt_la_join_chunks = value;
    } else if (!strcmp(name, "join_step")) {
        if (value < 1 || value > 64) return qlb_fail(h, -3, "join_step must be in [1,64]%s");
        h->opt_join_step = value;
    } else if (!strcmp(name, "chunk_cols")) {
        if (value < 128 || value % 128) return qlb_fail(h, -3, "chunk_cols must be a positive multiple of 128%s");
        h->opt_chunk_cols = value;
    } else if (!strcmp(name, "pair_sync")) {
        if (value < 0 || value > 1024) return qlb_fail(h, -3, "pair_sync must be in [0,1024]%s");
        h->opt_pair_sync = value;
        h->graph_key[0] = 0;                      // a captured graph holds the old launch shape
    } else if (!strcmp(name, "batch_chunk_mb")) {
        if (value < 0 || value > 4096) return qlb_fail(h, -3, "batch_chunk_mb must be in [0,4096]%s");
        h->opt_batch_chunk_mb = value;
    } else if (!strcmp(name, "stream_in")) {
        h->opt_stream_in = value ? 1 : 0;
    } else if (!strcmp(name, "graph")) {
        h->opt_graph = value ? 1 : 0;
    } else if (!strcmp(name, "bvar")) {
        h->opt_bvar = value;
    } else if (!strcmp(name, "bgrid")) {
        h->opt_bgrid = value < 0 ? 0 : value;
    } else if (!strcmp(name, "lookahead")) {
        if (value < 0 || value > 2) return qlb_fail(h, -3, "lookahead must be 0 (off), 1 (on) or 2 (auto)%s");
        h->opt_lookahead = value;
    } else if (!strcmp(name, "la_graph")) {
        h->opt_la_graph = value ? 1 : 0;
        h->graph_key[0] = 0;
        h->graph_seen_key[0] = 0;
    } else if (!strcmp(name, "la_min")) {
        h->opt_la_min = value < 0 ? 0 : value;
    } else if (!strcmp(name, "host_split")) {
        if (value < 0 || value > 2) return qlb_fail(h, -3, "host_split must be 0, 1 or 2%s");
        h->opt_host_split = (int)value;
        h->h2d_gbps = 0.0;
    } else if (!strcmp(name, "host_split_min")) {
        if (value < 512) return qlb_fail(h, -3, "host_split_min must be >= 512%s");
        h->opt_host_split_min = value;
    } else if (!strcmp(name, "host_split_chunk")) {
        if (value < 0) return qlb_fail(h, -3, "host_split_chunk must be >= 0%s");
        h->opt_host_split_chunk = value;
    } else if (!strcmp(name, "fuse_chunk")) {
        h->opt_fuse_chunk = value < 0 ? 0 : value;
    } else if (!strcmp(name, "la_gemm_all")) {
        h->opt_la_gemm_all = value ? 1 : 0;
    } else if (!strcmp(name, "strip")) {
        if (value != 8 && value != 16 && value != 32) return qlb_fail(h, -3, "strip must be 8, 16 or 32%s");
        h->opt_strip = value;
    } else {
        return qlb_fail(h, -2, "unknown option %s", name);
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// host-pointer batched calls: three-stream chunk pipeline (upload | compute | download)
// ---------------------------------------------------------------------------------------------
static int qlb_cuda(qlb200_ctx *h, cudaError_t e) {
    if (e == cudaSuccess) return 0;
    return qlb_fail(h, 1000 + (int)e, "batched host pipeline: %s", cudaGetErrorString(e));
}
// matrices per chunk: about "batch_chunk_mb" (32) MB of operands, at least 4 chunks for batches worth splitting
static int64_t batched_chunk(qlb200_ctx *h, int64_t bytes_per_matrix, int64_t batch) {
    if (h->opt_batch_chunk_mb == 0) return batch;          // one chunk: upload, compute, download back to back
    int64_t cm = (h->opt_batch_chunk_mb << 20) / (bytes_per_matrix > 0 ? bytes_per_matrix : 1);
    if (cm < 1) cm = 1;
    if (cm * 4 > batch && batch * bytes_per_matrix >= ((int64_t)8 << 20)) cm = (batch + 3) / 4;
    return cm < batch ? cm : batch;
}
// the staging buffer may still be read by earlier work on the compute stream
static int pipeline_begin(qlb200_ctx *h) {
    int rc = qlb_cuda(h, cudaEventRecord(h->ev_chunk[31], h->stream));
    if (!rc) rc = qlb_cuda(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_chunk[31], 0));
    return rc;
}
// Events are reused round-robin: a wait refers to the record that precedes it in host order, so re-recording an event
// later does not disturb waits already enqueued.
static int pipeline_uploaded(qlb200_ctx *h, int64_t c) {
    cudaEvent_t ev = h->ev_chunk[(2 * c) % 30];
    int rc = qlb_cuda(h, cudaEventRecord(ev, h->h2d_stream));
    if (!rc) rc = qlb_cuda(h, cudaStreamWaitEvent(h->stream, ev, 0));
    return rc;
}
static int pipeline_computed(qlb200_ctx *h, int64_t c) {
    cudaEvent_t ev = h->ev_chunk[(2 * c + 1) % 30];
    int rc = qlb_cuda(h, cudaEventRecord(ev, h->stream));
    if (!rc) rc = qlb_cuda(h, cudaStreamWaitEvent(h->copy_stream, ev, 0));
    return rc;
}
// every stream is drained before the caller's host arrays are handed back, also on the error path
static int pipeline_end(qlb200_ctx *h, int rc) {
    cudaError_t e0 = cudaStreamSynchronize(h->h2d_stream);
    cudaError_t e1 = cudaStreamSynchronize(h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->copy_stream);
    if (rc) return rc;
    if (e0 != cudaSuccess) return qlb_cuda(h, e0);
    if (e1 != cudaSuccess) return qlb_cuda(h, e1);
    return qlb_cuda(h, e2);
}

// ---------------------------------------------------------------------------------------------
// batched: argument checks shared by the typed entry points
// ---------------------------------------------------------------------------------------------
template <typename T>
static int check_batched(qlb200_ctx *h, int64_t n, const T *A, int64_t lda, int64_t strideA, const T *tau,
                         int64_t strideTau, int64_t batch) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 1 && n <= 64, 2, "n must be in [1,64]");
    QLB_REQUIRE(h, A != nullptr || batch == 0, 3, "A is null");
    QLB_REQUIRE(h, lda >= n, 4, "lda < n");
    QLB_REQUIRE(h, strideA >= lda * (n - 1) + n || batch <= 1, 5, "strideA too small (matrices overlap)");
    QLB_REQUIRE(h, tau != nullptr || batch == 0, 6, "tau is null");
    QLB_REQUIRE(h, strideTau >= n || batch <= 1, 7, "strideTau < n");
    QLB_REQUIRE(h, batch >= 0, 8, "batch < 0");
    return 0;
}

template <typename T>
static int geqlf_batched_dev(qlb200_ctx *h, int64_t n, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau,
                             int64_t batch);
template <>
int geqlf_batched_dev<double>(qlb200_ctx *h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                              int64_t strideTau, int64_t batch) {
    return qlb_geqlf_batched_f64(h, n, A, lda, strideA, tau, strideTau, batch);
}
template <>
int geqlf_batched_dev<float>(qlb200_ctx *h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                             int64_t strideTau, int64_t batch) {
    return qlb_geqlf_batched_f32(h, n, A, lda, strideA, tau, strideTau, batch);
}
static int apply_batched_dev(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                             int64_t strideA, const double *tau, int64_t strideTau, double *B, int64_t ldb,
                             int64_t strideB, int64_t batch, int32_t *info) {
    return qlb_apply_batched_f64(h, mode, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB, batch, info);
}
static int apply_batched_dev(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                             int64_t strideA, const float *tau, int64_t strideTau, float *B, int64_t ldb,
                             int64_t strideB, int64_t batch, int32_t *info) {
    return qlb_apply_batched_f32(h, mode, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB, batch, info);
}

template <typename T>
static int geqlf_batched_any(qlb200_ctx *h, bool host, int64_t n, T *A, int64_t lda, int64_t strideA, T *tau,
                             int64_t strideTau, int64_t batch) {
    int rc = check_batched(h, n, A, lda, strideA, tau, strideTau, batch);
    if (rc) return rc;
    if (batch == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) return geqlf_batched_dev<T>(h, n, A, lda, strideA, tau, strideTau, batch);
    // host pointers: the strided slab is staged on the device and processed in chunks of matrices; chunk c+1 uploads
    // (h2d_stream) while chunk c is factored (h->stream) and chunk c-1 downloads (copy_stream), so with pinned host
    // memory both PCIe directions are busy at once.  Pageable memory makes the copies synchronous: same result, no overlap.
    const size_t a_elems = (size_t)((batch - 1) * strideA + lda * (n - 1) + n);
    const size_t t_elems = (size_t)((batch - 1) * strideTau + n);
    const size_t a_bytes = (a_elems * sizeof(T) + 255) & ~(size_t)255;
    rc = qlb_reserve(h, &h->stage, &h->stage_bytes, a_bytes + t_elems * sizeof(T));
    if (rc) return rc;
    T *dA = (T *)h->stage, *dT = (T *)((char *)h->stage + a_bytes);
    const int64_t cm = batched_chunk(h, strideA * (int64_t)sizeof(T), batch);
    rc = pipeline_begin(h);
    for (int64_t m0 = 0, c = 0; m0 < batch && !rc; m0 += cm, ++c) {
        const int64_t mc = (batch - m0 < cm) ? batch - m0 : cm;
        const size_t ea = (size_t)((mc - 1) * strideA + lda * (n - 1) + n), et = (size_t)((mc - 1) * strideTau + n);
        T *dAc = dA + m0 * strideA, *dTc = dT + m0 * strideTau;
        rc = qlb_cuda(h, cudaMemcpyAsync(dAc, A + m0 * strideA, ea * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = pipeline_uploaded(h, c);
        if (!rc) rc = geqlf_batched_dev<T>(h, n, dAc, lda, strideA, dTc, strideTau, mc);
        if (!rc) rc = pipeline_computed(h, c);
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(A + m0 * strideA, dAc, ea * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
        // tau was never uploaded: only the n entries of each matrix come back (a contiguous copy would overwrite the caller's
        // gaps between the tau vectors, strideTau > n, with staging-buffer contents)
        if (!rc) {
            if (strideTau == n) rc = qlb_cuda(h, cudaMemcpyAsync(tau + m0 * strideTau, dTc, et * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
            else rc = qlb_cuda(h, cudaMemcpy2DAsync(tau + m0 * strideTau, (size_t)strideTau * sizeof(T), dTc, (size_t)strideTau * sizeof(T),
                                                    (size_t)n * sizeof(T), (size_t)mc, cudaMemcpyDeviceToHost, h->copy_stream));
        }
    }
    return pipeline_end(h, rc);
}

// mode 0: lmul!(Q,B); 1: lmul!(Q',B); 2: ldiv!(F,B)
template <typename T>
static int apply_batched_any(qlb200_ctx *h, bool host, int mode, int64_t n, int64_t nrhs, const T *A, int64_t lda,
                             int64_t strideA, const T *tau, int64_t strideTau, T *B, int64_t ldb, int64_t strideB,
                             int64_t batch, int32_t *info_out, int argbase) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 1 && n <= 64, argbase, "n must be in [1,64]");
    QLB_REQUIRE(h, nrhs >= 0, argbase + 1, "nrhs < 0");
    QLB_REQUIRE(h, A != nullptr || batch == 0, argbase + 2, "A is null");
    QLB_REQUIRE(h, lda >= n, argbase + 3, "lda < n");
    QLB_REQUIRE(h, strideA >= lda * (n - 1) + n || batch <= 1, argbase + 4, "strideA too small");
    QLB_REQUIRE(h, tau != nullptr || batch == 0, argbase + 5, "tau is null");
    QLB_REQUIRE(h, strideTau >= n || batch <= 1, argbase + 6, "strideTau < n");
    QLB_REQUIRE(h, B != nullptr || batch == 0 || nrhs == 0, argbase + 7, "B is null");
    QLB_REQUIRE(h, ldb >= n, argbase + 8, "ldb < n (reference: DimensionMismatch)");
    QLB_REQUIRE(h, nrhs == 0 || strideB >= ldb * (nrhs - 1) + n || batch <= 1, argbase + 9, "strideB too small");
    QLB_REQUIRE(h, batch >= 0, argbase + 10, "batch < 0");
    if (batch == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) return apply_batched_dev(h, mode, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB, batch, info_out);
    const size_t a_elems = (size_t)((batch - 1) * strideA + lda * (n - 1) + n);
    const size_t t_elems = (size_t)((batch - 1) * strideTau + n);
    const size_t b_elems = (size_t)((batch - 1) * strideB + ldb * (nrhs - 1) + n);
    const size_t a_bytes = (a_elems * sizeof(T) + 255) & ~(size_t)255;
    const size_t t_bytes = (t_elems * sizeof(T) + 255) & ~(size_t)255;
    const size_t b_bytes = (b_elems * sizeof(T) + 255) & ~(size_t)255;
    int rc = qlb_reserve(h, &h->stage, &h->stage_bytes, a_bytes + t_bytes + b_bytes + (size_t)batch * 4);
    if (rc) return rc;
    T *dA = (T *)h->stage, *dT = (T *)((char *)h->stage + a_bytes), *dB = (T *)((char *)h->stage + a_bytes + t_bytes);
    int32_t *dI = (int32_t *)((char *)h->stage + a_bytes + t_bytes + b_bytes);
    const int64_t per = (strideA + strideB) * (int64_t)sizeof(T);
    const int64_t cm = batched_chunk(h, per, batch);
    rc = pipeline_begin(h);
    for (int64_t m0 = 0, c = 0; m0 < batch && !rc; m0 += cm, ++c) {
        const int64_t mc = (batch - m0 < cm) ? batch - m0 : cm;
        const size_t ea = (size_t)((mc - 1) * strideA + lda * (n - 1) + n), et = (size_t)((mc - 1) * strideTau + n);
        const size_t eb = (size_t)((mc - 1) * strideB + ldb * (nrhs - 1) + n);
        T *dAc = dA + m0 * strideA, *dTc = dT + m0 * strideTau, *dBc = dB + m0 * strideB;
        rc = qlb_cuda(h, cudaMemcpyAsync(dAc, A + m0 * strideA, ea * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(dTc, tau + m0 * strideTau, et * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(dBc, B + m0 * strideB, eb * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = pipeline_uploaded(h, c);
        if (!rc) rc = apply_batched_dev(h, mode, n, nrhs, dAc, lda, strideA, dTc, strideTau, dBc, ldb, strideB, mc, (mode == 2) ? dI + m0 : nullptr);
        if (!rc) rc = pipeline_computed(h, c);
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(B + m0 * strideB, dBc, eb * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
        if (!rc && mode == 2 && info_out)
            rc = qlb_cuda(h, cudaMemcpyAsync(info_out + m0, dI + m0, (size_t)mc * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    return pipeline_end(h, rc);
}

// Fused factor + apply (BASELINE configs[2] and [4] as ONE call): A <- ql!(A), then B <- Q'B / Q B / L^{-1} Q'B with the
// factors just computed.  Device pointers: the two kernels back to back (option `fuse_chunk` > 0 walks the batch in chunks
// so that the apply finds the factors in L2 -- measured slower, see common.cuh).  Host pointers: one upload | compute |
// download pipeline -- the factors cross PCIe once in each direction instead of up, down and up again.
template <typename T>
static int geqlf_apply_batched_any(qlb200_ctx *h, bool host, int mode, int64_t n, int64_t nrhs, T *A, int64_t lda, int64_t strideA,
                                   T *tau, int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch, int32_t *info_out) {
    int rc = check_batched(h, n, A, lda, strideA, tau, strideTau, batch);
    if (rc) return rc;
    QLB_REQUIRE(h, nrhs >= 1, 4, "nrhs < 1");
    QLB_REQUIRE(h, B != nullptr || batch == 0, 10, "B is null");
    QLB_REQUIRE(h, ldb >= n, 11, "ldb < n (reference: DimensionMismatch)");
    QLB_REQUIRE(h, strideB >= ldb * (nrhs - 1) + n || batch <= 1, 12, "strideB too small");
    if (batch == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) {
        const int64_t cm = h->opt_fuse_chunk > 0 ? h->opt_fuse_chunk : batch;
        for (int64_t m0 = 0; m0 < batch; m0 += cm) {
            const int64_t mc = (batch - m0 < cm) ? batch - m0 : cm;
            rc = geqlf_batched_dev<T>(h, n, A + m0 * strideA, lda, strideA, tau + m0 * strideTau, strideTau, mc);
            if (!rc) rc = apply_batched_dev(h, mode, n, nrhs, A + m0 * strideA, lda, strideA, tau + m0 * strideTau, strideTau, B + m0 * strideB, ldb,
                                            strideB, mc, (mode == 2 && info_out) ? info_out + m0 : nullptr);
            if (rc) return rc;
        }
        return 0;
    }
    const size_t a_elems = (size_t)((batch - 1) * strideA + lda * (n - 1) + n);
    const size_t t_elems = (size_t)((batch - 1) * strideTau + n);
    const size_t b_elems = (size_t)((batch - 1) * strideB + ldb * (nrhs - 1) + n);
    const size_t a_bytes = (a_elems * sizeof(T) + 255) & ~(size_t)255;
    const size_t t_bytes = (t_elems * sizeof(T) + 255) & ~(size_t)255;
    const size_t b_bytes = (b_elems * sizeof(T) + 255) & ~(size_t)255;
    rc = qlb_reserve(h, &h->stage, &h->stage_bytes, a_bytes + t_bytes + b_bytes + (size_t)batch * 4);
    if (rc) return rc;
    T *dA = (T *)h->stage, *dT = (T *)((char *)h->stage + a_bytes), *dB = (T *)((char *)h->stage + a_bytes + t_bytes);
    int32_t *dI = (int32_t *)((char *)h->stage + a_bytes + t_bytes + b_bytes);
    const int64_t cm = batched_chunk(h, (strideA + strideB) * (int64_t)sizeof(T), batch);
    rc = pipeline_begin(h);
    for (int64_t m0 = 0, c = 0; m0 < batch && !rc; m0 += cm, ++c) {
        const int64_t mc = (batch - m0 < cm) ? batch - m0 : cm;
        const size_t ea = (size_t)((mc - 1) * strideA + lda * (n - 1) + n), et = (size_t)((mc - 1) * strideTau + n);
        const size_t eb = (size_t)((mc - 1) * strideB + ldb * (nrhs - 1) + n);
        T *dAc = dA + m0 * strideA, *dTc = dT + m0 * strideTau, *dBc = dB + m0 * strideB;
        rc = qlb_cuda(h, cudaMemcpyAsync(dAc, A + m0 * strideA, ea * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(dBc, B + m0 * strideB, eb * sizeof(T), cudaMemcpyHostToDevice, h->h2d_stream));
        if (!rc) rc = pipeline_uploaded(h, c);
        if (!rc) rc = geqlf_batched_dev<T>(h, n, dAc, lda, strideA, dTc, strideTau, mc);
        if (!rc) rc = apply_batched_dev(h, mode, n, nrhs, dAc, lda, strideA, dTc, strideTau, dBc, ldb, strideB, mc, (mode == 2) ? dI + m0 : nullptr);
        if (!rc) rc = pipeline_computed(h, c);
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(A + m0 * strideA, dAc, ea * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
        if (!rc) {
            if (strideTau == n) rc = qlb_cuda(h, cudaMemcpyAsync(tau + m0 * strideTau, dTc, et * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
            else rc = qlb_cuda(h, cudaMemcpy2DAsync(tau + m0 * strideTau, (size_t)strideTau * sizeof(T), dTc, (size_t)strideTau * sizeof(T),
                                                    (size_t)n * sizeof(T), (size_t)mc, cudaMemcpyDeviceToHost, h->copy_stream));
        }
        if (!rc) rc = qlb_cuda(h, cudaMemcpyAsync(B + m0 * strideB, dBc, eb * sizeof(T), cudaMemcpyDeviceToHost, h->copy_stream));
        if (!rc && mode == 2 && info_out)
            rc = qlb_cuda(h, cudaMemcpyAsync(info_out + m0, dI + m0, (size_t)mc * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    return pipeline_end(h, rc);
}

static int trans_mode(qlb200_ctx *h, char trans, int *mode) {
    if (trans == 'N' || trans == 'n') { *mode = 0; return 0; }
    if (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c') { *mode = 1; return 0; }
    return qlb_fail(h, -2, "trans must be 'N' or 'T'%s");
}

extern "C" {

int qlb200_dgeqlf_batched(qlb200_handle h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                          int64_t strideTau, int64_t batch) {
    return geqlf_batched_any<double>(h, true, n, A, lda, strideA, tau, strideTau, batch);
}
int qlb200_sgeqlf_batched(qlb200_handle h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                          int64_t strideTau, int64_t batch) {
    return geqlf_batched_any<float>(h, true, n, A, lda, strideA, tau, strideTau, batch);
}
int qlb200_dgeqlf_batched_dev(qlb200_handle h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                              int64_t strideTau, int64_t batch) {
    return geqlf_batched_any<double>(h, false, n, A, lda, strideA, tau, strideTau, batch);
}
int qlb200_sgeqlf_batched_dev(qlb200_handle h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                              int64_t strideTau, int64_t batch) {
    return geqlf_batched_any<float>(h, false, n, A, lda, strideA, tau, strideTau, batch);
}

#define QLB_ORMQL_BATCHED(NAME, T, HOST)                                                                           \
    int NAME(qlb200_handle h, char trans, int64_t n, int64_t nrhs, const T *A, int64_t lda, int64_t strideA,       \
             const T *tau, int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch) {                 \
        if (!h) return -1;                                                                                         \
        int mode;                                                                                                  \
        int rc = trans_mode(h, trans, &mode);                                                                      \
        if (rc) return rc;                                                                                         \
        return apply_batched_any<T>(h, HOST, mode, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB,      \
                                    batch, nullptr, 3);                                                            \
    }
QLB_ORMQL_BATCHED(qlb200_dormql_batched, double, true)
QLB_ORMQL_BATCHED(qlb200_sormql_batched, float, true)
QLB_ORMQL_BATCHED(qlb200_dormql_batched_dev, double, false)
QLB_ORMQL_BATCHED(qlb200_sormql_batched_dev, float, false)

/* fused: op 'N' -> B <- Q B, 'T'/'C' -> B <- Q'B, 'S' -> B <- L^{-1} Q'B (ldiv!); info (n entries, may be null) only for 'S' */
#define QLB_GEQLF_APPLY_BATCHED(NAME, T, HOST)                                                                     \
    int NAME(qlb200_handle h, char op, int64_t n, int64_t nrhs, T *A, int64_t lda, int64_t strideA, T *tau,         \
             int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch, int32_t *info) {                 \
        if (!h) return -1;                                                                                         \
        int mode = 2;                                                                                              \
        if (!(op == 'S' || op == 's')) {                                                                           \
            int rc = trans_mode(h, op, &mode);                                                                     \
            if (rc) return rc;                                                                                     \
        }                                                                                                          \
        return geqlf_apply_batched_any<T>(h, HOST, mode, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB, batch, info); \
    }
QLB_GEQLF_APPLY_BATCHED(qlb200_dgeqlf_apply_batched, double, true)
QLB_GEQLF_APPLY_BATCHED(qlb200_sgeqlf_apply_batched, float, true)
QLB_GEQLF_APPLY_BATCHED(qlb200_dgeqlf_apply_batched_dev, double, false)
QLB_GEQLF_APPLY_BATCHED(qlb200_sgeqlf_apply_batched_dev, float, false)

#define QLB_QLSOLVE_BATCHED(NAME, T, HOST)                                                                         \
    int NAME(qlb200_handle h, int64_t n, int64_t nrhs, const T *A, int64_t lda, int64_t strideA, const T *tau,     \
             int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch, int32_t *info_out) {            \
        return apply_batched_any<T>(h, HOST, 2, n, nrhs, A, lda, strideA, tau, strideTau, B, ldb, strideB, batch,  \
                                    info_out, 2);                                                                  \
    }
QLB_QLSOLVE_BATCHED(qlb200_dqlsolve_batched, double, true)
QLB_QLSOLVE_BATCHED(qlb200_sqlsolve_batched, float, true)
QLB_QLSOLVE_BATCHED(qlb200_dqlsolve_batched_dev, double, false)
QLB_QLSOLVE_BATCHED(qlb200_sqlsolve_batched_dev, float, false)

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// single matrix: argument checks (LAPACK order: the handle is argument 1)
// ---------------------------------------------------------------------------------------------
static int check_fact(qlb200_ctx *h, int64_t m, int64_t n, const double *A, int64_t lda, const double *tau) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 3, "n out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 5, "lda < max(1,m)");
    QLB_REQUIRE(h, tau != nullptr || m * n == 0, 6, "tau is null");
    return 0;
}

// stage a column-major host matrix into a device buffer with an even leading dimension
static int stage_in(qlb200_ctx *h, void **buf, size_t *have, const double *A, int64_t m, int64_t n, int64_t lda,
                    double **dA, int64_t *ldd) {
    *ldd = (m + 1) & ~(int64_t)1;
    if (*ldd < 2) *ldd = 2;
    int rc = qlb_reserve(h, buf, have, (size_t)(*ldd) * (n > 0 ? n : 1) * sizeof(double));
    if (rc) return rc;
    *dA = (double *)*buf;
    if (m > 0 && n > 0)
        QLB_CUDA(h, cudaMemcpy2DAsync(*dA, (size_t)(*ldd) * 8, A, (size_t)lda * 8, (size_t)m * 8, (size_t)n,
                                      cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int stage_out(qlb200_ctx *h, double *A, int64_t m, int64_t n, int64_t lda, const double *dA, int64_t ldd) {
    if (m > 0 && n > 0)
        QLB_CUDA(h, cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)ldd * 8, (size_t)m * 8, (size_t)n,
                                      cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

typedef int (*fact_fn)(qlb200_ctx *, int64_t, int64_t, double *, int64_t, double *);

// ql! on device pointers, replayed as a CUDA graph when the same call comes again.  First sight of a key: plain launches
// (workspace allocation and function attributes happen here).  Second sight: the launch sequence is captured on the
// handle's stream and instantiated.  From then on: one cudaGraphLaunch.  Skipped while profiling, on the legacy default
// stream and inside somebody else's capture.
static int geqlf_dev_graph(qlb200_ctx *h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau) {
    const int64_t key[10] = {m, n, (int64_t)(uintptr_t)dA, lda, (int64_t)(uintptr_t)dtau, (int64_t)qlb_effective_nb(h, m, n), h->opt_strip_rpt * 16 + h->opt_lookahead,
                             h->opt_gemm_cfg * 16 + h->opt_trail_splits, (int64_t)(uintptr_t)h->stream, 1 + 2 * h->opt_overlap_t + 4 * h->opt_tcache + 16 * h->opt_la_graph + 32 * h->opt_la_gemm_all};
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    bool usable = h->opt_graph && !h->opt_profile && (h->opt_la_graph || !qlb_lookahead_active(h, m, n)) && h->stream != nullptr && m * n >= 64 * 64;
    if (usable && (cudaStreamIsCapturing(h->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)) {
        (void)cudaGetLastError();
        usable = false;
    }
    if (!usable) return qlb_dgeqlf_dev(h, m, n, dA, lda, dtau);
    if (h->graph_exec && !memcmp(key, h->graph_key, sizeof(key))) {
        QLB_CUDA(h, cudaGraphLaunch(h->graph_exec, h->stream));
        h->launches += h->graph_nodes;
        if (h->opt_tcache && h->tcache) {          // the replay refills the T cache of exactly this factorization
            const int64_t k = m < n ? m : n;
            h->tc_key[0] = (int64_t)(uintptr_t)(dA + (n - k) * lda);
            h->tc_key[1] = lda;
            h->tc_key[2] = (int64_t)(uintptr_t)dtau;
            h->tc_key[3] = m;
            h->tc_key[4] = k;
            h->tc_key[5] = qlb_effective_nb(h, m, n);
            h->tc_valid = true;
            QLB_CUDA(h, cudaEventRecord(h->ev_tc, h->stream));
        }
        return 0;
    }
    if (memcmp(key, h->graph_seen_key, sizeof(key))) {           // first sight: run it plainly
        memcpy(h->graph_seen_key, key, sizeof(key));
        return qlb_dgeqlf_dev(h, m, n, dA, lda, dtau);
    }
    if (h->graph_exec) {
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
    }
    const int64_t l0 = h->launches;
    QLB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = qlb_dgeqlf_dev(h, m, n, dA, lda, dtau);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc || e != cudaSuccess || !graph) {
        (void)cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        h->launches = l0;
        if (rc) return rc;
        return qlb_dgeqlf_dev(h, m, n, dA, lda, dtau);          // capture refused: plain launches
    }
    e = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        h->graph_exec = nullptr;
        h->launches = l0;
        return qlb_dgeqlf_dev(h, m, n, dA, lda, dtau);
    }
    h->graph_nodes = h->launches - l0;
    memcpy(h->graph_key, key, sizeof(key));
    QLB_CUDA(h, cudaGraphLaunch(h->graph_exec, h->stream));
    if (h->tc_valid) QLB_CUDA(h, cudaEventRecord(h->ev_tc, h->stream));
    return 0;
}

// Upload rate of the first piece of a host-pointer ql! (events ev_up[0..1] on the H2D stream, read after the call's final sync).
static void note_upload_rate(qlb200_ctx *h) {
    float ms = 0.f;
    if (h->up_bytes > 0.0 && cudaEventElapsedTime(&ms, h->ev_up[0], h->ev_up[1]) == cudaSuccess && ms > 0.f)
        h->h2d_gbps = h->up_bytes / (ms * 1e6);
    else
        cudaGetLastError();
    h->up_bytes = 0.0;
}

static int fact_host(qlb200_ctx *h, fact_fn f, int64_t m, int64_t n, double *A, int64_t lda, double *tau) {
    int rc = check_fact(h, m, n, A, lda, tau);
    if (rc) return rc;
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    // ql! of a large matrix: the matrix streams IN chunk by chunk while the right-hand panels are already being factored,
    // and the finished panels stream OUT while the rest is still being factored -- both PCIe transfers disappear behind the
    // computation (except the first chunk and the last panel).  Three schedules, option "host_split":
    // "host_split" = 2 (default): the chunks join the LOOK-AHEAD factorization itself (blocked.cu): one ql! with the
    // device-resident schedule, the upload hidden behind it except for the first chunk, finished panels streaming home.
    // (1 = the split schedule below, 0 = the chunk-joining schedule without look-ahead: both kept for comparison;
    // n = 16384 from pinned memory: 251 / 274 / 284 ms on an idle host, 325-330 / 382 / 348 ms with 8 GPUs uploading at once)
    if (f == geqlf_dev_graph && h->opt_host_split == 2 && h->gc_ok && h->opt_lookahead == 2 && m >= n && n >= h->opt_host_split_min &&
        h->opt_tcache && qlb_lookahead_active(h, m, n)) {
        const int64_t ldd = (m + 1) & ~(int64_t)1;
        rc = qlb_reserve(h, &h->stage, &h->stage_bytes, (size_t)ldd * n * sizeof(double));
        if (!rc) rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * sizeof(double));
        if (rc) return rc;
        double *dA = (double *)h->stage, *dtau = (double *)h->stage2;
        // everything up to "all panels are home" on the handle's streams (the copy stream joins the main stream at the end)
        auto enqueue = [&]() -> int {
            h->h2d.host = A; h->h2d.host_ld = lda; h->h2d.la = true;
            h->d2h.host = A; h->d2h.host_ld = lda; h->d2h.la = true; h->d2h.rows = m;
            h->up_bytes = 0.0;
            int r = qlb_dgeqlf_dev(h, m, n, dA, ldd, dtau);
            h->h2d.host = nullptr; h->h2d.la = false;
            h->d2h.host = nullptr; h->d2h.la = false; h->d2h.rows = 0;
            if (r) return r;
            QLB_CUDA(h, cudaEventRecord(h->ev_home, h->copy_stream));
            QLB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_home, 0));
            return 0;
        };
        auto finish = [&]() -> int {
            QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            QLB_CUDA(h, cudaStreamSynchronize(h->stream));
            return 0;
        };
        auto plain = [&]() -> int {
            int r = enqueue();
            if (r) {
                cudaStreamSynchronize(h->h2d_stream);
                cudaStreamSynchronize(h->copy_stream);
                return r;
            }
            r = finish();
            if (!r) note_upload_rate(h);
            return r;
        };
        // The same call again (same pinned host matrix, same options): the whole schedule -- uploads, both partitions' kernels,
        // downloads -- is replayed as ONE CUDA graph; ~1000 launches on two streams otherwise cost ~11 ms of a 262 ms call.
        cudaPointerAttributes pa;
        bool pinned = cudaPointerGetAttributes(&pa, A) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        (void)cudaGetLastError();
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        bool usable = pinned && h->opt_graph && h->opt_la_graph && !h->opt_profile && h->stream != nullptr;
        if (usable && (cudaStreamIsCapturing(h->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)) {
            (void)cudaGetLastError();
            usable = false;
        }
        if (!usable) return plain();
        auto make_key = [&](int64_t *key) {
            const int64_t v[14] = {m, n, (int64_t)(uintptr_t)A, lda, (int64_t)(uintptr_t)dA, (int64_t)(uintptr_t)dtau, (int64_t)qlb_effective_nb(h, m, n),
                                   h->opt_strip_rpt * 16 + h->opt_lookahead, h->opt_gemm_cfg * 16 + h->opt_trail_splits, (int64_t)(uintptr_t)h->stream,
                                   1 + 2 * h->opt_overlap_t + 4 * h->opt_tcache + 32 * h->opt_la_gemm_all, h->opt_chunk_cols,
                                   h->opt_la_join_ramp * 64 + h->opt_la_join_chunks, (int64_t)(uintptr_t)h->tcache};
            memcpy(key, v, sizeof(v));
        };
        int64_t key[14];
        make_key(key);
        auto mark_cache = [&]() -> int {       // the replay refills the T cache of exactly this factorization
            if (h->opt_tcache && h->tcache) {
                h->tc_key[0] = (int64_t)(uintptr_t)dA; h->tc_key[1] = ldd; h->tc_key[2] = (int64_t)(uintptr_t)dtau;
                h->tc_key[3] = m; h->tc_key[4] = k; h->tc_key[5] = qlb_effective_nb(h, m, n);
                h->tc_valid = true;
                QLB_CUDA(h, cudaEventRecord(h->ev_tc, h->stream));
            }
            return 0;
        };
        if (h->hgraph_exec && !memcmp(key, h->hgraph_key, sizeof(key))) {
            QLB_CUDA(h, cudaGraphLaunch(h->hgraph_exec, h->stream));
            h->launches += h->hgraph_nodes;
            rc = mark_cache();
            return rc ? rc : finish();
        }
        if (memcmp(key, h->hgraph_seen_key, sizeof(key))) {          // first sight: plain launches (allocations, attributes, upload rate)
            rc = plain();
            if (!rc) make_key(h->hgraph_seen_key);                  // (the T cache may have been allocated by this very call)
            return rc;
        }
        if (h->hgraph_exec) {
            cudaGraphExecDestroy(h->hgraph_exec);
            h->hgraph_exec = nullptr;
        }
        const int64_t l0 = h->launches;
        QLB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue();
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (rc || e != cudaSuccess || !graph) {
            (void)cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            h->launches = l0;
            if (rc) return rc;
            return plain();                                          // capture refused
        }
        e = cudaGraphInstantiate(&h->hgraph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            h->hgraph_exec = nullptr;
            h->launches = l0;
            return plain();
        }
        h->hgraph_nodes = h->launches - l0;
        memcpy(h->hgraph_key, key, sizeof(key));
        QLB_CUDA(h, cudaGraphLaunch(h->hgraph_exec, h->stream));
        rc = mark_cache();
        return rc ? rc : finish();
    }
    // "host_split" = 1, SPLIT schedule.  The right quarter of the columns is uploaded and factored first (look-ahead ql! of the
    // m x wR block) while the rest is still on its way; then Q_R' is applied to the left part (blocked ormql with the cached T
    // factors, all SMs), then the leading (m - wR) x wL block is factored (look-ahead again).  Same arithmetic as one ql!
    // (LAPACK's right-to-left order restricted to column blocks); finished panels stream home in both factorizations.
    const bool split_now = h->opt_host_split == 1;
    if (f == geqlf_dev_graph && split_now && h->gc_ok && h->opt_lookahead == 2 && m >= n && n >= h->opt_host_split_min && h->opt_tcache &&
        (h->opt_host_split_chunk == 0 || n <= 31 * h->opt_host_split_chunk)) {      // (one event per chunk)
        const int64_t ldd = (m + 1) & ~(int64_t)1;
        rc = qlb_reserve(h, &h->stage, &h->stage_bytes, (size_t)ldd * n * sizeof(double));
        if (!rc) rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * sizeof(double));
        if (rc) return rc;
        double *dA = (double *)h->stage, *dtau = (double *)h->stage2;
        int64_t wR = ((n / 4 + 127) / 128) * 128;
        const int64_t wL = n - wR;
        QLB_CUDA(h, cudaEventRecord(h->ev_up[0], h->h2d_stream));
        QLB_CUDA(h, cudaMemcpy2DAsync(dA + wL * ldd, (size_t)ldd * 8, A + wL * lda, (size_t)lda * 8, (size_t)m * 8, (size_t)wR,
                                      cudaMemcpyHostToDevice, h->h2d_stream));
        QLB_CUDA(h, cudaEventRecord(h->ev_up[1], h->h2d_stream));
        h->up_bytes = (double)m * 8.0 * (double)wR;
        QLB_CUDA(h, cudaEventRecord(h->ev_chunk[0], h->h2d_stream));
        // the left part follows in column chunks, right to left: Q_R' is applied to a chunk as soon as it has arrived, so a slow
        // upload (several GPUs sharing the host's memory system) hides behind the applies instead of stalling them
        const int64_t CW = h->opt_host_split_chunk > 0 ? ((h->opt_host_split_chunk + 127) / 128) * 128 : wL;
        const int nchunks = (int)((wL + CW - 1) / CW);
        for (int c = 0; c < nchunks; ++c) {
            const int64_t hi = wL - (int64_t)c * CW, lo = hi - CW > 0 ? hi - CW : 0;
            QLB_CUDA(h, cudaMemcpy2DAsync(dA + lo * ldd, (size_t)ldd * 8, A + lo * lda, (size_t)lda * 8, (size_t)m * 8, (size_t)(hi - lo),
                                          cudaMemcpyHostToDevice, h->h2d_stream));
            QLB_CUDA(h, cudaEventRecord(h->ev_chunk[1 + c], h->h2d_stream));
        }
        QLB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[0], 0));
        h->d2h.la = true;
        h->d2h.rows = m;
        h->d2h.host_ld = lda;
        h->d2h.host = A + wL * lda;
        rc = qlb_dgeqlf_dev(h, m, wR, dA + wL * ldd, ldd, dtau + wL);
        h->d2h.host = nullptr;
        for (int c = 0; c < nchunks && !rc; ++c) {
            const int64_t hi = wL - (int64_t)c * CW, lo = hi - CW > 0 ? hi - CW : 0;
            QLB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[1 + c], 0));
            rc = qlb_dormql_dev(h, 'L', 'T', m, hi - lo, wR, dA + wL * ldd, ldd, dtau + wL, dA + lo * ldd, ldd, 2);
        }
        if (!rc) {
            h->d2h.host = A;
            rc = qlb_dgeqlf_dev(h, m - wR, wL, dA, ldd, dtau);
            h->d2h.host = nullptr;
        }

        h->d2h.la = false;
        h->d2h.rows = 0;
        h->tc_valid = false;                 // the cache now holds the T factors of the leading block only
        if (rc) {
            cudaStreamSynchronize(h->h2d_stream);
            cudaStreamSynchronize(h->copy_stream);
            return rc;
        }
        QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        QLB_CUDA(h, cudaStreamSynchronize(h->stream));
        QLB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
        note_upload_rate(h);
        return 0;
    }
    const bool stream_back = (f == geqlf_dev_graph) && m * n >= ((int64_t)1 << 22) && h->opt_lookahead != 1;
    const bool stream_in = stream_back && h->opt_stream_in && h->opt_tcache;
    double *dA;
    int64_t ldd;
    if (stream_in) {
        ldd = (m + 1) & ~(int64_t)1;
        rc = qlb_reserve(h, &h->stage, &h->stage_bytes, (size_t)ldd * n * sizeof(double));
        if (rc) return rc;
        dA = (double *)h->stage;
    } else {
        rc = stage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, &dA, &ldd);
        if (rc) return rc;
    }
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * sizeof(double));
    if (rc) return rc;
    double *dtau = (double *)h->stage2;
    if (stream_back) {
        h->d2h.host = A;
        h->d2h.host_ld = lda;
        if (stream_in) {
            h->h2d.host = A;
            h->h2d.host_ld = lda;
            h->up_bytes = 0.0;              // (set by the driver once the first chunk is on its way)
        }
        rc = qlb_dgeqlf_dev(h, m, n, dA, ldd, dtau);
        h->d2h.host = nullptr;
        h->h2d.host = nullptr;
        if (rc) {
            cudaStreamSynchronize(h->h2d_stream);
            cudaStreamSynchronize(h->copy_stream);
            return rc;
        }
        QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        QLB_CUDA(h, cudaStreamSynchronize(h->stream));
        QLB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
        if (stream_in) note_upload_rate(h);
    } else {
        rc = f(h, m, n, dA, ldd, dtau);
        if (rc) return rc;
        rc = stage_out(h, A, m, n, lda, dA, ldd);
        if (rc) return rc;
        QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

static int check_ormql(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                       const double *tau, const double *C, int64_t ldc) {
    if (!h) return -1;
    const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
    QLB_REQUIRE(h, left || right, 2, "side must be 'L' or 'R'");
    QLB_REQUIRE(h, trans == 'N' || trans == 'n' || trans == 'T' || trans == 't' || trans == 'C' || trans == 'c', 3,
                "trans must be 'N' or 'T'");
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 4, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 5, "n out of range");
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, A != nullptr || k == 0, 7, "A is null");
    QLB_REQUIRE(h, lda >= (nq > 1 ? nq : 1), 8, "lda < order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, tau != nullptr || k == 0, 9, "tau is null");
    QLB_REQUIRE(h, C != nullptr || m * n == 0, 10, "C is null");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    return 0;
}

extern "C" {

int qlb200_dgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau) {
    int rc = check_fact(h, m, n, dA, lda, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return geqlf_dev_graph(h, m, n, dA, lda, dtau);
}
int qlb200_dgeqlf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau) {
    return fact_host(h, geqlf_dev_graph, m, n, A, lda, tau);
}
int qlb200_dgerqf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau) {
    int rc = check_fact(h, m, n, dA, lda, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dgerqf_dev(h, m, n, dA, lda, dtau);
}
int qlb200_dgerqf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau) {
    return fact_host(h, qlb_dgerqf_dev, m, n, A, lda, tau);
}

int qlb200_dormql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *dA,
                      int64_t lda, const double *dtau, double *dC, int64_t ldc) {
    int rc = check_ormql(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dormql_dev(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc, 0);
}
int qlb200_dormql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                  const double *tau, double *C, int64_t ldc) {
    int rc = check_ormql(h, side, trans, m, n, k, A, lda, tau, C, ldc);
    if (rc) return rc;
    if (m == 0 || n == 0 || k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    const int64_t nq = (side == 'L' || side == 'l') ? m : n;
    double *dA, *dC;
    int64_t ldda, lddc;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, nq, k, lda, &dA, &ldda);
    if (rc) return rc;
    // C and tau share the second staging buffer
    const int64_t ldc2 = ((m + 1) & ~(int64_t)1) < 2 ? 2 : ((m + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)ldc2 * n + (size_t)k + 2) * sizeof(double));
    if (rc) return rc;
    dC = (double *)h->stage2;
    lddc = ldc2;
    double *dtau = dC + (size_t)ldc2 * n;
    QLB_CUDA(h, cudaMemcpy2DAsync(dC, (size_t)lddc * 8, C, (size_t)ldc * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)k * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dormql_dev(h, side, trans, m, n, k, dA, ldda, dtau, dC, lddc, h->opt_tcache == 1 ? 0 : -1);      // the staged copy is validated by checksum (never by its address)
    if (rc) return rc;
    rc = stage_out(h, C, m, n, ldc, dC, lddc);
    if (rc) return rc;
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int check_solve(qlb200_ctx *h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau,
                       const double *B, int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 2, "n out of range");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 30), 3, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (n > 1 ? n : 1), 5, "lda < max(1,n)");
    QLB_REQUIRE(h, tau != nullptr || n == 0, 6, "tau is null");
    QLB_REQUIRE(h, B != nullptr || n * nrhs == 0, 7, "B is null");
    QLB_REQUIRE(h, ldb >= (n > 1 ? n : 1), 8, "ldb < max(1,n) (DimensionMismatch)");
    return 0;
}

int qlb200_dqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const double *dtau,
                        double *dB, int64_t ldb) {
    int rc = check_solve(h, n, nrhs, dA, lda, dtau, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dqlsolve_dev(h, n, nrhs, dA, lda, dtau, dB, ldb, nullptr, 0);
}
int qlb200_dqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                    int64_t ldb) {
    int rc = check_solve(h, n, nrhs, A, lda, tau, B, ldb);
    if (rc) return rc;
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, n, n, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddb = ((n + 1) & ~(int64_t)1) < 2 ? 2 : ((n + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddb * nrhs + (size_t)n + 2) * sizeof(double));
    if (rc) return rc;
    dB = (double *)h->stage2;
    double *dtau = dB + (size_t)lddb * nrhs;
    QLB_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)lddb * 8, B, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dqlsolve_dev(h, n, nrhs, dA, ldda, dtau, dB, lddb, h->dinfo, h->opt_tcache == 1 ? 0 : -1);
    if (rc) return rc;
    rc = stage_out(h, B, n, nrhs, ldb, dB, lddb);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;       // i > 0: L[i,i] == 0 (SingularException)
}


// ---- the tall least-squares solve the reference's own branch gets wrong (@test_broken, test/test_ql.jl:148).
// A = Q [0; L] (m >= n), so min ||A x - b|| has x = L^{-1} (Q'b)[m-n+1:m]; the first m-n entries of Q'b are the residual's coordinates.
static int check_lstsq(qlb200_ctx *h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau,
                       const double *B, int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n <= m, 3, "n must satisfy 0 <= n <= m (tall or square)");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 30), 4, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 5, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 6, "lda < max(1,m)");
    QLB_REQUIRE(h, tau != nullptr || n == 0, 7, "tau is null");
    QLB_REQUIRE(h, B != nullptr || m * nrhs == 0, 8, "B is null");
    QLB_REQUIRE(h, ldb >= (m > 1 ? m : 1), 9, "ldb < max(1,m) (DimensionMismatch)");
    return 0;
}
static int qllstsq_dev(qlb200_ctx *h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const double *dtau,
                       double *dB, int64_t ldb, int32_t *dinfo, int cache_mode) {
    if (n == 0 || nrhs == 0) return 0;
    int rc = qlb_dormql_dev(h, 'L', 'T', m, nrhs, n, dA, lda, dtau, dB, ldb, cache_mode);
    if (rc) return rc;
    rc = qlb_trsm_lower_dev(h, (int)n, (int)nrhs, dA + (m - n), lda, dB + (m - n), ldb, dinfo);
    if (rc || m == n) return rc;
    // The reference's `\` returns X[1:n, :] of the buffer ldiv! worked on (src/ql.jl:496-501, _cut_B(X, 1:n)), the ?gels
    // convention: rotate every column so that the solution comes first and the residual's coordinates follow.
    return qlb_rotate_rows_dev(h, m, nrhs, dB, ldb, m - n);
}
// wide minimum-norm solve (m < n); B is n x nrhs
static int check_minnorm(qlb200_ctx *h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau,
                         const double *B, int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 2, "m out of range");
    QLB_REQUIRE(h, n >= m && n < (1 << 30), 3, "minimum-norm solve needs m <= n");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 30), 4, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 5, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 6, "lda < max(1,m)");
    QLB_REQUIRE(h, tau != nullptr || m == 0, 7, "tau is null");
    QLB_REQUIRE(h, B != nullptr || n * nrhs == 0, 8, "B is null");
    QLB_REQUIRE(h, ldb >= (n > 1 ? n : 1), 9, "ldb < max(1,n) (DimensionMismatch)");
    return 0;
}
int qlb200_dqlminnorm_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const double *dtau,
                          double *dB, int64_t ldb) {
    int rc = check_minnorm(h, m, n, nrhs, dA, lda, dtau, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dqlminnorm_dev(h, m, n, nrhs, dA, lda, dtau, dB, ldb, nullptr);
}
int qlb200_dqlminnorm(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                      int64_t ldb) {
    int rc = check_minnorm(h, m, n, nrhs, A, lda, tau, B, ldb);
    if (rc) return rc;
    if (m == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddb = ((n + 1) & ~(int64_t)1) < 2 ? 2 : ((n + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddb * nrhs + (size_t)m + 2) * sizeof(double));
    if (rc) return rc;
    dB = (double *)h->stage2;
    double *dtau = dB + (size_t)lddb * nrhs;
    QLB_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)lddb * 8, B, (size_t)ldb * 8, (size_t)m * 8, (size_t)nrhs, cudaMemcpyHostToDevice, h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dqlminnorm_dev(h, m, n, nrhs, dA, ldda, dtau, dB, lddb, h->dinfo);
    if (rc) return rc;
    rc = stage_out(h, B, n, nrhs, ldb, dB, lddb);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;       // i > 0: rank deficient
}
int qlb200_dqllstsq_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const double *dtau,
                        double *dB, int64_t ldb) {
    int rc = check_lstsq(h, m, n, nrhs, dA, lda, dtau, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qllstsq_dev(h, m, n, nrhs, dA, lda, dtau, dB, ldb, nullptr, 0);
}
int qlb200_dqllstsq(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                    int64_t ldb) {
    int rc = check_lstsq(h, m, n, nrhs, A, lda, tau, B, ldb);
    if (rc) return rc;
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddb = ((m + 1) & ~(int64_t)1) < 2 ? 2 : ((m + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddb * nrhs + (size_t)n + 2) * sizeof(double));
    if (rc) return rc;
    dB = (double *)h->stage2;
    double *dtau = dB + (size_t)lddb * nrhs;
    QLB_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)lddb * 8, B, (size_t)ldb * 8, (size_t)m * 8, (size_t)nrhs, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qllstsq_dev(h, m, n, nrhs, dA, ldda, dtau, dB, lddb, h->dinfo, h->opt_tcache == 1 ? 0 : -1);
    if (rc) return rc;
    rc = stage_out(h, B, m, nrhs, ldb, dB, lddb);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;       // i > 0: L[i,i] == 0 (rank deficient)
}


int qlb200_dormrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *dA, int64_t lda,
                      const double *dtau, double *dC, int64_t ldc) {
    if (!h) return -1;
    const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
    QLB_REQUIRE(h, left || right, 2, "side must be 'L' or 'R'");
    QLB_REQUIRE(h, trans == 'N' || trans == 'n' || trans == 'T' || trans == 't' || trans == 'C' || trans == 'c', 3,
                "trans must be 'N' or 'T'");
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 4, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 5, "n out of range");
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, dA != nullptr || k == 0, 7, "A is null");
    QLB_REQUIRE(h, lda >= (k > 1 ? k : 1), 8, "lda < max(1,k)");
    QLB_REQUIRE(h, dtau != nullptr || k == 0, 9, "tau is null");
    QLB_REQUIRE(h, dC != nullptr || m * n == 0, 10, "C is null");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dormrq_dev(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
}
int qlb200_dormrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                  const double *tau, double *C, int64_t ldc) {
    if (!h) return -1;
    if (m == 0 || n == 0 || k == 0) return qlb200_dormrq_dev(h, side, trans, m, n, k, A, lda, tau, C, ldc);   // checks only
    const bool left = side == 'L' || side == 'l';
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, (left || side == 'R' || side == 'r'), 2, "side must be 'L' or 'R'");
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, lda >= (k > 1 ? k : 1), 8, "lda < max(1,k)");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dC;
    int64_t ldda;
    int rc = stage_in(h, &h->stage, &h->stage_bytes, A, k, nq, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddc = ((m + 1) & ~(int64_t)1) < 2 ? 2 : ((m + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddc * n + (size_t)k + 2) * sizeof(double));
    if (rc) return rc;
    dC = (double *)h->stage2;
    double *dtau = dC + (size_t)lddc * n;
    QLB_CUDA(h, cudaMemcpy2DAsync(dC, (size_t)lddc * 8, C, (size_t)ldc * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)k * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb200_dormrq_dev(h, side, trans, m, n, k, dA, ldda, dtau, dC, lddc);
    if (rc) return rc;
    rc = stage_out(h, C, m, n, ldc, dC, lddc);
    if (rc) return rc;
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* ---- QR through index reversal -------------------------------------------------- */
// geqrf(A) = J geqlf(J A J) J, tau reversed; the QL in the middle is the same call as ql! (graph replay on repeated calls)
static int geqrf_dev(qlb200_ctx *h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau) {
    int rc = qlb_qr_flip_in(h, m, n, dA, lda);
    if (!rc) rc = geqlf_dev_graph(h, m, n, dA, lda, dtau);
    if (!rc) rc = qlb_qr_flip_out(h, m, n, dA, lda, dtau);
    return rc;
}
int qlb200_dgeqrf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau) {
    int rc = check_fact(h, m, n, dA, lda, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return geqrf_dev(h, m, n, dA, lda, dtau);
}
int qlb200_dgeqrf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau) {
    return fact_host(h, geqrf_dev, m, n, A, lda, tau);
}
int qlb200_dormqr_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *dA, int64_t lda,
                      const double *dtau, double *dC, int64_t ldc) {
    int rc = check_ormql(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dormqr_dev(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
}
int qlb200_dormqr(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                  const double *tau, double *C, int64_t ldc) {
    int rc = check_ormql(h, side, trans, m, n, k, A, lda, tau, C, ldc);
    if (rc) return rc;
    if (m == 0 || n == 0 || k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    const int64_t nq = (side == 'L' || side == 'l') ? m : n;
    double *dA, *dC;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, nq, k, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddc = ((m + 1) & ~(int64_t)1) < 2 ? 2 : ((m + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddc * n + (size_t)k + 2) * sizeof(double));
    if (rc) return rc;
    dC = (double *)h->stage2;
    double *dtau = dC + (size_t)lddc * n;
    QLB_CUDA(h, cudaMemcpy2DAsync(dC, (size_t)lddc * 8, C, (size_t)ldc * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)k * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dormqr_dev(h, side, trans, m, n, k, dA, ldda, dtau, dC, lddc);
    if (rc) return rc;
    rc = stage_out(h, C, m, n, ldc, dC, lddc);
    if (rc) return rc;
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int qlb200_dqrsolve_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const double *dtau,
                        double *dB, int64_t ldb) {
    int rc = check_lstsq(h, m, n, nrhs, dA, lda, dtau, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dqrsolve_dev(h, m, n, nrhs, dA, lda, dtau, dB, ldb, nullptr);
}
int qlb200_dqrsolve(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                    int64_t ldb) {
    int rc = check_lstsq(h, m, n, nrhs, A, lda, tau, B, ldb);
    if (rc) return rc;
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddb = ((m + 1) & ~(int64_t)1) < 2 ? 2 : ((m + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddb * nrhs + (size_t)n + 2) * sizeof(double));
    if (rc) return rc;
    dB = (double *)h->stage2;
    double *dtau = dB + (size_t)lddb * nrhs;
    QLB_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)lddb * 8, B, (size_t)ldb * 8, (size_t)m * 8, (size_t)nrhs, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dqrsolve_dev(h, m, n, nrhs, dA, ldda, dtau, dB, lddb, h->dinfo);
    if (rc) return rc;
    rc = stage_out(h, B, m, nrhs, ldb, dB, lddb);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    // the zero was found on the diagonal of L = J R J: report R's index
    return info == 0x7fffffff ? 0 : (int)(n + 1 - info);
}

/* ---- UL ------------------------------------------------------------------------------------------- */
static int check_ul(qlb200_ctx *h, int64_t m, int64_t n, const double *A, int64_t lda, const int64_t *ipiv) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 3, "n out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 5, "lda < max(1,m)");
    QLB_REQUIRE(h, ipiv != nullptr || m * n == 0, 6, "ipiv is null");
    return 0;
}
int qlb200_dulfact_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, int64_t *dipiv, int pivot,
                       int32_t *dinfo_singular) {
    int rc = check_ul(h, m, n, dA, lda, dipiv);
    if (rc) return rc;
    QLB_REQUIRE(h, dinfo_singular != nullptr, 8, "info_singular is null");
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dulfact_dev(h, m, n, dA, lda, dipiv, pivot, dinfo_singular);
}
int qlb200_dulfact(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, int64_t *ipiv, int pivot,
                   int32_t *info_singular) {
    int rc = check_ul(h, m, n, A, lda, ipiv);
    if (rc) return rc;
    if (info_singular) *info_singular = 0;
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA;
    int64_t ldd;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, &dA, &ldd);
    if (rc) return rc;
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * sizeof(int64_t));
    if (rc) return rc;
    int64_t *dipiv = (int64_t *)h->stage2;
    rc = qlb_dulfact_dev(h, m, n, dA, ldd, dipiv, pivot, h->dinfo);
    if (rc) return rc;
    rc = stage_out(h, A, m, n, lda, dA, ldd);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpyAsync(ipiv, dipiv, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (info_singular) *info_singular = info;
    return 0;
}
static int check_ulsolve(qlb200_ctx *h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const int64_t *ipiv, const double *B,
                         int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 2, "n out of range");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 30), 3, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (n > 1 ? n : 1), 5, "lda < max(1,n)");
    QLB_REQUIRE(h, ipiv != nullptr || n == 0, 6, "ipiv is null");
    QLB_REQUIRE(h, B != nullptr || n * nrhs == 0, 7, "B is null");
    QLB_REQUIRE(h, ldb >= (n > 1 ? n : 1), 8, "ldb < max(1,n) (DimensionMismatch)");
    return 0;
}
int qlb200_dulsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const double *dA, int64_t lda, const int64_t *dipiv, double *dB,
                        int64_t ldb) {
    int rc = check_ulsolve(h, n, nrhs, dA, lda, dipiv, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_dulsolve_dev(h, n, nrhs, dA, lda, dipiv, dB, ldb, nullptr);
}
int qlb200_dulsolve(qlb200_handle h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const int64_t *ipiv, double *B,
                    int64_t ldb) {
    int rc = check_ulsolve(h, n, nrhs, A, lda, ipiv, B, ldb);
    if (rc) return rc;
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    int64_t ldda;
    rc = stage_in(h, &h->stage, &h->stage_bytes, A, n, n, lda, &dA, &ldda);
    if (rc) return rc;
    const int64_t lddb = ((n + 1) & ~(int64_t)1) < 2 ? 2 : ((n + 1) & ~(int64_t)1);
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, ((size_t)lddb * nrhs + (size_t)n + 2) * sizeof(double));
    if (rc) return rc;
    dB = (double *)h->stage2;
    int64_t *dipiv = (int64_t *)(dB + (size_t)lddb * nrhs);
    QLB_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)lddb * 8, B, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, cudaMemcpyHostToDevice,
                                  h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(dipiv, ipiv, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_dulsolve_dev(h, n, nrhs, dA, ldda, dipiv, dB, lddb, h->dinfo);
    if (rc) return rc;
    rc = stage_out(h, B, n, nrhs, ldb, dB, lddb);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;       // i > 0: L[i,i] == 0 (SingularException)
}

/* ---- batched work sharded over several devices (host pointers) --------------------------------------
 * Matrices are independent: handle g of G (each bound to its own device) gets the contiguous index range
 * [g*batch/G, (g+1)*batch/G) (+ remainder on the first ranks), one host thread per device, no data-path
 * collective.  Returns the first non-zero info of any shard. */
static void shard_range(int64_t total, int g, int G, int64_t *lo, int64_t *hi) {
    const int64_t base = total / G, rem = total % G;
    *lo = g * base + (g < rem ? g : rem);
    *hi = *lo + base + (g < rem ? 1 : 0);
}
}  // extern "C"
template <typename F>
static int run_sharded(qlb200_handle *hs, int nh, int64_t batch, F fn) {
    if (!hs || nh < 1) return -1;
    for (int g = 0; g < nh; ++g)
        if (!hs[g]) return -1;
    std::vector<int> rcs((size_t)nh, 0);
    std::vector<std::thread> th;
    for (int g = 0; g < nh; ++g) {
        int64_t lo, hi;
        shard_range(batch, g, nh, &lo, &hi);
        th.emplace_back([&, g, lo, hi]() { rcs[(size_t)g] = (hi > lo) ? fn(hs[g], lo, hi) : 0; });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < nh; ++g)
        if (rcs[(size_t)g]) return rcs[(size_t)g];
    return 0;
}

extern "C" {
#define QLB_GEQLF_MG(NAME, T, SINGLE)                                                                              \
    int NAME(qlb200_handle *hs, int nh, int64_t n, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau,   \
             int64_t batch) {                                                                                       \
        return run_sharded(hs, nh, batch, [&](qlb200_handle h, int64_t lo, int64_t hi) {                            \
            return SINGLE(h, n, A + lo * strideA, lda, strideA, tau + lo * strideTau, strideTau, hi - lo);          \
        });                                                                                                         \
    }
QLB_GEQLF_MG(qlb200_dgeqlf_batched_mg, double, qlb200_dgeqlf_batched)
QLB_GEQLF_MG(qlb200_sgeqlf_batched_mg, float, qlb200_sgeqlf_batched)

#define QLB_ORMQL_MG(NAME, T, SINGLE)                                                                              \
    int NAME(qlb200_handle *hs, int nh, char trans, int64_t n, int64_t nrhs, const T *A, int64_t lda,              \
             int64_t strideA, const T *tau, int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch) { \
        return run_sharded(hs, nh, batch, [&](qlb200_handle h, int64_t lo, int64_t hi) {                            \
            return SINGLE(h, trans, n, nrhs, A + lo * strideA, lda, strideA, tau + lo * strideTau, strideTau,       \
                          B + lo * strideB, ldb, strideB, hi - lo);                                                 \
        });                                                                                                         \
    }
QLB_ORMQL_MG(qlb200_dormql_batched_mg, double, qlb200_dormql_batched)
QLB_ORMQL_MG(qlb200_sormql_batched_mg, float, qlb200_sormql_batched)

#define QLB_QLSOLVE_MG(NAME, T, SINGLE)                                                                            \
    int NAME(qlb200_handle *hs, int nh, int64_t n, int64_t nrhs, const T *A, int64_t lda, int64_t strideA,         \
             const T *tau, int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int64_t batch,                    \
             int32_t *info_out) {                                                                                   \
        return run_sharded(hs, nh, batch, [&](qlb200_handle h, int64_t lo, int64_t hi) {                            \
            return SINGLE(h, n, nrhs, A + lo * strideA, lda, strideA, tau + lo * strideTau, strideTau,              \
                          B + lo * strideB, ldb, strideB, hi - lo, info_out ? info_out + lo : nullptr);             \
        });                                                                                                         \
    }
QLB_QLSOLVE_MG(qlb200_dqlsolve_batched_mg, double, qlb200_dqlsolve_batched)
QLB_QLSOLVE_MG(qlb200_sqlsolve_batched_mg, float, qlb200_sqlsolve_batched)

// BLAS-style test/benchmark hook for the DMMA GEMM (column-major, device pointers):
// C = alpha*op(A)*op(B) + beta*C, op = 'N' or 'T'.
int qlb200_dgemm_dev(qlb200_handle h, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                     const double *dA, int64_t lda, const double *dB, int64_t ldb, double beta, double *dC, int64_t ldc) {
    if (!h) return -1;
    const bool ta = (transa == 'T' || transa == 't'), tb = (transb == 'T' || transb == 't');
    QLB_REQUIRE(h, ta || transa == 'N' || transa == 'n', 2, "transa must be 'N' or 'T'");
    QLB_REQUIRE(h, tb || transb == 'N' || transb == 'n', 3, "transb must be 'N' or 'T'");
    QLB_REQUIRE(h, m >= 0 && n >= 0 && k >= 0, 4, "negative dimension");
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_gemm(h, h->stream, -1, /*akc=*/ta, /*bkc=*/!tb, (int)m, (int)n, (int)k, alpha, dA, lda, dB, ldb, beta, dC, ldc,
                    1, 0);
}

}  // extern "C"

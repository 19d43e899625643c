// common.cuh -- handle, error plumbing and sm_100a PTX helpers (mbarrier, TMA bulk copies, DMMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/qlb200.h"

#define QLB200_VERSION 1000

struct qlb200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;     // created with the handle
    cudaStream_t stream = nullptr;         // stream `_dev` calls run on
    cudaStream_t side_stream = nullptr;    // look-ahead panel stream
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // host-pointer ql!: finished panels are copied back while the factorization of the rest is still running
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_panel[8] = {nullptr};   // ring of "panel j is final" events
    struct {
        double *host = nullptr;            // destination (column-major, ld = host_ld); null = disabled
        int64_t host_ld = 0;
        int64_t rows = 0;                  // rows copied per finished column (0: the m of the call; the split host path copies whole columns of the parent matrix)
        bool la = false;                   // the caller wants the look-ahead schedule (device-resident input, finished panels streamed out)
    } d2h;
    // host-pointer ql!: the matrix arrives in column chunks (right to left) while the panels on the right are already
    // being factored; a chunk joins the trailing update when it has arrived and caught up with the finished panels
    cudaStream_t h2d_stream = nullptr;
    cudaEvent_t ev_chunk[32] = {nullptr};
    cudaEvent_t ev_up[2] = {nullptr, nullptr};   // timing events around the first upload of a host-pointer ql! (host_split = 2 policy)
    double up_bytes = 0.0, h2d_gbps = 0.0;      // bytes between ev_up[0..1] of the call in flight; upload rate the last call saw (0: none yet)
    struct {
        const double *host = nullptr;      // source (column-major, ld = host_ld); null = the matrix is already on the device
        int64_t host_ld = 0;
        bool la = false;                   // chunks join the LOOK-AHEAD schedule (tall/square, SM partitions, T cache)
    } h2d;
    cudaGraphExec_t hgraph_exec = nullptr;  // host-pointer ql! (look-ahead joining schedule, pinned host memory) as a CUDA graph
    int64_t hgraph_key[14] = {0}, hgraph_seen_key[14] = {0}, hgraph_nodes = 0;
    cudaEvent_t ev_rest[2] = {nullptr, nullptr}, ev_join = nullptr;   // look-ahead: "rest update j done" (ring of 2), "late chunks joined"
    int64_t opt_la_pnarrow_min = 8192;     // look-ahead: the panel partition updates the next panel's columns itself while that panel starts at or right of this column
    cudaEvent_t ev_home = nullptr;         // copy stream -> main stream join at the end of a host-pointer ql!
    void *joinV = nullptr;                 // clean reflectors of an already factored panel while a late chunk catches up
    size_t joinV_bytes = 0;
    int64_t opt_host_split = 2;            // host-pointer ql!, large tall/square matrices: 2 chunks join the look-ahead ql!, 1 split schedule, 0 chunk-joining without look-ahead (capi.cu fact_host)
    int64_t opt_host_split_chunk = 0;      // ... left part uploaded and multiplied in chunks of this many columns (0: one piece)
    int64_t opt_host_split_min = 8192;     // ... from this number of columns on
    int64_t opt_stream_in = 1;
    int64_t opt_chunk_cols = 1024;         // columns per H2D chunk
    int64_t opt_batch_chunk_mb = 32;       // host-pointer batched calls: operand megabytes per pipeline chunk (0: no chunking)
    int64_t opt_la_join_ramp = 4, opt_la_join_chunks = 2;   // look-ahead host schedule: one chunk per panel for the first .. panels, then .. per panel
    int64_t opt_join_step = 1;             // panels factored per arriving chunk before the next chunk joins
    void *ws = nullptr;                    // device workspace (grown on demand)
    size_t ws_bytes = 0;
    void *ws2 = nullptr;                   // second workspace (transposed copy for RQ)
    size_t ws2_bytes = 0;
    void *stage = nullptr;                 // device staging for host-pointer entry points
    size_t stage_bytes = 0;
    void *stage2 = nullptr;                // second staging buffer (C / B operands)
    size_t stage2_bytes = 0;
    int32_t *dinfo = nullptr;              // device-side status word
    void *tall = nullptr;                  // scratch of the tall-strip fallback (partial dots, strip Gram matrix)
    size_t tall_bytes = 0;
    void *zws = nullptr;                   // complex RQ: the conjugate-transposed reflectors (complex.cu)
    size_t zws_bytes = 0;
    void *cvt[3] = {nullptr, nullptr, nullptr};   // Float32 / ComplexF32 entry points: the widened operands (mixed.cu)
    size_t cvt_bytes[3] = {0, 0, 0};
    void *redo = nullptr;                  // batched: counter + list of matrices for the generic kernel
    size_t redo_bytes = 0;
    int64_t launches = 0;
    int64_t opt_nb = 0, opt_strip = 16;     // panel width; 0 = auto (qlb_effective_nb: 192 for look-ahead ql! with min(m,n) >= 14336, else 128)
    // Look-ahead (panel j+1 is factored while the trailing update of panel j runs): 0 off, 1 on, 2 (default) = on for
    // device-resident ql! with min(m,n) >= 2048 when the SM partitions below exist.
    int64_t opt_lookahead = 2;
    int64_t opt_la_graph = 1;              // replay the look-ahead ql! (two streams joined by events) as a CUDA graph too
    int64_t opt_la_min = 4096;             // look-ahead (auto) from this min(m,n) on
    int64_t opt_la_gemm_all = 0;           // (experiment) look-ahead GEMMs on the caller's stream (all SMs) instead of the GEMM partition
    // Two CUDA green contexts split the SMs into a small partition for the latency-bound panel kernels (the strip kernel is
    // a 16-CTA cluster of register-filling CTAs: it can only start on 16 completely free SMs of one GPC) and the rest for the
    // trailing-update GEMMs.  Kernels launched into gc_gemm never occupy the panel SMs, so the panel kernels start at once.
    bool gc_ok = false;
    void *gc_ctx[2] = {nullptr, nullptr};  // CUgreenCtx
    cudaStream_t gc_panel = nullptr, gc_gemm = nullptr;
    int gc_panel_sms = 0, gc_gemm_sms = 0;
    cudaEvent_t ev_gc[4] = {nullptr};
    int64_t opt_gemm_cfg = -1;             // -1 auto, else forces a tile configuration (benchmarking)
    int64_t opt_strip_rpt = 1;             // minimum rows per thread in the strip kernel (1, 2 or 4); 1 = as many CTAs of the cluster as the rows fill (measured best at every n, tools/sweep_geqlf.py: 2048: 6.9 vs 8.1 ms with 4; 8192: 44.3 vs 47.4; 16384: 232 vs 235)
    // ql! replayed as a CUDA graph: the launch sequence of one (m, n, A, lda, tau, options, stream) call is captured the
    // second time it is seen and replayed afterwards (same kernels, same order: bitwise identical results)
    int64_t opt_graph = 1;
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_key[10] = {0};
    int64_t graph_seen_key[10] = {0};
    int64_t graph_nodes = 0;
    // T-factor cache: ql! keeps the nb x nb compact-WY T of every panel, so that lmul!/rmul!/ldiv! on the factors it just
    // produced do not rebuild them (SURVEY.md A.3 allows exactly this).  Option "tcache": 1 (default) = VALIDATED: ql! leaves
    // a 64-bit position-dependent checksum of ALL reflector columns and tau on the device; lmul!/rmul!/ldiv! recompute it for
    // the factors they are given (one pass over V, 8 bytes read back) and use the cache only on a match -- factors edited in
    // place, or other factors at the same addresses, rebuild T.  2 = TRUSTED (opt-in fast path): the cache is keyed on the
    // pointers alone; the caller promises not to modify the factors between the calls.  0 = off.
    int64_t opt_tcache = 1;
    void *tcache = nullptr;
    size_t tcache_bytes = 0;
    bool tc_valid = false;
    int64_t tc_key[6] = {0};               // reflector pointer (first of the last k columns), lda, tau, m, k, nb
    unsigned long long *tc_sum_dev = nullptr;   // [0] checksum left by ql!, [1] checksum of the factors under validation
    cudaEvent_t ev_tc = nullptr;           // recorded after ql! filled the cache (readers on other streams wait for it)
    int64_t opt_overlap_t = 1;             // build the panel's T on the side stream while W = C'V runs
    int64_t opt_pair_sync = 8;             // GEMMs with exactly two N-tiles: cluster of the pair, barrier every this many k-tiles (0: off)
    int64_t opt_ul_cluster = 1;            // UL strips factored by the cluster kernel (0: single-CTA kernel)
    // fused batched factor + apply on device pointers: matrices per chunk (0 = whole batch, the default).  Chunks small enough
    // for the factors to stay in L2 between the two kernels were measured and LOSE (262144 x 32 x 32 f64 + Q'B: 2.13 ms unchunked,
    // 2.39 / 2.71 / 3.05 ms with 32768 / 16384 / 8192-matrix chunks): neither kernel is HBM-bound, every launch pays a fill/drain.
    int64_t opt_fuse_chunk = 0;
    int64_t opt_bvar = 0;                  // batched kernel variant (benchmarking)
    int64_t opt_bgrid = 0;                 // batched kernels: cap on the number of CTAs (0: one resident wave; benchmarking)
    int64_t opt_trail_splits = 0;          // 0 auto, else split-K count of the trailing W = C'V GEMM
    bool strip_attr_done[8] = {false};
    bool larft_attr_done = false, trsm_attr_done = false, zlarft_attr_done = false, zstep_attr_done = false;
    bool ul_cluster_attr_done = false, ul_trsm_attr_done = false;
    bool gemm_attr_done[64] = {false};
    // batched kernels: dynamic-shared-memory attribute and occupancy, looked up once per (handle, kernel, size)
    struct KernelSetup {
        const void *fn;
        size_t smem;
        int occ;
    };
    KernelSetup ksetup[256] = {};
    int n_ksetup = 0;
    // optional device-side timing of the dominant kernels (trailing-update GEMMs): event pairs around
    // each launch, summed by qlb200_profile_collect (bench.py's roofline leg)
    int64_t opt_profile = 0;
    cudaEvent_t *prof_ev = nullptr;        // 2 * prof_cap events
    double *prof_flops = nullptr;
    int prof_cap = 0, prof_n = 0;
    char err[512] = {0};
};

// Brackets one launch with an event pair when profiling is on.
struct QlbProfScope {
    qlb200_ctx *h;
    cudaStream_t st;
    int slot;
    QlbProfScope(qlb200_ctx *h_, cudaStream_t st_, double flops) : h(h_), st(st_), slot(-1) {
        if (h->opt_profile && h->prof_n < h->prof_cap) {
            slot = h->prof_n++;
            h->prof_flops[slot] = flops;
            cudaEventRecord(h->prof_ev[2 * slot], st);
        }
    }
    ~QlbProfScope() {
        if (slot >= 0) cudaEventRecord(h->prof_ev[2 * slot + 1], st);
    }
};

static inline int qlb_fail(qlb200_ctx *h, int code, const char *fmt, const char *detail = "") {
    if (h) snprintf(h->err, sizeof(h->err), fmt, detail);
    return code;
}

#define QLB_CUDA(h, call)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            if (h) snprintf((h)->err, sizeof((h)->err), "%s:%d %s -> %s", __FILE__, __LINE__, \
                            #call, cudaGetErrorString(e__));                                  \
            return 1000 + (int)e__;                                                           \
        }                                                                                     \
    } while (0)

#define QLB_LAUNCH_CHECK(h)            \
    do {                               \
        (h)->launches++;               \
        QLB_CUDA(h, cudaGetLastError()); \
    } while (0)

// Grow-only device buffers owned by the handle.
static inline int qlb_reserve(qlb200_ctx *h, void **buf, size_t *have, size_t need) {
    if (need <= *have) return 0;
    if (h->graph_exec) {                   // a cached graph holds pointers into the buffers: drop it on any reallocation
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
    }
    memset(h->graph_key, 0, sizeof(h->graph_key));
    memset(h->graph_seen_key, 0, sizeof(h->graph_seen_key));
    if (h->hgraph_exec) {                  // ... and so does the host-pointer schedule's graph
        cudaGraphExecDestroy(h->hgraph_exec);
        h->hgraph_exec = nullptr;
    }
    memset(h->hgraph_key, 0, sizeof(h->hgraph_key));
    memset(h->hgraph_seen_key, 0, sizeof(h->hgraph_seen_key));
    if (buf == &h->tcache) h->tc_valid = false;
    if (*buf) {
        QLB_CUDA(h, cudaStreamSynchronize(h->stream));
        QLB_CUDA(h, cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    size_t sz = (need + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    QLB_CUDA(h, cudaMalloc(buf, sz));
    *have = sz;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// device-side helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace qlb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Ampere-style 16-byte async copy global -> shared (SASS: LDGSTS), zero-fill when !pred.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool pred) {
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src, bool pred) {
    int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// FP64 tensor-core MMA (DMMA): D(8x8) += A(8x4, row) * B(4x8, col).
// Fragment layout (g = lane>>2, t = lane&3): a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// m16n8k8 (sm_90+): a0=A[g][t] a1=A[g+8][t] a2=A[g][t+4] a3=A[g+8][t+4]; b0=B[t][g] b1=B[t+4][g];
// c0,c1 = C[g][2t,2t+1]; c2,c3 = C[g+8][2t,2t+1].
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
          "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

}  // namespace qlb
#endif

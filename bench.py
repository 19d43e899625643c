#!/usr/bin/env python
"""bench.py -- headline benchmark of the QL hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one synthetic batch:
  * workload (every N): `ql!` of a 16384 x 16384 Float64 randn matrix (BASELINE.json configs[1]) --
    metric = GFLOP/s with the 4/3*n^3 flop count.  A single large matrix does not shard
    (north_star: "runs on 1 GPU only"), so --gpus N runs N independent replicas, one process per
    GPU ("replicas only"); value = total GFLOP/s over all replicas, weak scaling.
  * `batched` (extra object on the same line): 262144 x (32 x 32) Float64 QL (configs[2]) sharded by
    contiguous matrix-index ranges over the N ranks, no data-path collective; matrices/s over all ranks.

`value` is device-timed with the input resident in HBM (each step restores the matrix from a second
device buffer -- the reference's `ql(A)` = copy + ql!, src/ql.jl:181-186 -- then factors in place;
2 GiB per matrix, far larger than the 126 MB L2).  `e2e` is the same metric through the public host
API (NumPy array in pinned host memory -> libqlb200's host-pointer entry point: H2D, factor, D2H).
`--impl reference` times the routine the reference itself runs for this path (LAPACK dgeqlf from the
SciPy OpenBLAS, oracle/lapack.py; Julia is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BIG = 16384
BATCH, NB = 262144, 32
METRIC = "ql! GFLOP/s (4/3*n^3) FP64 at n=16384"


def flops_ql(n):
    return 4.0 / 3.0 * n ** 3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(n):
    """`config` of the bench line -- the SAME dict for both arms (the reference arm runs this workload with LAPACK dgeqlf on the host
    cores; where its input lives and how it is threaded is stated in its cpu_baseline.sample)."""
    return {"workload": f"ql! Float64 n={n} square (configs[1]); copy + in-place blocked QL per step",
            "parallelism": "replicas only (a single matrix runs on 1 GPU); batched work shards by matrix index",
            "l2": "inputs (2 GiB/matrix) are larger than the 126 MB L2", "nb": "auto (192 at this size)",
            "lookahead": "panel j+1 on a 16-SM green-context partition while the 132-SM partition applies panel j",
            "residency": "GPU arm: input resident in HBM for `value`, in pinned host memory for `e2e`; reference arm: host memory"}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: LAPACK dgeqlf (src/ql.jl:72) on the host cores."""
    if rank != 0:
        return
    import numpy as np
    from oracle import lapack

    cores = os.cpu_count() or 1
    lapack.set_num_threads(cores)
    n = args.n                                  # the SAME configuration as the GPU arm: n = 16384 (about 30 s per step on 16 cores)
    rng = np.random.default_rng(1234321)
    A0 = np.asfortranarray(rng.standard_normal((n, n)))
    A = np.empty_like(A0)
    lapack.geqlf(np.asfortranarray(A0[:2048, :2048]).copy(order="F"), inplace=True)      # warm the thread pool (a full warm-up step costs 30 s)
    ts = []
    budget = float(os.environ.get("QLB200_REF_BUDGET_S", "600"))      # the whole arm must end within minutes: stop taking steps past this
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        A[...] = A0
        t0 = time.perf_counter()
        lapack.geqlf(A, inplace=True)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin + ts[-1] > budget:
            break
    t = sum(ts) / len(ts)
    val = flops_ql(n) / t / 1e9
    sample = (f"dgeqlf (scipy-openblas {lapack.get_num_threads()} threads; the routine ql! calls, src/ql.jl:72) on randn {n}x{n}, "
              f"{len(ts)} step(s) of {t:.1f} s; 4/3*n^3 flops / time")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "steps_timed": len(ts), "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n),
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline():
    import numpy as np
    from oracle import lapack

    cores = os.cpu_count() or 1
    lapack.set_num_threads(cores)
    n = N_BIG if cores >= 8 else 8192           # the workload itself (one dgeqlf of 16384 x 16384 is ~30 s on 16 cores)
    A = np.asfortranarray(np.random.default_rng(7).standard_normal((n, n)))
    lapack.geqlf(A[:1024, :1024].copy(order="F"))
    t0 = time.perf_counter()
    lapack.geqlf(A, inplace=True)
    t = time.perf_counter() - t0
    out = {"value": flops_ql(n) / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
           "sample": f"one LAPACK dgeqlf (scipy-openblas, {lapack.get_num_threads()} threads; the routine ql! calls, src/ql.jl:72) "
                     f"on randn {n}x{n}, {t:.2f} s"}
    # batched reference-equivalent loop `for i: ql!(view(A,:,:,i))` on one core
    lapack.set_num_threads(1)
    nb = 16384
    A3 = np.random.default_rng(8).standard_normal((nb, NB, NB))
    t0 = time.perf_counter()
    lapack.geqlf_loop(A3)
    tb = time.perf_counter() - t0
    lapack.set_num_threads(cores)
    out["batched"] = {"value": nb / tb, "unit": "matrices/s", "cores": 1,
                      "sample": f"loop of dgeqlf over {nb} 32x32 matrices, 1 thread, {tb:.2f} s"}
    # configs[0]: ql(randn(1024,1024)) + F \\ b the way the reference computes it: dgeqlf (all threads) + the scalar loops of
    # src/ql.jl:350-377,440-447,467 (single thread, C restatement)
    from oracle import cport

    A1 = np.asfortranarray(np.random.default_rng(9).standard_normal((1024, 1024)))
    b1 = np.asfortranarray(np.random.default_rng(10).standard_normal((1024, 1)))
    t0 = time.perf_counter()
    F1, tau1 = lapack.geqlf(A1)
    t1 = time.perf_counter()
    cport.ldiv(F1, tau1, b1)
    t2 = time.perf_counter()
    out["ql_1024_plus_solve"] = {"ql_ms": (t1 - t0) * 1e3, "solve_ms": (t2 - t1) * 1e3, "cores": cores,
                                 "sample": "dgeqlf 1024x1024 (all threads) + the reference's scalar lmul!/ldiv! loops (1 thread)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_BIG, help="(debug) matrix size; the benchmark contract is the default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    q = load_package()
    h = q.Handle(local)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    h.set_stream(st.cuda_stream)
    n = args.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- measured FP64 GEMM peak on this GPU (cuBLAS via torch, yardstick for the roofline only) ----
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    peak_tflops = 2 * 8192 ** 3 / best / 1e12
    del a, b, c

    # ---- the workload: A0 resident in HBM, every step = copy + ql! ----
    g = torch.Generator(device="cuda").manual_seed(1234321 + rank)
    A0 = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T        # column-major n x n
    A = torch.empty((n, n), dtype=torch.float64, device="cuda").T
    tau = torch.empty(n, dtype=torch.float64, device="cuda")

    def step():
        A.T.copy_(A0.T)                                       # contiguous 2 GiB device copy (ql = copy + ql!)
        h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_total = e0.elapsed_time(e1) * 1e-3
    launches = h.launches - l0
    # roofline leg (outside the timed region, so that the event pairs do not perturb `value`): the dominant kernel's
    # launches of two more steps are bracketed by CUDA events on the launching stream
    h.set_option("profile", 1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = 2
    p0.record()
    for _ in range(prof_steps):
        step()
    p1.record()
    torch.cuda.synchronize()
    t_prof = p0.elapsed_time(p1) * 1e-3
    gemm_ms, gemm_flops, gemm_launches = h.profile_collect()
    h.set_option("profile", 0)
    tt = torch.tensor([t_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    value = world * args.steps * flops_ql(n) / t_max / 1e9

    # parity spot check on the timed output (size-independent property): || Q'Q - I || on a few columns is
    # covered by tests/; here only finiteness so that a broken run cannot report a number
    if not bool(torch.isfinite(tau).all().item()):
        raise SystemExit("non-finite tau in the benchmark output")

    # ---- e2e through the host API: pinned NumPy array in, factors + tau out ----
    hp = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    src = A0.T.contiguous().cpu()
    Ah = hp.numpy().T                                        # column-major view of the pinned buffer
    e2e_steps, e2e_warm = max(1, min(args.steps, 3)), 3     # (call 1 plain launches, call 2 captures the schedule, then replays)
    h.reset_stream()
    te = 0.0
    for i in range(e2e_steps + e2e_warm):
        hp.copy_(src)
        barrier()
        t0 = time.perf_counter()
        F = q.ql_(Ah, handle=h)                              # H2D + factor + D2H inside
        dt = time.perf_counter() - t0
        if i >= e2e_warm:
            te += dt
    tt = torch.tensor([te / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * flops_ql(n) / float(tt.item()) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": n * n * 8,
           "d2h_bytes_per_step": n * n * 8 + n * 8, "steps": e2e_steps, "warmup": e2e_warm,
           "schedule": "column chunks join the look-ahead ql! (host_split=2), replayed as one CUDA graph from the pinned matrix"}
    h.set_stream(st.cuda_stream)
    del hp, src, A0, A

    # ---- batched 32x32 Float64 QL sharded by matrix index (no collective) ----
    lo, hi = q.shard.shard_range(BATCH, rank, world)
    mine = hi - lo
    gB = torch.Generator(device="cuda").manual_seed(1234321)
    B0 = torch.randn((mine, NB, NB), dtype=torch.float64, device="cuda", generator=gB)
    Bw = torch.empty_like(B0)
    tb = torch.empty((mine, NB), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > L2, written between iterations
    times = []
    for i in range(args.warmup + max(args.steps, 5)):
        Bw.copy_(B0)
        flush.fill_(i & 255)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        q.ql_batched_(Bw, tb, handle=h)
        b1.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            times.append(b0.elapsed_time(b1) * 1e-3)
    tbm = torch.tensor([sorted(times)[len(times) // 2]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tbm, op=dist.ReduceOp.MAX)
    tbat = float(tbm.item())
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    bytes_per = (2 * NB * NB + NB) * 8
    batched = {"metric": "batched 32x32 Float64 QL matrices/s", "value": BATCH / tbat, "unit": "matrices/s",
               "total_matrices": BATCH, "per_rank": mine, "scaling": "strong", "ms": tbat * 1e3,
               "roofline": {"bound": "hbm", "achieved": mine * bytes_per / tbat / 1e9, "peak": hbm_peak or 6650.0,
                            "unit": "GB/s", "frac": mine * bytes_per / tbat / 1e9 / (hbm_peak or 6650.0),
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6.65 TB/s",
                            "bytes_per_matrix": bytes_per, "flush": "256 MiB written between iterations"}}

    # The same workload through HOST pointers (the call a Julia user makes on Arrays): pinned host arrays, the library's
    # upload | compute | download chunk pipeline inside the timed region (PCIe-bound: 2 x 2.1 GB per step at N = 1).
    hA = torch.empty((mine, NB, NB), dtype=torch.float64).pin_memory()
    hT = torch.empty((mine, NB), dtype=torch.float64).pin_memory()
    hsrc = B0.cpu()
    he = []
    for i in range(3):
        hA.copy_(hsrc)
        barrier()
        t0 = time.perf_counter()
        q.ql_batched_(hA.numpy(), hT.numpy(), handle=h)
        he.append(time.perf_counter() - t0)
    hem = torch.tensor([min(he[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hem, op=dist.ReduceOp.MAX)
    batched["e2e"] = {"value": BATCH / float(hem.item()), "unit": "matrices/s", "ms": float(hem.item()) * 1e3,
                      "h2d_bytes_per_step": mine * NB * NB * 8, "d2h_bytes_per_step": mine * (NB * NB + NB) * 8,
                      "timing": "host wall clock around the blocking host-pointer call, max over ranks"}
    del hA, hT, hsrc

    # The same kernel with the FULL 262144-matrix batch on every rank (weak scaling): the strong-scaling line above shrinks
    # the per-GPU shard to 0.2 ms at 8 GPUs, where the fill/drain of the two persistent kernels is a visible share.
    if world > 1:
        W0 = torch.randn((BATCH, NB, NB), dtype=torch.float64, device="cuda", generator=gB)
        Ww = torch.empty_like(W0)
        tw = torch.empty((BATCH, NB), dtype=torch.float64, device="cuda")
        wt = []
        for i in range(args.warmup + max(args.steps, 5)):
            Ww.copy_(W0)
            flush.fill_(i & 255)
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            q.ql_batched_(Ww, tw, handle=h)
            b1.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                wt.append(b0.elapsed_time(b1) * 1e-3)
        twm = torch.tensor([sorted(wt)[len(wt) // 2]], dtype=torch.float64, device="cuda")
        dist.all_reduce(twm, op=dist.ReduceOp.MAX)
        tweak = float(twm.item())
        batched["weak"] = {"value": world * BATCH / tweak, "unit": "matrices/s", "per_rank": BATCH, "total_matrices": world * BATCH,
                           "ms": tweak * 1e3, "hbm_frac_per_gpu": BATCH * bytes_per / tweak / 1e9 / (hbm_peak or 6650.0)}
        del W0, Ww, tw

    # ---- the other BASELINE.json configs (secondary lines; same timing hygiene: events, flush between iterations) ----
    def timed_dev(fn, reset, iters=5):
        ts = []
        for i in range(iters + 2):
            reset()
            flush.fill_(i & 255)
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            fn()
            b1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(b0.elapsed_time(b1) * 1e-3)
        tm = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    hb = hbm_peak or 6650.0
    extras = {}
    # configs[2] second half: lmul!(Q', B) with B 32 x 8 per matrix, on the factors just computed
    RH = torch.randn((mine, 8, NB), dtype=torch.float64, device="cuda", generator=gB)
    RHw = torch.empty_like(RH)
    t_l = timed_dev(lambda: q.lmul_batched_(Bw, tb, RHw, adjoint=True, handle=h), lambda: RHw.copy_(RH))
    by_l = (NB * NB + NB + 2 * NB * 8) * 8
    extras["batched_32x32_f64_lmul_QtB_nrhs8"] = {"value": BATCH / t_l, "unit": "matrices/s", "ms": t_l * 1e3,
                                                  "hbm_frac": mine * by_l / t_l / 1e9 / hb, "bytes_per_matrix": by_l}
    # configs[2] as ONE call: ql! + lmul!(Q',B) (qlb200_dgeqlf_apply_batched_dev; factors stay in L2 between the two kernels)
    def _reset_fused():
        Bw.copy_(B0)
        RHw.copy_(RH)
    t_fu = timed_dev(lambda: q.ql_apply_batched_(Bw, RHw, op="T", tau=tb, handle=h), _reset_fused)
    by_fu = (2 * NB * NB + NB + 2 * NB * 8) * 8          # 20 736 B compulsory per matrix
    extras["batched_32x32_f64_ql_plus_lmul_QtB_fused_call"] = {
        "value": BATCH / t_fu, "unit": "matrices/s", "ms": t_fu * 1e3, "hbm_frac_fused_bytes": mine * by_fu / t_fu / 1e9 / hb,
        "bytes_per_matrix": by_fu, "two_calls_ms": (tbat + t_l) * 1e3,
        "note": "same kernels as the two separate calls, one C call; both kernels are bound by their shared-memory / FP64 pipes, "
                "not by HBM, so re-reading the factors from L2 (fuse_chunk) buys nothing (DESIGN.md 4.1)"}
    del RH, RHw, B0, Bw, tb
    # configs[4]: 1048576 x (16 x 16) Float32 QL + ldiv! on 8 right-hand sides
    B5, N5 = 1048576, 16
    lo5, hi5 = q.shard.shard_range(B5, rank, world)
    m5 = hi5 - lo5
    A5 = torch.randn((m5, N5, N5), dtype=torch.float32, device="cuda", generator=gB)
    A5w = torch.empty_like(A5)
    t5 = torch.empty((m5, N5), dtype=torch.float32, device="cuda")
    R5 = torch.randn((m5, 8, N5), dtype=torch.float32, device="cuda", generator=gB)
    R5w = torch.empty_like(R5)
    t_f = timed_dev(lambda: q.ql_batched_(A5w, t5, handle=h), lambda: A5w.copy_(A5))
    t_s = timed_dev(lambda: q.ldiv_batched_(A5w, t5, R5w, handle=h), lambda: R5w.copy_(R5))
    by_f, by_s = (2 * N5 * N5 + N5) * 4, (N5 * N5 + N5 + 2 * N5 * 8) * 4
    extras["batched_16x16_f32_ql_plus_ldiv_nrhs8"] = {
        "value": B5 / (t_f + t_s), "unit": "matrices/s", "ql_ms": t_f * 1e3, "ldiv_ms": t_s * 1e3,
        "ql_matrices_per_s": B5 / t_f, "ql_hbm_frac": m5 * by_f / t_f / 1e9 / hb, "ldiv_hbm_frac": m5 * by_s / t_s / 1e9 / hb,
        "bytes_per_matrix": by_f + by_s, "total_matrices": B5, "per_rank": m5}
    del A5, A5w, t5, R5, R5w
    # configs[3]: rq of a tall-skinny 65536 x 1024 Float64 matrix (replicas only; rank 0 reports)
    if rank == 0:
        mr, nr = 65536, 1024
        Ar = torch.randn((nr, mr), dtype=torch.float64, device="cuda", generator=gB).T        # column-major 65536 x 1024
        Arw = torch.empty((nr, mr), dtype=torch.float64, device="cuda").T
        tr_ = torch.empty(nr, dtype=torch.float64, device="cuda")
        ts = []
        for i in range(4):
            Arw.T.copy_(Ar.T)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            h.call("qlb200_dgerqf_dev", mr, nr, Arw.data_ptr(), Arw.stride(1), tr_.data_ptr())
            b1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(b0.elapsed_time(b1) * 1e-3)
        t_r = min(ts)
        fl_r = 2.0 * nr * nr * (mr - nr / 3.0)
        extras["rq_65536x1024_f64"] = {"value": fl_r / t_r / 1e9, "unit": "GFLOP/s", "ms": t_r * 1e3,
                                        "flops": "2 n^2 (m - n/3), gerqf(A) = geqlf(A')'", "frac_of_dgemm_peak": fl_r / t_r / 1e12 / peak_tflops}
        del Ar, Arw, tr_
        # configs[3], UL half: ul! of a square 16384 x 16384 matrix, and of the tall-skinny 65536 x 1024 one -- for which the
        # reference's loop (src/ul.jl:155-191, k = min(m,n) ... 1 over rows/cols 1:k) factors only the top 1024 x 1024 block and
        # never reads the other rows (degenerate): both lines are stated for what they are
        for (mu_, nu_, tag) in ((16384, 16384, "ul_16384x16384_f64"), (65536, 1024, "ul_65536x1024_f64_degenerate_top_block")):
            Au = torch.randn((nu_, mu_), dtype=torch.float64, device="cuda", generator=gB).T
            Auw = torch.empty((nu_, mu_), dtype=torch.float64, device="cuda").T
            ipv = torch.zeros(min(mu_, nu_), dtype=torch.int64, device="cuda")
            dinf = torch.zeros(1, dtype=torch.int32, device="cuda")
            ts = []
            for i in range(3):
                Auw.T.copy_(Au.T)
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                h.call("qlb200_dulfact_dev", mu_, nu_, Auw.data_ptr(), Auw.stride(1), ipv.data_ptr(), 1, dinf.data_ptr())
                b1.record()
                torch.cuda.synchronize()
                if i:
                    ts.append(b0.elapsed_time(b1) * 1e-3)
            t_u = min(ts)
            ku = min(mu_, nu_)
            fl_u = 2.0 / 3.0 * ku ** 3
            extras[tag] = {"value": fl_u / t_u / 1e9, "unit": "GFLOP/s", "ms": t_u * 1e3, "flops": "2/3 k^3, k = min(m,n)",
                           "frac_of_dgemm_peak": fl_u / t_u / 1e12 / peak_tflops}
            del Au, Auw, ipv, dinf
        # configs[0]: ql(randn(1024,1024)) + F \\ b -- device-resident (copy + ql! + ldiv!) and through the host API
        n1 = 1024
        A1 = torch.randn((n1, n1), dtype=torch.float64, device="cuda", generator=gB).T
        A1w = torch.empty((n1, n1), dtype=torch.float64, device="cuda").T
        b1v = torch.randn((1, n1), dtype=torch.float64, device="cuda", generator=gB).T
        b1w = torch.empty((1, n1), dtype=torch.float64, device="cuda").T
        tq, tsv = [], []
        for i in range(6):
            A1w.T.copy_(A1.T)
            b1w.T.copy_(b1v.T)
            c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            c0.record()
            F1 = q.ql_(A1w, handle=h)
            c1.record()
            q.ldiv_(F1, b1w)
            c2.record()
            torch.cuda.synchronize()
            if i >= 2:
                tq.append(c0.elapsed_time(c1) * 1e-3)
                tsv.append(c1.elapsed_time(c2) * 1e-3)
        res1 = float((A1 @ b1w - b1v).abs().max().item())
        A1h = np.asfortranarray(A1.cpu().numpy())
        b1h = np.asfortranarray(b1v.cpu().numpy())
        th = []
        for i in range(4):
            t0 = time.perf_counter()
            Fh = q.ql(A1h, handle=h)
            xh = Fh.solve(b1h)
            th.append(time.perf_counter() - t0)
        extras["ql_1024_plus_solve_f64"] = {"ql_ms": min(tq) * 1e3, "solve_ms": min(tsv) * 1e3, "value": flops_ql(n1) / min(tq) / 1e9,
                                            "unit": "GFLOP/s (ql! alone, 4/3 n^3)", "host_api_ql_plus_solve_ms": min(th[1:]) * 1e3,
                                            "max_abs_residual": res1,
                                            "note": "latency-bound on a GPU: 1024 dependent reflector steps (SURVEY.md 8(d)); parity anchor, not a throughput target"}
        del A1, A1w, b1v, b1w
        # the other BlasFloat element types of the same hook (src/ql.jl:72): ComplexF64 (complex panel kernels
        # + the FP64 DMMA GEMMs on the real block form), Float32 / ComplexF32 (promoted to the FP64 path); real-equivalent flops
        for (tag, tdt, ne, mult) in (("ql_8192_c128", torch.complex128, 8192, 4.0), ("ql_8192_f32", torch.float32, 8192, 1.0),
                                     ("ql_4096_c64", torch.complex64, 4096, 4.0)):
            Ae = torch.empty_strided((ne, ne), (1, ne), dtype=tdt, device="cuda")
            Ae.copy_(torch.randn((ne, ne), dtype=tdt, device="cuda", generator=gB))
            Aew = torch.empty_strided((ne, ne), (1, ne), dtype=tdt, device="cuda")
            ts = []
            for i in range(3):
                Aew.copy_(Ae)
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                q.ql_(Aew, handle=h)
                b1.record()
                torch.cuda.synchronize()
                if i:
                    ts.append(b0.elapsed_time(b1) * 1e-3)
            t_e = min(ts)
            extras[tag] = {"value": mult * flops_ql(ne) / t_e / 1e9, "unit": "GFLOP/s (real-equivalent: 4/3 n^3, x4 for complex)", "ms": t_e * 1e3,
                           "frac_of_dgemm_peak": mult * flops_ql(ne) / t_e / 1e12 / peak_tflops}
            del Ae, Aew
    # the batched orders other than 8/16/32: n = 24 (zero-padded 32 x 32 register tile), n = 48, 64 (shared-memory resident
    # matrix, panel in registers, DMMA compact-WY update: batched_wy.cuh), Float64, factor only.  n = 64 is bound by the FP64
    # pipe, not HBM (350 kflop per 66 KB): fp64_frac = useful flops / the DFMA-DMMA datapath peak of 37.2 TFLOP/s
    for ng in (24, 48, 64):
        bg = 65536
        log_, hig_ = q.shard.shard_range(bg, rank, world)
        mg = hig_ - log_
        Ag = torch.randn((mg, ng, ng), dtype=torch.float64, device="cuda", generator=gB)
        Agw = torch.empty_like(Ag)
        tg = torch.empty((mg, ng), dtype=torch.float64, device="cuda")
        t_g = timed_dev(lambda: q.ql_batched_(Agw, tg, handle=h), lambda: Agw.copy_(Ag), iters=3)
        by_g = (2 * ng * ng + ng) * 8
        extras[f"batched_{ng}x{ng}_f64_ql"] = {"value": bg / t_g, "unit": "matrices/s", "ms": t_g * 1e3, "total_matrices": bg,
                                               "hbm_frac": mg * by_g / t_g / 1e9 / hb, "bytes_per_matrix": by_g,
                                               "fp64_frac": world * mg * (4.0 / 3.0 * ng ** 3) / t_g / 37.2e12 / world}
        del Ag, Agw, tg

    traffic = {}
    try:
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        traffic = json.load(open(tp if os.path.exists(tp) else os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        pass
    tb_ = traffic.get("batched_geqlf_n32_f64")
    if tb_:
        batched["roofline"]["traffic"] = tb_["bytes"] / tb_["matrices"] * mine          # ncu DRAM bytes, scaled to this launch
        batched["roofline"]["traffic_source"] = tb_["source"]
    if rank == 0:
        gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        tg_ = traffic.get("gemm_trailing_update")
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(n),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": (gemm_tflops / peak_tflops) if gemm_tflops else None,
                         "traffic": (tg_["CtV_bytes"] + tg_["rankk_bytes"]) if tg_ else None,
                         "traffic_note": ("ncu DRAM bytes of the two GEMM launches of the FIRST panel (largest shape); algorithmic "
                                          f"{tg_['algorithmic_CtV_bytes'] + tg_['algorithmic_rankk_bytes']} B; " + tg_["source"]) if tg_ else None,
                         "kernel": "qlb::gemm_kernel (FP64 DMMA trailing update: W = C'V and C -= V W2')",
                         "launches_timed": gemm_launches, "share_of_step": gemm_ms * 1e-3 / t_prof,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                        "B200 datasheet FP64 tensor = 40 TFLOP/s)",
                         "whole_ql_frac_of_peak": value / 1e3 / world / peak_tflops,
                         "frac_of_partition_peak": (gemm_tflops / peak_tflops * 148.0 / 132.0) if gemm_tflops else None,
                         "partition_note": "the trailing-update GEMMs run on the 132-SM green-context partition (16 SMs factor the next "
                                           "panel meanwhile); `frac` divides by the 148-SM cuBLAS peak, frac_of_partition_peak by 132/148 of it"},
            "batched": batched, "other_configs": extras,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

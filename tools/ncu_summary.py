"""Summarise an ncu report (read here, no GPU needed): per captured kernel the launch geometry, duration,
DRAM traffic, pipe utilisation, and the warp-stall breakdown from the source page.
usage: ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xxx.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
base = ["ncu", "-i", rep, "--csv"]
sel = ["--kernel-name", "regex:" + kern] if kern else []
rows = list(csv.reader(subprocess.run(base + ["--page", "raw"] + sel, capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
print(f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none; serialised, cold-cache launches)")
for r in rows[2:]:
    print()
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:82s} {r[i]} {rows[1][i]}")
srows = list(csv.reader(subprocess.run(base + ["--page", "source"] + sel, capture_output=True, text=True).stdout.splitlines()))
# the source page concatenates kernels: split on header rows
blocks, cur, name = [], None, None
for r in srows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    elif r and r[0] == "Address":
        cur = {"name": name, "hdr": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)


def I(x):
    try:
        return int(x)
    except Exception:
        return 0


seen = set()
for b in blocks:
    key = (b["name"], len(b["rows"]), sum(len(r) for r in b["rows"][:50]), b["rows"][0][0] if b["rows"] else "")
    if key in seen:
        continue
    seen.add(key)
    h = b["hdr"]
    idx = {}
    for i, x in enumerate(h):
        idx.setdefault(x, i)
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot, samp, cnt = collections.Counter(), collections.Counter(), collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for r in b["rows"]:
        if len(r) < len(h):
            continue
        src = r[idx["Source"]].split()
        op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
        cnt[op] += I(r[idx["Instructions Executed"]])
        samp[op] += I(r[idx["# Samples"]])
        for s in stalls:
            v = I(r[idx[s]])
            tot[s] += v
            per[op][s] += v
    T = sum(tot.values()) or 1
    print(f"\n## source page: {b['name'][:120]}")
    print(f"static SASS instructions: {len(b['rows'])} ({len(b['rows']) * 16 / 1024:.1f} KB); warp-state samples: {T}")
    print("stall breakdown: " + "  ".join(f"{s[6:]}={100 * v / T:.1f}%" for s, v in tot.most_common(10)))
    print("top opcodes by samples (warp-instructions executed, share of samples, main stall reasons):")
    for op, v in samp.most_common(12):
        print(f"  {op:18s} exec={cnt[op]:>12d}  {100 * v / T:5.1f}%  " + " ".join(f"{s[6:]}={x}" for s, x in per[op].most_common(3)))

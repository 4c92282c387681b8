#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, --set full --import-source on) into a small text
file for profiles/: duration, pipe utilisation, occupancy, DRAM bytes, executed-instruction mix
and warp-stall breakdown from the source page.   usage: ncu_summary.py report.ncu-rep > out.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    print(f"# {rep}")
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:72s} {vals[i]:>18s} {units[i]}")
    src = page(rep, "source")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = src[2:]
    tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(int(r[ix["# Samples"]]) for r in data) or 1
    ops = collections.Counter()
    stalls = collections.Counter()
    scols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
        for s in scols:
            stalls[s] += int(r[ix[s]] or 0)
    print(f"\nexecuted warp-instructions: {tot_inst}")
    print("instruction mix (share of executed warp-instructions):")
    for o, c in ops.most_common(16):
        print(f"   {o:10s} {c / tot_inst:6.3f}")
    print("warp-stall sampling (share of samples):")
    for s, c in stalls.most_common(8):
        print(f"   {s:26s} {c / tot_samp:6.3f}")
    # hot loops: maximal runs of consecutive SASS lines with the same execution count
    regions, cur = [], None
    for i, r in enumerate(data):
        k = int(r[ix["Instructions Executed"]])
        if cur is not None and cur["count"] == k:
            cur["end"] = i
        else:
            cur = dict(count=k, start=i, end=i, samples=0, st=collections.Counter())
            regions.append(cur)
        cur["samples"] += int(r[ix["# Samples"]])
        for c in scols:
            cur["st"][c] += int(r[ix[c]] or 0)
    regions = [g for g in regions if g["count"] > 0 and g["end"] - g["start"] >= 40]
    regions.sort(key=lambda g: -g["count"] * (g["end"] - g["start"] + 1))
    print("\nhot loops (runs of SASS lines with one execution count; share of executed warp-instructions):")
    for g in regions[:6]:
        n = g["end"] - g["start"] + 1
        tot = sum(g["st"].values()) or 1
        top = ", ".join(f"{k[6:]} {v / tot:.2f}" for k, v in g["st"].most_common(4))
        print(f"   {n:4d} instructions x {g['count']:9d} executions = {n * g['count'] / tot_inst:5.3f}; "
              f"samples {g['samples'] / tot_samp:5.3f} of all ({top}); first: {data[g['start']][ix['Source']].strip()[:40]}")


if __name__ == "__main__":
    main()

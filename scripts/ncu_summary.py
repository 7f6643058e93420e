#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep, captured on the B200 box with
`ncu --set full --clock-control none --import-source on`) into a small tracked text file:
per kernel launch the duration, DRAM traffic, pipe utilisation, occupancy, top stall reasons and
the hottest SASS instructions.   usage: ncu_summary.py REPORT.ncu-rep OUT.md [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe active % (elapsed)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per issued inst (cycles)"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    lines = [f"# {title}", "", f"source report: `{rep}` (ncu --set full --clock-control none; per-launch values; "
             "times under the profiler are cold-cache and serialised)", ""]
    for r in rows:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## {d['Kernel Name'][:110]}")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k, name in KEYS:
            if k in d and d[k] != "":
                lines.append(f"| {name} (`{k}`) | {d[k]} {u.get(k, '')} |")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v),
                        reverse=True)[:6]
        lines.append("| top stall reasons (warps per issue) | " + ", ".join(
            f"{k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, k in stalls) + " |")
        lines.append("")
    # hottest instructions per kernel from the source page
    try:
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
        heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
        for n, hi in enumerate(heads):
            end = heads[n + 1] - 1 if n + 1 < len(heads) else len(src)
            h = src[hi]
            ix = {k: i for i, k in enumerate(h)}
            data = [r for r in src[hi + 1:end] if len(r) == len(h) and r[0] != "Address"]
            tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
            kname = src[hi - 1][1][:100] if hi > 0 and len(src[hi - 1]) > 1 else f"kernel {n}"
            lines.append(f"### hottest SASS, launch {n}: {kname}")
            lines.append("")
            lines.append("| # | instruction | samples % | executed (warp) | top stall |")
            lines.append("|---|---|---|---|---|")
            stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
            for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]]))[:14]:
                s = int(r[ix["# Samples"]])
                top = max(stall_cols, key=lambda k: int(r[ix[k]] or 0))
                lines.append(f"| {i} | `{r[ix['Source']].strip()[:70]}` | {100.0 * s / tot:.1f} | {r[ix['Instructions Executed']]} | {top[6:]} |")
            lines.append("")
    except Exception as exc:  # source page missing
        lines.append(f"(no source page: {exc})")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

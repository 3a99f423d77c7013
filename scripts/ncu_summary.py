#!/usr/bin/env python
"""Summarise an .ncu-rep (read offline with `ncu -i`) into profiles/<name>.md:
headline metrics, pipe utilisation, stall reasons and the hottest source lines."""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], ""]
    for d in rows[2:]:
        lines += ["## %s" % d[hdr.index("Kernel Name")], "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                lines.append("| %s | %s | %s |" % (w, d[hdr.index(w)], units[hdr.index(w)]))
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    agg, samp, thr, stalls = (collections.Counter() for _ in range(4))
    cur, h = None, None
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            h = r
            continue
        if h is None or len(r) < len(h):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        k = (cur, ln, r[1].strip()[:100])
        try:
            agg[k] += int(r[h.index("Instructions Executed")] or 0)
            samp[k] += int(r[h.index("# Samples")] or 0)
            thr[k] += int(r[h.index("Thread Instructions Executed")] or 0)
        except ValueError:
            pass
        for i, name in enumerate(h):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    stalls[name] += int(r[i] or 0)
                except ValueError:
                    pass
    tot, tots, st = sum(agg.values()) or 1, sum(samp.values()) or 1, sum(stalls.values()) or 1
    lines += ["## warp-stall sampling (all launches in the report)", "",
              ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / st) for k, v in stalls.most_common(10)), "",
              "## hottest source lines (share of executed warp instructions; active lanes per instruction)", "",
              "| inst % | samples % | lanes | where | source |", "|---|---|---|---|---|"]
    for k, v in agg.most_common(25):
        lines.append("| %.1f | %.1f | %.1f | %s:%d | `%s` |" % (100 * v / tot, 100 * samp[k] / tots, thr[k] / max(v, 1),
                                                              k[0], k[1], k[2].replace("|", "\\|")))
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (read offline with `ncu -i`) into profiles/<name>.md: headline metrics, the dynamic SASS
opcode histogram, warp-stall sampling (with and without the barrier-polling instructions of waiting warps, which
otherwise drown everything else) and the hottest source lines."""
import collections
import csv
import io
import re
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
POLL = re.compile(r"SYNCS\.PHASECHK|NANOSLEEP")


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


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

    # ---- SASS page: opcode histogram + stall sampling --------------------------------------------------------------
    sass = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h = next((r for r in sass if "Source" in r and "# Samples" in r), None)
    if h is not None:
        ix = {n: i for i, n in enumerate(h)}
        ops, stalls_all, stalls_work = collections.Counter(), collections.Counter(), collections.Counter()
        n_all = n_poll = 0
        for r in sass:
            if len(r) < len(h) or r is h:
                continue
            text = r[ix["Source"]].strip()
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", text)
            n = num(r[ix["Instructions Executed"]])
            if m is None or (n == 0 and num(r[ix["# Samples"]]) == 0):
                continue
            ops[m.group(2)] += n
            n_all += n
            poll = bool(POLL.search(text))
            n_poll += n if poll else 0
            for name in h:
                if name.startswith("stall_") and "Not Issued" not in name:
                    v = num(r[ix[name]])
                    stalls_all[name[6:]] += v
                    if not poll:
                        stalls_work[name[6:]] += v
        tot = max(n_all, 1)
        lines += ["## executed SASS instructions by opcode (warp level, all launches in the report)", "",
                  ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(24)), "",
                  "barrier polling (SYNCS.PHASECHK / NANOSLEEP of waiting warps): %.1f %% of the executed instructions" % (
                      100 * n_poll / tot), ""]
        for title, st in (("all instructions", stalls_all), ("without the barrier-polling instructions", stalls_work)):
            s = max(sum(st.values()), 1)
            lines += ["## warp-stall sampling, %s" % title, "",
                      ", ".join("%s %.1f%%" % (k, 100 * v / s) for k, v in st.most_common(10)), ""]

    # ---- hottest source lines ----------------------------------------------------------------------------------------
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    agg, samp, thr = (collections.Counter() for _ in range(3))
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
        agg[k] += num(r[h.index("Instructions Executed")])
        samp[k] += num(r[h.index("# Samples")])
        thr[k] += num(r[h.index("Thread Instructions Executed")])
    tot, tots = sum(agg.values()) or 1, sum(samp.values()) or 1
    lines += ["## hottest source lines (share of executed warp instructions; active lanes per instruction)", "",
              "| inst % | samples % | lanes | where | source |", "|---|---|---|---|---|"]
    for k, v in agg.most_common(25):
        lines.append("| %.1f | %.1f | %.1f | %s:%d | `%s` |" % (100 * v / tot, 100 * samp[k] / tots, thr[k] / max(v, 1),
                                                              k[0], k[1], k[2].replace("|", "\\|")))
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches <launches.csv>          per-kernel totals and shares of a launch list
    python tools/summarize_ncu.py raw <report.ncu-rep> [regex]     headline metrics of each captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s n=%5d total=%11.1f us avg=%9.1f us share=%5.1f%%" % (k, n, t, t / n, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        print("## " + name)
        for k in KEYS:
            if k in hdr:
                print("  %-80s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for i, h in enumerate(hdr):
            if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  top stalls (warps per issue): " + ", ".join("%s=%.2f" % (h, v) for v, h in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
